# A/B of the GPS_B200_OPT switches (1 MN-major weight planes, 2 merged attention backward, 4 early edge BN backward)
for o in ${AB_OPTS:-7 0 7 0}; do echo "OPT=$o"; GPS_B200_OPT=$o timeout 100 python bench.py --steps 30 --warmup 5 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['eager']['ms_per_step'], d['e2e']['ms_per_step'])"; done
timeout 120 python tools/profile_step.py pcqm4m-small > gpurun_out/prof_opt7.txt 2>&1
