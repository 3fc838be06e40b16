# End-of-round evidence: parity suite, A/B of the newest switch, bench lines for every workload, ncu launch list, timeline.
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -2 | tee gpurun_out/final_pytest.txt
best=15; bestv=1e9
for o in 15 7; do
  v=$(GPS_B200_OPT=$o timeout 100 python bench.py --steps 30 --warmup 5 2>/dev/null | tail -1 | python -c "import sys,json; print(json.loads(sys.stdin.read())['ms_per_step'])")
  echo "OPT=$o ms_per_step=$v" | tee -a gpurun_out/final_ab.txt
  if python -c "import sys; sys.exit(0 if float('$v') < float('$bestv') else 1)"; then best=$o; bestv=$v; fi
done
echo "best OPT=$best" | tee -a gpurun_out/final_ab.txt
export GPS_B200_OPT=$best
timeout 60 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1 | tee gpurun_out/final_smoke.txt
timeout 200 python bench.py 2>/dev/null | tail -1 > gpurun_out/final_bench_pcqm4m-small.json
timeout 200 python bench.py --impl reference 2>/dev/null | tail -1 > gpurun_out/final_bench_reference.json
for w in zinc-gine zinc-gatedgcn pcqm4m-medium-performer code2; do
  timeout 150 python bench.py --workload $w --steps 20 --warmup 3 2>/dev/null | tail -1 > gpurun_out/final_bench_$w.json
done
timeout 150 python bench.py --precision bf16 --steps 20 --warmup 3 2>/dev/null | tail -1 > gpurun_out/final_bench_pcqm4m-small_bf16.json
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 2 --warmup 3 --no-graph > gpurun_out/final_ncu_bench.log 2>&1
timeout 120 python tools/profile_step.py pcqm4m-small > gpurun_out/final_timeline.txt 2>&1
