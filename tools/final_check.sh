# Last GPU call of the round: full parity suite on the committed state, smoke, bench lines, memcheck of the new kernels.
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/fc_pytest.txt; tail -3 gpurun_out/fc_pytest.txt
timeout 60 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1 | tee gpurun_out/fc_smoke.txt
timeout 150 python bench.py 2>/dev/null | tail -1 > gpurun_out/fc_bench_pcqm4m-small.json
timeout 100 python bench.py --workload zinc-gcn --steps 20 --warmup 3 2>/dev/null | tail -1 > gpurun_out/fc_bench_zinc-gcn.json
timeout 170 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_layer_gpu.py -q -x \
  -k "(golden and fp32 and (gcn_transformer_relu or gatedgcn_transformer_relu)) or gcn_self_loops" > gpurun_out/fc_memcheck.txt 2>&1
echo "memcheck rc=$?" | tee -a gpurun_out/fc_memcheck.txt
