mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
(timeout 900 python -m pytest tests/test_gpu_parity.py -k "median or every_variant or bit_identical" -x -q 2>&1 | tail -3) > gpurun_out/gpu_tests_subset.log 2>&1; cat gpurun_out/gpu_tests_subset.log
timeout 300 python bench.py --no-e2e --no-cpu > gpurun_out/bench_check.json 2>gpurun_out/bench_check.err; cut -c1-260 gpurun_out/bench_check.json
