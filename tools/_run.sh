mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6) > gpurun_out/gpu_tests.log 2>&1
cat gpurun_out/gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 300 python bench.py --no-cpu > gpurun_out/bench_check.json 2>gpurun_out/bench_check.err; cut -c1-330 gpurun_out/bench_check.json
