(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6) > gpurun_out/pytest_gpu.log 2>&1
cat gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
tail -c 600 gpurun_out/bench_final.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_final_ref.json 2>/dev/null
timeout 600 python tools/default_sweep.py --gb 12 > gpurun_out/default_sweep_12gb_final.log 2>&1
cat gpurun_out/default_sweep_12gb_final.log
timeout 300 python bench.py --no-cpu --no-raw --no-e2e --steps 2 --warmup 3 > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches_bench_final.csv python bench.py --no-cpu --no-raw --no-e2e --steps 2 --warmup 3 > gpurun_out/ncu_launch.log 2>&1
