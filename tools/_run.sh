(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6) > gpurun_out/pytest_gpu.log 2>&1
cat gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
tail -c 800 gpurun_out/bench_final.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_final_ref.json 2>/dev/null
tail -c 300 gpurun_out/bench_final_ref.json
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 300 python bench.py --no-cpu --no-raw --no-e2e --steps 2 --warmup 3 > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches_bench_final.csv python bench.py --no-cpu --no-raw --no-e2e --steps 2 --warmup 3 > gpurun_out/ncu_launch.log 2>&1
tail -n 3 gpurun_out/ncu_launch.log
