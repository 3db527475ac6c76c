(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6) > gpurun_out/pytest_gpu.log 2>&1
cat gpurun_out/pytest_gpu.log
timeout 600 python tools/default_sweep.py --gb 12 > gpurun_out/default_sweep_12gb_whole.log 2>&1
cat gpurun_out/default_sweep_12gb_whole.log
