(timeout 900 python -m pytest tests/test_gpu_parity.py -k "bluestein or arbitrary or non_power" -x -q 2>&1 | tail -15) > gpurun_out/bluestein_tests.log 2>&1
cat gpurun_out/bluestein_tests.log
timeout 600 python tools/default_sweep.py --gb 4 --nffts 100,1000,3000,5000,8000 > gpurun_out/bluestein_sweep_4gb.log 2>&1
cat gpurun_out/bluestein_sweep_4gb.log
timeout 600 python tools/default_sweep.py --gb 4 --nffts 100,1000,3000,5000,8000 --variant bluestein_r2 > gpurun_out/bluestein_r2_sweep_4gb.log 2>&1
cat gpurun_out/bluestein_r2_sweep_4gb.log
