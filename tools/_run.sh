(timeout 900 python -m pytest tests/test_gpu_parity.py -k "arbitrary or non_power" -x -q 2>&1 | tail -15) > gpurun_out/bluestein_tests.log 2>&1
cat gpurun_out/bluestein_tests.log
timeout 600 python tools/default_sweep.py --gb 4 --nffts 100,1000,1200,3000,5000,8000,10000,15000 > gpurun_out/mixed_sweep_4gb.log 2>&1
cat gpurun_out/mixed_sweep_4gb.log
