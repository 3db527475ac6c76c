(timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "minmax or gather or plot_data" 2>&1 | tail -25) > gpurun_out/n4_tests.log 2>&1
cat gpurun_out/n4_tests.log
