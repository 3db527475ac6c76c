mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q -x -k "compile_time_mixed or arbitrary_nfft or non_power_of_two or drop_in_accepts_raw or mode_r_multi_column" 2>&1 | tail -30) > gpurun_out/r02_gpu_tests_mixct.log 2>&1; tail -5 gpurun_out/r02_gpu_tests_mixct.log
L=1000,1200,1500,2000,2400,3000,3600,4000,4800,5000,6000,8000,10000
timeout 600 python tools/default_sweep.py --gb 4 --nffts $L > gpurun_out/r02_mixct_sweep_4GB.log 2>&1; cat gpurun_out/r02_mixct_sweep_4GB.log
timeout 600 python tools/default_sweep.py --gb 4 --nffts $L --variant mixed_rt > gpurun_out/r02_mixed_rt_sweep_4GB.log 2>&1; cat gpurun_out/r02_mixed_rt_sweep_4GB.log
