mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15)
timeout 300 python tools/kernel_sweep.py --gb 2 --reps 5 --raw int16 2>&1 | grep nfft | tee gpurun_out/sweep_i16.log
timeout 300 python tools/kernel_sweep.py --gb 1 --reps 5 --raw int8 --only tma 2>&1 | grep nfft | tee gpurun_out/sweep_i8.log
