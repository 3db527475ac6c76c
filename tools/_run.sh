mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5)
for o in _tq _tp; do timeout 300 python tools/kernel_sweep.py --gb 4 --reps 7 --only $o 2>&1 | grep nfft; done | tee gpurun_out/sweep_tq2.log
