mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8)
for o in ldg5 ldg6 ldg7; do timeout 300 python tools/kernel_sweep.py --gb 4 --reps 5 --only $o 2>&1 | grep nfft; done | tee gpurun_out/sweep_tiny.log
