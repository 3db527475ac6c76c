mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -x -q -k "every_variant" 2>&1 | tail -5)
for o in ldg8 ldg9 10_ 11_; do timeout 300 python tools/kernel_sweep.py --gb 4 --reps 5 --only $o 2>&1 | grep nfft; done | tee gpurun_out/sweep_small2.log
