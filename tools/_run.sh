mkdir -p gpurun_out
for mb in 128 256 512 1024 2048; do echo "slot $mb MB"; timeout 300 python tools/kernel_sweep.py --gb 4 --reps 5 --big --slot-mb $mb 2>&1 | grep nfft; done | tee gpurun_out/sweep_pipe2.log
