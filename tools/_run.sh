mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_parity.py -k "16x16x2x16" -x -q 2>&1 | tail -8) > gpurun_out/fuse2_tests.log 2>&1
cat gpurun_out/fuse2_tests.log
timeout 300 python tools/kernel_sweep.py --gb 4 --reps 7 --only tma13 --out gpurun_out/sweep_tma13_4gb.json > gpurun_out/sweep_tma13_4gb.log 2>&1
cat gpurun_out/sweep_tma13_4gb.log
timeout 300 python tools/kernel_sweep.py --gb 12 --reps 7 --only tma13_16x --out gpurun_out/sweep_tma13_12gb.json > gpurun_out/sweep_tma13_12gb.log 2>&1
cat gpurun_out/sweep_tma13_12gb.log
timeout 300 python tools/kernel_sweep.py --gb 4 --reps 7 --only tma13_16x --raw int16 --out gpurun_out/sweep_tma13_i16.json > gpurun_out/sweep_tma13_i16.log 2>&1
cat gpurun_out/sweep_tma13_i16.log
timeout 200 python tools/default_sweep.py --gb 4 --nffts 8192 --variant tma13_16x16x2x16_f1_s2x1_tq > gpurun_out/fuse2_plain.log 2>&1 && \
timeout 400 ncu --set full --clock-control none --import-source on -k regex:sti_fused --launch-skip 2 -c 1 -f -o gpurun_out/fuse2_8192 \
  python tools/default_sweep.py --gb 4 --nffts 8192 --variant tma13_16x16x2x16_f1_s2x1_tq > gpurun_out/fuse2_ncu.log 2>&1
tail -3 gpurun_out/fuse2_ncu.log
