mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gpu_parity.py -k "whole_f" -x -q 2>&1 | tail -12) > gpurun_out/whole_f_tests.log 2>&1
cat gpurun_out/whole_f_tests.log
for v in whole whole_f; do
  timeout 120 python tools/default_sweep.py --gb 4 --nffts 16384 --variant $v > gpurun_out/w16_${v}_4gb.log 2>&1; cat gpurun_out/w16_${v}_4gb.log
  timeout 120 python tools/default_sweep.py --gb 12 --nffts 16384 --variant $v > gpurun_out/w16_${v}_12gb.log 2>&1; cat gpurun_out/w16_${v}_12gb.log
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sti_whole16 --launch-skip 2 -c 1 -f -o gpurun_out/whole16 \
  python tools/default_sweep.py --gb 4 --nffts 16384 --variant whole_f > gpurun_out/whole16_ncu.log 2>&1
tail -2 gpurun_out/whole16_ncu.log
