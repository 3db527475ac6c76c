(timeout 400 python -m pytest tests/test_gpu_parity.py -k "whole" -x -q 2>&1 | tail -5) > gpurun_out/whole_tests.log 2>&1
cat gpurun_out/whole_tests.log
grep -q "passed" gpurun_out/whole_tests.log && ! grep -q "failed" gpurun_out/whole_tests.log && timeout 300 python tools/big_nfft_probe.py --gb 12 --nffts 16384,32768,65536 --variants default,whole_r2,whole_r4 > gpurun_out/whole_pfl2_probe_12gb.log 2>&1
cat gpurun_out/whole_pfl2_probe_12gb.log
