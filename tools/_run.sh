(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6) > gpurun_out/pytest_gpu.log 2>&1
cat gpurun_out/pytest_gpu.log
timeout 300 python tools/big_nfft_probe.py --gb 4 --variants cluster_ldg,cluster,cluster_dsmem,split,default > gpurun_out/big_nfft_probe_4gb.log 2>&1
timeout 300 python tools/big_nfft_probe.py --gb 12 --variants cluster_ldg,cluster,cluster_dsmem,split,default > gpurun_out/big_nfft_probe_12gb.log 2>&1
cat gpurun_out/big_nfft_probe_4gb.log gpurun_out/big_nfft_probe_12gb.log | cut -c1-110
