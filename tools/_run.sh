mkdir -p gpurun_out
timeout 600 python tools/accuracy_report.py 2>&1 | tee gpurun_out/accuracy.log
