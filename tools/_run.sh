mkdir -p gpurun_out
timeout 300 python tools/mode_r_ab.py > gpurun_out/mode_r_ab.log 2>&1; cat gpurun_out/mode_r_ab.log
