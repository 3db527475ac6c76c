mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5)
timeout 600 python tools/default_sweep.py --gb 12 2>&1 | tee gpurun_out/default_sweep_12gb.log
timeout 600 python tools/default_sweep.py --gb 4 2>&1 | tee gpurun_out/default_sweep_4gb.log
timeout 300 python tools/mode_r_probe.py 2>&1 | tee gpurun_out/mode_r_probe2.log
