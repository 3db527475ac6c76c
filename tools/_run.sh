mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6)
timeout 300 python tools/mode_r_probe.py 2>&1 | tee gpurun_out/mode_r_probe4.log
