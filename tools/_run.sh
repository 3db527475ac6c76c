mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5)
timeout 300 python bench.py --no-cpu --no-raw --no-e2e --steps 10 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms'], d['gpu_launches'])"
timeout 300 python tools/mode_r_probe.py 2>&1 | tee gpurun_out/mode_r_probe2.log
