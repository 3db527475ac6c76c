(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4)
for i in 1 2; do timeout 300 python bench.py --no-cpu --no-raw --no-e2e --steps 20 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms'], d['gpu_launches'])"; done
timeout 600 python tools/default_sweep.py --gb 4 2>&1 | tail -12
