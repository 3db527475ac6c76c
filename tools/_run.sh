mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3)
timeout 300 python bench.py --no-cpu --no-raw --no-e2e --steps 10 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms'])"
for o in _tq ldg9_8x8x8_f1_tp ldg8_16x16_f8; do timeout 300 python tools/kernel_sweep.py --gb 4 --reps 7 --only $o 2>&1 | grep nfft; done | tee gpurun_out/sweep_wfold.log
