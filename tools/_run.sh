(timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "streams_in_column or host or golden" 2>&1 | tail -25) > gpurun_out/host_tests.log 2>&1
cat gpurun_out/host_tests.log
for i in 1 2; do timeout 300 python bench.py --no-cpu --steps 10 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e_raw_int16']['value'])"; done
