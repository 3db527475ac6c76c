mkdir -p gpurun_out
timeout 300 python tools/strided_probe.py 2>&1 | tee gpurun_out/strided_probe.log
timeout 600 python bench.py --no-cpu > gpurun_out/bench4.json 2> gpurun_out/bench4.err; echo rc=$?; python -c "
import json; d=json.load(open('gpurun_out/bench4.json')); print(d['value'], d['e2e'], d.get('e2e_raw_int16'))"; tail -3 gpurun_out/bench4.err
