mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/bench8.json 2> gpurun_out/bench8.err; echo rc=$?; python -c "
import json; d=json.load(open('gpurun_out/bench8.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['e2e_raw_int16']['value'], d['cpu_baseline']['value'], d['clocks'])"; tail -3 gpurun_out/bench8.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke()"
