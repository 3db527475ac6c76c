mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q -k "gui_maximum or median" 2>&1 | tail -5)
timeout 600 python bench.py --no-cpu --no-raw --no-e2e > gpurun_out/bench7.json 2> gpurun_out/bench7.err; echo rc=$?; python -c "
import json; d=json.load(open('gpurun_out/bench7.json')); print(d['value'], d['roofline']['frac'], d['gpu_launches'], d['clocks'])"; tail -3 gpurun_out/bench7.err
