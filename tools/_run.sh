mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5)
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/bench3.json 2> gpurun_out/bench3.err; echo rc=$?; cat gpurun_out/bench3.json; tail -3 gpurun_out/bench3.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench3_ref.json 2> gpurun_out/bench3_ref.err; echo rc=$?; cat gpurun_out/bench3_ref.json; tail -3 gpurun_out/bench3_ref.err
