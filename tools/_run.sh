mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_parity.py -k "median or minmax or proc_data or drop_in or golden" -x -q 2>&1 | tail -6) > gpurun_out/median_tests.log 2>&1
cat gpurun_out/median_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29511 tools/dist_configs_probe.py --out gpurun_out/dist_configs_n1_newmedian.json > gpurun_out/dist_configs_n1_newmedian.log 2>&1
grep '^{' gpurun_out/dist_configs_n1_newmedian.log | cut -c1-420
timeout 300 python bench.py --no-e2e --no-cpu > gpurun_out/bench_newmedian.json 2>gpurun_out/bench_newmedian.err; cut -c1-330 gpurun_out/bench_newmedian.json
