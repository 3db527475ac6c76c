mkdir -p gpurun_out
# 1. launch list of the bench command (per-launch gpu time; cold-cache, serialised)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_bench.log 2>&1
# 2. one full capture of the dominant kernel (skip warm-up launches)
ncu --set full --clock-control none --import-source on -k regex:sti_fused_kernel -s 3 -c 1 -o gpurun_out/r01_full_tma12_cfg2 -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
