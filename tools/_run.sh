mkdir -p gpurun_out
python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-raw > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-raw > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sti_fused_kernel -s 3 -c 1 -o gpurun_out/r01_full_tma12_tq_cfg2_final -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-raw > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
