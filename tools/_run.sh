mkdir -p gpurun_out
N=${NGPU:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/dist_configs_probe.py --out gpurun_out/dist_configs_n$N.json > gpurun_out/dist_configs_n$N.log 2>&1
grep '^{' gpurun_out/dist_configs_n$N.log | cut -c1-900
