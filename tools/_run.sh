mkdir -p gpurun_out
for n in 8 4 2; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2954$n bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err
echo "N=$n rc=$?"; python -c "
import json
for l in open('gpurun_out/bench_n$n.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['n_gpus'], round(d['value']), d['ms_per_step'], round(d['roofline']['frac'],4), d['e2e'].get('value'), d['e2e'].get('ms_per_step'), d['clocks']['sm_mhz'], d['clocks']['reasons'])"; tail -2 gpurun_out/bench_n$n.err | cut -c1-300
done
