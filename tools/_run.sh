mkdir -p gpurun_out
for mb in 16 32 48 96 128; do echo "scratch $mb MB"; timeout 300 python tools/kernel_sweep.py --gb 4 --reps 5 --big --scratch-mb $mb 2>&1 | grep -v "^$"; done > gpurun_out/sweep_big_scratch.log 2>&1
cat gpurun_out/sweep_big_scratch.log
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 60 --csv --log-file gpurun_out/ncu_split16.csv python tools/kernel_sweep.py --gb 1 --reps 1 --big --only split > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/ncu_split16.csv')) if len(r)>10]
hdr=rows[0]; ix={h:i for i,h in enumerate(hdr)}
from collections import OrderedDict
cur=None
for r in rows[1:]:
    print(r[ix['ID']], r[ix['Kernel Name']][:40], r[ix['Metric Name']], r[ix['Metric Value']], r[ix.get('Grid Size','Grid Size')] if 'Grid Size' in ix else '')
PY
