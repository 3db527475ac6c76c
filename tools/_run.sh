mkdir -p gpurun_out
: > gpurun_out/r02_mixct_alts2.log
L=1000,1200,1500,2000,2400,3000,3600,4000,4800,5000,6000,8000,10000
PSG_MIXCT_ALT=0 timeout 300 python tools/default_sweep.py --gb 4 --nffts $L >> gpurun_out/r02_mixct_alts2.log 2>&1
for a in 1 2 3 4 5 6; do PSG_MIXCT_ALT=$a timeout 120 python tools/default_sweep.py --gb 4 --nffts 1000 >> gpurun_out/r02_mixct_alts2.log 2>&1; done
for a in 1 2 3; do PSG_MIXCT_ALT=$a timeout 120 python tools/default_sweep.py --gb 4 --nffts 2000 >> gpurun_out/r02_mixct_alts2.log 2>&1; done
for a in 1 2; do PSG_MIXCT_ALT=$a timeout 120 python tools/default_sweep.py --gb 4 --nffts 5000,8000,10000 >> gpurun_out/r02_mixct_alts2.log 2>&1; done
cat gpurun_out/r02_mixct_alts2.log
