#!/usr/bin/env python
"""Time the fused kernel on the reference's (rows, ntime) array layout (sample axis slowest:
sample_stride = ntime) against the frame-contiguous layout of the same data (tuning aid)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pyspectrogram_b200 import engine

def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))

for nfft, ntime, nint in ((1024, 100, 97), (4096, 1000, 32), (1024, 1000, 128)):
    rows = nfft * nint
    d1 = torch.empty((rows, ntime), dtype=torch.complex64, device="cuda")
    torch.view_as_real(d1).normal_(0, 1e-2)
    plan = engine.StiPlan(nfft)
    cols = torch.arange(ntime, dtype=torch.int64, device="cuda")
    t_str = timeit(lambda: plan.run(d1, cols, nint, nfft, sample_stride=ntime))
    v_str = plan.variant
    dt = d1.t().contiguous()
    t_tr = timeit(lambda: d1.t().contiguous())
    starts = cols * rows
    t_con = timeit(lambda: plan.run(dt, starts, nint, nfft))
    ns = rows * ntime
    print(f"nfft={nfft} ntime={ntime} nint={nint}: strided {t_str:.3f} ms ({ns/t_str/1e6:.1f} Gs/s, {v_str}); "
          f"contiguous {t_con:.3f} ms ({ns/t_con/1e6:.1f} Gs/s, {plan.variant}); torch transpose {t_tr:.3f} ms", flush=True)
