#!/usr/bin/env python
"""Time-median kernel on the image shapes of BASELINE configs 2 / 3 / 4 (device-resident image, dB median out):
python tools/median_bench.py [--generic]   (--generic: the CTA-wide bisection kernel it replaced, for comparison)"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from pyspectrogram_b200 import engine
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak = json.load(open(pk))["hbm_gbs"] if os.path.exists(pk) else 6650.0
    engine.set_force_generic("--generic" in sys.argv)
    plan = engine.StiPlan(256)
    for name, ncol, nfft in (("cfg2", 1000, 4096), ("cfg3", 3600, 16384), ("cfg4", 3600, 65536), ("cfg4/8", 450, 65536),
                             ("cfg1", 100, 1024)):
        img = torch.empty((1, ncol, nfft), dtype=torch.float32, device="cuda")
        img.exponential_(1.0)
        img *= 1e-4
        for _ in range(3):
            plan.median(img, want_lin=False, want_db=True)
        torch.cuda.synchronize()
        ts = []
        reps = 20  # back to back: the launch path (~25 us of Python + ctypes per call) stays off the device timeline
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                plan.median(img, want_lin=False, want_db=True)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) / reps)
        ms = float(np.median(ts))
        nbytes = 4 * ncol * nfft + 4 * nfft
        print(f"{name:7s} ncol={ncol:5d} nfft={nfft:6d} image {nbytes / 1e6:7.1f} MB  {ms * 1e3:8.1f} us  {nbytes / ms / 1e6:7.0f} GB/s "
              f"{100 * nbytes / ms / 1e6 / peak:5.1f}% of {peak:.0f}", flush=True)
    engine.set_force_generic(False)


if __name__ == "__main__":
    main()
