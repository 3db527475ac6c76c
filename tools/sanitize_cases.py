#!/usr/bin/env python
"""Small run of every default kernel for compute-sanitizer (memcheck / racecheck / synccheck).

    compute-sanitizer --tool racecheck python tools/sanitize_cases.py [--nffts 512,8192] [--variant whole_f]
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from pyspectrogram_b200 import engine
    ap = argparse.ArgumentParser()
    ap.add_argument("--nffts", default="64,256,512,1024,2048,4096,8192,16384")
    ap.add_argument("--variant", default=None, help="force a kernel variant / path for the contiguous runs")
    args = ap.parse_args()
    rng = np.random.default_rng(0)
    for nfft in [int(v) for v in args.nffts.split(",")]:
        nfr, ncol = 3, 4
        n = nfft * nfr * ncol + nfft + 7
        x = ((rng.standard_normal(n) + 1j * rng.standard_normal(n)) * 1e-2).astype(np.complex64)
        starts = (np.arange(ncol) * nfft * nfr + np.arange(ncol) % 2).astype(np.int64)  # odd and even starts
        plan = engine.StiPlan(nfft)
        dx, ds = torch.from_numpy(x).cuda(), torch.from_numpy(starts).cuda()
        engine.set_variant(args.variant)
        try:
            lin, db = plan.run(dx, ds, nfr, nfft, want_lin=True, want_db=True)
            used = plan.variant
        finally:
            engine.set_variant(None)
        med, _ = plan.median(lin)
        # strided layout (LDG loader): two interleaved sub-channels
        x2 = np.stack([x, x[::-1]], axis=1).copy()
        lin2, _ = plan.run(torch.from_numpy(x2).cuda(), ds * 2, nfr, nfft, sample_stride=2, sub_stride=1, nsub=2)
        torch.cuda.synchronize()
        a, b = lin.cpu().numpy()[0], lin2.cpu().numpy()[0]
        assert np.allclose(a, b, rtol=1e-4, atol=1e-12), nfft
        print(f"nfft={nfft:6d} {used:36s} ok  sum={float(a.sum()):.6e}", flush=True)


if __name__ == "__main__":
    main()
