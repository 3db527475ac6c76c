#!/usr/bin/env python
"""Large nfft with few frames per column (Mode R and short integrations): whole-frame kernels vs the
cluster kernels / split path, to place the defaults.  python tools/short_int_probe.py [nfft] [variants]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pyspectrogram_b200 import engine


def main():
    dev = torch.device("cuda")
    nfft = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    variants = sys.argv[2].split(",") if len(sys.argv) > 2 else ["whole", "cluster_ldg", "cluster_dsmem"]
    tot = 16000 * 16384 // nfft
    for nint, ncol in ((1, tot), (2, tot // 2), (4, tot // 4), (8, tot // 8), (16, tot // 16), (32, tot // 32)):
        n = nfft * nint * ncol
        iq = torch.empty(n + 8, dtype=torch.complex64, device=dev)
        torch.view_as_real(iq).normal_(0, 1e-2)
        starts = torch.arange(ncol, device=dev, dtype=torch.int64) * nfft * nint
        out = torch.empty((1, ncol, nfft), dtype=torch.float32, device=dev)
        for var in variants:
            plan = engine.StiPlan(nfft)
            try:
                engine.set_variant(var)
                run = lambda: plan.run(iq, starts, nint, nfft, want_lin=False, want_db=True, out_db=out)
                run(); run(); torch.cuda.synchronize()
                ts = []
                for _ in range(7):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(); run(); e1.record(); torch.cuda.synchronize()
                    ts.append(e0.elapsed_time(e1))
            finally:
                engine.set_variant(None)
            ms = float(np.median(ts))
            nbytes = (8 * nint + 4) * nfft * ncol
            print(f"nfft={nfft} nint={nint:3d} ncol={ncol:6d} {var:14s} {ms:8.3f} ms {nfft * nint * ncol / ms / 1e6:7.1f} Gs/s "
                  f"{nbytes / ms / 1e6:6.0f} GB/s  {plan.variant}", flush=True)
        del iq, out


if __name__ == "__main__":
    main()
