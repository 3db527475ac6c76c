#!/usr/bin/env python
"""Mode R (one frame per column) columns/s, default kernel vs a forced base variant, for the sizes whose default
changed to a plan with the radix-2 pass in registers.

    python tools/mode_r_ab.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pyspectrogram_b200 import engine


def main():
    dev = torch.device("cuda")
    for nfft, ncol, variants in ((512, 400000, (None, "tma9_2x16x16_f1_s1x2_tq")), (8192, 25000, (None, "tma13_16x8x8x8_f1_s2x1_tq")),
                                 (4096, 100000, (None,))):
        iq = torch.empty(nfft * ncol + 8, dtype=torch.complex64, device=dev)
        torch.view_as_real(iq).normal_(0, 1e-2)
        starts = torch.arange(ncol, device=dev, dtype=torch.int64) * nfft
        plan = engine.StiPlan(nfft)
        out = torch.empty((1, ncol, nfft), dtype=torch.float32, device=dev)
        for var in variants:
            engine.set_variant(var)
            try:
                for _ in range(2):
                    plan.run(iq, starts, 1, nfft, want_lin=False, want_db=True, out_db=out)
                torch.cuda.synchronize()
                ts = []
                for _ in range(7):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    plan.run(iq, starts, 1, nfft, want_lin=False, want_db=True, out_db=out)
                    e1.record()
                    torch.cuda.synchronize()
                    ts.append(e0.elapsed_time(e1))
            finally:
                engine.set_variant(None)
            ms = float(np.median(ts))
            print(f"Mode R nfft={nfft:6d} ncol={ncol:7d}: {ms:8.4f} ms {ncol / ms / 1e3:9.1f} Mcols/s {12 * nfft * ncol / ms / 1e6:7.0f} GB/s "
                  f"(8 B in + 4 B out per sample)  {plan.variant}", flush=True)
        del iq, out, plan


if __name__ == "__main__":
    main()
