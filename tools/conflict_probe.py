#!/usr/bin/env python
"""Which accesses of the 4096-point kernel produce the shared-memory bank-conflict wavefronts ncu reports
(68 M of 540 M on cfg2)?  Three launches on the same recording that differ only in the parity of
the frame starts -- np.linspace starts (half of them odd: the bulk copy then starts at the 16-byte boundary below
the frame and the stage is read with an element skew), all even, all odd -- to be run under
  ncu --metrics l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,gpu__time_duration.sum
and, without ncu, timed with CUDA events."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nfft", type=int, default=4096)
    ap.add_argument("--gb", type=float, default=4.0)
    ap.add_argument("--reps", type=int, default=7)
    args = ap.parse_args()
    import torch
    from pyspectrogram_b200 import engine
    dev = torch.device("cuda")
    n = int(args.gb * 1e9 / 8)
    nfft, ntime = args.nfft, 1000
    iq = torch.empty(n + 8, dtype=torch.complex64, device=dev)
    torch.view_as_real(iq).normal_(0.0, 1e-2)
    nint = n // ntime // nfft
    base = engine.frame_starts(0, n, nfft, nint, ntime).astype(np.int64)
    plan = engine.StiPlan(nfft)
    out = torch.empty((1, ntime, nfft), dtype=torch.float32, device=dev)
    for name, st in (("linspace", base), ("even", base & ~1), ("odd", (base & ~1) + 1)):
        st = np.minimum(st, n - nint * nfft)
        starts = torch.from_numpy(st).to(dev)
        ts = []
        for _ in range(args.reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            plan.run(iq, starts, nint, nfft, want_lin=False, want_db=True, out_db=out)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = float(np.median(ts))
        print(f"{name:9s} odd starts {int((st & 1).sum()):4d}/{ntime}  {ms:7.3f} ms  {nfft * nint * ntime / ms / 1e6:7.1f} Gs/s  {plan.variant}", flush=True)


if __name__ == "__main__":
    main()
