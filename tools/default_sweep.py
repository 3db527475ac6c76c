#!/usr/bin/env python
"""Default kernel of every FFT length on device-resident IQ (Mode A, full coverage, dB image out):
the table DESIGN.md quotes.  python tools/default_sweep.py [--gb 4]"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gb", type=float, default=4.0)
    ap.add_argument("--ntime", type=int, default=1000)
    ap.add_argument("--nffts", default="64,256,512,1000,1024,2048,4096,8192,16384,32768,65536")
    ap.add_argument("--variant", default=None, help="force a kernel variant / path (psg_debug_set_variant)")
    args = ap.parse_args()
    import torch
    from pyspectrogram_b200 import engine
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    dev = torch.device("cuda")
    n = int(args.gb * 1e9 / 8)
    iq = torch.empty(n + 8, dtype=torch.complex64, device=dev)
    torch.view_as_real(iq).normal_(0.0, 1e-2)
    for nfft in [int(v) for v in args.nffts.split(",")]:
        nint = n // args.ntime // nfft
        starts = torch.from_numpy(engine.frame_starts(0, n, nfft, nint, args.ntime).astype(np.int64)).to(dev)
        plan = engine.StiPlan(nfft)
        engine.set_variant(args.variant)
        out = torch.empty((1, args.ntime, nfft), dtype=torch.float32, device=dev)
        for _ in range(2):
            plan.run(iq, starts, nint, nfft, want_lin=False, want_db=True, out_db=out)
        torch.cuda.synchronize()
        ts = []
        for _ in range(7):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            plan.run(iq, starts, nint, nfft, want_lin=False, want_db=True, out_db=out)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = float(np.median(ts))
        nbytes = 8 * nfft * nint * args.ntime + 4 * nfft * args.ntime
        print(f"nfft={nfft:6d} nint={nint:6d} {ms:8.3f} ms {nfft * nint * args.ntime / ms / 1e6:7.1f} Gs/s "
              f"{nbytes / ms / 1e6:6.0f} GB/s {100 * nbytes / ms / 1e6 / peak:5.1f}% of {peak:.0f}  {plan.variant}", flush=True)
        del plan, out


if __name__ == "__main__":
    main()
