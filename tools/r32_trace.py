#!/usr/bin/env python
"""Phase timeline of the radix-32 kernel (tuning aid): PSG_R32_OPT with bit 2 set makes CTA 0 record clock64 at
the phase boundaries of four frames; this prints, per warp, the duration of every phase and when it started
relative to the frame's first event.

PSG_R32_OPT=4 python tools/r32_trace.py [--nfft 16384] [--gb 2]"""
import argparse
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

EV = ["top", "S full", "P0 loaded", "P0 math", "M free", "P0 stored", "barrier", "P1 loaded", "P1 math", "P1 stored", "P2 a", "P2 end"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nfft", type=int, default=16384)
    ap.add_argument("--gb", type=float, default=2.0)
    args = ap.parse_args()
    assert int(os.environ.get("PSG_R32_OPT", "0")) & 4, "set PSG_R32_OPT with bit 2 (4)"
    import torch
    from pyspectrogram_b200 import _lib, engine
    dev = torch.device("cuda")
    n = int(args.gb * 1e9 / 8)
    iq = torch.empty(n + 8, dtype=torch.complex64, device=dev)
    torch.view_as_real(iq).normal_(0.0, 1e-2)
    ntime = 1000
    nint = n // ntime // args.nfft
    starts = torch.from_numpy(engine.frame_starts(0, n, args.nfft, nint, ntime).astype(np.int64)).to(dev)
    plan = engine.StiPlan(args.nfft)
    if args.nfft == 8192:
        engine.set_variant("r32")
    for _ in range(2):
        plan.run(iq, starts, nint, args.nfft, want_lin=False, want_db=True)
    torch.cuda.synchronize()
    lib = _lib.load()
    nfr, nw, nev = 4, 16, len(EV)
    buf = np.zeros(nfr * nw * nev, np.int64)
    got = lib.psg_r32_trace_dump(buf.ctypes.data_as(C.c_void_p), buf.size)
    assert got == buf.size, got
    tr = buf.reshape(nfr, nw, nev)
    print(plan.variant, "OPT", os.environ["PSG_R32_OPT"])
    for f in range(nfr):
        t0 = tr[f, :, 0].min()
        end = tr[f, :, -1].max()
        nxt = tr[f + 1, :, 0].min() if f + 1 < nfr else end
        print(f"frame {f}: first top -> last P2 end {end - t0} clk; next frame's first top at +{nxt - t0}")
        print("  warp " + " ".join(f"{e[:9]:>9s}" for e in EV[1:]))
        for w in range(nw):
            d = np.diff(tr[f, w])
            print(f"  {w:4d} " + " ".join(f"{int(v):9d}" for v in d) + f"   start +{tr[f, w, 0] - t0}")
        print("  mean " + " ".join(f"{int(v):9d}" for v in np.diff(tr[f], axis=1).mean(axis=0)))
    per = np.diff(tr[:, :, 0], axis=0).mean()
    print(f"frame period (top to top): {per:.0f} clk")


if __name__ == "__main__":
    main()
