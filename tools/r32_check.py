#!/usr/bin/env python
"""Radix-32 whole-frame kernels (sti_r32.cuh) on the GPU: parity against the float64 oracle on the cases that
exercise the persistent frame pipeline (item switches, split columns, odd frame starts, integer IQ, more items
than CTAs), then device-resident timing next to the kernels they replace.

python tools/r32_check.py [--nffts 16384,32768,65536] [--gb 4] [--skip-time]"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def check(nfft, torch, engine, variant=None):
    from oracle import np_oracle
    from tests.parity import psd_errors
    ok = True
    rng = np.random.default_rng(nfft)
    dev = torch.device("cuda")
    cases = [  # (ncol, nfr, hop, odd starts, dtype)
        (3, 1, nfft, False, "c64"),      # Mode R, fewer items than CTAs
        (5, 4, nfft, True, "c64"),       # Mode A, odd frame starts (16-byte skew)
        (2, 40, nfft, False, "c64"),     # long columns -> split into chunks (partial sums + finalize)
        (700 if nfft <= 16384 else 200, 1, nfft, True, "c64"),  # more items than resident CTAs: item switching every frame
        (40, 3, nfft - nfft // 8, False, "c64"),  # Mode S hop
        (4, 3, nfft, True, "i16"),
        (4, 2, nfft, False, "i8"),
    ]
    for ncol, nfr, hop, odd, dt in cases:
        span = (nfr - 1) * hop + nfft
        n = ncol * span + 16
        if dt == "c64":
            x = ((rng.standard_normal(n) + 1j * rng.standard_normal(n)) * 1e-2).astype(np.complex64)
            xd = torch.from_numpy(x).to(dev)
            xo, scale = x, 1.0
        else:
            amp = 2000 if dt == "i16" else 100
            raw = rng.integers(-amp, amp, size=(n, 2)).astype(np.int16 if dt == "i16" else np.int8)
            xd = torch.from_numpy(raw).to(dev)
            scale = 1.0 / (32768.0 if dt == "i16" else 128.0)
            xo = (raw[:, 0].astype(np.float64) + 1j * raw[:, 1].astype(np.float64)) * scale
        starts = (np.arange(ncol) * span + (np.arange(ncol) % 2 if odd else 0) * (1 if dt != "i8" else 1)).astype(np.int64)
        plan = engine.StiPlan(nfft)
        engine.set_variant(variant)
        try:
            lin, db = plan.run(xd, torch.from_numpy(starts).to(dev), nfr, hop, in_scale=scale, want_lin=True, want_db=True)
            torch.cuda.synchronize()
        finally:
            engine.set_variant(None)
        got = lin.cpu().numpy()[0]
        gdb = db.cpu().numpy()[0]
        pick = sorted(set([0, 1, ncol // 2, ncol - 1]))
        ref = np.stack([np_oracle.column_power(xo[s:], nfft, nfr, hop) for s in starts[pick]])
        e = psd_errors(got[pick].T, ref.T)
        ddb = float(np.abs(gdb[pick] - 10 * np.log10(ref + 1e-15)).max())
        good = e["col"] <= 1e-5 and e["bin_p999"] <= 1e-5 and e["db_max"] <= 1e-3 and ddb <= 1e-3
        # every column against the checked ones' statistics: no column may be garbage (mean power of noise is flat)
        means = got.mean(axis=1)
        good = good and bool(np.all(np.isfinite(got))) and float(means.max() / means.min()) < 1.5
        ok = ok and good
        print(f"  nfft={nfft} ncol={ncol} nfr={nfr} hop={hop} odd={odd} {dt}: {plan.variant} col={e['col']:.2e} p999={e['bin_p999']:.2e} "
              f"max={e['bin_max']:.2e} dB={e['db_max']:.2e}/{ddb:.2e} {'ok' if good else 'FAIL'}", flush=True)
        del plan
    return ok


def bench(nfft, gb, ntime, variant, torch, engine, peak, reps=5):
    dev = torch.device("cuda")
    n = int(gb * 1e9 / 8)
    iq = torch.empty(n + 8, dtype=torch.complex64, device=dev)
    torch.view_as_real(iq).normal_(0.0, 1e-2)
    nint = n // ntime // nfft
    starts = torch.from_numpy(engine.frame_starts(0, n, nfft, nint, ntime).astype(np.int64)).to(dev)
    plan = engine.StiPlan(nfft)
    engine.set_variant(variant)
    out = torch.empty((1, ntime, nfft), dtype=torch.float32, device=dev)
    try:
        for _ in range(2):
            plan.run(iq, starts, nint, nfft, want_lin=False, want_db=True, out_db=out)
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            plan.run(iq, starts, nint, nfft, want_lin=False, want_db=True, out_db=out)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
    finally:
        engine.set_variant(None)
    ms = float(np.median(ts))
    nbytes = 8 * nfft * nint * ntime + 4 * nfft * ntime
    print(f"  nfft={nfft:6d} ntime={ntime} nint={nint:5d} {ms:8.3f} ms {nfft * nint * ntime / ms / 1e6:7.1f} Gs/s "
          f"{100 * nbytes / ms / 1e6 / peak:5.1f}% of {peak:.0f} GB/s  {plan.variant}", flush=True)
    del plan, out, iq


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nffts", default="16384,32768,65536")
    ap.add_argument("--gb", type=float, default=4.0)
    ap.add_argument("--skip-time", action="store_true")
    ap.add_argument("--skip-check", action="store_true")
    ap.add_argument("--old", default="whole,whole,split", help="variant override of the kernels being replaced, per nfft")
    ap.add_argument("--variant", default=None, help="variant override for the kernel under test (e.g. r32 at 8192)")
    args = ap.parse_args()
    import torch
    from pyspectrogram_b200 import engine
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak = json.load(open(pk))["hbm_gbs"] if os.path.exists(pk) else 6650.0
    nffts = [int(v) for v in args.nffts.split(",")]
    olds = args.old.split(",")
    allok = True
    for i, nfft in enumerate(nffts):
        if not args.skip_check:
            print(f"parity nfft={nfft}", flush=True)
            allok = check(nfft, torch, engine, args.variant) and allok
        if not args.skip_time:
            print(f"timing nfft={nfft}", flush=True)
            bench(nfft, args.gb, 1000, args.variant, torch, engine, peak)
            bench(nfft, args.gb, 1000, olds[min(i, len(olds) - 1)], torch, engine, peak)
            nt = int(min(args.gb, 2.0) * 1e9 / 8) // nfft  # one frame per column (Mode R rate)
            bench(nfft, min(args.gb, 2.0), nt, args.variant, torch, engine, peak)
    print("ALL OK" if allok else "FAILURES")
    return 0 if allok else 1


if __name__ == "__main__":
    sys.exit(main())
