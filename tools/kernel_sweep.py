#!/usr/bin/env python
"""Time every registered kernel variant on device-resident IQ (tuning aid; GPU box only).

    python tools/kernel_sweep.py [--gb 2] [--reps 5] [--only tma12] [--out gpurun_out/sweep.json]

Workload per variant: one channel of ``gb`` GB complex64, ntime=1000 bins, nint = full coverage
(Mode A), dB image out.  Reports Gsamples/s and the fraction of the measured HBM peak using the
algorithmic bytes of SURVEY.md section 8(d).
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gb", type=float, default=2.0)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--ntime", type=int, default=1000)
    ap.add_argument("--only", default="")
    ap.add_argument("--odd", action="store_true", help="odd (8-byte aligned) frame starts")
    ap.add_argument("--big", action="store_true", help="only the large-nfft split path (8192..65536)")
    ap.add_argument("--raw", default="", choices=["", "int16", "int8"], help="raw integer IQ ingest variants")
    ap.add_argument("--scratch-mb", type=int, default=0, help="split path scratch size")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "sweep.json"))
    args = ap.parse_args()
    import torch
    from pyspectrogram_b200 import engine
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        peak = 6650.0
    dev = torch.device("cuda")
    ebytes = {"": 8, "int16": 4, "int8": 2}[args.raw]
    n = int(args.gb * 1e9 / ebytes)
    if args.raw:
        dt = torch.int16 if args.raw == "int16" else torch.int8
        iq = torch.randint(-100, 100, (n + 8, 2), dtype=dt, device=dev)
    else:
        iq = torch.empty(n + 8, dtype=torch.complex64, device=dev)
        torch.view_as_real(iq).normal_(0.0, 1e-2)
    suffix = {"": "", "int16": "_i16", "int8": "_i8"}[args.raw]
    rows = []
    if args.scratch_mb:
        engine.set_split_scratch(args.scratch_mb << 20)
    todo = engine.variants() + [("generic", 12)]
    if args.big:
        todo = [("split", 13), ("split", 14), ("split", 15), ("split", 16)]
    for name, logn in todo:
        if args.only and args.only not in name:
            continue
        if name not in ("split", "generic") and name.endswith(("_i16", "_i8")) != bool(suffix):
            continue
        if suffix and name not in ("split", "generic") and not name.endswith(suffix):
            continue
        nfft = 1 << logn
        nint = n // args.ntime // nfft
        starts = engine.frame_starts(0, n, nfft, nint, args.ntime).astype(np.int64)
        if args.odd:
            starts |= 1
        ds = torch.from_numpy(starts).to(dev)
        plan = engine.StiPlan(nfft)
        out = torch.empty((1, args.ntime, nfft), dtype=torch.float32, device=dev)
        try:
            if name == "generic":
                engine.set_force_generic(True)
            else:
                engine.set_variant(name)
            for _ in range(2):
                plan.run(iq, ds, nint, nfft, want_lin=False, want_db=True, out_db=out)
            torch.cuda.synchronize()
            ts = []
            for _ in range(args.reps):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                plan.run(iq, ds, nint, nfft, want_lin=False, want_db=True, out_db=out)
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            used = plan.variant
        finally:
            engine.set_variant(None)
            engine.set_force_generic(False)
        ms = float(np.median(ts))
        nbytes = ebytes * nfft * nint * args.ntime + 4 * nfft * args.ntime
        row = {"variant": used, "nfft": nfft, "nint": nint, "ms": ms, "best_ms": float(min(ts)),
               "gsamples_s": nfft * nint * args.ntime / ms / 1e6, "gbs": nbytes / ms / 1e6,
               "frac": nbytes / ms / 1e6 / peak}
        rows.append(row)
        print(f"{used:32s} nfft={nfft:6d} nint={nint:5d} {ms:8.3f} ms  {row['gsamples_s']:7.1f} Gs/s  "
              f"{row['gbs']:7.0f} GB/s  {100 * row['frac']:5.1f}% of {peak:.0f}", flush=True)
        del plan
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(rows, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
