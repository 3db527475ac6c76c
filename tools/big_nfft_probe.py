#!/usr/bin/env python
"""Large-nfft paths side by side on device-resident IQ (Mode A, full coverage, dB image out):
cluster kernel (sti_cluster.cuh) vs the three-launch split path (and the fused 8192 kernel).
python tools/big_nfft_probe.py [--gb 4] [--ntime 1000] [--variants cluster,split]"""
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
    ap.add_argument("--variants", default="cluster_dsmem,cluster_ldg,split")
    ap.add_argument("--nffts", default="8192,16384,32768,65536")
    ap.add_argument("--reps", type=int, default=7)
    args = ap.parse_args()
    import torch
    from pyspectrogram_b200 import engine
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak = json.load(open(pk))["hbm_gbs"] if os.path.exists(pk) else 6650.0
    dev = torch.device("cuda")
    n = int(args.gb * 1e9 / 8)
    iq = torch.empty(n + 8, dtype=torch.complex64, device=dev)
    torch.view_as_real(iq).normal_(0.0, 1e-2)
    for nfft in [int(v) for v in args.nffts.split(",")]:
        nint = n // args.ntime // nfft
        starts = torch.from_numpy(engine.frame_starts(0, n, nfft, nint, args.ntime).astype(np.int64)).to(dev)
        out = torch.empty((1, args.ntime, nfft), dtype=torch.float32, device=dev)
        ref = None
        for var in args.variants.split(","):
            plan = engine.StiPlan(nfft)
            try:
                engine.set_variant(None if var == "default" else var)
                try:
                    for _ in range(2):
                        plan.run(iq, starts, nint, nfft, want_lin=False, want_db=True, out_db=out)
                except (NotImplementedError, ValueError, RuntimeError) as e:
                    print(f"nfft={nfft:6d} {var:13s} not available: {e}", flush=True)
                    continue
                torch.cuda.synchronize()
                ts = []
                for _ in range(args.reps):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    plan.run(iq, starts, nint, nfft, want_lin=False, want_db=True, out_db=out)
                    e1.record()
                    torch.cuda.synchronize()
                    ts.append(e0.elapsed_time(e1))
            finally:
                engine.set_variant(None)
            ms = float(np.median(ts))
            nbytes = 8 * nfft * nint * args.ntime + 4 * nfft * args.ntime
            dmax = 0.0
            if ref is None:
                ref = out.clone()
            else:
                dmax = float((out - ref).abs().max())
            print(f"nfft={nfft:6d} nint={nint:5d} {var:13s} {ms:8.3f} ms {nfft * nint * args.ntime / ms / 1e6:7.1f} Gs/s "
                  f"{nbytes / ms / 1e6:6.0f} GB/s {100 * nbytes / ms / 1e6 / peak:5.1f}% of {peak:.0f}  {plan.variant}  "
                  f"max|dB diff vs first|={dmax:.2e}", flush=True)
            del plan


if __name__ == "__main__":
    main()
