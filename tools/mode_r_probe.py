#!/usr/bin/env python
"""Columns/s of the shipped semantics (Mode R: one frame per STI column, drfProc.py:364-403) on
device-resident IQ, plus cfg1 (nfft=1024, 100 columns) through the drop-in call with host arrays."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pyspectrogram_b200 import drfProc as dp
from pyspectrogram_b200 import engine


def main():
    dev = torch.device("cuda")
    for nfft, ncol in ((1024, 100), (1024, 100000), (4096, 100000), (16384, 20000), (65536, 3600)):
        n = nfft * ncol
        iq = torch.empty(n + 8, dtype=torch.complex64, device=dev)
        torch.view_as_real(iq).normal_(0, 1e-2)
        starts = torch.arange(ncol, device=dev, dtype=torch.int64) * nfft
        plan = engine.StiPlan(nfft)
        out = torch.empty((1, ncol, nfft), dtype=torch.float32, device=dev)
        run = lambda: plan.run(iq, starts, 1, nfft, want_lin=False, want_db=True, out_db=out)
        run(); torch.cuda.synchronize()
        ts = []
        for _ in range(7):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); run(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = float(np.median(ts))
        nbytes = 12 * nfft * ncol
        print(f"Mode R nfft={nfft:6d} ncol={ncol:6d}: {ms:8.4f} ms  {ncol / ms / 1e3:9.1f} Mcols/s  "
              f"{nbytes / ms / 1e6:7.0f} GB/s (8 B in + 4 B out per sample)  {plan.variant}", flush=True)
        del iq, out
    # short integrations (a few frames per column): multi-column twin kernels on / off
    from pyspectrogram_b200 import _lib
    for nfft, nint, ncol in ((1024, 4, 50000), (4096, 4, 25000), (4096, 8, 12500), (65536, 4, 1800)):
        n = nfft * nint * ncol
        iq = torch.empty(n + 8, dtype=torch.complex64, device=dev)
        torch.view_as_real(iq).normal_(0, 1e-2)
        starts = torch.arange(ncol, device=dev, dtype=torch.int64) * (nfft * nint)
        plan = engine.StiPlan(nfft)
        out = torch.empty((1, ncol, nfft), dtype=torch.float32, device=dev)
        line = f"Mode A nfft={nfft:6d} nint={nint} ncol={ncol:6d}:"
        for multi in (1, 0):
            _lib.check(_lib.load().psg_debug_set_mode_r_multi(multi))
            run = lambda: plan.run(iq, starts, nint, nfft, want_lin=False, want_db=True, out_db=out)
            run(); torch.cuda.synchronize()
            ts = []
            for _ in range(7):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); run(); e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            ms = float(np.median(ts))
            line += f"  [{plan.variant}] {ms:.4f} ms {n / ms / 1e6:6.1f} Gs/s"
        _lib.check(_lib.load().psg_debug_set_mode_r_multi(1))
        print(line, flush=True)
        del iq, out
    # cfg1 through the drop-in API (host arrays, as the reference is called)
    rng = np.random.default_rng(0)
    d1 = ((rng.standard_normal((1024 * 97, 100)) + 1j * rng.standard_normal((1024 * 97, 100))) * 1e-2).astype(np.complex64)
    for integ in (False, True):
        dp.sti_proc_data_db(d1, 1.0e6, 1024, integrate=integ)
        t0 = time.perf_counter()
        for _ in range(5):
            dp.sti_proc_data_db(d1, 1.0e6, 1024, integrate=integ)
        dt = (time.perf_counter() - t0) / 5
        print(f"cfg1 drop-in sti_proc_data_db(integrate={integ}): {dt * 1e3:.3f} ms per call "
              f"({100 / dt:.0f} cols/s; host array in, dB image + median out)", flush=True)


if __name__ == "__main__":
    main()
