#!/usr/bin/env python
"""Measured error of the GPU path (the DEFAULT kernel psg_plan_variant reports) against the float64 oracle
(tests/parity.py metrics) per FFT length.  Table 1: -40 dBFS noise + a -20 dBFS tone, Mode R (single frame, the
hardest case) and Mode A.  Table 2: pure noise, Modes R / A / S -- the per-bin criterion of SURVEY.md section
8(c): p99.9 relative error <= 1e-5 and <= 1e-3 dB on EVERY bin."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle import np_oracle
from pyspectrogram_b200 import engine
from tests.parity import psd_errors


def main():
    rng = np.random.default_rng(1)
    print("nfft   mode  variant                                   col(max err/peak)  bin p99.9   dB max (bins within 60 dB of peak)")
    for nfft in (32, 256, 1000, 1024, 4096, 8192, 16384, 32768, 65536):
        for mode, nfr in (("R", 1), ("A", 8)):
            ncol = 6
            n = nfft * nfr * ncol + 64
            x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)) * (1e-2 / np.sqrt(2))
            x += 0.1 * np.exp(2j * np.pi * 0.123 * np.arange(n))
            x = x.astype(np.complex64)
            starts = (np.arange(ncol) * nfft * nfr + np.arange(ncol) % 2).astype(np.int64)
            plan = engine.StiPlan(nfft)
            lin, _ = plan.run(torch.from_numpy(x).cuda(), torch.from_numpy(starts).cuda(), nfr, nfft)
            ref = np.stack([np_oracle.column_power(x[s:], nfft, nfr, nfft) for s in starts])
            e = psd_errors(lin.cpu().numpy()[0].T, ref.T)
            print(f"{nfft:6d} {mode}     {plan.variant:40s}  {e['col']:.2e}          {e['bin_p999']:.2e}    {e['db_max_strong']:.2e}", flush=True)


def noise_table():
    rng = np.random.default_rng(2)
    print()
    print("pure noise (-40 dBFS), every bin counted")
    print("nfft   mode  variant                                   col(max err/peak)  bin p99.9   bin max     dB max (all bins)")
    for nfft in (32, 64, 128, 256, 512, 1000, 1024, 2048, 4096, 8192, 16384, 32768, 65536):
        for mode, nfr, hop in (("R", 1, nfft), ("A", 5, nfft), ("S", 4, nfft - nfft // 8)):
            ncol = 4 if nfft >= 16384 else 12
            span = (nfr - 1) * hop + nfft
            n = ncol * span + 8
            x = ((rng.standard_normal(n) + 1j * rng.standard_normal(n)) * (1e-2 / np.sqrt(2))).astype(np.complex64)
            starts = (np.arange(ncol) * span + np.arange(ncol) % 2).astype(np.int64)
            plan = engine.StiPlan(nfft)
            lin, db = plan.run(torch.from_numpy(x).cuda(), torch.from_numpy(starts).cuda(), nfr, hop, want_lin=True, want_db=True)
            ref = np.stack([np_oracle.column_power(x[s:], nfft, nfr, hop) for s in starts])
            e = psd_errors(lin.cpu().numpy()[0].T, ref.T)
            ddb = float(np.abs(db.cpu().numpy()[0].astype(np.float64) - 10 * np.log10(ref + 1e-15)).max())
            print(f"{nfft:6d} {mode}     {plan.variant:40s}  {e['col']:.2e}          {e['bin_p999']:.2e}    {e['bin_max']:.2e}    {max(ddb, e['db_max']):.2e}",
                  flush=True)


if __name__ == "__main__":
    main()
    noise_table()
