#!/usr/bin/env python
"""Generate tests/golden/*.npz from the UNMODIFIED reference functions.

Runs only in the build container (needs /root/reference and scipy).  The
reference module cannot be imported here (PyQt5 / digital_rf / ipdb are not
installed, drfProc.py:41-56), so the three pure functions on the hot path --
``sti_proc_data`` (drfProc.py:364-403), ``proc_data`` (drfProc.py:406-453) and
``get_ref`` (drfProc.py:182-201) -- are pulled out of the source by AST and
executed verbatim with their two real dependencies (numpy, scipy.signal).
Nothing from the reference is copied into this repository: only the numerical
inputs/outputs are stored.

Usage:  python tools/make_golden.py  [--ref /root/reference/drfProc.py]
"""
from __future__ import annotations

import argparse
import ast
import json
import os
from fractions import Fraction

import numpy as np
import scipy
import scipy.signal as sig

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "..", "tests", "golden")
WANTED = ("sti_proc_data", "proc_data", "get_ref")


def load_reference(path):
    tree = ast.parse(open(path).read())
    body = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in WANTED]
    assert len(body) == len(WANTED), [n.name for n in body]
    ns = {"np": np, "sig": sig}
    exec(compile(ast.Module(body=body, type_ignores=[]), path, "exec"), ns)
    return ns


def iq(rng, shape, sigma=1e-2, tone=None, dtype=np.complex64):
    x = (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)) * (sigma / np.sqrt(2))
    if tone is not None:
        amp, cyc_per_sample = tone
        n = np.arange(shape[0]).reshape((-1,) + (1,) * (len(shape) - 1))
        x = x + amp * np.exp(2j * np.pi * cyc_per_sample * n)
    return x.astype(dtype)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference/drfProc.py")
    args = ap.parse_args()
    ref = load_reference(args.ref)
    os.makedirs(OUT, exist_ok=True)
    rng = np.random.default_rng(20240112)
    meta = {"numpy": np.__version__, "scipy": scipy.__version__, "cases": {}}

    def save(name, **arrays):
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **arrays)
        meta["cases"][name] = sorted(arrays)

    # --- sti_proc_data, Mode R as shipped -------------------------------------------------
    sti_cases = {
        # name: (nfft, rows, ntime, nsub or None, sr, dtype, tone)
        "sti_r_64x7": (64, 64 * 3, 7, None, 1.0e4, np.complex64, (0.1, 0.123)),
        "sti_r_256x10x3": (256, 256, 10, 3, 2.5e6, np.complex64, (0.1, -0.31)),
        "sti_r_1024x100": (1024, 1024, 100, None, 1.0e6, np.complex64, (0.1, 0.123)),
        "sti_r_4096x4": (4096, 4096 * 2, 4, 1, 25.0e6, np.complex64, None),
        "sti_r_96x5_nonpow2": (96, 96 * 2, 5, None, 48000.0, np.complex64, (0.5, 0.25)),
        "sti_r_128x6_c128": (128, 128, 6, 2, 1.0e3, np.complex128, (1.0, 10 / 128)),
    }
    for name, (nfft, rows, ntime, nsub, sr, dt, tone) in sti_cases.items():
        shape = (rows, ntime) if nsub is None else (rows, ntime, nsub)
        d1 = iq(rng, shape, tone=tone, dtype=dt)
        f, sxx, med = ref["sti_proc_data"](d1, sr, nfft)
        save(name, d1=d1, sr=np.float64(sr), nfft=np.int64(nfft), f=f, sxx=sxx, med=med,
             sxx_db=10 * np.log10(sxx + 1e-15), med_db=10 * np.log10(med + 1e-15))

    # sample rate given as a Fraction, as DrfInput stores it (drfProc.py:77-79)
    d1 = iq(rng, (64, 4), dtype=np.complex64)
    f, sxx, med = ref["sti_proc_data"](d1, Fraction(1000000, 3), 64)
    save("sti_r_fraction_sr", d1=d1, sr_num=np.int64(1000000), sr_den=np.int64(3),
         nfft=np.int64(64), f=f, sxx=sxx, med=med)

    # analytic known answers through the reference function
    imp = np.zeros((16, 2), np.complex64)
    imp[0, :] = 1 + 2j
    f, sxx, med = ref["sti_proc_data"](imp, 1.0, 16)
    save("sti_r_impulse16", d1=imp, sr=np.float64(1.0), nfft=np.int64(16), f=f, sxx=sxx, med=med)
    n = np.arange(1024)
    tone = np.exp(2j * np.pi * 37 * n / 1024).astype(np.complex64)[:, None]
    f, sxx, med = ref["sti_proc_data"](tone, 1024.0, 1024)
    save("sti_r_tone1024", d1=tone, sr=np.float64(1024.0), nfft=np.int64(1024), f=f, sxx=sxx, med=med)
    z = np.zeros((32, 3), np.complex64)
    f, sxx, med = ref["sti_proc_data"](z, 1.0, 32)
    save("sti_r_zeros32", d1=z, sr=np.float64(1.0), nfft=np.int64(32), f=f, sxx=sxx, med=med,
         sxx_db=10 * np.log10(sxx + 1e-15))

    # high dynamic range: unit tone + noise 60 dB down
    hd = iq(rng, (2048, 3), sigma=1e-3, tone=(1.0, 0.2), dtype=np.complex64)
    f, sxx, med = ref["sti_proc_data"](hd, 1.0e6, 2048)
    save("sti_r_hdr2048", d1=hd, sr=np.float64(1.0e6), nfft=np.int64(2048), f=f, sxx=sxx, med=med)

    # --- Mode A: the welch(noverlap=0) call periodogram forwards to, without truncation ---
    # (scipy-derived, not a reference function: the reference has no averaging STI entry point)
    for name, (nfft, nint, ntime, nsub) in {"sti_a_128x5x6x2": (128, 5, 6, 2),
                                            "sti_a_512x9x4": (512, 9, 4, None)}.items():
        shape = (nfft * nint + 17, ntime) if nsub is None else (nfft * nint + 17, ntime, nsub)
        d1 = iq(rng, shape, tone=(0.1, 0.123), dtype=np.complex64)
        w = sig.get_window(("kaiser", 1.7), nfft)
        f, p = sig.welch(d1, 1.0e6, window=w, nperseg=nfft, noverlap=0, nfft=nfft, detrend=False,
                         return_onesided=False, scaling="spectrum", axis=0)
        sxx = np.fft.fftshift(p, axes=0)
        save(name, d1=d1, sr=np.float64(1.0e6), nfft=np.int64(nfft), f=np.fft.fftshift(f),
             sxx=sxx, med=np.median(sxx, axis=1))

    # --- proc_data (Mode S) -----------------------------------------------------------------
    for name, (nfft, nsamp, sr, dtv) in {"proc_256": (256, 256 * 60, 1.0e4, 0.1),
                                         "proc_1024": (1024, 1024 * 40 + 333, 1.0e6, 0.005)}.items():
        x = iq(rng, (nsamp,), tone=(0.1, 0.123), dtype=np.complex64)
        t_out, f, sxx, med = ref["proc_data"](x, sr, nfft, dtv)
        save(name, x=x, sr=np.float64(sr), nfft=np.int64(nfft), dt=np.float64(dtv),
             t_out=t_out, f=f, sxx=sxx, med=med)

    # --- sti_proc_data on pure noise at the large FFT lengths (per-bin criterion of tests/parity.py: every bin of
    # a noise-like column within 1e-5 at the 99.9th percentile and 1e-3 dB).  A generator of their own, so that the
    # fixtures above stay bit-identical when cases are added here.
    rng_big = np.random.default_rng(20261018)
    for name, (nfft, ntime) in {"sti_r_noise8192x4": (8192, 4), "sti_r_noise16384x3": (16384, 3),
                                "sti_r_noise32768x2": (32768, 2), "sti_r_noise65536x2": (65536, 2)}.items():
        d1 = iq(rng_big, (nfft, ntime), dtype=np.complex64)
        f, sxx, med = ref["sti_proc_data"](d1, 25.0e6, nfft)
        save(name, d1=d1, sr=np.float64(25.0e6), nfft=np.int64(nfft), f=f, sxx=sxx, med=med)

    # --- round FFT lengths (the numbers people type into the viewer's nfft box, drfview.py:474-479; scipy transforms any
    # length): pure noise through the reference for the compile-time mixed-radix plans (1000, 5000), the run-time
    # mixed-radix kernel with radices 7 / 11 / 13 (1001, 7000) and Bluestein (1009 is prime).  Their own generator.
    rng_round = np.random.default_rng(20261019)
    for name, (nfft, ntime) in {"sti_r_noise1000x4": (1000, 4), "sti_r_noise5000x3": (5000, 3), "sti_r_noise1001x4": (1001, 4),
                                "sti_r_noise7000x2": (7000, 2), "sti_r_noise1009x3": (1009, 3)}.items():
        d1 = iq(rng_round, (nfft, ntime), dtype=np.complex64)
        f, sxx, med = ref["sti_proc_data"](d1, 1.0e6, nfft)
        save(name, d1=d1, sr=np.float64(1.0e6), nfft=np.int64(nfft), f=f, sxx=sxx, med=med)

    # --- get_ref ----------------------------------------------------------------------------
    props = [
        {"H5Tget_class": 1, "H5Tget_precision": 32, "H5Tget_size": 4},
        {"H5Tget_class": 0, "H5Tget_precision": 16, "H5Tget_size": 2},
        {"H5Tget_class": 0, "H5Tget_precision": 8, "H5Tget_size": 1},
        {"H5Tget_class": 0, "H5Tget_precision": 32, "H5Tget_size": 4},
        {"H5Tget_class": 0, "H5Tget_precision": 12, "H5Tget_size": 2},
    ]
    meta["get_ref"] = [{"props": p, "ref": float(ref["get_ref"](p))} for p in props]

    # --- frame starts: np.linspace(..., dtype=int) exactly as drfProc.py:158-159 evaluates it
    starts = []
    for st, en, nfft, nint, ntime in [(0, 10_000_000, 1024, 97, 100), (0, 10_000_000, 1024, 1, 100),
                                      (170000000000000000, 170000000360000000 + 4096 * 366, 4096, 366, 1000),
                                      (123, 123 + 50 * 256, 256, 2, 3), (5, 5 + 64, 64, 1, 1)]:
        n_sample = nint * nfft
        n_st = np.linspace(st, en - n_sample, ntime, dtype=int)
        starts.append({"st": st, "en": en, "nfft": nfft, "nint": nint, "ntime": ntime,
                       "first": [int(v) for v in n_st[:4]], "last": [int(v) for v in n_st[-4:]],
                       "sum": int(np.sum(n_st.astype(object)))})
    meta["frame_starts"] = starts

    # Kaiser table at the sizes the path uses (float64, from scipy.signal.get_window)
    save("kaiser_tables", **{f"n{n}": sig.get_window(("kaiser", 1.7), n) for n in (16, 96, 256, 1024, 4096)})

    with open(os.path.join(OUT, "meta.json"), "w") as fh:
        json.dump(meta, fh, indent=1)
    total = sum(os.path.getsize(os.path.join(OUT, p)) for p in os.listdir(OUT))
    print(f"wrote {len(meta['cases'])} cases, {total/1e6:.2f} MB -> {os.path.normpath(OUT)}")


if __name__ == "__main__":
    main()
