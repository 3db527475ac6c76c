#!/usr/bin/env python
"""Measured A/B for north_star's "DFT-as-GEMM option" at small nfft (BASELINE config 5).

The windowed DFT of a batch of frames is one real GEMM: [2N x 2N] (cos/sin blocks, window folded
into the columns) times [2N x B] (re/im of B frames).  cuBLAS through torch.matmul stands in for the
best tensor-core GEMM stage one could write (it runs tcgen05 kernels on B200), in three precisions:
bf16 and tf32 (too coarse for the 1e-5 parity bar, shown as the speed ceiling) and fp32.  |X|^2 and
the STI accumulation are not even included.  Printed beside the fused shared-memory FFT kernel on
the same number of samples.  GPU box only; tuning aid, not product code.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pyspectrogram_b200 import engine


def timeit(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def main():
    dev = torch.device("cuda")
    nsamp = 1 << 28  # 2 GiB of complex64
    for nfft in (256, 512, 1024, 2048):
        nfr = nsamp // nfft
        iq = torch.empty(nsamp, dtype=torch.complex64, device=dev)
        torch.view_as_real(iq).normal_(0, 1e-2)
        plan = engine.StiPlan(nfft)
        ntime = 1024
        nint = nfr // ntime
        starts = torch.arange(ntime, device=dev, dtype=torch.int64) * (nint * nfft)
        out = torch.empty((1, ntime, nfft), dtype=torch.float32, device=dev)
        t_fft = timeit(lambda: plan.run(iq, starts, nint, nfft, want_lin=False, want_db=True, out_db=out))
        n = np.arange(nfft)
        w = plan.window_table().astype(np.float64)
        ang = -2 * np.pi * np.outer(n, n) / nfft
        c, s = np.cos(ang) * w[None, :], np.sin(ang) * w[None, :]
        dft = np.block([[c, -s], [s, c]])  # [re; im] -> [Re X; Im X]
        frames = torch.view_as_real(iq).reshape(nfr, nfft, 2).permute(2, 1, 0).reshape(2 * nfft, nfr)  # [2N, B] view
        line = f"nfft={nfft:5d}: fused FFT kernel {t_fft:7.3f} ms ({nsamp / t_fft / 1e6:6.1f} Gs/s, {plan.variant})"
        for name, dt, tf32 in (("bf16", torch.bfloat16, False), ("tf32", torch.float32, True), ("fp32", torch.float32, False)):
            torch.backends.cuda.matmul.allow_tf32 = tf32
            a = torch.from_numpy(dft).to(dev, dt)
            nb = min(nfr, (1 << 27) // nfft)  # bound the operand copy
            b = frames[:, :nb].to(dt).contiguous()
            t = timeit(lambda: torch.matmul(a, b), reps=3)
            gs = nb * nfft / t / 1e6
            line += f" | GEMM {name} {gs:6.1f} Gs/s ({2 * (2 * nfft) ** 2 * nb / t / 1e9:7.1f} TFLOP/s)"
            del a, b
        print(line, flush=True)
        del iq, frames


if __name__ == "__main__":
    main()
