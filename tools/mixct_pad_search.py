#!/usr/bin/env python
"""Padding of the exchange buffer of the compile-time mixed-radix plans (sti_mixct.cuh): for every plan
(N, radices, T) count the shared-memory wavefronts of all 64-bit accesses of one frame under
pad(pos) = pos + PA * (pos // PQ) and print the best (PQ, PA).  Model: a warp's 64-bit access is served per
half-warp; a half-warp needs as many wavefronts as the largest number of distinct 8-byte words that fall into one
of the 16 64-bit banks."""
import itertools
import sys

import numpy as np

PLANS = {
    1000: ((10, 10, 10), 100), 1200: ((10, 10, 12), 120), 1500: ((10, 10, 15), 150), 2000: ((10, 10, 20), 200),
    2400: ((10, 15, 16), 240), 3000: ((10, 15, 20), 300), 3600: ((15, 15, 16), 240), 4000: ((10, 10, 10, 4), 400),
    4800: ((15, 16, 20), 320), 5000: ((10, 10, 10, 5), 500), 6000: ((15, 20, 20), 400), 8000: ((20, 20, 20), 400),
    10000: ((10, 10, 10, 10), 500),
    1600: ((10, 10, 16), 160), 2500: ((5, 10, 10, 5), 250), 3200: ((10, 16, 20), 320), 6400: ((16, 20, 20), 400),
    1800: ((10, 12, 15), 180), 2700: ((12, 15, 15), 225), 4500: ((15, 15, 20), 300),
}


def accesses(N, radices, T):
    """yield arrays of positions, one per (warp-instruction): shape [lanes<=32] (inactive lanes dropped)"""
    S = N
    out = []
    for p, R in enumerate(radices):
        S //= R
        nbf = N // R
        nb = -(-nbf // T)
        for i in range(nb):
            bf = np.arange(T) + i * T
            act = bf < nbf
            blk, npr = bf // S, bf % S
            base = blk * R * S + npr
            kinds = (["w"] if p == 0 else ["r"] if p == len(radices) - 1 else ["r", "w"])
            for _ in kinds:
                for n in range(R):
                    pos = base + n * S
                    for w0 in range(0, T, 32):
                        sel = act[w0:w0 + 32]
                        if sel.any():
                            out.append(pos[w0:w0 + 32][sel])
    return out


def wavefronts(acc, PQ, PA, PQ2=0, PA2=0):
    tot = 0
    for pos in acc:
        addr = pos + (PA * (pos // PQ) if PQ else 0) + (PA2 * (pos // PQ2) if PQ2 else 0)
        # half-warps by lane position inside the warp-instruction (inactive lanes are at the end of a group only)
        for h in (addr[:16], addr[16:]):
            if h.size == 0:
                continue
            u = np.unique(h)
            tot += np.bincount(u % 16, minlength=16).max()
    return tot


def main():
    """Candidates are the paddings that stay LINEAR inside every butterfly (MixPlan::lin_ok): PQ = the last radix, PQ2 a
    product of trailing radices; smallest buffer among equals."""
    for N, (radices, T) in PLANS.items():
        acc = accesses(N, radices, T)
        ideal = sum((1 if a.size <= 16 else 2) for a in acc)
        RL = radices[-1]
        q2s, q = [], RL
        for r in reversed(radices[:-1]):
            q *= r
            if q < N:
                q2s.append(q)
        best = None
        for pa in range(0, 4):
            for q2 in q2s + [0]:
                for pa2 in (range(0, 17) if q2 else [0]):
                    w = wavefronts(acc, RL if pa else 0, pa, q2 if pa2 else 0, pa2)
                    size = N + pa * (N // RL) + (pa2 * (N // q2) if q2 else 0)
                    if best is None or w < best[0] or (w == best[0] and size < best[1]):
                        best = (w, size, RL if pa else 0, pa, q2 if pa2 else 0, pa2)
        w0 = wavefronts(acc, 0, 0)
        print(f"N={N:6d} {'x'.join(map(str, radices)):12s} T={T:4d} ideal {ideal:6d}  unpadded {w0 / ideal:.2f}x  "
              f"best PQ={best[2]:3d} PA={best[3]} PQ2={best[4]:4d} PA2={best[5]:2d}: {best[0] / ideal:.2f}x")


if __name__ == "__main__":
    main()
