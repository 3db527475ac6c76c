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
    2400: ((10, 15, 16), 240), 3000: ((10, 15, 20), 300), 3600: ((15, 15, 16), 240), 4000: ((10, 20, 20), 400),
    4800: ((15, 16, 20), 320), 5000: ((10, 10, 10, 5), 500), 6000: ((15, 20, 20), 400), 8000: ((20, 20, 20), 400),
    10000: ((10, 10, 10, 10), 500),
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


def wavefronts(acc, PQ, PA):
    tot = 0
    for pos in acc:
        addr = pos + (PA * (pos // PQ) if PQ else 0)
        # half-warps by lane position inside the warp-instruction (inactive lanes are at the end of a group only)
        for h in (addr[:16], addr[16:]):
            if h.size == 0:
                continue
            u = np.unique(h)
            tot += np.bincount(u % 16, minlength=16).max()
    return tot


def main():
    for N, (radices, T) in PLANS.items():
        acc = accesses(N, radices, T)
        ideal = sum((1 if a.size <= 16 else 2) for a in acc)
        best = None
        cands = [(0, 0)] + [(pq, pa) for pq in sorted({radices[-1], 16, 32, radices[-1] * 2, N // radices[0], 8, 10, 20, 64})
                            for pa in (1, 2, 3)]
        for pq, pa in cands:
            w = wavefronts(acc, pq, pa)
            if best is None or w < best[0]:
                best = (w, pq, pa)
        w0 = wavefronts(acc, 0, 0)
        print(f"N={N:6d} {'x'.join(map(str, radices)):12s} T={T:4d} ideal {ideal:6d}  unpadded {w0:6d} ({w0 / ideal:.2f}x)  "
              f"best PQ={best[1]:3d} PA={best[2]}: {best[0]:6d} ({best[0] / ideal:.2f}x)")


if __name__ == "__main__":
    main()
