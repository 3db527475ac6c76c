"""numpy restatement of the tuned kernels' index algebra (pyspectrogram_b200/csrc/sti_kernels.cuh).

Runs the same pass structure -- in-place mixed-radix DIF, per-pass twiddle tables
``twp[(k-1)*S_p + n']``, digit-reversed last pass mapped by ``low_freq`` -- in float64 and checks it
against ``np.fft.fft`` for every radix set the library registers.  This pins the addressing on the
CPU so GPU time is spent on numerics and speed, not on index bugs.
"""
import os

import numpy as np
import pytest

from pyspectrogram_b200 import _lib


def pass_tables(n, radices):
    tabs, s = [], n
    for p, r in enumerate(radices[:-1]):
        s //= r
        m = r * s
        k = np.arange(1, r)[:, None]
        i = np.arange(s)[None, :]
        tabs.append(np.exp(-2j * np.pi * ((i * k) % m) / m).reshape(-1))
    return tabs


def low_freq(b, n, radices):
    P = len(radices)
    RL = radices[-1]
    S = [n // int(np.prod(radices[:p + 1])) for p in range(P)]
    rem, k = b.copy(), np.zeros_like(b)
    mult = 1
    for p in range(P - 1):
        s = S[p] // RL
        k += (rem // s) * mult
        rem = rem % s
        mult *= radices[p]
    return k


def model_fft(x, radices):
    n = x.shape[0]
    buf = x.astype(np.complex128).copy()
    tabs = pass_tables(n, radices)
    s = n
    P = len(radices)
    for p, r in enumerate(radices):
        s //= r
        m = r * s
        b = np.arange(n // r)
        npr = b & (s - 1)
        base = (b // s) * m + npr
        idx = base[:, None] + np.arange(r)[None, :] * s        # [butterfly][n]
        a = buf[idx]
        dft = np.exp(-2j * np.pi * np.outer(np.arange(r), np.arange(r)) / r)
        out = a @ dft.T                                         # out[:, k] = sum_n a[:, n] W_r^{nk}
        if p < P - 1:
            tw = np.ones((n // r, r), np.complex128)
            tw[:, 1:] = tabs[p].reshape(r - 1, s)[:, npr].T
            buf[idx] = out * tw
        else:
            freq = low_freq(b, n, radices)[:, None] + (n // r) * np.arange(r)[None, :]
            res = np.empty(n, np.complex128)
            res[freq] = out
            return res


def registered_radix_sets():
    lib = _lib.load()
    seen = set()
    for i in range(lib.psg_variant_count()):
        name = lib.psg_variant_name(i).decode()
        logn = lib.psg_variant_logn(i)
        radices = tuple(int(v) for v in name.split("_")[1].split("x"))
        seen.add((logn, radices))
    return sorted(seen)


@pytest.mark.parametrize("logn,radices", registered_radix_sets())
def test_variant_index_algebra(logn, radices):
    n = 1 << logn
    assert int(np.prod(radices)) == n
    rng = np.random.default_rng(logn)
    x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    got = model_fft(x, list(radices))
    ref = np.fft.fft(x)
    assert np.abs(got - ref).max() <= 1e-9 * np.abs(ref).max()


def conflict_degree(n, radices, E=16):
    """Worst number of lanes of one shared-memory phase that share a bank group, per pass, under
    pos + 2*(pos >> 4).  64-bit accesses: 16 lanes per phase, 16 groups of 8 bytes.  The last pass
    (S == 1) reads float4: 8 lanes per phase, 8 groups of 16 bytes."""
    T = n // E
    s = n
    out = []
    for p, r in enumerate(radices):
        s //= r
        m = r * s
        wide = (p == len(radices) - 1) and s == 1 and r >= 2
        lanes = 8 if wide else 16
        worst = 0
        for i in range(E // r):
            for ph in range(0, min(T, 32), lanes):
                t = np.arange(ph, min(ph + lanes, T))
                b = t + i * T
                base = b if p == 0 else (b // s) * m + (b & (s - 1))  # pass 0: its stores
                for nn in range(0, r, 2 if wide else 1):
                    pos = base + nn * s
                    addr = pos + 2 * (pos >> 4)
                    if wide:
                        assert not (addr & 1).any()  # 16-byte aligned
                        grp = (addr // 2) % 8
                    else:
                        grp = addr % 16
                    worst = max(worst, int(np.bincount(grp).max()))
        out.append(worst)
    return out


def test_padding_is_conflict_free_for_default_variants():
    """pos + 2 * (pos >> 4): in every pass of every default plan the 16 lanes of a half-warp hit 16
    distinct 8-byte banks."""
    lib = _lib.load()
    seen = set()
    for i in range(lib.psg_variant_count()):
        name = lib.psg_variant_name(i).decode()
        key = (lib.psg_variant_logn(i), name[:3])
        if key in seen:
            continue  # only the first variant per (nfft, loader) is a default
        seen.add(key)
        radices = [int(v) for v in name.split("_")[1].split("x")]
        assert conflict_degree(1 << key[0], radices) == [1] * len(radices), name


def test_fused_radix2_pair_exchange_8192():
    """8192 = 16 x 16 x 2 x 16 with the radix-2 pass done in registers (smem_pass_fused_r2): lane l of a
    warp holds the 16 outputs of pass 1 for n' = l; lanes l and l ^ 16 swap half of them (upper lanes
    send k = 0..7, lower lanes k = 8..15), finish the radix-2 butterflies of the half they keep with
    w2 = +-W_32^(l & 15) and store to k*32 + {0, 16} + (l & 15).  Simulated lane by lane and compared
    with the plain four-pass model and with np.fft."""
    n, radices = 8192, [16, 16, 2, 16]
    rng = np.random.default_rng(13)
    x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    buf = x.astype(np.complex128).copy()
    tabs = pass_tables(n, radices)
    # pass 0 (radix 16, stride 512), as in model_fft
    s0 = 512
    b = np.arange(s0)
    idx = b[:, None] + np.arange(16)[None, :] * s0
    dft16 = np.exp(-2j * np.pi * np.outer(np.arange(16), np.arange(16)) / 16)
    tw = np.ones((s0, 16), np.complex128)
    tw[:, 1:] = tabs[0].reshape(15, s0)[:, b].T
    buf[idx] = (buf[idx] @ dft16.T) * tw
    # pass 1 + fused radix-2, thread by thread (b = t; lane = t & 31)
    s1 = 32
    out = buf.copy()
    regs = np.empty((s0, 16), np.complex128)
    for t in range(s0):
        npr = t & 31
        pos = (t // s1) * 512 + npr + np.arange(16) * s1
        a = dft16 @ buf[pos]
        a[1:] *= tabs[1].reshape(15, s1)[:, npr]
        regs[t] = a
    w32 = np.exp(-2j * np.pi * np.arange(16) / 32)
    for t in range(s0):
        npr = t & 31
        hi = bool(npr & 16)
        partner = t ^ 16
        w2 = -w32[npr & 15] if hi else w32[npr & 15]
        q = (t // s1) * 512 + (npr & 15) + (8 * s1 if hi else 0)
        for i in range(8):
            keep = regs[t][8 + i] if hi else regs[t][i]
            # the partner sends (it is hi ? a[i] : a[8 + i]); it sits in the other half-warp
            recv = regs[partner][8 + i] if hi else regs[partner][i]
            out[q + i * s1] = keep + recv
            out[q + i * s1 + 16] = (keep - recv) * w2
    # reference: the plain pass 1 then pass 2 of the four-pass model
    ref = buf.copy()
    for p, (r, s) in ((1, (16, 32)), (2, (2, 16))):
        m = r * s
        bb = np.arange(n // r)
        npr = bb & (s - 1)
        ii = ((bb // s) * m + npr)[:, None] + np.arange(r)[None, :] * s
        d = np.exp(-2j * np.pi * np.outer(np.arange(r), np.arange(r)) / r)
        twp = np.ones((n // r, r), np.complex128)
        twp[:, 1:] = tabs[p].reshape(r - 1, s)[:, npr].T
        ref[ii] = (ref[ii] @ d.T) * twp
    assert np.abs(out - ref).max() <= 1e-12 * np.abs(ref).max()
    # and the whole plan is an FFT
    got = model_fft(x, radices)
    assert np.abs(got - np.fft.fft(x)).max() <= 1e-9 * np.abs(x).sum()
    # the 512 positions a warp writes are the ones its threads read in the last pass (stride 1):
    # __syncwarp is enough between them
    for wq in range(16):
        ts = np.arange(32 * wq, 32 * wq + 32)
        written = set()
        for t in ts:
            npr = t & 31
            q = (t // s1) * 512 + (npr & 15) + (8 * s1 if npr & 16 else 0)
            for i in range(8):
                written.update((q + i * s1, q + i * s1 + 16))
        read = set((ts[:, None] * 16 + np.arange(16)[None, :]).reshape(-1).tolist())
        assert written == read


def test_whole16_plan_16x2x16x2x16():
    """The 16384-point whole-frame kernel with both radix-2 passes in registers (sti_whole16.cuh): the
    five-pass plan is an FFT, and the kernel's closed-form frequency map of the last pass
    (butterfly b = k0*64 + k1*32 + k2*2 + k3 -> k0 + 16 k1 + 32 k2 + 512 k3 + 1024 j) is low_freq."""
    n, radices = 16384, [16, 2, 16, 2, 16]
    rng = np.random.default_rng(16)
    x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    got = model_fft(x, radices)
    ref = np.fft.fft(x)
    assert np.abs(got - ref).max() <= 1e-9 * np.abs(ref).max()
    b = np.arange(n // 16)
    klow = (b >> 6) + 16 * ((b >> 5) & 1) + 32 * ((b >> 1) & 15) + 512 * (b & 1)
    assert np.array_equal(klow, low_freq(b, n, radices))
    # slab geometry of the first pass: the four slabs of 256 threads cover every n' exactly once
    seen = np.zeros(1024, int)
    for m in range(4):
        for u in range(256):
            lane = u & 31
            npp = 128 * m + 16 * (u >> 5) + (lane & 15)
            seen[npp + 512 * (lane >> 4)] += 1
    assert (seen == 1).all()


# ---------------------------------------------------------------------------------------------
# radix-32 whole-frame kernels (sti_r32.cuh): CPU replay with the kernel's own math header
# ---------------------------------------------------------------------------------------------
def test_r32_kernels_replayed_on_the_cpu(tmp_path):
    """tests/c/r32_emu.cu includes csrc/r32_math.cuh (butterflies, twiddle recurrences, address functions and the
    accumulator-to-bin map are __host__ __device__ and bit-identical on the host) and walks the three passes of
    every geometry (8192 ... 65536, one CTA to a cluster of four) thread by thread: bins against a float64 FFT,
    every address against the generic swizzle, every warp-wide shared-memory access bank-conflict free."""
    import shutil
    import subprocess
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    here = os.path.dirname(os.path.abspath(__file__))
    exe = str(tmp_path / "r32_emu")
    subprocess.run([nvcc, "-O2", "-std=c++17", "-Wno-deprecated-gpu-targets", "-o", exe, os.path.join(here, "c", "r32_emu.cu")],
                   check=True, capture_output=True)
    res = subprocess.run([exe], capture_output=True, text=True)
    assert res.returncode == 0 and res.stdout.strip().endswith("OK"), res.stdout + res.stderr


# ---------------------------------------------------------------------------------------------
# compile-time mixed-radix plans for round lengths (sti_mixct.cuh): CPU replay with the kernel's own header
# ---------------------------------------------------------------------------------------------
def test_mixct_plans_replayed_on_the_cpu(tmp_path):
    """tests/c/mixct_emu.cu includes csrc/sti_mixct.cuh and csrc/mixct_plans.inc: the prime-factor butterflies
    (6 / 10 / 12 / 15 / 20 from 2 / 3 / 4 / 5-point DFTs) against a float64 DFT, and for every plan of the list the
    whole pass structure -- index algebra, twiddles (register form and rebuilt from powers), padded addresses, the
    position -> frequency permutation of the epilogue -- against a float64 DFT of the same input."""
    import shutil
    import subprocess
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    here = os.path.dirname(os.path.abspath(__file__))
    exe = str(tmp_path / "mixct_emu")
    subprocess.run([nvcc, "-O2", "-std=c++17", "-Wno-deprecated-gpu-targets", "-I", os.path.join(here, "..", "pyspectrogram_b200", "csrc"),
                    "-o", exe, os.path.join(here, "c", "mixct_emu.cu")], check=True, capture_output=True)
    res = subprocess.run([exe], capture_output=True, text=True)
    assert res.returncode == 0 and res.stdout.strip().endswith("ALL OK"), res.stdout + res.stderr
    assert res.stdout.count("plan ") >= 20
