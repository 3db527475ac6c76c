"""GPU parity tests (run with ``-m gpu`` on a B200): the CUDA path through the C ABI against the
golden fixtures made from the unmodified reference functions, and against the CPU oracle on the
same seeded inputs.  Tolerances are the ones in tests/parity.py (north_star: PSD 1e-5 relative,
1e-3 dB, frame indexing and bin order exact)."""
import json
import os
from fractions import Fraction

import numpy as np
import pytest

from tests.parity import assert_db_close, assert_psd_close, psd_errors

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


@pytest.fixture(scope="module")
def dp():
    from pyspectrogram_b200 import drfProc
    return drfProc


@pytest.fixture(scope="module")
def torch():
    import torch
    assert torch.cuda.is_available()
    return torch


# ---------------------------------------------------------------------------------------------
# golden fixtures from the reference's own functions
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,noise_like", [
    ("sti_r_64x7", False), ("sti_r_256x10x3", False), ("sti_r_1024x100", False), ("sti_r_4096x4", True),
    ("sti_r_hdr2048", False), ("sti_r_impulse16", True), ("sti_r_tone1024", False),
    # pure noise through the reference at the large FFT lengths: every bin is held to the per-bin criterion
    ("sti_r_noise8192x4", True), ("sti_r_noise16384x3", True), ("sti_r_noise32768x2", True), ("sti_r_noise65536x2", True),
    # round lengths: compile-time mixed-radix plans (1000, 5000), radices 7 / 11 / 13 (1001, 7000), Bluestein (1009: prime)
    ("sti_r_noise1000x4", True), ("sti_r_noise5000x3", True), ("sti_r_noise1001x4", True), ("sti_r_noise7000x2", True),
    ("sti_r_noise1009x3", True)])
def test_sti_proc_data_matches_reference_golden(dp, name, noise_like):
    g = load(name)
    f, sxx, med = dp.sti_proc_data(g["d1"], float(g["sr"]), int(g["nfft"]))
    assert np.array_equal(f, g["f"])
    assert sxx.dtype == g["sxx"].dtype and med.dtype == g["med"].dtype
    assert sxx.shape == g["sxx"].shape and med.shape == g["med"].shape
    assert_psd_close(sxx, g["sxx"], noise_like=noise_like, what=name)
    assert_psd_close(med, g["med"], noise_like=noise_like, what=name + " median")
    if "sxx_db" in g:
        f2, sdb, mdb = dp.sti_proc_data_db(g["d1"], float(g["sr"]), int(g["nfft"]))
        assert_db_close(sdb, g["sxx_db"], ref_lin=g["sxx"], what=name + " dB")
        assert_db_close(mdb, g["med_db"], ref_lin=g["med"], what=name + " median dB")


def test_complex128_input_returns_float64(dp):
    g = load("sti_r_128x6_c128")
    f, sxx, med = dp.sti_proc_data(g["d1"], float(g["sr"]), int(g["nfft"]))
    assert sxx.dtype == np.float64 and med.dtype == np.float64 and np.array_equal(f, g["f"])
    assert_psd_close(sxx, g["sxx"], noise_like=False, what="c128")


def test_fraction_sample_rate(dp):
    g = load("sti_r_fraction_sr")
    f, sxx, med = dp.sti_proc_data(g["d1"], Fraction(int(g["sr_num"]), int(g["sr_den"])), int(g["nfft"]))
    assert np.array_equal(f, g["f"])
    assert_psd_close(sxx, g["sxx"], what="fraction sr")


def test_zero_input_hits_db_floor(dp):
    g = load("sti_r_zeros32")
    f, sxx, med = dp.sti_proc_data(g["d1"], 1.0, 32)
    assert np.array_equal(sxx, g["sxx"]) and not sxx.any()
    f, sdb, mdb = dp.sti_proc_data_db(g["d1"], 1.0, 32)
    assert np.abs(sdb - g["sxx_db"]).max() <= 1e-4  # -150.00002 dB
    assert sdb.dtype == np.float32


def test_non_power_of_two_matches_reference_golden(dp):
    """The viewer allows any FFT length (drfview.py:474-479); non powers of two run Bluestein."""
    g = load("sti_r_96x5_nonpow2")
    f, sxx, med = dp.sti_proc_data(g["d1"], float(g["sr"]), 96)
    assert np.array_equal(f, g["f"]) and sxx.shape == g["sxx"].shape and sxx.dtype == g["sxx"].dtype
    assert_psd_close(sxx, g["sxx"], noise_like=False, what="nfft=96")
    assert_psd_close(med, g["med"], noise_like=False, what="nfft=96 median")


@pytest.mark.parametrize("nfft", [3, 5, 7, 33, 100, 169, 1000, 1001, 1023, 1025, 3000, 4004, 4095, 5000, 7000, 8191, 10000, 12000, 15000,
                                  20000, 100000])
@pytest.mark.parametrize("mode", ["R", "A"])
def test_arbitrary_nfft_against_float64_oracle(torch, nfft, mode):
    """Odd, even, prime and large non power-of-two lengths (work buffer in shared memory up to
    M = 16384, in global scratch above), fftshift for odd N, oracle = float64 numpy."""
    from pyspectrogram_b200 import engine
    rng = np.random.default_rng(nfft)
    nfr = 1 if mode == "R" else 3
    ncol = 4 if nfft <= 20000 else 2
    n = nfft * nfr * ncol + nfft + 3
    x = _recording(rng, n)
    starts = (np.arange(ncol) * nfft * nfr + np.arange(ncol) % 2).astype(np.int64)
    plan = engine.StiPlan(nfft)
    lin, db = plan.run(torch.from_numpy(x).cuda(), torch.from_numpy(starts).cuda(), nfr, nfft, want_lin=True, want_db=True)
    # 2^a 3^b 5^c 7^d 11^e 13^f that fit shared memory run the direct mixed-radix transform; the rest Bluestein:
    # convolution lengths up to 16384 with mixed-radix passes, longer ones with the radix-2 kernel
    smooth = nfft
    for f in (2, 3, 5, 7, 11, 13):
        while smooth % f == 0:
            smooth //= f
    if smooth == 1 and nfft <= 15000:
        assert plan.variant.startswith((f"mixed{nfft}_", f"mixct{nfft}_")), plan.variant
    else:
        assert plan.variant.startswith("bluestein")
        assert plan.variant.endswith("_r2") == (2 * nfft - 1 > 16384), plan.variant
    ref = _oracle_columns(x, starts, nfft, nfr, nfft)
    assert_psd_close(lin.cpu().numpy()[0].T, ref.T, noise_like=False, what=f"nfft={nfft} {plan.variant}")
    assert_db_close(db.cpu().numpy()[0].T, 10 * np.log10(ref.T.astype(np.float32) + np.float32(1e-15)),
                    ref_lin=ref.T, what=f"nfft={nfft} dB")


@pytest.mark.parametrize("nfft,nfr,ncol", [(1000, 700, 2), (96, 5, 300), (6000, 2, 3), (45, 9, 40), (7500, 3, 5), (1001, 4, 9), (7000, 2, 3),
                                           (77, 6, 50)])
def test_arbitrary_nfft_kernels_agree(torch, nfft, nfr, ncol):
    """The three kernels for non powers of two on the same input: direct mixed-radix transform (default for
    2^a 3^b 5^c), Bluestein with mixed-radix passes ("bluestein"), Bluestein radix 2 ("bluestein_r2"),
    including columns split into frame chunks and more items than resident CTAs."""
    from pyspectrogram_b200 import engine
    rng = np.random.default_rng(nfft + nfr)
    x = torch.from_numpy(_recording(rng, nfft * nfr * ncol + 5)).cuda()
    starts = torch.from_numpy((np.arange(ncol) * nfft * nfr + np.arange(ncol) % 2).astype(np.int64)).cuda()
    plan = engine.StiPlan(nfft)
    res = {}
    try:
        for var, want in ((None, ("mixed", "mixct")), ("mixed_rt", "mixed"), ("bluestein", "bluestein_m"), ("bluestein_r2", "bluestein_m")):
            engine.set_variant(var)
            lin, _ = plan.run(x, starts, nfr, nfft)
            torch.cuda.synchronize()
            assert plan.variant.startswith(want) and plan.variant.endswith("_r2") == (var == "bluestein_r2"), plan.variant
            res[var] = lin.cpu().numpy()[0]
    finally:
        engine.set_variant(None)
    ref = _oracle_columns(x.cpu().numpy(), starts.cpu().numpy()[:2], nfft, nfr, nfft)
    for var, a in res.items():
        assert_psd_close(a[:2].T, ref.T, noise_like=False, what=f"{var or 'mixed'} nfft={nfft}")
        assert_psd_close(a.T, res["bluestein_r2"].T.astype(np.float64), noise_like=False, what=f"{var or 'mixed'} vs radix-2 nfft={nfft}")


def test_short_input_raises_value_error(dp):
    with pytest.raises(ValueError):
        dp.sti_proc_data(np.zeros((100, 4), np.complex64), 1.0, 128)


@pytest.mark.parametrize("name", ["sti_a_128x5x6x2", "sti_a_512x9x4"])
def test_mode_a_matches_welch_golden(dp, name):
    g = load(name)
    f, sxx, med = dp.sti_proc_data(g["d1"], float(g["sr"]), int(g["nfft"]), integrate=True)
    assert np.array_equal(f, g["f"]) and sxx.shape == g["sxx"].shape and sxx.dtype == np.float32
    assert_psd_close(sxx, g["sxx"], noise_like=False, what=name)
    assert_psd_close(med, g["med"], noise_like=False, what=name + " median")


@pytest.mark.parametrize("name", ["proc_256", "proc_1024"])
def test_proc_data_matches_reference_golden(dp, name):
    g = load(name)
    t_out, f, sxx, med = dp.proc_data(g["x"], float(g["sr"]), int(g["nfft"]), float(g["dt"]))
    assert np.array_equal(t_out, g["t_out"]) and np.array_equal(f, g["f"])
    assert sxx.shape == g["sxx"].shape and sxx.dtype == g["sxx"].dtype
    assert_psd_close(sxx, g["sxx"], noise_like=False, what=name)
    assert_psd_close(med, g["med"], noise_like=False, what=name + " median")


# ---------------------------------------------------------------------------------------------
# every kernel variant against the float64 oracle, device-resident path
# ---------------------------------------------------------------------------------------------
def _recording(rng, n, tone=0.123):
    x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)) * (1e-2 / np.sqrt(2))
    x += 0.1 * np.exp(2j * np.pi * tone * np.arange(n))
    return x.astype(np.complex64)


def _oracle_columns(x, starts, nfft, nfr, hop):
    from oracle import np_oracle
    return np.stack([np_oracle.column_power(x[s:], nfft, nfr, hop) for s in starts])


def _all_variants():
    from pyspectrogram_b200 import engine
    return engine.variants()


def _variant_ids():
    try:
        return [v[0] for v in _all_variants()]
    except Exception:
        return []


@pytest.mark.parametrize("variant", _variant_ids() or ["none"])
def test_every_variant_matches_float64_oracle(torch, variant):
    from pyspectrogram_b200 import engine
    logn = dict(_all_variants())[variant]
    nfft = 1 << logn
    rng = np.random.default_rng(logn * 7 + 1)
    nfr, ncol = (1, 23) if variant.endswith("_m") else (5, 7)  # "_m": one-frame-per-column kernels
    n = nfft * (nfr * ncol + 3) + 11
    x = _recording(rng, n)
    starts = np.sort(rng.choice(n - nfr * nfft, ncol, replace=False)).astype(np.int64)
    starts[0] = 0
    starts[1] |= 1  # an odd (8-byte aligned only) start
    starts[-1] = n - nfr * nfft  # a column that ends at the last sample
    plan = engine.StiPlan(nfft)
    in_scale = 1.0
    feed = x
    if variant.endswith(("_i16", "_i8")):  # raw integer ingest variants: quantise the recording
        amp, dt = (20000.0, np.int16) if variant.endswith("_i16") else (100.0, np.int8)
        feed = np.stack([np.round(x.real * amp * 8), np.round(x.imag * amp * 8)], axis=1).astype(dt)
        in_scale = 1.0 / (amp * 8)
        x = ((feed[:, 0].astype(np.float32) + 1j * feed[:, 1].astype(np.float32)) * np.float32(in_scale)).astype(np.complex64)
    dx, ds = torch.from_numpy(feed).cuda(), torch.from_numpy(starts).cuda()
    try:
        engine.set_variant(variant)
        lin, db = plan.run(dx, ds, nfr, nfft, in_scale=in_scale, want_lin=True, want_db=True)
        torch.cuda.synchronize()
        assert plan.variant == variant
    finally:
        engine.set_variant(None)
    ref = _oracle_columns(x, starts, nfft, nfr, nfft)
    got = lin.cpu().numpy()[0]
    assert_psd_close(got.T, ref.T, noise_like=False, what=variant)
    assert_db_close(db.cpu().numpy()[0].T, 10 * np.log10(ref.T.astype(np.float32) + np.float32(1e-15)),
                    ref_lin=ref.T, what=variant + " dB")


@pytest.mark.parametrize("nfft", [2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384, 32768, 65536])
@pytest.mark.parametrize("mode", ["R", "A", "S"])
def test_default_path_all_sizes_and_modes(torch, nfft, mode):
    from pyspectrogram_b200 import engine
    rng = np.random.default_rng(nfft + ord(mode))
    ncol = 5 if nfft >= 16384 else 9
    nfr = {"R": 1, "A": 4, "S": 3}[mode]
    hop = nfft - nfft // 8 if mode == "S" else nfft
    n = (nfr * nfft) * ncol + 5 * nfft + 3
    x = _recording(rng, n)
    span = (nfr - 1) * hop + nfft
    starts = engine.frame_starts(1, n, nfft, -(-span // nfft), ncol).astype(np.int64)
    plan = engine.StiPlan(nfft)
    lin, _ = plan.run(torch.from_numpy(x).cuda(), torch.from_numpy(starts).cuda(), nfr, hop)
    ref = _oracle_columns(x, starts, nfft, nfr, hop)
    assert_psd_close(lin.cpu().numpy()[0].T, ref.T, noise_like=False, what=f"nfft={nfft} mode {mode} {plan.variant}")


@pytest.mark.parametrize("nfft,nfr,ncol,scratch_mb", [
    (16384, 3, 7, 1),     # 8 scratch frames: 2 columns per chunk, 4 chunks (ragged tail)
    (16384, 21, 2, 1),    # a column's frames exceed the scratch: frame blocks 8+8+5 carried and summed
    (65536, 5, 3, 2),     # 4 scratch frames < 5 frames: carried, one column per chunk
    (8192, 6, 5, 64),     # the split path cross-checks the fused 8192 kernel's size
    (32768, 2, 4, 64)])
def test_split_path_chunking(torch, nfft, nfr, ncol, scratch_mb):
    """Large-nfft two-phase path (first pass -> L2 scratch -> 4096-point fused kernel -> interleave)
    with the scratch shrunk so that every chunking branch runs; oracle = float64 numpy."""
    from pyspectrogram_b200 import engine
    rng = np.random.default_rng(nfft // 1024 + nfr)
    n = nfft * nfr * ncol + 2 * nfft + 5
    x = _recording(rng, n)
    starts = (np.arange(ncol) * nfft * nfr + np.arange(ncol) % 3).astype(np.int64)
    plan = engine.StiPlan(nfft)
    try:
        engine.set_split_scratch(scratch_mb << 20)
        engine.set_variant("split")
        lin, db = plan.run(torch.from_numpy(x).cuda(), torch.from_numpy(starts).cuda(), nfr, nfft,
                           want_lin=True, want_db=True)
        torch.cuda.synchronize()
        assert plan.variant.startswith(f"split{nfft // 4096}x4096")
    finally:
        engine.set_variant(None)
        engine.set_split_scratch(2048 << 20)
    ref = _oracle_columns(x, starts, nfft, nfr, nfft)
    assert_psd_close(lin.cpu().numpy()[0].T, ref.T, noise_like=False, what=f"split {nfft}")
    assert_db_close(db.cpu().numpy()[0].T, 10 * np.log10(ref.T.astype(np.float32) + np.float32(1e-15)),
                    ref_lin=ref.T, what=f"split {nfft} dB")


@pytest.mark.parametrize("nfft,nfr,ncol,nsub,kind", [
    (16384, 1, 9, 1, "c64"),      # Mode R: every frame is a finished item
    (16384, 5, 7, 2, "c64"),      # two sub-channels, odd starts (TMA skew)
    (16384, 640, 2, 1, "c64"),    # two long columns: split into frame chunks, fp64 sum of the splits
    (32768, 4, 5, 1, "c64"),
    (65536, 3, 5, 1, "c64"),
    (65536, 1, 40, 1, "c64"),     # more items than resident clusters: several items per cluster
    (65536, 33, 1, 1, "c64"),     # one column, ragged last chunk
    (8192, 6, 5, 1, "c64"),       # cluster of two
    (16384, 4, 6, 1, "i16"),
    (65536, 2, 3, 1, "i8"),
    (32768, 3, 4, 1, "ldg"),
    (65536, 9, 4, 1, "ldg"),      # rows loaded straight to registers instead of by bulk copy
    (8192, 6, 5, 1, "dsmem"),     # exchange through distributed shared memory (st.async)
    (16384, 1, 9, 1, "dsmem"),
    (16384, 5, 7, 2, "dsmem"),
    (16384, 640, 2, 1, "dsmem"),
    (32768, 4, 5, 1, "dsmem"),
    (65536, 3, 5, 1, "dsmem"),
    (65536, 1, 40, 1, "dsmem"),
    (65536, 33, 1, 1, "dsmem"),
    (16384, 4, 6, 1, "dsmem_i16"),
    (65536, 2, 3, 1, "dsmem_i8")])
def test_cluster_path(torch, nfft, nfr, ncol, nsub, kind):
    """Large-nfft cluster kernel (r0 = nfft/4096 CTAs per frame, exchange through L2, one cluster
    barrier per frame) against the float64 oracle: modes, sub-channels, skewed frame starts, item
    scheduling (fewer / more items than clusters, split columns) and raw integer ingest."""
    from pyspectrogram_b200 import engine
    rng = np.random.default_rng(nfft // 1024 + nfr + ncol)
    per_sub = nfft * nfr * ncol + 2 * nfft + 8
    per_sub += (-per_sub) % 8  # sub-channels start 16-byte aligned whatever the sample type
    x = _recording(rng, per_sub * nsub)
    starts = (np.arange(ncol) * nfft * nfr + np.arange(ncol) % 3).astype(np.int64)
    in_scale, feed = 1.0, x
    if kind.endswith(("i16", "i8")):
        amp, dt = (20000.0, np.int16) if kind.endswith("i16") else (100.0, np.int8)
        feed = np.stack([np.round(x.real * amp * 8), np.round(x.imag * amp * 8)], axis=1).astype(dt)
        in_scale = 1.0 / (amp * 8)
        x = ((feed[:, 0].astype(np.float32) + 1j * feed[:, 1].astype(np.float32)) * np.float32(in_scale)).astype(np.complex64)
    plan = engine.StiPlan(nfft)
    try:
        engine.set_variant("cluster_ldg" if kind == "ldg" else "cluster_dsmem" if kind.startswith("dsmem") else "cluster")  # "cluster": rows by bulk copy
        lin, db = plan.run(torch.from_numpy(feed).cuda(), torch.from_numpy(starts).cuda(), nfr, nfft, sub_stride=per_sub,
                           nsub=nsub, in_scale=in_scale, want_lin=True, want_db=True)
        torch.cuda.synchronize()
        assert plan.variant.startswith(f"cluster{nfft // 4096}x4096"), plan.variant
    finally:
        engine.set_variant(None)
    for s in range(nsub):
        ref = _oracle_columns(x[s * per_sub:], starts, nfft, nfr, nfft)
        assert_psd_close(lin.cpu().numpy()[s].T, ref.T, noise_like=False, what=f"cluster {nfft} sub {s}")
        assert_db_close(db.cpu().numpy()[s].T, 10 * np.log10(ref.T.astype(np.float32) + np.float32(1e-15)),
                        ref_lin=ref.T, what=f"cluster {nfft} dB")


@pytest.mark.parametrize("nfft,nfr,ncol,nsub,kind", [
    (16384, 1, 9, 1, "whole"),       # Mode R: one frame per item
    (16384, 5, 7, 2, "whole"),       # two sub-channels, odd starts (TMA skew)
    (16384, 640, 2, 1, "whole"),     # two long columns: split into frame chunks, fp64 sum of the splits
    (16384, 3, 500, 1, "whole"),     # more items than SMs (several waves)
    (16384, 5, 7, 1, "whole_s2"),    # two-stage ring
    (8192, 6, 5, 1, "whole"),        # radix-2 first pass
    (8192, 6, 5, 1, "whole_s8"),
    (16384, 4, 6, 1, "whole_i16"),
    (16384, 2, 3, 1, "whole_i8"),
    (32768, 1, 9, 1, "whole"),       # cluster of two CTAs, half of the first-pass outputs through DSMEM
    (32768, 5, 7, 2, "whole"),
    (32768, 320, 2, 1, "whole"),
    (32768, 3, 200, 1, "whole"),
    (32768, 4, 6, 1, "whole_i16"),
    (65536, 1, 9, 1, "whole"),       # cluster of four
    (65536, 5, 5, 2, "whole"),
    (65536, 160, 2, 1, "whole"),
    (65536, 3, 100, 1, "whole"),
    (65536, 2, 3, 1, "whole_i8"),
    (16384, 1, 9, 1, "whole_f"),     # 16 x 2 x 16 x 2 x 16 with both radix-2 passes in registers (sti_whole16.cuh)
    (16384, 5, 7, 2, "whole_f"),
    (16384, 640, 2, 1, "whole_f"),
    (16384, 3, 500, 1, "whole_f"),
    (16384, 4, 6, 1, "whole_f_i16"),
    (16384, 2, 3, 1, "whole_f_i8"),
    (16384, 1, 9, 1, "whole_r2"),    # two rows per CTA, 256 threads, two CTAs per SM: clusters of 2 / 4 / 8
    (16384, 5, 7, 2, "whole_r2"),
    (16384, 640, 2, 1, "whole_r2"),
    (16384, 3, 500, 1, "whole_r2"),
    (16384, 4, 6, 1, "whole_r2_i16"),
    (32768, 5, 7, 2, "whole_r2"),
    (32768, 3, 200, 1, "whole_r2"),
    (65536, 5, 5, 2, "whole_r2"),
    (65536, 160, 2, 1, "whole_r2"),
    (65536, 2, 3, 1, "whole_r2_i8")])
def test_whole_frame_path(torch, nfft, nfr, ncol, nsub, kind):
    """Whole-frame kernel (sti_whole.cuh: 8192 / 16384 points resident in one SM, first pass fed in
    slabs through a bulk-copy ring refilled by the last reader) against the float64 oracle: modes,
    sub-channels, skewed frame starts, split columns, several waves, raw integer ingest."""
    from pyspectrogram_b200 import engine
    rng = np.random.default_rng(nfft // 1024 + nfr + ncol)
    per_sub = nfft * nfr * ncol + 2 * nfft + 8
    per_sub += (-per_sub) % 8
    x = _recording(rng, per_sub * nsub)
    starts = (np.arange(ncol) * nfft * nfr + np.arange(ncol) % 3).astype(np.int64)
    in_scale, feed = 1.0, x
    if kind.endswith(("i16", "i8")):
        amp, dt = (20000.0, np.int16) if kind.endswith("i16") else (100.0, np.int8)
        feed = np.stack([np.round(x.real * amp * 8), np.round(x.imag * amp * 8)], axis=1).astype(dt)
        in_scale = 1.0 / (amp * 8)
        x = ((feed[:, 0].astype(np.float32) + 1j * feed[:, 1].astype(np.float32)) * np.float32(in_scale)).astype(np.complex64)
    plan = engine.StiPlan(nfft)
    try:
        engine.set_variant(kind if kind in ("whole_s2", "whole_s8") else "whole_r2" if kind.startswith("whole_r2")
                           else "whole_f" if kind.startswith("whole_f") else "whole")
        lin, db = plan.run(torch.from_numpy(feed).cuda(), torch.from_numpy(starts).cuda(), nfr, nfft, sub_stride=per_sub,
                           nsub=nsub, in_scale=in_scale, want_lin=True, want_db=True)
        torch.cuda.synchronize()
        if kind.startswith("whole_f"):
            assert plan.variant.startswith("whole16x2x16x2x16"), plan.variant
        else:
            assert plan.variant.startswith(f"whole{nfft // 4096}x4096"), plan.variant
            assert ("r2" in plan.variant) == kind.startswith("whole_r2"), plan.variant
    finally:
        engine.set_variant(None)
    for s in range(nsub):
        ref = _oracle_columns(x[s * per_sub:], starts, nfft, nfr, nfft)
        assert_psd_close(lin.cpu().numpy()[s].T, ref.T, noise_like=False, what=f"whole {nfft} sub {s}")
        assert_db_close(db.cpu().numpy()[s].T, 10 * np.log10(ref.T.astype(np.float32) + np.float32(1e-15)),
                        ref_lin=ref.T, what=f"whole {nfft} dB")


def test_large_nfft_defaults_and_fallback(torch):
    """16384 / 32768 / 65536 run the three-pass radix-32 kernels (one CTA, a pair, a cluster of four: the measured
    defaults); a recording whose base is not 16-byte aligned cannot use bulk copies and takes the split path at
    every size."""
    from pyspectrogram_b200 import engine
    for nfft, want in ((16384, "r32_32x32x16_t512c1"), (32768, "r32_32x32x32_t512c2"), (65536, "r32_32x32x2x32_t512c4")):
        x = torch.from_numpy(_recording(np.random.default_rng(nfft), nfft * 9)).cuda()
        starts = torch.from_numpy(np.arange(4, dtype=np.int64) * 2 * nfft).cuda()
        plan = engine.StiPlan(nfft)
        lin, _ = plan.run(x, starts, 2, nfft)
        torch.cuda.synchronize()
        assert plan.variant.startswith(want), plan.variant
        ref = _oracle_columns(x.cpu().numpy(), starts.cpu().numpy(), nfft, 2, nfft)
        assert_psd_close(lin.cpu().numpy()[0].T, ref.T, noise_like=False, what=f"default {nfft}")
        lin2, _ = plan.run(x.view(torch.float32)[2:].view(torch.complex64), starts, 2, nfft)
        torch.cuda.synchronize()
        assert plan.variant.startswith("split"), plan.variant
        ref = _oracle_columns(x.cpu().numpy()[1:], starts.cpu().numpy(), nfft, 2, nfft)
        assert_psd_close(lin2.cpu().numpy()[0].T, ref.T, noise_like=False, what=f"split fallback {nfft}")


@pytest.mark.parametrize("nfft", [256, 512, 1024, 2048, 4096, 8192, 16384, 32768, 65536, (16384, "cluster"), (32768, "cluster_dsmem"),
                                  (65536, "cluster_ldg"), (65536, "cluster_dsmem"), (16384, "whole"), (8192, "whole"), (32768, "whole"), (65536, "whole"), (16384, "whole_r2"), (65536, "whole_r2"), (16384, "whole_f"),
                                  (8192, "r32")])
def test_repeated_runs_are_bit_identical(torch, nfft):
    """Race canary (compute-sanitizer is not available on the GPU pool): no atomic touches data (the
    whole-frame kernel counts stage readers with one, which only decides WHO issues the next copy) and
    sums run in a fixed order, so 12 back-to-back runs on two streams must agree bit for bit; a
    missing barrier in an exchange shows up as run-to-run differences."""
    from pyspectrogram_b200 import engine
    variant = None
    if isinstance(nfft, tuple):
        nfft, variant = nfft
    rng = np.random.default_rng(nfft)
    nfr = 37 if nfft <= 4096 else 9
    ncol = 301 if nfft <= 4096 else 24
    n = nfft * nfr * ncol + 64
    x = torch.from_numpy(_recording(rng, n)).cuda()
    starts = torch.from_numpy((np.arange(ncol) * nfft * nfr + np.arange(ncol) % 2).astype(np.int64)).cuda()
    plan = engine.StiPlan(nfft)
    ref = None
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    torch.cuda.synchronize()
    try:
        engine.set_variant(variant)
        for i in range(12):
            with torch.cuda.stream(streams[i % 2]):
                lin, db = plan.run(x, starts, nfr, nfft, want_lin=True, want_db=True)
            torch.cuda.synchronize()
            if ref is None:
                ref = (lin.clone(), db.clone())
            else:
                assert torch.equal(lin, ref[0]) and torch.equal(db, ref[1]), (nfft, i, plan.variant)
    finally:
        engine.set_variant(None)


# ---------------------------------------------------------------------------------------------
# raw integer IQ ingest
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("nfft", [64, 256, 512, 1024, 2048, 4096, 8192, 16384])
@pytest.mark.parametrize("kind", ["int16", "int8"])
def test_raw_integer_iq_matches_reference_on_normalised_samples(torch, nfft, kind):
    """The reference casts the stored integers to complex64 and divides by get_ref's full scale
    (drfProc.py:124-129, :182-201) before the path; the GPU reads the integers and folds 1/ref into
    the epilogue.  Oracle: float64 PSD of x/ref.  Contiguous (TMA) and interleaved (LDG) layouts."""
    from pyspectrogram_b200 import engine
    from pyspectrogram_b200.drfProc import get_ref
    rng = np.random.default_rng(nfft + len(kind))
    dt, amp, props = (np.int16, 3000, dict(H5Tget_class=0, H5Tget_precision=16, H5Tget_size=2)) if kind == "int16" else \
                     (np.int8, 40, dict(H5Tget_class=0, H5Tget_precision=8, H5Tget_size=1))
    ref = get_ref(props)
    nfr, ncol = 3, 6
    n = nfft * nfr * ncol + nfft + 13
    tone = np.exp(2j * np.pi * 0.123 * np.arange(n)) * amp * 3
    xc = (rng.standard_normal(n) + 1j * rng.standard_normal(n)) * amp + tone
    raw = np.stack([np.clip(np.round(xc.real), -32000, 32000), np.clip(np.round(xc.imag), -32000, 32000)], axis=1)
    if kind == "int8":
        raw = np.clip(raw, -127, 127)
    raw = raw.astype(dt)
    x = ((raw[:, 0].astype(np.float32) + 1j * raw[:, 1].astype(np.float32)) / np.float32(ref)).astype(np.complex64)
    starts = (np.arange(ncol) * nfft * nfr + np.array([0, 1, 2, 3, 5, 7])).astype(np.int64)
    plan = engine.StiPlan(nfft)
    expect = _oracle_columns(x, starts, nfft, nfr, nfft)
    lin, db = plan.run(torch.from_numpy(raw).cuda(), torch.from_numpy(starts).cuda(), nfr, nfft,
                       in_scale=1.0 / ref, want_lin=True, want_db=True)
    assert_psd_close(lin.cpu().numpy()[0].T, expect.T, noise_like=False, what=f"{kind} contiguous {plan.variant}")
    assert_db_close(db.cpu().numpy()[0].T, 10 * np.log10(expect.T.astype(np.float32) + np.float32(1e-15)),
                    ref_lin=expect.T, what=f"{kind} dB")
    if nfft >= 256 and nfft <= 8192:
        assert plan.variant.endswith("_i16" if kind == "int16" else "_i8"), plan.variant
    # two interleaved sub-channels [sample][sub][re,im]: strided loader
    raw2 = np.stack([raw, raw[::-1]], axis=1).copy()
    lin2, _ = plan.run(torch.from_numpy(raw2).cuda(), torch.from_numpy(starts * 2).cuda(), nfr, nfft,
                       sample_stride=2, sub_stride=1, nsub=2, in_scale=1.0 / ref)
    assert_psd_close(lin2.cpu().numpy()[0].T, expect.T, noise_like=False, what=f"{kind} strided {plan.variant}")
    # host entry point with the structured dtype Digital RF uses
    if kind == "int16":
        st = np.empty(n, dtype=np.dtype([("r", np.int16), ("i", np.int16)]))
        st["r"], st["i"] = raw[:, 0], raw[:, 1]
        res = plan.host(st, starts, nfr, nfft, in_scale=1.0 / ref, want=("lin",))
        assert_psd_close(res["lin"][0].T, expect.T, noise_like=False, what="int16 structured host")


def test_drop_in_accepts_raw_iq_and_processor_raw_ingest(dp):
    """sti_proc_data(raw, ..., ref=...) == sti_proc_data(raw_as_complex/ref, ...), and a DrfProcessor
    with raw_ingest=True emits the same dB arrays as the cast-and-divide path."""
    from tests.fake_drf import FakeReader
    rng = np.random.default_rng(4)
    nfft, nint, ntime, nsub = 512, 2, 12, 2
    raw = np.empty((nfft * nint, ntime, nsub), dtype=np.dtype([("r", np.int16), ("i", np.int16)]))
    raw["r"] = rng.integers(-9000, 9000, raw.shape)
    raw["i"] = rng.integers(-9000, 9000, raw.shape)
    ref = 2 ** 15.5
    xc = ((raw["r"].astype(np.float32) + 1j * raw["i"].astype(np.float32)) / np.float32(ref)).astype(np.complex64)
    f0, s0, m0 = dp.sti_proc_data(xc, 1.0e6, nfft)
    f1, s1, m1 = dp.sti_proc_data(raw, 1.0e6, nfft, ref=ref)
    assert np.array_equal(f0, f1) and s1.shape == s0.shape and s1.dtype == np.float32
    assert_psd_close(s1, s0, what="raw drop-in")
    assert_psd_close(m1, m0, what="raw drop-in median")
    # plain integer array with a trailing (re, im) axis
    plain = np.stack([raw["r"], raw["i"]], axis=-1)
    f2, s2, m2 = dp.sti_proc_data(plain, 1.0e6, nfft, ref=ref, integrate=True, raw_pairs=True)
    f3, s3, m3 = dp.sti_proc_data(xc, 1.0e6, nfft, integrate=True)
    assert_psd_close(s2, s3, what="raw drop-in mode A")
    # without the keyword a plain integer array is REAL samples, as for the reference (scipy casts it): the
    # (rows, ntime, 2) array above is two real sub-channels, float64 out (drfProc.py:387-396)
    from oracle import ref_port
    real2 = np.ascontiguousarray(plain[:, :, 0, :])
    f4, s4, m4 = dp.sti_proc_data(real2, 1.0e6, nfft)
    f5, s5, m5 = ref_port.sti_mode_r(real2, 1.0e6, nfft)
    assert s4.shape == s5.shape == (nfft, ntime, 2) and s4.dtype.kind == "f"  # (scipy's own result dtype for int16 input varies with its version)
    assert_psd_close(s4, s5, noise_like=False, what="plain int array = real samples")
    # the plot-data entry takes the raw forms too (same dispatch)
    pf_a, sx_a, md_a = dp.sti_plot_data(raw, 1.0e6, nfft, (-200.0, 200.0), ref=ref)
    pf_b, sx_b, md_b = dp.sti_plot_data(xc, 1.0e6, nfft, (-200.0, 200.0))
    pf_c, sx_c, md_c = dp.sti_plot_data(plain, 1.0e6, nfft, (-200.0, 200.0), ref=ref, raw_pairs=True)
    assert np.array_equal(pf_a, pf_b) and sx_a.shape == sx_b.shape == sx_c.shape
    assert np.abs(sx_a - sx_b).max() <= 1e-3 and np.abs(md_a - md_b).max() <= 1e-3
    assert np.array_equal(sx_a, sx_c) and np.array_equal(md_a, md_c)

    n = 1 << 17
    data = np.round((rng.standard_normal((n, nsub)) + 1j * rng.standard_normal((n, nsub))) * 3000).astype(np.complex64)
    out = []
    for raw_ingest in (False, True):
        reader = FakeReader({"ch0": data}, sample_rate=1000000, first_sample=1_700_000_000 * 1000000, int16=True)
        proc = dp.DrfProcessor("file", "/nonexistent", 1, 1024.0, 2.0, 16.0, reader=reader, raw_ingest=raw_ingest)
        out.append(proc.iterate_once(0))
    (t0, fa, sa, ma), (t1, fb, sb, mb) = out
    assert np.array_equal(fa, fb) and sa.shape == sb.shape == (1024, 16, nsub)
    assert np.abs(sa - sb).max() <= 1e-3 and np.abs(ma - mb).max() <= 1e-3


@pytest.mark.parametrize("nfft", [32, 256, 1024, 4096, 8192])
def test_mode_r_multi_column_kernels_match_single_column_kernels(torch, nfft):
    """Mode R launches use kernels that run several column blocks per CTA; bit-identical to the
    one-column-per-CTA kernels (same arithmetic, same order), ragged last block included."""
    ncol = 200003 if nfft <= 256 else 20011
    from pyspectrogram_b200 import _lib, engine
    rng = np.random.default_rng(nfft + 5)
    x = torch.from_numpy(_recording(rng, ncol * 3 + nfft + 7)).cuda()
    starts = torch.from_numpy((np.arange(ncol) * 3).astype(np.int64)).cuda()  # overlapping, odd and even starts
    plan = engine.StiPlan(nfft)
    lin_m, db_m = plan.run(x, starts, 1, nfft, want_lin=True, want_db=True)
    name_m = plan.variant
    try:
        _lib.check(_lib.load().psg_debug_set_mode_r_multi(0))
        lin_s, db_s = plan.run(x, starts, 1, nfft, want_lin=True, want_db=True)
        name_s = plan.variant
    finally:
        _lib.check(_lib.load().psg_debug_set_mode_r_multi(1))
    if name_m.startswith("r32_"):
        # 8192 points, one frame per column: the persistent radix-32 kernel is the default (its frame pipeline runs
        # across columns by construction); a different transform, so agreement within the parity tolerance
        assert nfft == 8192 and name_s.startswith("tma13_"), (name_m, name_s)
        a, b = lin_m[0].double(), lin_s[0].double()
        assert float(((a - b).abs().amax(dim=1) / b.amax(dim=1)).max()) <= 2e-6
        strong = lin_s[0] >= lin_s[0].amax(dim=1, keepdim=True) * 1e-6  # bins within 60 dB of the column peak (tests/parity.py)
        assert float((db_m[0] - db_s[0]).abs()[strong].max()) <= 1e-3
        return
    # small nfft packs many columns into one CTA already: too few column blocks here to batch further
    assert name_m == name_s + "_m" or (nfft < 1024 and name_m == name_s), (name_m, name_s)
    assert torch.equal(lin_m, lin_s) and torch.equal(db_m, db_s)


def test_generic_kernel_cross_checks_tuned(torch):
    from pyspectrogram_b200 import engine
    rng = np.random.default_rng(5)
    nfft, nfr, ncol = 4096, 6, 12
    x = _recording(rng, nfft * nfr * ncol + 17)
    starts = (np.arange(ncol) * nfft * nfr + (np.arange(ncol) % 2)).astype(np.int64)
    plan = engine.StiPlan(nfft)
    dx, ds = torch.from_numpy(x).cuda(), torch.from_numpy(starts).cuda()
    a, _ = plan.run(dx, ds, nfr, nfft)
    try:
        engine.set_force_generic(True)
        b, _ = plan.run(dx, ds, nfr, nfft)
        assert plan.variant == "generic_radix2"
    finally:
        engine.set_force_generic(False)
    e = psd_errors(a.cpu().numpy()[0].T, b.cpu().numpy()[0].T)
    assert e["col"] <= 2e-6, e


def test_split_columns_and_strided_subchannels(torch):
    """Long integration (column split over several CTAs + finalize) on an interleaved
    [sample][nsub] recording (strided LDG loader), against the oracle."""
    from pyspectrogram_b200 import engine
    from oracle import np_oracle
    rng = np.random.default_rng(11)
    nfft, nfr, ncol, nsub = 1024, 700, 3, 2
    n = nfft * nfr * ncol + 9
    x = (rng.standard_normal((n, nsub)) + 1j * rng.standard_normal((n, nsub))).astype(np.complex64) * np.float32(1e-2)
    starts = (np.arange(ncol) * nfft * nfr + np.array([0, 5, 9])).astype(np.int64)
    plan = engine.StiPlan(nfft)
    lin, _ = plan.run(torch.from_numpy(x).cuda(), torch.from_numpy(starts * nsub).cuda(), nfr, nfft,
                      sample_stride=nsub, sub_stride=1, nsub=nsub, in_scale=0.5)
    got = lin.cpu().numpy()
    for s in range(nsub):
        ref = np.stack([np_oracle.column_power(x[st:, s] * 0.5, nfft, nfr) for st in starts])
        assert_psd_close(got[s].T, ref.T, what=f"split sub {s}")


def test_median_is_exact_order_statistic(torch):
    """np.median(sxx, axis=1) bit for bit: odd/even column counts, repeated values, every tile
    width of the kernel (32/16/8 bins per CTA), bins that do not fill the last CTA, and the
    untiled fallback for more columns than fit in shared memory."""
    from pyspectrogram_b200 import engine
    rng = np.random.default_rng(3)
    plan = engine.StiPlan(256)
    for nsub, ncol, nfft in ((2, 1, 256), (2, 2, 256), (2, 3, 256), (2, 100, 256), (2, 101, 256), (2, 1000, 256),
                             (1, 1000, 8192), (1, 3600, 512), (3, 7001, 40), (1, 6, 5)):
        img = rng.random((nsub, ncol, nfft), dtype=np.float32) ** 4
        img[0, : ncol // 2, 3] = img[0, 0, 3]  # repeated values
        lin, db = plan.median(torch.from_numpy(img).cuda(), want_lin=True, want_db=True)
        ref = np.median(img, axis=1)
        assert np.array_equal(lin.cpu().numpy(), ref), (nsub, ncol, nfft)
        assert np.abs(db.cpu().numpy() - 10 * np.log10(ref + np.float32(1e-15))).max() <= 1e-4
    # the bisection starts below the leading bits all keys of a tile share: narrow value ranges (long
    # common prefix), a constant image (no differing bit at all), mixed signs (no common prefix), one
    # outlier per row, with column counts around the 4-column vector loads
    for ncol in (4, 5, 7, 8, 64, 99, 3600):
        for kind in ("narrow", "const", "signed", "outlier"):
            img = rng.random((2, ncol, 72), dtype=np.float32)
            if kind == "narrow":
                img = np.float32(1.0) + img * np.float32(1e-6)
            elif kind == "const":
                img[:] = np.float32(0.3)
            elif kind == "signed":
                img = img - np.float32(0.5)
            else:
                img = np.float32(2.0) + img * np.float32(1e-3)
                img[:, 0, :] = np.float32(1e-30)
            lin, _ = plan.median(torch.from_numpy(np.ascontiguousarray(img)).cuda(), want_lin=True, want_db=False)
            assert np.array_equal(lin.cpu().numpy(), np.median(img, axis=1)), (ncol, kind)


def test_epoch_sized_frame_starts_are_reproduced(torch):
    """np.linspace(..., dtype=int) at ~1.7e17 quantises starts to multiples of 32 (SURVEY section 0,
    trap 2): the table the kernel consumes is numpy's own, and columns land exactly there."""
    from pyspectrogram_b200 import engine
    meta = json.load(open(os.path.join(GOLDEN, "meta.json")))
    case = [c for c in meta["frame_starts"] if c["st"] > 1e15][0]
    n_st = engine.frame_starts(case["st"], case["en"], case["nfft"], case["nint"], case["ntime"])
    assert [int(v) for v in n_st[:4]] == case["first"] and [int(v) for v in n_st[-4:]] == case["last"]
    # a small recording addressed with epoch-sized absolute indices
    nfft, nint, ntime = 256, 2, 16
    st = 170000000000000000
    en = st + 40 * nfft * nint + 77
    starts = engine.frame_starts(st, en, nfft, nint, ntime)
    rng = np.random.default_rng(9)
    x = _recording(rng, en - st)
    plan = engine.StiPlan(nfft)
    rel = (starts - st).astype(np.int64)
    lin, _ = plan.run(torch.from_numpy(x).cuda(), torch.from_numpy(rel).cuda(), nint, nfft)
    ref = _oracle_columns(x, rel, nfft, nint, nfft)
    assert_psd_close(lin.cpu().numpy()[0].T, ref.T, noise_like=False, what="epoch starts")


def test_drfprocessor_iteration_with_fake_reader(dp):
    from tests.fake_drf import FakeReader
    from oracle import ref_port
    rng = np.random.default_rng(21)
    n, nsub = 1 << 18, 2
    data = ((rng.standard_normal((n, nsub)) + 1j * rng.standard_normal((n, nsub))) * 3000).astype(np.complex64)
    reader = FakeReader({"ch0": data}, sample_rate=1000000, first_sample=1_700_000_000 * 1000000, int16=True)
    proc = dp.DrfProcessor("file", "/nonexistent", 3, 1024.0, 2.0, 20.0, reader=reader)
    got = {}
    proc.signals.iterated.connect(lambda i, tab, t, f, s, m: got.update(i=i, tab=tab, t=t, f=f, s=s, m=m))
    time_ar, f, sdb, mdb = proc.iterate_once(0)
    assert got["tab"] == 3 and got["s"] is sdb and sdb.shape == (1024, 20, nsub) and mdb.shape == (1024, nsub)
    # the reference pipeline on the CPU: read_sti -> sti_proc_data -> dB
    sr = proc.drfIn.sr_dict["ch0"]
    s_samp = dp._time_to_sample(proc.bnds[0], sr)
    e_samp = dp._time_to_sample(proc.bnds[1], sr)
    oin = ref_port.read_sti_from_array(data, s_samp, e_samp, 1024, 2, 20, ref=2 ** 15.5,
                                       first_sample=reader.first)
    fr, sr_, mr = ref_port.sti_mode_r(oin[1].astype(np.complex64), Fraction(1000000, 1), 1024)
    assert np.array_equal(f, fr)
    assert_db_close(sdb, ref_port.to_dbfs(sr_), ref_lin=sr_, what="processor dB")
    assert_db_close(mdb, ref_port.to_dbfs(mr), ref_lin=mr, what="processor median dB")


@pytest.mark.parametrize("nsub", [1, 2])
@pytest.mark.parametrize("raw_ingest", [False, True])
def test_resident_recording_cache_matches_per_bin_reads(dp, nsub, raw_ingest):
    """the worker loop on a device-resident window gives the same dB arrays
    as the reference's read_sti gather, reads every sample once, and reads nothing when the same
    window is processed again."""
    from tests.fake_drf import FakeReader
    rng = np.random.default_rng(31 + nsub)
    n = 1 << 17
    shape = (n, nsub) if nsub > 1 else (n,)
    data = np.round((rng.standard_normal(shape) + 1j * rng.standard_normal(shape)) * 3000).astype(np.complex64)
    for integrate in (False, True):
        outs = []
        for resident in (False, True):
            reader = FakeReader({"ch0": data}, sample_rate=1000000, first_sample=1_700_000_000 * 1000000, int16=True)
            proc = dp.DrfProcessor("file", "/nonexistent", 1, 512.0, 4.0, 50.0, reader=reader, integrate=integrate,
                                   raw_ingest=raw_ingest, resident=resident)
            outs.append(proc.iterate_once(0))
            if resident:
                cache = proc._cache["ch0"]
                assert cache.samples_read == cache.hi - cache.lo <= n
                nreads = len(reader.reads)
                from pyspectrogram_b200 import engine
                launches = engine.launch_count()
                again = proc.iterate_once(1)
                assert len(reader.reads) == nreads and cache.samples_read == cache.hi - cache.lo
                assert np.array_equal(again[2], outs[-1][2]) and np.array_equal(again[3], outs[-1][3])
                # incremental recompute: same window, same settings -> no kernel launched at all
                assert engine.launch_count() == launches and proc.recomputes_skipped == 1
                proc.updatesettings_slot(256.0, 4.0, 50.0, proc.drfIn.time_bnds[0], proc.drfIn.time_bnds[1])
                changed = proc.iterate_once(2)
                assert changed[2].shape[0] == 256 and engine.launch_count() > launches and proc.recomputes_skipped == 1
        (t0, f0, s0, m0), (t1, f1, s1, m1) = outs
        assert np.array_equal(f0, f1) and np.array_equal(t0, t1)
        assert s1.shape == (512, 50, nsub) and m1.shape == (512, nsub)
        s0 = s0.reshape(s1.shape)
        m0 = m0.reshape(m1.shape)
        assert np.abs(s0 - s1).max() <= 1e-3 and np.abs(m0 - m1).max() <= 1e-3, (integrate, nsub, raw_ingest)
        # oracle leg: the reference pipeline on the CPU (read_sti -> sti_proc_data / per-bin averaging -> dB)
        from oracle import ref_port
        sr = proc.drfIn.sr_dict["ch0"]
        s_samp = dp._time_to_sample(proc.drfIn.time_bnds[0], sr)
        e_samp = dp._time_to_sample(proc.drfIn.time_bnds[1], sr)
        _, dout = ref_port.read_sti_from_array(data, s_samp, e_samp, 512, 4, 50, ref=2 ** 15.5, first_sample=reader.first)
        fr, sr_, mr = (ref_port.sti_mode_a if integrate else ref_port.sti_mode_r)(dout.astype(np.complex64), sr, 512)
        assert np.array_equal(f1, fr)
        assert_db_close(s1, ref_port.to_dbfs(sr_).reshape(s1.shape), ref_lin=sr_.reshape(s1.shape), what="resident dB vs oracle")
        assert_db_close(m1, ref_port.to_dbfs(mr).reshape(m1.shape), ref_lin=mr.reshape(m1.shape), what="resident median dB vs oracle")


def test_resident_cache_sliding_window_reads_only_the_new_tail(torch):
    from pyspectrogram_b200 import engine
    rng = np.random.default_rng(2)
    rec = (rng.standard_normal(50000) + 1j * rng.standard_normal(50000)).astype(np.complex64)
    calls = []

    def read(st, n):
        calls.append((st, n))
        return rec[st:st + n]

    cache = engine.RecordingCache(0)
    buf, base = cache.ensure(read, 1000, 21000)
    assert base == 1000 and calls == [(1000, 20000)]
    assert np.array_equal(buf[:20000, 0].cpu().numpy(), rec[1000:21000])
    buf, base = cache.ensure(read, 5000, 20000)  # inside the window: nothing read
    assert base == 1000 and len(calls) == 1
    buf, base = cache.ensure(read, 9000, 30000)  # slid forward: only [21000, 30000) is read
    assert base == 9000 and calls[-1] == (21000, 9000) and cache.samples_read == 29000
    assert np.array_equal(buf[:21000, 0].cpu().numpy(), rec[9000:30000])
    buf, base = cache.ensure(read, 40000, 45000)  # disjoint: re-read
    assert base == 40000 and calls[-1] == (40000, 5000)
    assert np.array_equal(buf[:5000, 0].cpu().numpy(), rec[40000:45000])
    # a small slide fits the slack behind the resident samples: appended in place, same storage, same base
    ptr, before = buf.data_ptr(), cache.appended_in_place
    buf, base = cache.ensure(read, 41000, 46000)
    assert base == 40000 and calls[-1] == (45000, 1000) and buf.data_ptr() == ptr and cache.appended_in_place == before + 1
    assert np.array_equal(buf[1000:6000, 0].cpu().numpy(), rec[41000:46000])


def test_checked_run_clamps_and_flags_a_bad_offset_table(torch):
    """psg_sti_run_checked (the recording's extent travels with it): a good table gives
    the unchecked result bit for bit; a table that reaches outside the array is clamped on the device and
    flagged (IndexError here), never read out of bounds; a column longer than the array is refused."""
    from pyspectrogram_b200 import engine
    rng = np.random.default_rng(77)
    nfft, nfr = 1024, 3
    x = _recording(rng, nfft * nfr * 5 + 100)
    xd = torch.from_numpy(x).cuda()
    good = torch.from_numpy((np.arange(5) * nfft * nfr + np.array([0, 3, 1, 100, 7])).astype(np.int64)).cuda()
    plan = engine.StiPlan(nfft)
    lin0, db0 = plan.run(xd, good, nfr, nfft, want_lin=True, want_db=True)
    lin1, db1 = plan.run(xd, good, nfr, nfft, want_lin=True, want_db=True, validate=True)
    assert torch.equal(lin0, lin1) and torch.equal(db0, db1)
    for bad_off in (x.size - nfft * nfr + 1, -5, 1 << 40):
        bad = good.clone()
        bad[2] = bad_off
        with pytest.raises(IndexError):
            plan.run(xd, bad, nfr, nfft, validate=True)
    with pytest.raises(ValueError):  # one column needs more samples than the array holds
        plan.run(xd[: nfft * 2], good[:1], nfr, nfft, validate=True)
    with pytest.raises(ValueError):  # a sliced view: the kernels address storage, not the tensor's strides
        plan.run(xd[::2], good[:1], 1, nfft)
    # raw integer pairs: the extent is counted in complex elements
    raw = torch.from_numpy(np.stack([x.real, x.imag], axis=1) * 1000).to(torch.int16).cuda()
    lin2, _ = plan.run(raw, good, nfr, nfft, validate=True)
    bad = good.clone()
    bad[4] = x.size - nfft * nfr + 1
    with pytest.raises(IndexError):
        plan.run(raw, bad, nfr, nfft, validate=True)


def test_seven_worker_threads_share_the_library(dp):
    """The viewer runs up to 7 workers at once (drfview.py:177-178): concurrent calls from Python
    threads -- same nfft (one shared, locked plan) and different nfft (separate plans) -- give the
    single-threaded results."""
    import threading
    rng = np.random.default_rng(17)
    cases = []
    for k, nfft in enumerate([256, 1024, 1024, 4096, 4096, 1000, 16384]):
        d1 = ((rng.standard_normal((nfft * 2, 12, 2)) + 1j * rng.standard_normal((nfft * 2, 12, 2))) * 1e-2).astype(np.complex64)
        cases.append((nfft, d1, k % 2 == 1))
    expect = [dp.sti_proc_data_db(d1, 1.0e6, nfft, integrate=integ) for nfft, d1, integ in cases]
    got = [None] * len(cases)
    errs = []

    def work(i):
        try:
            nfft, d1, integ = cases[i]
            for _ in range(5):
                got[i] = dp.sti_proc_data_db(d1, 1.0e6, nfft, integrate=integ)
        except Exception as exc:  # pragma: no cover
            errs.append((i, exc))

    threads = [threading.Thread(target=work, args=(i,)) for i in range(len(cases))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errs, errs
    for (f0, s0, m0), (f1, s1, m1) in zip(expect, got):
        assert np.array_equal(f0, f1) and np.array_equal(s0, s1) and np.array_equal(m0, m1)


def test_device_path_workers_at_one_nfft_do_not_share_scratch(dp, torch):
    """the device-path calls (run / median / gather: enqueue, no sync) of two viewer workers at the SAME
    nfft must not race on a shared plan's scratch.  Plans are per thread (engine.get_plan): threads mixing
    sti_plot_data, the resident worker loop's path and sti_proc_data_db at one nfft -- a long integration whose
    columns are split over CTAs (partial-sum scratch) -- reproduce the single-threaded results bit for bit."""
    import threading
    from pyspectrogram_b200 import engine
    rng = np.random.default_rng(23)
    nfft, nint, ntime = 4096, 96, 6   # few columns x many frames: split columns + finalize through plan scratch
    cases = []
    for k in range(6):
        d1 = ((rng.standard_normal((nfft * nint, ntime)) + 1j * rng.standard_normal((nfft * nint, ntime))) * 1e-2).astype(np.complex64)
        cases.append(d1)

    def call(k):
        d1 = cases[k]
        if k % 3 == 0:
            return dp.sti_plot_data(d1, 1.0e6, nfft, (-400.0, 400.0), integrate=True)
        if k % 3 == 1:
            return dp.sti_proc_data_db(d1, 1.0e6, nfft, integrate=True)
        # the device path on a side stream, as a worker with its own stream would use it
        plan = engine.get_plan(nfft)
        st = torch.cuda.Stream()
        with torch.cuda.stream(st):
            x = torch.from_numpy(np.ascontiguousarray(d1.T).reshape(-1)).cuda()
            offs = torch.arange(ntime, dtype=torch.int64, device="cuda") * (nfft * nint)
            lin, db = plan.run(x, offs, nint, nfft, want_lin=True, want_db=True)
            med, _ = plan.median(lin)
            out = (lin.cpu().numpy(), db.cpu().numpy(), med.cpu().numpy())
        return out

    expect = [call(k) for k in range(len(cases))]
    got = [None] * len(cases)
    errs = []
    plans = {}

    def work(k):
        try:
            plans[k] = engine.get_plan(nfft)
            for _ in range(6):
                got[k] = call(k)
        except Exception as exc:  # pragma: no cover
            errs.append((k, exc))

    threads = [threading.Thread(target=work, args=(k,)) for k in range(len(cases))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errs, errs
    assert len({id(p) for p in plans.values()}) == len(cases)  # one plan per thread
    for e, g in zip(expect, got):
        for a, b in zip(e, g):
            assert np.array_equal(a, b)


def test_gui_maximum_counts(torch):
    """The viewer's spin boxes go up to nint = 100000 and ntime = 100000 (drfview.py:488-503): a
    column of 100000 frames (split over CTAs, fp64 finalize) and an image of 100000 columns."""
    from pyspectrogram_b200 import engine
    rng = np.random.default_rng(8)
    nfft = 32
    plan = engine.StiPlan(nfft)
    # nint = 100000, three columns
    nfr, ncol = 100000, 3
    x = _recording(rng, nfft * nfr * ncol + 5)
    starts = (np.arange(ncol) * nfft * nfr + np.array([0, 1, 3])).astype(np.int64)
    lin, _ = plan.run(torch.from_numpy(x).cuda(), torch.from_numpy(starts).cuda(), nfr, nfft)
    ref = _oracle_columns(x, starts, nfft, nfr, nfft)
    assert_psd_close(lin.cpu().numpy()[0].T, ref.T, noise_like=False, what="nint=100000")
    # ntime = 100000 columns of one frame (Mode R), overlapping starts one sample apart
    ncol = 100000
    x = _recording(rng, ncol + nfft)
    starts = np.arange(ncol, dtype=np.int64)
    lin, db = plan.run(torch.from_numpy(x).cuda(), torch.from_numpy(starts).cuda(), 1, nfft, want_lin=True, want_db=True)
    pick = np.array([0, 1, 2, 4999, 50000, 99998, 99999])
    ref = _oracle_columns(x, starts[pick], nfft, 1, nfft)
    assert_psd_close(lin.cpu().numpy()[0][pick].T, ref.T, noise_like=False, what="ntime=100000")
    med, _ = plan.median(lin)
    assert np.array_equal(med.cpu().numpy(), np.median(lin.cpu().numpy(), axis=1))


def test_c_abi_from_plain_c(tmp_path):
    """The boundary is a C ABI: a program compiled by gcc from include/psg_b200.h alone (no CUDA
    headers, no torch, no Python) drives the path through psg_sti_host and checks it against a
    naive float64 DFT."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "abi_smoke")
    libdir = os.path.join(root, "pyspectrogram_b200")
    subprocess.run(["gcc", "-O2", "-I", os.path.join(root, "include"), os.path.join(root, "tests", "c", "abi_smoke.c"),
                    "-o", exe, "-L", libdir, "-lpsgb200", "-lm", f"-Wl,-rpath,{libdir}"], check=True)
    res = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "c-abi smoke" in res.stdout and "6_8x8" in res.stdout


@pytest.mark.parametrize("nchan,nfft,ntime,nint,world", [(1, 65536, 12, 2, 4), (8, 16384, 6, 3, 8), (3, 1024, 50, 4, 4)])
def test_sharded_columns_equal_the_single_gpu_image(torch, nchan, nfft, ntime, nint, world):
    """SURVEY.md section 8(e) on one GPU: every rank's shard (whole channels when nchan >= world, as in
    BASELINE config 3; contiguous time-bin ranges otherwise, config 4) computed on its own and
    concatenated in rank order is bit-identical to the image of one launch over all columns, and the
    time-median of the assembled image equals numpy's."""
    from pyspectrogram_b200 import dist as pdist
    from pyspectrogram_b200 import engine
    rng = np.random.default_rng(nfft + world)
    n = nfft * nint * ntime + 3 * nfft
    chans = [torch.from_numpy(_recording(rng, n)).cuda() for _ in range(nchan)]
    starts = engine.frame_starts(5, n, nfft, nint, ntime).astype(np.int64)
    plan = engine.StiPlan(nfft)
    whole = torch.stack([plan.run(ch, torch.from_numpy(starts).cuda(), nint, nfft)[0][0] for ch in chans])  # [chan][t][f]
    pieces = pdist.shard_plan(nchan, ntime, world)
    assert sum(hi - lo for r in pieces for (_, lo, hi) in r) == nchan * ntime
    parts = []
    for rank_pieces in pieces:
        for (c, lo, hi) in rank_pieces:
            lin, _ = plan.run(chans[c], torch.from_numpy(starts[lo:hi]).cuda(), nint, nfft)
            parts.append(lin[0])
    assembled = torch.cat(parts, dim=0).reshape(nchan, ntime, nfft)
    assert torch.equal(assembled, whole)
    med, _ = plan.median(assembled)
    assert np.array_equal(med.cpu().numpy(), np.median(whole.cpu().numpy(), axis=1))


def test_large_workload_properties(torch):
    """Size-independent properties at a bench-like size (1 GiB of IQ, nfft=4096, nint=128):
    Parseval (sum of the PSD column == mean windowed frame energy * N / sum(w)^2), a unit tone
    lands at exactly bin nfft/2+k with 0 dBFS, and linearity in the input scale."""
    from pyspectrogram_b200 import engine
    nfft, nint, ntime = 4096, 128, 256
    n = nfft * nint * ntime
    dev = torch.device("cuda")
    gen = torch.Generator(device=dev).manual_seed(1)
    iq = torch.empty(n, dtype=torch.complex64, device=dev)
    v = torch.view_as_real(iq)
    v.normal_(0.0, 1e-2, generator=gen)
    k = 37
    ph = (torch.arange(n, device=dev, dtype=torch.int64) * k % nfft).to(torch.float32) * (2 * np.pi / nfft)
    v[:, 0] += torch.cos(ph)
    v[:, 1] += torch.sin(ph)
    starts = torch.arange(ntime, device=dev, dtype=torch.int64) * (nfft * nint)
    plan = engine.StiPlan(nfft)
    lin, db = plan.run(iq, starts, nint, nfft, want_lin=True, want_db=True)
    w = torch.from_numpy(plan.window_table().astype(np.float64)).to(dev)  # w / sum(w)
    energy = (torch.view_as_real(iq).double().pow(2).sum(-1).reshape(ntime, nint, nfft) * w.pow(2)).sum(-1).mean(-1)
    pars = lin[0].double().sum(-1)
    assert float(((pars - nfft * energy).abs() / (nfft * energy)).max()) <= 1e-5
    assert int(lin[0].argmax(-1).unique().item()) == nfft // 2 + k
    assert float((db[0, :, nfft // 2 + k]).abs().max()) <= 2e-3  # unit tone + noise -> 0 dBFS
    lin2, _ = plan.run(iq, starts, nint, nfft, in_scale=0.25)
    assert float(((lin2 - lin * 0.0625).abs() / lin).max()) <= 1e-6


# ---------------------------------------------------------------------------------------------
# viewer-side reductions on the device
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("nsub,ncol,nfft", [(1, 1, 33), (2, 7, 100), (1, 100, 1024), (3, 1000, 256), (1, 9000, 64)])
def test_minmax_over_time_is_exact(torch, nsub, ncol, nfft):
    """np.min / np.max over the time axis of the same image, bit for bit (order statistics)."""
    from pyspectrogram_b200 import engine
    rng = np.random.default_rng(nsub * 1000 + ncol)
    img = rng.random((nsub, ncol, nfft), dtype=np.float32) ** 8
    plan = engine.StiPlan(max(nfft, 2))
    mn, mx, mn_db, mx_db = plan.minmax(torch.from_numpy(img).cuda(), want_db=True)
    assert np.array_equal(mn.cpu().numpy(), img.min(axis=1)) and np.array_equal(mx.cpu().numpy(), img.max(axis=1))
    assert np.abs(mn_db.cpu().numpy() - 10 * np.log10(img.min(axis=1) + np.float32(1e-15))).max() <= 1e-3
    assert np.abs(mx_db.cpu().numpy() - 10 * np.log10(img.max(axis=1) + np.float32(1e-15))).max() <= 1e-3
    img[0, ncol // 2, 5] = np.nan  # np.min / np.max propagate NaN
    mn, mx, _, _ = plan.minmax(torch.from_numpy(img).cuda())
    assert np.array_equal(mn.cpu().numpy(), img.min(axis=1), equal_nan=True)
    assert np.array_equal(mx.cpu().numpy(), img.max(axis=1), equal_nan=True)


def test_proc_data_minmax_matches_reference_golden(dp):
    """proc_data(minmax=True): the first four outputs are the reference's, min / max are np.min / np.max
    of the reference's own STI within the PSD tolerance and exact on the returned STI."""
    from oracle import ref_port
    g = load("proc_1024")
    t_out, f, sxx, med, mn, mx = dp.proc_data(g["x"], float(g["sr"]), int(g["nfft"]), float(g["dt"]), minmax=True)
    assert np.array_equal(t_out, g["t_out"]) and np.array_equal(f, g["f"])
    assert_psd_close(sxx, g["sxx"], noise_like=False, what="proc_1024 minmax sxx")
    assert_psd_close(med, g["med"], noise_like=False, what="proc_1024 minmax median")
    rmn, rmx = ref_port.proc_data_min_max(g["sxx"])
    assert_psd_close(mn, rmn, noise_like=False, what="proc_1024 min")
    assert_psd_close(mx, rmx, noise_like=False, what="proc_1024 max")
    assert np.array_equal(mn, sxx.min(axis=-1)) and np.array_equal(mx, sxx.max(axis=-1))
    assert mn.dtype == g["sxx"].dtype and mn.shape == (int(g["nfft"]),)


@pytest.mark.parametrize("shape,count", [((3, 40, 1024), 100), ((1, 1, 65536), 32768), ((2, 2048), 7), ((5,), 5)])
def test_gather_bins_and_clip_are_exact(torch, shape, count):
    from oracle import ref_port
    from pyspectrogram_b200 import engine
    rng = np.random.default_rng(count)
    img = (rng.standard_normal(shape) * 40 - 60).astype(np.float32)
    idx = np.sort(rng.choice(shape[-1], count, replace=False))
    plan = engine.StiPlan(64)
    t = torch.from_numpy(img).cuda()
    assert np.array_equal(plan.gather_bins(t, idx).cpu().numpy(), img[..., idx])
    got = plan.gather_bins(t, idx, clamp=(-90.0, -30.0)).cpu().numpy()
    assert np.array_equal(got, ref_port.clip_to_colour_range(img[..., idx], (-90.0, -30.0)))
    assert np.array_equal(plan.gather_bins(t, [-1, 0]).cpu().numpy(), img[..., [-1, 0]])
    with pytest.raises(IndexError):
        plan.gather_bins(t, [shape[-1]])


@pytest.mark.parametrize("nfft,cfrange,maxn", [(1024, (-400.0, 400.0), 2 ** 15), (65536, (-500.0, 500.0), 2 ** 15),
                                              (65536, (-100.0, 250.0), 1000), (4096, (12.0, 13.0), 2 ** 15)])
def test_sti_plot_data_matches_viewer_selection(dp, nfft, cfrange, maxn):
    """The reduced arrays equal the full dB result indexed with the viewer's plotindices
    (drfview.py:1005-1023) and clipped like the PNG export (drfview.py:1515-1518)."""
    from oracle import ref_port
    rng = np.random.default_rng(nfft)
    ntime = 6
    d1 = _recording(rng, nfft * ntime).reshape(nfft, ntime)
    sr = 1.0e6
    f, sdb, mdb = dp.sti_proc_data_db(d1, sr, nfft)
    pidx, pfreqs, fscale = ref_port.plot_indices(f, cfrange, maxn)
    idx, freqs, fs2 = dp.plot_indices(f, cfrange, maxn)
    assert list(idx) == list(pidx) and np.array_equal(freqs, pfreqs) and fs2 == fscale and len(idx) <= maxn
    pf, sxx, med = dp.sti_plot_data(d1, sr, nfft, cfrange, max_nfreqs=maxn)
    assert np.array_equal(pf, pfreqs) and np.array_equal(sxx, sdb[pidx]) and np.array_equal(med, mdb[pidx])
    pf, sxx, med = dp.sti_plot_data(d1, sr, nfft, cfrange, max_nfreqs=maxn, crange=(-100.0, -50.0))
    assert np.array_equal(sxx, ref_port.clip_to_colour_range(sdb[pidx], (-100.0, -50.0)))
    assert np.array_equal(med, ref_port.clip_to_colour_range(mdb[pidx], (-100.0, -50.0)))


# ---------------------------------------------------------------------------------------------
# streamed host path (recordings larger than the staging chunk)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("nfft,nfr,ncol,nsub,kind", [(1024, 4, 50, 1, "c64"), (4096, 3, 21, 1, "c64"), (256, 5, 64, 2, "c64"),
                                                     (2048, 2, 33, 1, "i16"), (16384, 2, 9, 1, "c64")])
def test_host_path_streams_in_column_chunks(nfft, nfr, ncol, nsub, kind):
    """psg_sti_host with the chunk size shrunk so that the recording crosses PCIe in several column
    groups (copy of group j+1 overlapping the kernels of group j): same image, median and dB as the
    float64 oracle, and the same as the one-piece path within fp32 summation-order noise."""
    from pyspectrogram_b200 import engine
    rng = np.random.default_rng(nfft + ncol)
    n = nfft * nfr * ncol + nfft + 7
    x = _recording(rng, n * nsub).reshape(n, nsub)  # interleaved sub-channels (sample_stride = nsub)
    starts = (np.arange(ncol) * nfft * nfr + np.arange(ncol) % 2).astype(np.int64)
    feed, in_scale = x, 1.0
    if kind == "i16":
        amp = 20000.0 * 8
        feed = np.stack([np.round(x.real * amp), np.round(x.imag * amp)], axis=-1).astype(np.int16)
        in_scale = 1.0 / amp
        x = ((feed[..., 0].astype(np.float32) + 1j * feed[..., 1].astype(np.float32)) * np.float32(in_scale)).astype(np.complex64)
    plan = engine.StiPlan(nfft)
    kw = dict(sample_stride=nsub, sub_stride=1, nsub=nsub, in_scale=in_scale, want=("lin", "db", "med", "med_db"))
    flat = feed.reshape(-1) if kind == "c64" else feed.reshape(-1, 2)
    whole = plan.host(flat, starts * nsub, nfr, nfft, **kw)
    try:
        engine.set_host_chunk(max(1 << 16, x.nbytes // 7))
        got = plan.host(flat, starts * nsub, nfr, nfft, **kw)
    finally:
        engine.set_host_chunk(1 << 30)
    for s in range(nsub):
        ref = _oracle_columns(x[:, s], starts, nfft, nfr, nfft)
        assert_psd_close(got["lin"][s].T, ref.T, noise_like=False, what=f"streamed host sub {s}")
        assert_db_close(got["db"][s].T, 10 * np.log10(ref.T.astype(np.float32) + np.float32(1e-15)), ref_lin=ref.T,
                        what=f"streamed host dB sub {s}")
        assert_psd_close(got["med"][s], np.median(ref, axis=0), noise_like=False, what="streamed host median")
    assert np.allclose(got["lin"], whole["lin"], rtol=1e-5, atol=0) and np.allclose(got["med"], whole["med"], rtol=1e-5, atol=0)
