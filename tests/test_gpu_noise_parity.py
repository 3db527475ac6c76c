"""Per-bin parity of every DEFAULT kernel on noise-like input (run with ``-m gpu`` on a B200).

SURVEY.md section 8(c) states the criterion for noise-like columns: relative error per bin at the
99.9th percentile <= 1e-5 and <= 1e-3 dB on EVERY bin (tests/parity.py, ``noise_like=True``).  The
sweeps of tests/test_gpu_parity.py feed noise plus a tone 20 dB above it and therefore only check the
bins within 60 dB of the tone; here the input is pure noise, so nothing is masked: every default
kernel from nfft = 32 to 65536, Modes R / A / S, the integer-ingest twins and non powers of two.
Oracle: ``oracle.np_oracle`` (float64 restatement of drfProc.py:364-403 / :406-453)."""
import numpy as np
import pytest

from tests.parity import BIN_P999_TOL, COL_TOL, DB_TOL, psd_errors

pytestmark = pytest.mark.gpu

POW2 = [32, 64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384, 32768, 65536]


@pytest.fixture(scope="module")
def torch():
    import torch
    assert torch.cuda.is_available()
    return torch


def _noise(rng, n):
    return ((rng.standard_normal(n) + 1j * rng.standard_normal(n)) * (1e-2 / np.sqrt(2))).astype(np.complex64)


def _oracle(x, starts, nfft, nfr, hop):
    from oracle import np_oracle
    return np.stack([np_oracle.column_power(x[s:], nfft, nfr, hop) for s in starts])


def _assert_noise_like(got, got_db, ref, what):
    """got / got_db / ref: [ncol][nfft]"""
    e = psd_errors(got.T, ref.T)
    ddb = float(np.abs(got_db.astype(np.float64) - 10 * np.log10(ref + 1e-15)).max())
    assert e["col"] <= COL_TOL and e["bin_p999"] <= BIN_P999_TOL and e["db_max"] <= DB_TOL and ddb <= DB_TOL, (what, e, ddb)
    return e


def _run(torch, nfft, x_dev, starts, nfr, hop, in_scale=1.0):
    from pyspectrogram_b200 import engine
    plan = engine.StiPlan(nfft)
    lin, db = plan.run(x_dev, torch.from_numpy(starts).cuda(), nfr, hop, in_scale=in_scale, want_lin=True, want_db=True)
    torch.cuda.synchronize()
    return lin.cpu().numpy()[0], db.cpu().numpy()[0], plan.variant


@pytest.mark.parametrize("nfft", POW2)
@pytest.mark.parametrize("mode", ["R", "A", "S"])
def test_default_kernel_per_bin_on_noise(torch, nfft, mode):
    rng = np.random.default_rng(1000 * nfft + ord(mode))
    nfr, hop = {"R": (1, nfft), "A": (5, nfft), "S": (4, nfft - nfft // 8)}[mode]
    ncol = 4 if nfft >= 16384 else 12
    span = (nfr - 1) * hop + nfft
    x = _noise(rng, ncol * span + 8)
    starts = (np.arange(ncol) * span + np.arange(ncol) % 2).astype(np.int64)  # odd starts: 8-byte aligned only
    got, gdb, variant = _run(torch, nfft, torch.from_numpy(x).cuda(), starts, nfr, hop)
    _assert_noise_like(got, gdb, _oracle(x, starts, nfft, nfr, hop), f"nfft={nfft} mode {mode} {variant}")


@pytest.mark.parametrize("nfft", [64, 1024, 4096, 8192, 16384, 32768, 65536])
@pytest.mark.parametrize("kind", ["int16", "int8"])
def test_integer_ingest_twins_per_bin_on_noise(torch, nfft, kind):
    """Raw (re, im) integer pairs, 1 / full scale folded into the epilogue (drfProc.py:124-129, :182-201)."""
    rng = np.random.default_rng(nfft + len(kind))
    amp, dt, full = (3000, np.int16, 32768.0) if kind == "int16" else (100, np.int8, 128.0)
    nfr, ncol = 3, 4
    n = ncol * nfr * nfft + 16
    raw = rng.integers(-amp, amp, size=(n, 2)).astype(dt)
    x = (raw[:, 0].astype(np.float64) + 1j * raw[:, 1].astype(np.float64)) / full
    starts = (np.arange(ncol) * nfr * nfft + 2 * (np.arange(ncol) % 2)).astype(np.int64)
    got, gdb, variant = _run(torch, nfft, torch.from_numpy(raw).cuda(), starts, nfr, nfft, in_scale=1.0 / full)
    assert variant.endswith("_i16" if kind == "int16" else "_i8"), variant
    _assert_noise_like(got, gdb, _oracle(x, starts, nfft, nfr, nfft), f"nfft={nfft} {kind} {variant}")


@pytest.mark.parametrize("nfft", [96, 1000, 1001, 3000, 5000, 7000, 9100, 10000])
@pytest.mark.parametrize("mode", ["R", "A"])
def test_non_power_of_two_per_bin_on_noise(torch, nfft, mode):
    """The lengths people type into the viewer (drfview.py:474-479): direct mixed-radix transforms (compile-time plans,
    and the run-time kernel with radices 2 ... 16, 3, 5, 7, 11, 13)."""
    rng = np.random.default_rng(nfft * 7 + ord(mode))
    nfr = 1 if mode == "R" else 4
    ncol = 6
    x = _noise(rng, ncol * nfr * nfft + 8)
    starts = (np.arange(ncol) * nfr * nfft + np.arange(ncol) % 2).astype(np.int64)
    got, gdb, variant = _run(torch, nfft, torch.from_numpy(x).cuda(), starts, nfr, nfft)
    _assert_noise_like(got, gdb, _oracle(x, starts, nfft, nfr, nfft), f"nfft={nfft} mode {mode} {variant}")


MIXCT_LENGTHS = [1000, 1200, 1500, 1600, 1800, 2000, 2400, 2500, 2700, 3000, 3200, 3600, 4000, 4500, 4800, 5000, 6000, 6400, 8000, 10000]


@pytest.mark.parametrize("nfft", MIXCT_LENGTHS)
@pytest.mark.parametrize("case", ["R", "A", "S", "long", "int16_strided"])
def test_compile_time_mixed_radix_plans(torch, nfft, case):
    """Round lengths with a compile-time plan (sti_mixct.cuh: prime-factor butterflies 6 / 10 / 12 / 15 / 20,
    register-resident twiddles / window / accumulators): per-bin bar on noise against the float64 oracle in Modes
    R / A / S, long columns split over CTAs (partial sums + finalize, several frame groups per CTA), integer
    samples in an interleaved two-sub-channel layout -- and agreement with the run-time mixed-radix kernel."""
    from pyspectrogram_b200 import engine
    rng = np.random.default_rng(nfft * 11 + len(case))
    nfr, hop, ncol = {"R": (1, nfft, 9), "A": (5, nfft, 7), "S": (4, nfft - nfft // 8, 6), "long": (301, nfft, 2),
                      "int16_strided": (3, nfft, 5)}[case]
    span = (nfr - 1) * hop + nfft
    n = ncol * span + 8
    starts = (np.arange(ncol) * span + np.arange(ncol) % 2).astype(np.int64)
    plan = engine.StiPlan(nfft)
    if case == "int16_strided":
        raw = rng.integers(-3000, 3000, size=(n, 2, 2)).astype(np.int16)  # [sample][sub][re, im]
        x = (raw[:, 1, 0].astype(np.float64) + 1j * raw[:, 1, 1].astype(np.float64)) / 32768.0
        lin, db = plan.run(torch.from_numpy(raw).cuda(), torch.from_numpy(starts * 2).cuda(), nfr, hop, sample_stride=2,
                           sub_stride=1, nsub=2, in_scale=1.0 / 32768.0, want_lin=True, want_db=True)
        got, gdb = lin.cpu().numpy()[1], db.cpu().numpy()[1]
    else:
        x = _noise(rng, n)
        xd = torch.from_numpy(x).cuda()
        lin, db = plan.run(xd, torch.from_numpy(starts).cuda(), nfr, hop, want_lin=True, want_db=True)
        got, gdb = lin.cpu().numpy()[0], db.cpu().numpy()[0]
    assert plan.variant.startswith(f"mixct{nfft}_"), plan.variant
    _assert_noise_like(got, gdb, _oracle(x, starts, nfft, nfr, hop), f"nfft={nfft} {case} {plan.variant}")
    if case in ("A", "long"):
        try:
            engine.set_variant("mixed_rt")
            lin_rt, _ = plan.run(xd, torch.from_numpy(starts).cuda(), nfr, hop)
            assert plan.variant.startswith(f"mixed{nfft}_"), plan.variant
        finally:
            engine.set_variant(None)
        e = psd_errors(got.T, lin_rt.cpu().numpy()[0].T.astype(np.float64))
        assert e["col"] <= 2e-6 and e["bin_p999"] <= 1e-5, (nfft, case, e)


# ---------------------------------------------------------------------------------------------
# the persistent frame pipeline of the radix-32 kernels (sti_r32.cuh)
# ---------------------------------------------------------------------------------------------
R32_CASES = {
    # name: (ncol, nfr, hop as a fraction of nfft in eighths, odd starts)
    "mode_r_fewer_items_than_ctas": (3, 1, 8, False),
    "mode_a_odd_starts": (5, 4, 8, True),
    "long_columns_split_into_chunks": (2, 40, 8, False),
    "more_items_than_ctas": (400, 1, 8, True),
    "mode_s_hop": (40, 3, 7, False),
}


@pytest.mark.parametrize("nfft", [8192, 16384, 32768, 65536])
@pytest.mark.parametrize("case", sorted(R32_CASES))
def test_r32_frame_pipeline(torch, nfft, case):
    """Item switches every frame, columns split into chunks with partial sums, more items than resident CTAs
    (clusters), odd frame starts (16-byte skew of the bulk copies): the cases where the continuous pipeline of a
    persistent CTA can go wrong.  8192 points: the 256-thread form (two CTAs per SM) is forced."""
    from pyspectrogram_b200 import engine
    ncol, nfr, hop8, odd = R32_CASES[case]
    if nfft >= 32768:
        ncol = min(ncol, 150)
    hop = nfft * hop8 // 8
    rng = np.random.default_rng(nfft + len(case))
    span = (nfr - 1) * hop + nfft
    x = _noise(rng, ncol * span + 8)
    starts = (np.arange(ncol) * span + (np.arange(ncol) % 2 if odd else 0)).astype(np.int64)
    try:
        engine.set_variant("r32")
        got, gdb, variant = _run(torch, nfft, torch.from_numpy(x).cuda(), starts, nfr, hop)
    finally:
        engine.set_variant(None)
    assert variant.startswith("r32_"), variant
    pick = sorted(set([0, 1, ncol // 2, ncol - 2, ncol - 1]))
    _assert_noise_like(got[pick], gdb[pick], _oracle(x, starts[pick], nfft, nfr, hop), f"nfft={nfft} {case} {variant}")
    # every other column: finite and of the level noise has (a column written by the wrong item would not be)
    means = got.mean(axis=1)
    assert np.isfinite(got).all() and means.max() / means.min() < 1.5, (nfft, case)
