"""Pin the CPU oracle (oracle/) against the golden fixtures made from the unmodified reference
functions (tools/make_golden.py) and against scipy's upstream known-answer tests."""
import json
import os
from fractions import Fraction

import numpy as np
import pytest

from oracle import np_oracle, ref_port
from tests.parity import assert_psd_close

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
META = json.load(open(os.path.join(GOLDEN, "meta.json")))


def load(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


STI_R = [n for n in META["cases"] if n.startswith("sti_r_") and n != "sti_r_fraction_sr"]


@pytest.mark.parametrize("name", STI_R)
def test_ref_port_mode_r_matches_reference(name):
    g = load(name)
    f, sxx, med = ref_port.sti_mode_r(g["d1"], float(g["sr"]), int(g["nfft"]))
    # same library calls as the reference -> identical arrays
    assert sxx.dtype == g["sxx"].dtype and med.dtype == g["med"].dtype
    np.testing.assert_array_equal(f, g["f"])
    np.testing.assert_array_equal(sxx, g["sxx"])
    np.testing.assert_array_equal(med, g["med"])
    if "sxx_db" in g:
        np.testing.assert_array_equal(ref_port.to_dbfs(sxx), g["sxx_db"])
        assert ref_port.to_dbfs(sxx).dtype == g["sxx_db"].dtype


@pytest.mark.parametrize("name", STI_R)
def test_np_oracle_mode_r_matches_reference(name):
    g = load(name)
    f, sxx, med = np_oracle.sti(g["d1"], float(g["sr"]), int(g["nfft"]))
    np.testing.assert_array_equal(f, g["f"])
    noise_like = name not in ("sti_r_hdr2048", "sti_r_tone1024", "sti_r_impulse16", "sti_r_zeros32")
    ref = g["sxx"].astype(np.float64)
    assert sxx.shape == ref.shape
    if name == "sti_r_zeros32":
        assert not sxx.any()
        return
    e = assert_psd_close(sxx, ref, noise_like=noise_like, what=name)
    # float64 restatement vs reference output stored in float32: one fp32 rounding
    assert e["col"] < 2e-7
    np.testing.assert_allclose(med, g["med"], rtol=3e-7, atol=1e-30)


def test_reference_output_dtypes_and_shapes():
    g = load("sti_r_256x10x3")
    assert g["sxx"].dtype == np.float32 and g["sxx"].shape == (256, 10, 3)
    assert g["med"].shape == (256, 3) and g["f"].dtype == np.float64
    g = load("sti_r_128x6_c128")
    assert g["sxx"].dtype == np.float64


def test_fraction_sample_rate():
    g = load("sti_r_fraction_sr")
    sr = Fraction(int(g["sr_num"]), int(g["sr_den"]))
    f, sxx, med = ref_port.sti_mode_r(g["d1"], sr, 64)
    np.testing.assert_array_equal(f, g["f"])
    np.testing.assert_array_equal(sxx, g["sxx"])
    np.testing.assert_array_equal(np_oracle.freq_axis(64, sr), g["f"])


def test_mode_r_ignores_integration_rows():
    """scipy truncates every time bin to its first nfft samples (scipy:_spectral_py.py:498-503)."""
    g = load("sti_r_64x7")
    d1 = g["d1"]
    assert d1.shape[0] == 3 * 64
    _, full, _ = ref_port.sti_mode_r(d1, 1.0e4, 64)
    _, first, _ = ref_port.sti_mode_r(d1[:64], 1.0e4, 64)
    np.testing.assert_array_equal(full, first)
    np.testing.assert_array_equal(full, g["sxx"])


def test_short_input_raises_like_reference():
    d1 = np.zeros((32, 3), np.complex64)
    with pytest.raises(ValueError):
        ref_port.sti_mode_r(d1, 1.0, 64)
    with pytest.raises(ValueError):
        np_oracle.sti(d1, 1.0, 64)


@pytest.mark.parametrize("name", ["sti_a_128x5x6x2", "sti_a_512x9x4"])
def test_mode_a(name):
    g = load(name)
    nfft = int(g["nfft"])
    f, sxx, med = ref_port.sti_mode_a(g["d1"], float(g["sr"]), nfft)
    np.testing.assert_array_equal(sxx, g["sxx"])
    np.testing.assert_array_equal(med, g["med"])
    f2, sxx2, med2 = np_oracle.sti(g["d1"], float(g["sr"]), nfft, integrate=True)
    np.testing.assert_array_equal(f2, g["f"])
    e = assert_psd_close(sxx2, g["sxx"].astype(np.float64), what=name)
    assert e["col"] < 2e-7
    # Mode A == mean of per-frame Mode R
    nint = g["d1"].shape[0] // nfft
    acc = 0
    for k in range(nint):
        acc = acc + ref_port.sti_mode_r(g["d1"][k * nfft:(k + 1) * nfft], float(g["sr"]), nfft)[1].astype(np.float64)
    assert_psd_close(acc / nint, g["sxx"].astype(np.float64), what=name + " mean of R")


@pytest.mark.parametrize("name", ["proc_256", "proc_1024"])
def test_mode_s_proc_data(name):
    g = load(name)
    nfft = int(g["nfft"])
    t, f, sxx, med = ref_port.sti_mode_s(g["x"], float(g["sr"]), nfft, float(g["dt"]))
    np.testing.assert_array_equal(t, g["t_out"])
    np.testing.assert_array_equal(f, g["f"])
    np.testing.assert_array_equal(sxx, g["sxx"])
    np.testing.assert_array_equal(med, g["med"])
    assert sxx.dtype == np.float32
    t2, f2, sxx2, med2 = np_oracle.sti_overlap(g["x"], float(g["sr"]), nfft, float(g["dt"]))
    np.testing.assert_allclose(t2, g["t_out"], rtol=1e-15, atol=0)
    np.testing.assert_array_equal(f2, g["f"])
    assert sxx2.shape == g["sxx"].shape
    # scipy's legacy spectrogram helper works in complex64 for complex64 input
    # (scipy:_spectral_py.py:2272): compare at fp32-internal accuracy
    e = assert_psd_close(sxx2, g["sxx"].astype(np.float64), what=name)
    assert e["col"] < 2e-6


def test_get_ref():
    for item in META["get_ref"]:
        assert ref_port.full_scale_ref(item["props"]) == item["ref"]
    assert ref_port.full_scale_ref({"H5Tget_class": 0, "H5Tget_precision": 16, "H5Tget_size": 2}) == 2 ** 15.5


def test_frame_starts_bit_exact():
    for item in META["frame_starts"]:
        n_st = ref_port.sti_frame_starts(item["st"], item["en"], item["nfft"], item["nint"], item["ntime"])
        assert n_st.dtype == np.int64 and len(n_st) == item["ntime"]
        assert [int(v) for v in n_st[:4]] == item["first"]
        assert [int(v) for v in n_st[-4:]] == item["last"]
        assert int(np.sum(n_st.astype(object))) == item["sum"]


def test_frame_starts_are_quantised_for_epoch_indices():
    """float64 linspace: at ~1.7e17 the spacing is 32 samples (trap 2)."""
    item = [i for i in META["frame_starts"] if i["st"] > 10 ** 16][0]
    n_st = ref_port.sti_frame_starts(item["st"], item["en"], item["nfft"], item["nint"], item["ntime"])
    assert np.all(n_st % 32 == 0)


def test_read_sti_from_array_layout():
    rng = np.random.default_rng(5)
    rec = (rng.standard_normal((5000, 2)) + 1j * rng.standard_normal((5000, 2))).astype(np.complex64)
    starts, d = ref_port.read_sti_from_array(rec, 1000, 1000 + 5000, 64, 3, 7, ref=2.0, first_sample=1000)
    assert d.shape == (192, 7, 2) and d.dtype == np.complex64
    for c, s0 in enumerate(starts):
        np.testing.assert_array_equal(d[:, c, :], rec[s0 - 1000:s0 - 1000 + 192] / 2.0)


# ---- scipy upstream known answers for the calls on the path -------------

def test_kaiser_kat_and_tables():
    # scipy:tests/test_windows.py:492-520  kaiser(6, 2.7, sym=False)
    want = [0.2603047507678832, 0.5985765418119844, 0.8868495172060835, 1.0,
            0.8868495172060835, 0.5985765418119844]
    np.testing.assert_allclose(np_oracle.kaiser_periodic(6, 2.7), want, rtol=1e-13)
    tabs = load("kaiser_tables")
    for key in tabs.files:
        n = int(key[1:])
        np.testing.assert_allclose(np_oracle.kaiser_periodic(n), tabs[key], rtol=1e-13)
        np.testing.assert_array_equal(ref_port.kaiser_window(n), tabs[key])
    w = np_oracle.kaiser_periodic(1024)
    assert abs(w[0] - 0.53649) < 1e-5 and abs(w.sum() - 854.9537) < 1e-3


def test_periodogram_impulse_kat():
    # scipy:tests/test_spectral.py:94-101 (boxcar: 5/16 flat); with the Kaiser window the
    # impulse at n=0 gives the flat value |1+2j|^2 * w[0]^2 / sum(w)^2
    g = load("sti_r_impulse16")
    w = np_oracle.kaiser_periodic(16)
    want = 5.0 * w[0] ** 2 / w.sum() ** 2
    np.testing.assert_allclose(g["sxx"], want, rtol=1e-6)
    _, sxx, _ = np_oracle.sti(g["d1"], 1.0, 16)
    np.testing.assert_allclose(sxx, want, rtol=1e-13)


def test_unit_tone_is_0_dbfs_at_shifted_bin():
    g = load("sti_r_tone1024")
    col = g["sxx"][:, 0]
    assert int(np.argmax(col)) == 512 + 37
    assert abs(col[512 + 37] - 1.0) < 1e-6
    assert g["f"][512 + 37] == 37.0 and g["f"][0] == -512.0


def test_zero_input_floor_is_minus_150_db():
    g = load("sti_r_zeros32")
    assert g["sxx_db"].dtype == np.float32
    np.testing.assert_allclose(g["sxx_db"], -150.0, atol=1e-4)
    assert np.all(np_oracle.to_db(g["sxx"]) == g["sxx_db"])


def test_median_even_count_is_mean_of_middle_pair_in_float32():
    g = load("sti_r_256x10x3")
    srt = np.sort(g["sxx"], axis=1)
    want = ((srt[:, 4, :] + srt[:, 5, :]) * np.float32(0.5)).astype(np.float32)
    np.testing.assert_array_equal(g["med"], want)


@pytest.mark.parametrize("nfft,cfrange,maxn", [(1024, (-400.0, 400.0), 2 ** 15), (65536, (-500.0, 500.0), 2 ** 15),
                                              (65536, (-100.0, 250.0), 1000), (4096, (12.0, 13.0), 2 ** 15),
                                              (1000, (-1e9, 1e9), 7), (33, (0.0, 0.0), 4)])
def test_plot_indices_host_logic_matches_viewer_restatement(nfft, cfrange, maxn):
    """pyspectrogram_b200.drfProc.plot_indices (closed form) against the statement-by-statement
    restatement of drfview.py:1005-1023; pure host logic, no GPU."""
    from oracle import ref_port
    from pyspectrogram_b200.drfProc import plot_indices
    f = np.fft.fftshift(np.fft.fftfreq(nfft, 1 / 1.0e6))
    pidx, pfreqs, fscale = ref_port.plot_indices(f, cfrange, maxn)
    idx, freqs, fs2 = plot_indices(f, cfrange, maxn)
    assert list(idx) == list(pidx) and np.array_equal(freqs, pfreqs) and fs2 == fscale
    assert len(idx) <= maxn and idx.dtype == np.int64


def test_plot_indices_empty_selection_raises_like_the_viewer():
    from oracle import ref_port
    from pyspectrogram_b200.drfProc import plot_indices
    f = np.fft.fftshift(np.fft.fftfreq(64, 1 / 1.0e6))
    with pytest.raises(ValueError):
        ref_port.plot_indices(f, (900.0, 901.0))
    with pytest.raises(ValueError):
        plot_indices(f, (900.0, 901.0))
