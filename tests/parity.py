"""Parity metrics for PSD/STI images.

Tolerances (north_star: PSD 1e-5 relative in fp32, 1e-3 dB after the log, indexing exact):
  (i)   per column  max|got-ref| / max(ref column)      <= 1e-5   (every input class)
  (ii)  per bin     |got-ref|/ref, 99.9th percentile    <= 1e-5   (noise-like inputs; max is reported)
  (iii) max |dB(got)-dB(ref)| <= 1e-3 on noise-like inputs; on high-dynamic-range inputs only
        for bins within 60 dB of the column peak
  (iv)  shapes, dtypes, frequency axis, bin order: exact (checked by the callers)
"""
import numpy as np

COL_TOL = 1e-5
BIN_P999_TOL = 1e-5
DB_TOL = 1e-3


def psd_errors(got, ref, freq_axis=0):
    got = np.moveaxis(np.asarray(got, dtype=np.float64), freq_axis, 0).reshape(got.shape[freq_axis], -1)
    ref = np.moveaxis(np.asarray(ref, dtype=np.float64), freq_axis, 0).reshape(ref.shape[freq_axis], -1)
    peak = ref.max(axis=0, keepdims=True)
    peak = np.where(peak > 0, peak, 1.0)
    col = (np.abs(got - ref) / peak).max()
    nz = ref > 0
    rel = np.abs(got - ref)[nz] / ref[nz]
    p999 = float(np.quantile(rel, 0.999)) if rel.size else 0.0
    relmax = float(rel.max()) if rel.size else 0.0
    eps = 1e-15
    ddb = np.abs(10 * np.log10(got + eps) - 10 * np.log10(ref + eps))
    strong = ref >= peak * 1e-6
    return {"col": float(col), "bin_p999": p999, "bin_max": relmax,
            "db_max": float(ddb.max()), "db_max_strong": float(ddb[strong].max())}


def assert_psd_close(got, ref, freq_axis=0, noise_like=True, what=""):
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    e = psd_errors(got, ref, freq_axis)
    assert e["col"] <= COL_TOL, (what, e)
    if noise_like:
        assert e["bin_p999"] <= BIN_P999_TOL, (what, e)
        assert e["db_max"] <= DB_TOL, (what, e)
    else:
        assert e["db_max_strong"] <= DB_TOL, (what, e)
    return e


def assert_db_close(got_db, ref_db, ref_lin=None, what=""):
    """dB images: 1e-3 dB everywhere, or only on bins within 60 dB of each column's peak."""
    assert got_db.shape == ref_db.shape, (what, got_db.shape, ref_db.shape)
    d = np.abs(np.asarray(got_db, np.float64) - np.asarray(ref_db, np.float64))
    if ref_lin is not None:
        peak = np.asarray(ref_lin, np.float64).max(axis=0, keepdims=True)
        d = d[np.asarray(ref_lin) >= peak * 1e-6]
    assert d.max() <= DB_TOL, (what, float(d.max()))
    return float(d.max())
