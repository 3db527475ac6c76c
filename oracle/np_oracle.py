"""Independent float64 restatement of the PSD/STI arithmetic in plain numpy.

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.  Never imported by
``pyspectrogram_b200``.

The reference delegates its arithmetic to scipy (third-party, pinned
``scipy==1.6.0``, ``requirements.txt:5``; not under /root/reference).  This
module restates scipy's published algorithm for the calls the reference makes,
without importing scipy, in float64 throughout:

* Kaiser window  -- scipy:windows/_windows.py:1318-1320 with the periodic
  extension of ``get_window(..., fftbins=True)`` (scipy:windows/_windows.py:2551)
* spectrum scaling ``1/sum(w)**2`` -- scipy:_spectral_py.py:2277
* two-sided DFT, no detrend, ``|X|^2`` -- scipy:_spectral_py.py:2346-2397
* segment mean (welch) -- scipy:_spectral_py.py:658-670
* default spectrogram overlap ``nperseg//8`` and segment times
  -- scipy:_spectral_py.py:1129, :2324-2325

Results are returned in float64 ("truth"); ``cast_like_reference`` applies the
float32 cast the reference's outputs carry for complex64 input
(scipy:_spectral_py.py:969).

Parity pinning: ``tests/test_oracle.py`` checks it against the golden fixtures
made from the unmodified reference functions and against scipy's KATs.
"""
from __future__ import annotations

import numpy as np


def kaiser_periodic(nfft: int, beta: float = 1.7) -> np.ndarray:
    """w[n] = I0(beta*sqrt(1-((n-a)/a)^2))/I0(beta), a = nfft/2, n = 0..nfft-1.

    (Symmetric Kaiser of length nfft+1 with its last point dropped.)
    """
    if nfft == 1:
        return np.ones(1)
    n = np.arange(nfft, dtype=np.float64)
    a = nfft / 2.0
    r = (n - a) / a
    return np.i0(beta * np.sqrt(np.clip(1.0 - r * r, 0.0, None))) / np.i0(beta)


def _frames(col: np.ndarray, nfft: int, hop: int, count: int) -> np.ndarray:
    idx = np.arange(count)[:, None] * hop + np.arange(nfft)[None, :]
    return col[idx]


def column_power(col, nfft, frames_per_col=1, hop=None, beta=1.7):
    """fftshifted mean power of ``frames_per_col`` frames of one 1-D column."""
    hop = nfft if hop is None else hop
    w = kaiser_periodic(nfft, beta)
    x = _frames(np.asarray(col, dtype=np.complex128), nfft, hop, frames_per_col)
    spec = np.fft.fft(x * w[None, :], axis=1)
    p = (spec.real ** 2 + spec.imag ** 2) / (w.sum() ** 2)
    return np.fft.fftshift(p.mean(axis=0))


def freq_axis(nfft, sr):
    """fftshifted ``fftfreq(nfft, 1/fs)`` (drfProc.py:398)."""
    return np.fft.fftshift(np.fft.fftfreq(nfft, 1.0 / float(sr)))


def sti(d1, sr, nfft, integrate=False, beta=1.7):
    """STI image from the reference's ``(nint*nfft, ntime[, nsub])`` array.

    ``integrate=False`` -> Mode R (first nfft rows only, drfProc.py:364-403);
    ``integrate=True``  -> Mode A (mean of floor(rows/nfft) frames).
    Returns float64 ``(f, sxx, sxx_med)``.
    """
    d1 = np.asarray(d1)
    if d1.shape[0] < nfft:
        raise ValueError("fewer rows than nfft")
    squeeze = d1.ndim == 2
    d3 = d1[:, :, None] if squeeze else d1
    rows, ntime, nsub = d3.shape
    nfr = rows // nfft if integrate else 1
    img = np.empty((nfft, ntime, nsub))
    for t in range(ntime):
        for s in range(nsub):
            img[:, t, s] = column_power(d3[:, t, s], nfft, nfr, beta=beta)
    if squeeze:
        img = img[:, :, 0]
    return freq_axis(nfft, sr), img, np.median(img, axis=1)


def sti_overlap(x, sr, nfft, dt, beta=1.7):
    """Mode S (``proc_data``, drfProc.py:406-453) in float64."""
    x = np.asarray(x)
    hop = nfft - nfft // 8
    nseg = (x.shape[0] - nfft) // hop + 1
    t = (nfft / 2.0 + np.arange(nseg) * hop) / float(sr)
    per_col = int(dt / (t[1] - t[0]))
    edges = np.arange(0, nseg, per_col)
    ncol = len(edges) - 1
    img = np.empty((nfft, ncol))
    for c in range(ncol):
        img[:, c] = column_power(x[edges[c] * hop:], nfft, per_col, hop, beta)
    return t[edges][:-1], freq_axis(nfft, sr), img, np.median(img, axis=-1)


def cast_like_reference(arr, in_dtype):
    """Reference outputs are float32 for complex64 input, else float64."""
    return arr.astype(np.float32) if np.dtype(in_dtype) == np.complex64 else arr


def to_db(power, eps=1e-15):
    """10*log10(p+eps) in the dtype of ``power`` (drfProc.py:308-310)."""
    power = np.asarray(power)
    return 10 * np.log10(power + power.dtype.type(eps))
