"""Restatement of the reference CPU path (scipy/numpy call sequence).

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.  Never imported by
``pyspectrogram_b200``.

Each function names the reference lines it follows (``drfProc.py:N`` are lines
of ``/root/reference/drfProc.py``; ``scipy:`` are lines of the scipy 1.18.1
installed in this image -- third-party, pinned ``scipy==1.6.0`` by the
reference's ``requirements.txt:5``; the arithmetic is the same, only the
internal precision differs).

Parity pinning: checked against fixtures generated from the reference's own
unmodified function bodies (``tools/make_golden.py`` -> ``tests/golden``) by
``tests/test_oracle.py``.
"""
from __future__ import annotations

import numpy as np
import scipy.signal as _sig

KAISER_BETA = 1.7  # drfProc.py:386, :435
DB_EPS = 1e-15  # drfProc.py:308


def kaiser_window(nfft: int, beta: float = KAISER_BETA) -> np.ndarray:
    """Periodic Kaiser window the reference builds at drfProc.py:386 / :435."""
    return _sig.get_window(("kaiser", beta), nfft)


def sti_mode_r(d1, sr, nfft):
    """Mode R: the shipped ``sti_proc_data`` (drfProc.py:364-403).

    scipy.signal.periodogram keeps only rows ``[:nfft]`` of axis 0
    (scipy:_spectral_py.py:498-503), so extra integration rows are ignored.
    Returns ``(f, sxx, sxx_med)`` with linear power, fftshifted.
    """
    w = kaiser_window(nfft)
    freqs, p = _sig.periodogram(
        d1, sr, window=w, nfft=nfft, detrend=False,
        return_onesided=False, scaling="spectrum", axis=0)  # drfProc.py:387-396
    freqs = np.fft.fftshift(freqs)  # drfProc.py:398
    img = np.fft.fftshift(p, axes=0)  # drfProc.py:399
    return freqs, img, np.median(img, axis=1)  # drfProc.py:401


def sti_mode_a(d1, sr, nfft):
    """Mode A: mean of the ``nint`` back-to-back frames of every time bin.

    This is the averaging north_star describes and ``read_sti`` reads the data
    for (drfProc.py:158); it equals the call ``periodogram`` forwards to
    (scipy:_spectral_py.py:510-512, ``welch(noverlap=0)``) minus the truncation
    at scipy:_spectral_py.py:498-503.  A tail shorter than ``nfft`` is dropped.
    """
    w = kaiser_window(nfft)
    freqs, p = _sig.welch(
        d1, sr, window=w, nperseg=nfft, noverlap=0, nfft=nfft, detrend=False,
        return_onesided=False, scaling="spectrum", axis=0, average="mean")
    freqs = np.fft.fftshift(freqs)
    img = np.fft.fftshift(p, axes=0)
    return freqs, img, np.median(img, axis=1)


def sti_mode_s(d1, sr, nfft, dt):
    """Mode S: ``proc_data`` (drfProc.py:406-453).

    Spectrogram with scipy's default overlap ``nfft//8``
    (scipy:_spectral_py.py:1129), groups of ``n_int`` spectra averaged, the
    last (possibly partial) group always dropped (drfProc.py:440-447).
    Returns ``(t_out, f, sxx_int, sxx_med)``.
    """
    w = kaiser_window(nfft)
    freqs, t, s = _sig.spectrogram(
        d1, sr, window=w, detrend=False, return_onesided=False,
        scaling="spectrum")  # drfProc.py:436-438
    per_col = int(dt / (t[1] - t[0]))  # drfProc.py:439
    edges = np.arange(0, len(t), per_col)  # drfProc.py:440
    ncol = len(edges) - 1
    img = np.zeros((nfft, ncol), dtype=s.dtype)  # drfProc.py:442
    for c in range(ncol):
        img[:, c] = np.mean(s[:, edges[c]:edges[c + 1]], axis=-1)  # :443-445
    t_out = t[edges][:-1]  # drfProc.py:447
    freqs = np.fft.fftshift(freqs)
    img = np.fft.fftshift(img, axes=0)  # drfProc.py:448-449
    return t_out, freqs, img, np.median(img, axis=-1)  # drfProc.py:451


def to_dbfs(power, eps=DB_EPS):
    """dB conversion inlined in the worker loop (drfProc.py:308-310)."""
    return 10 * np.log10(power + eps)


def full_scale_ref(props: dict) -> float:
    """``get_ref`` (drfProc.py:182-201): 1.0 for float HDF5 class, otherwise
    ``2**((precision-1) + 0.5*(size_bytes-1))``."""
    if props["H5Tget_class"] == 1:
        return 1.0
    exponent = props["H5Tget_precision"] - 1.0
    exponent += 0.5 * (props["H5Tget_size"] - 1.0)
    return 2 ** exponent


def sti_frame_starts(st_sample, en_sample, nfft, nint, ntime):
    """First sample of every STI time bin (drfProc.py:158-159).

    float64 ``linspace`` then truncation to int -- quantised for epoch-sized
    indices, which is part of the contract (trap 2).
    """
    return np.linspace(st_sample, en_sample - nint * nfft, ntime, dtype=int)


def read_sti_from_array(recording, st_sample, en_sample, nfft, nint, ntime,
                        ref=1.0, first_sample=0):
    """``DrfInput.read_sti`` (drfProc.py:132-167) over an in-memory recording.

    ``recording`` is ``(N,)`` or ``(N, nsub)`` holding absolute samples
    ``first_sample .. first_sample+N``.  Each read is divided by ``ref`` like
    ``DrfInput.read`` does (drfProc.py:129), gets a new axis 1 and the reads
    are concatenated along it (drfProc.py:160-166).
    """
    starts = sti_frame_starts(st_sample, en_sample, nfft, nint, ntime)
    span = nint * nfft
    pieces = []
    for s0 in starts:
        lo = int(s0) - first_sample
        chunk = recording[lo:lo + span] / ref
        pieces.append(chunk[:, np.newaxis])
    return starts, np.concatenate(pieces, axis=1)


def plot_indices(freqs, cfrange_khz, max_nfreqs=2 ** 15):
    """The viewer's frequency selection and decimation (drfview.py:1005-1023), statement by statement.

    ``freqs``: the fftshifted frequency axis in Hz; ``cfrange_khz``: (low, high) in kHz.
    Returns ``(plotindices, plotfreqs, fscale)``.
    """
    keepvals = np.all((np.greater_equal(freqs, 1e3 * cfrange_khz[0]), np.less_equal(freqs, 1e3 * cfrange_khz[1])), axis=0)
    kept = freqs[keepvals]
    inds = np.argwhere(keepvals)
    fscale = int(np.ceil(len(kept) / max_nfreqs))
    relplotindices = range(int(np.floor(fscale / 2)), len(kept), fscale)
    plotindices = [inds[i][0] for i in relplotindices]
    plotfreqs = np.array([kept[i] for i in relplotindices])
    return plotindices, plotfreqs, fscale


def clip_to_colour_range(spectra, colorrange):
    """Colour-range clip of the PNG export (drfview.py:1515-1516), on a copy."""
    spectra = np.array(spectra, copy=True)
    spectra[spectra < colorrange[0]] = colorrange[0]
    spectra[spectra > colorrange[1]] = colorrange[1]
    return spectra


def proc_data_min_max(sxx_int):
    """The minimum and maximum across time that proc_data's docstring lists after the median
    (drfProc.py:430-433; the shipped function stops at the median, drfProc.py:451-453)."""
    return np.min(sxx_int, axis=-1), np.max(sxx_int, axis=-1)
