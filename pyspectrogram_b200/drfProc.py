"""Drop-in for the reference's ``drfProc`` module (the surface ``drfview.py:89`` imports as ``dp``).

Same names, call signatures and returned arrays as ``/root/reference/drfProc.py``; the PSD/STI
arithmetic runs on a B200 through ``libpsgb200.so`` instead of scipy.  There is no CPU fallback:
without the built library or a CUDA device these functions raise.

Reference anchors
* ``sti_proc_data(d1, sr, nfft)``      drfProc.py:364-403  (Mode R: only rows ``[:nfft]`` of each
  time bin are used, because ``scipy.signal.periodogram`` truncates, scipy:_spectral_py.py:498-503)
* ``proc_data(d1, sr, nfft, dt)``      drfProc.py:406-453  (Mode S: hop ``nfft - nfft//8``)
* dB conversion                        drfProc.py:308-310  (fused epilogue, ``want_db``)
* ``DrfInput`` / ``read_sti``          drfProc.py:59-179
* ``get_ref``                          drfProc.py:182-201
* ``DrfProcessor``                     drfProc.py:209-361, signals drfProc.py:458-465

Extensions (keyword-only, defaults reproduce the reference): ``integrate=True`` averages all
``nint`` frames of each time bin (Mode A, the averaging ``read_sti`` reads the data for,
drfProc.py:158); ``device=`` selects the GPU; ``proc_data(..., minmax=True)`` also returns the minimum
and maximum spectra its docstring lists (drfProc.py:430-433).  ``plot_indices`` / ``sti_plot_data``
 do the viewer's frequency selection, decimation and colour clip
(drfview.py:1005-1023, :1515-1518) on the GPU so that only the drawn bins are copied back.
"""
from __future__ import annotations

import time as timemodule
from fractions import Fraction
from pathlib import Path
from traceback import print_exc as trace_error

import numpy as np

from . import engine
from ._lib import PsgUnsupported

__all__ = ["DrfInput", "DrfProcessor", "ThreadProcessorSignals", "get_ref", "proc_data", "sti_proc_data",
           "sti_proc_data_db", "plot_indices", "sti_plot_data"]

_EPS = 1e-15  # drfProc.py:308


# ------------------------------------------------------------------------------------------------
# hot entry points
# ------------------------------------------------------------------------------------------------
def _as_c64(d1):
    """Input as C-contiguous complex64 plus the dtype the reference would return."""
    arr = np.asarray(d1)
    out_dtype = np.float32 if arr.dtype in (np.complex64, np.float32) else np.float64
    if arr.dtype != np.complex64:
        arr = arr.astype(np.complex64)
    return np.ascontiguousarray(arr), out_dtype


def _freq_axis(nfft, sr):
    # drfProc.py:398 -- fftshift of scipy's fftfreq(nfft, 1/fs) (scipy:_spectral_py.py:2307)
    return np.fft.fftshift(np.fft.fftfreq(nfft, 1 / sr))


def _is_raw_iq(d1, raw_pairs=False):
    """Raw integer IQ: Digital RF's structured ('r', 'i') int16 / int8 dtype (detected), or -- only when the
    caller says so with ``raw_pairs=True`` -- a plain int16 / int8 array whose last axis is (re, im).  A plain
    integer array without the keyword is REAL samples, as it is for the reference (scipy casts it)."""
    if not isinstance(d1, np.ndarray):
        return False
    if d1.dtype.fields is not None:
        return True
    if raw_pairs:
        if d1.dtype not in (np.int16, np.int8):
            raise TypeError("raw_pairs=True takes an int16 or int8 array with a last axis (re, im)")
        return True
    return False


def _raw_view(d1):
    """C-contiguous raw IQ array and its logical (complex) shape."""
    arr = np.ascontiguousarray(d1)
    shape = arr.shape if arr.dtype.fields is not None else arr.shape[:-1]
    if arr.dtype.fields is None and arr.shape[-1] != 2:
        raise ValueError("raw integer IQ must have a last axis of length 2 (re, im)")
    return arr, shape


def _sti(d1, sr, nfft, integrate, device, want, ref=1.0, raw_pairs=False):
    nfft = int(nfft)
    if _is_raw_iq(d1, raw_pairs):
        # extension: the samples go to the GPU as stored; 1/ref is applied
        # to the power in the kernel epilogue instead of x/ref on the host (drfProc.py:129)
        arr, shape = _raw_view(d1)
        out_dtype = np.float32
        in_scale = 1.0 / float(ref)
    else:
        arr, out_dtype = _as_c64(d1)
        shape = arr.shape
        in_scale = 1.0 / float(ref)
    return _sti_core(arr, shape, out_dtype, in_scale, sr, nfft, integrate, device, want)


def _sti_core(arr, shape, out_dtype, in_scale, sr, nfft, integrate, device, want):
    ndim = len(shape)
    if ndim not in (2, 3):
        raise ValueError("d1 must be (nfft*nint, ntime) or (nfft*nint, ntime, nsub)")
    rows, ntime = shape[0], shape[1]
    nsub = shape[2] if ndim == 3 else 1
    if rows < nfft:
        # the reference passes a length-nfft window to a shorter segment: scipy raises ValueError
        raise ValueError(f"window is longer than input signal ({rows} rows < nfft={nfft})")
    if ntime < 1 or nsub < 1:
        raise ValueError("empty time or sub-channel axis")
    frames = rows // nfft if integrate else 1
    plan = engine.get_plan(nfft, device)
    res = plan.host(arr.reshape(-1), np.arange(ntime, dtype=np.int64) * nsub, frames, nfft,
                    sample_stride=ntime * nsub, sub_stride=1, nsub=nsub, in_scale=in_scale, eps=_EPS, want=want)
    out = {}
    for key, val in res.items():
        if val.ndim == 3:  # [nsub][ntime][nfft] -> (nfft, ntime[, nsub]) as the viewer indexes it
            val = val.transpose(2, 1, 0)
            val = val if ndim == 3 else val[:, :, 0]
        else:  # [nsub][nfft] -> (nfft[, nsub])
            val = val.T if ndim == 3 else val[0]
        out[key] = val if out_dtype == np.float32 else val.astype(out_dtype)
    return _freq_axis(nfft, sr), out


def sti_proc_data(d1, sr, nfft, *, integrate=False, device=0, ref=1.0, raw_pairs=False):
    """STI of ``d1`` shaped ``(nfft*nint, ntime[, nsub])`` -> ``(f, sxx, sxx_med)``.

    ``sxx`` is ``(nfft, ntime[, nsub])`` linear power (float32 for complex64 input), fftshifted;
    ``sxx_med`` is its median over time (drfProc.py:364-403).  With the default
    ``integrate=False`` only the first ``nfft`` rows of each time bin are used, exactly like the
    reference; ``integrate=True`` averages ``floor(rows/nfft)`` back-to-back frames (Mode A).

    ``d1`` may also be raw integer IQ -- Digital RF's structured int16 / int8 ('r', 'i') dtype, or, with
    ``raw_pairs=True``, a plain integer array with a last axis ``(re, im)`` -- together with ``ref`` = the
    full-scale level of ``get_ref``: the result equals ``sti_proc_data(d1_as_complex / ref, ...)`` without the
    host-side cast and divide.  A plain integer array WITHOUT ``raw_pairs`` is real-valued samples, as in the
    reference.
    """
    f, out = _sti(d1, sr, nfft, integrate, device, ("lin", "med"), ref, raw_pairs)
    return f, out["lin"], out["med"]


def sti_proc_data_db(d1, sr, nfft, *, integrate=False, device=0, ref=1.0, raw_pairs=False):
    """``sti_proc_data`` plus the worker loop's dB step (drfProc.py:308-310, ``10*log10(x + 1e-15)``) fused on
    the GPU.

    Returns ``(f, sxx_dbfs, sxx_med_dbfs)``.
    """
    f, out = _sti(d1, sr, nfft, integrate, device, ("db", "med_db"), ref, raw_pairs)
    return f, out["db"], out["med_db"]


def proc_data(d1, sr, nfft, dt, *, device=0, minmax=False):
    """``proc_data`` (drfProc.py:406-453): overlapped spectrogram averaged in groups of ``n_int``.

    Returns ``(t_out, f, sxx_int, sxx_med)``; scipy's default overlap ``nfft//8``
    (scipy:_spectral_py.py:1129), segment times ``(nfft/2 + j*hop)/fs``
    (scipy:_spectral_py.py:2324-2325), last group always dropped (drfProc.py:440-447).
    ``minmax=True`` appends ``sxx_min, sxx_max`` (minimum / maximum across time, the two outputs
    the reference's docstring lists, drfProc.py:430-433, and its body never computes).
    """
    nfft = int(nfft)
    arr, out_dtype = _as_c64(d1)
    if arr.ndim != 1:
        raise ValueError("proc_data takes a 1-D sample vector")
    nsamp = arr.shape[0]
    if nsamp < nfft:
        raise ValueError("window is longer than input signal")
    noverlap = nfft // 8
    hop = nfft - noverlap
    t = np.arange(nfft / 2, nsamp - nfft / 2 + 1, hop) / float(sr)
    n_int = int(dt / (t[1] - t[0]))  # drfProc.py:439 (IndexError for a single segment, like the reference)
    n1 = np.arange(0, len(t), n_int)  # drfProc.py:440 (ZeroDivisionError when n_int == 0, like the reference)
    ncol = len(n1) - 1
    f = _freq_axis(nfft, sr)
    t_out = t[n1][:-1]
    if ncol < 1:
        empty = (t_out, f, np.zeros((nfft, 0), out_dtype), np.full(nfft, np.nan, out_dtype))
        return empty + (np.full(nfft, np.nan, out_dtype),) * 2 if minmax else empty
    plan = engine.get_plan(nfft, device)
    offs = n1[:-1].astype(np.int64) * hop
    if not minmax:
        res = plan.host(arr, offs, n_int, hop, want=("lin", "med"))
        sxx = res["lin"][0].T
        outs = [res["med"][0]]
    else:
        import torch
        dev = torch.device("cuda", device)
        lin, _ = plan.run(torch.from_numpy(arr).to(dev), torch.from_numpy(offs).to(dev), n_int, hop)
        med, _ = plan.median(lin)
        mn, mx, _, _ = plan.minmax(lin)
        sxx = lin[0].cpu().numpy().T
        outs = [v[0].cpu().numpy() for v in (med, mn, mx)]
    if out_dtype != np.float32:
        sxx, outs = sxx.astype(out_dtype), [v.astype(out_dtype) for v in outs]
    return (t_out, f, sxx, *outs)


def plot_indices(freqs, cfrange_khz, max_nfreqs=2 ** 15):
    """Bins the viewer draws (drfview.py:1005-1023): those with ``1e3*cfrange[0] <= f <= 1e3*cfrange[1]``,
    decimated by ``fscale = ceil(n / max_nfreqs)`` starting at ``floor(fscale / 2)``.

    Returns ``(plotindices int64 array, plotfreqs, fscale)`` -- the three values the viewer stores
    (``stats["plotindices"]``, ``data["plotfreqs"]``, ``stats["fscale"]``).  An empty selection raises
    ``ValueError`` exactly where the reference's ``range(..., step=0)`` does.
    """
    freqs = np.asarray(freqs)
    kept = np.flatnonzero((freqs >= 1e3 * cfrange_khz[0]) & (freqs <= 1e3 * cfrange_khz[1]))
    fscale = -(-len(kept) // int(max_nfreqs))
    if fscale == 0:
        raise ValueError("range() arg 3 must not be zero")  # drfview.py:1017 with no frequency in range
    sel = kept[fscale // 2::fscale]
    return sel.astype(np.int64), freqs[sel], int(fscale)


def sti_plot_data(d1, sr, nfft, cfrange_khz, *, max_nfreqs=2 ** 15, crange=None, integrate=False, device=0, eps=_EPS,
                  ref=1.0, raw_pairs=False):
    """What the viewer draws from one ``sti_proc_data`` call, reduced on the GPU before the copy back.

    The viewer converts the STI and its median to dB (drfProc.py:308-310), keeps ``plotindices``
    (drfview.py:1005-1023, indexing at :1289-1295) and, for the PNG export, clips to the colour range
    (drfview.py:1515-1518).  Returns ``(plotfreqs, sxx_db[:, plotindices...], med_db[plotindices...])``
    with the reference's orientation ``(nfreq, ntime[, nsub])`` / ``(nfreq[, nsub])``; ``crange=None``
    leaves the values unclipped (the on-screen plot clips through ``vmin``/``vmax``).  Raw integer IQ with
    ``ref`` / ``raw_pairs`` as in ``sti_proc_data``.
    """
    import torch
    nfft = int(nfft)
    if _is_raw_iq(d1, raw_pairs):
        arr, shape = _raw_view(d1)
        out_dtype = np.float32
        if arr.dtype.fields is not None:  # ('r', 'i') -> trailing (re, im) axis of the base integer type
            base = next(iter(arr.dtype.fields.values()))[0]
            arr = arr.view(base).reshape(shape + (2,))
    else:
        arr, out_dtype = _as_c64(d1)
        shape = arr.shape
    if len(shape) not in (2, 3):
        raise ValueError("d1 must be (rows, ntime) or (rows, ntime, nsub)")
    rows, ntime = int(shape[0]), int(shape[1])
    nsub = int(shape[2]) if len(shape) == 3 else 1
    if rows < nfft:
        raise ValueError("window is longer than input signal")
    f = _freq_axis(nfft, sr)
    idx, plotfreqs, _ = plot_indices(f, cfrange_khz, max_nfreqs)
    plan = engine.get_plan(nfft, device)
    dev = torch.device("cuda", device)
    nfr = rows // nfft if integrate else 1
    span = nfr * nfft
    # (rows, ntime, nsub) C-order: sample stride ntime*nsub, column offset c*nsub, sub-channel stride 1;
    # only the rows the columns touch are uploaded
    flat = arr[:span].reshape(-1)
    offs = torch.from_numpy(np.arange(ntime, dtype=np.int64) * nsub).to(dev)
    lin, db = plan.run(torch.from_numpy(flat).to(dev), offs, nfr, nfft, sample_stride=ntime * nsub, sub_stride=1, nsub=nsub,
                       in_scale=1.0 / float(ref), eps=eps, want_lin=True, want_db=True, validate=True)
    _, med_db = plan.median(lin, eps=eps, want_lin=False, want_db=True)
    sel_db = plan.gather_bins(db, idx, clamp=crange)
    sel_med = plan.gather_bins(med_db, idx, clamp=crange)
    sxx = np.moveaxis(sel_db.cpu().numpy(), (0, 1, 2), (2, 1, 0))  # [nsub][ntime][k] -> (k, ntime, nsub)
    med = sel_med.cpu().numpy().T
    if len(shape) == 2:
        sxx, med = sxx[..., 0], med[..., 0]
    if out_dtype != np.float32:
        sxx, med = sxx.astype(out_dtype), med.astype(out_dtype)
    return plotfreqs, sxx, med


def get_ref(prop_dict):
    """Full-scale reference level (drfProc.py:182-201): 1.0 for float data, otherwise
    ``2**((precision-1) + 0.5*(size_bytes-1))``."""
    if prop_dict["H5Tget_class"] == 1:
        return 1.0
    npow = prop_dict["H5Tget_precision"] - 1.0
    npow += 0.5 * (prop_dict["H5Tget_size"] - 1.0)
    return 2**npow


# ------------------------------------------------------------------------------------------------
# I/O + framing (drfProc.py:59-179)
# ------------------------------------------------------------------------------------------------
class DrfInput:
    """Digital RF reader wrapper with the reference's attributes and methods.

    ``reader`` lets a caller (or a test) supply any object with ``get_channels``,
    ``get_properties``, ``get_bounds`` and ``read_vector``; by default a
    ``digital_rf.DigitalRFReader`` is opened on ``drfdir`` (drfProc.py:61-63).
    """

    def __init__(self, drfdir, reader=None):
        if reader is None:
            import digital_rf as drf  # not needed for the hot path; imported only here
            reader = drf.DigitalRFReader(str(Path(drfdir).expanduser()))
        self.drf_Obj = reader
        self.chan_2sub = {}
        self.chan_entries = {}
        self.last_read = {}
        self.sr_dict = {}
        self.ref_dict = {}
        self.time_bnds = (np.inf, -np.inf)
        self.bnds = {}
        for ichan in self.drf_Obj.get_channels():
            props = self.drf_Obj.get_properties(ichan)
            sr_f = Fraction(props["sample_rate_numerator"], props["sample_rate_denominator"])
            bnds = self.drf_Obj.get_bounds(ichan)
            num_sub = props["num_subchannels"]
            self.chan_2sub[ichan] = np.arange(num_sub)
            self.bnds[ichan] = bnds
            self.time_bnds = (min(self.time_bnds[0], float(bnds[0] / sr_f)),
                              max(self.time_bnds[1], float(bnds[1] / sr_f)))
            self.sr_dict[ichan] = sr_f
            self.ref_dict[ichan] = get_ref(props)
            self.last_read[ichan] = (None, None)
            for isub in range(num_sub):
                self.chan_entries[ichan + ":" + str(isub)] = (ichan, isub)

    def _split(self, chan_entry):
        if ":" in chan_entry:
            return self.chan_entries[chan_entry]
        return chan_entry, None

    def read(self, st_sample, n_sample, chan_entry, adj_bnds=False):
        """Read ``n_sample`` samples from ``st_sample``, normalised to full scale (drfProc.py:94-130)."""
        ichan, isub = self._split(chan_entry)
        bnds = self.drf_Obj.get_bounds(ichan)
        ref = self.ref_dict[ichan]
        if adj_bnds:
            st_sample = max(st_sample, bnds[0])
            n_sample = min(bnds[1], n_sample + st_sample) - st_sample
        if isub is None:
            x = self.drf_Obj.read_vector(st_sample, n_sample, ichan)
        else:
            x = self.drf_Obj.read_vector(st_sample, n_sample, ichan, isub)
        self.bnds[ichan] = bnds
        self.last_read[ichan] = (st_sample, n_sample)
        return x / ref

    def read_raw(self, st_sample, n_sample, chan_entry):
        """``read`` without the cast to complex64 and without ``/ ref`` (drfProc.py:124-129): the
        samples as stored (``DigitalRFReader.read_vector_raw``).  Extension for the raw-ingest path."""
        ichan, isub = self._split(chan_entry)
        x = self.drf_Obj.read_vector_raw(st_sample, n_sample, ichan) if isub is None else \
            self.drf_Obj.read_vector_raw(st_sample, n_sample, ichan, isub)
        self.bnds[ichan] = self.drf_Obj.get_bounds(ichan)
        self.last_read[ichan] = (st_sample, n_sample)
        return x

    def read_sti_raw(self, st_sample, chan_entry, en_sample, nfft, nint, ntime):
        """``read_sti`` on raw samples: ``(n_st, dout_raw, ref)``; same framing (drfProc.py:158-166)."""
        ichan, _ = self._split(chan_entry)
        n_sample = nint * nfft
        n_st = engine.frame_starts(st_sample, en_sample, nfft, nint, ntime)
        dlist = [self.read_raw(ist, n_sample, chan_entry)[:, np.newaxis] for ist in n_st]
        return n_st, np.concatenate(dlist, axis=1), self.ref_dict[ichan]

    def read_sti(self, st_sample, chan_entry, en_sample, nfft, nint, ntime):
        """``(n_st, dout)`` with ``dout`` shaped ``(nfft*nint, ntime[, nsub])`` (drfProc.py:132-167)."""
        n_sample = nint * nfft
        n_st = engine.frame_starts(st_sample, en_sample, nfft, nint, ntime)
        dlist = [self.read(ist, n_sample, chan_entry)[:, np.newaxis] for ist in n_st]
        return n_st, np.concatenate(dlist, axis=1)

    def bnds_update(self):
        """Refresh channel bounds (drfProc.py:169-179)."""
        for ichan in list(self.chan_2sub.keys()):
            bnds = self.drf_Obj.get_bounds(ichan)
            sr_f = self.sr_dict[ichan]
            self.bnds[ichan] = bnds
            self.time_bnds = (min(self.time_bnds[0], float(bnds[0] / sr_f)),
                              max(self.time_bnds[1], float(bnds[1] / sr_f)))


# ------------------------------------------------------------------------------------------------
# worker (drfProc.py:209-361) -- Qt is optional so the module imports headless
# ------------------------------------------------------------------------------------------------
try:  # pragma: no cover - PyQt5 is not installed in the build image
    from PyQt5.Qt import QRunnable
    from PyQt5.QtCore import QObject, pyqtSignal, pyqtSlot

    _HAVE_QT = True
except Exception:  # headless: minimal stand-ins with the same connect/emit surface
    _HAVE_QT = False

    class QRunnable:  # noqa: D401
        def __init__(self, *a, **k):
            pass

    class QObject:
        def __init__(self, *a, **k):
            pass

    class _BoundSignal:
        def __init__(self):
            self._slots = []

        def connect(self, slot):
            self._slots.append(slot)

        def emit(self, *args):
            for s in list(self._slots):
                s(*args)

    class pyqtSignal:  # descriptor giving every instance its own signal
        def __init__(self, *types):
            self._types = types

        def __set_name__(self, owner, name):
            self._name = "_sig_" + name

        def __get__(self, obj, owner=None):
            if obj is None:
                return self
            sig = obj.__dict__.get(self._name)
            if sig is None:
                sig = obj.__dict__[self._name] = _BoundSignal()
            return sig

    def pyqtSlot(*a, **k):
        return lambda fn: fn


class ThreadProcessorSignals(QObject):
    """Signals of the worker, same signatures as drfProc.py:458-465."""

    iterated = pyqtSignal(int, int, np.ndarray, np.ndarray, np.ndarray, np.ndarray)
    statsupdated = pyqtSignal(int, Fraction, int, float, int, tuple)
    terminated = pyqtSignal(int, int)


def _time_to_sample(time_s, sr):
    """``digital_rf.util.time_to_sample`` for a UNIX time in seconds: floor(time * rate)."""
    try:
        import digital_rf as drf
        return drf.util.time_to_sample(time_s, sr)
    except ImportError:
        return int(np.floor(Fraction(time_s) * Fraction(sr)))


def _sample_to_datetime(sample, sr):
    try:
        import digital_rf as drf
        return drf.util.sample_to_datetime(sample, sr)
    except ImportError:
        import datetime
        return datetime.datetime.fromtimestamp(0, datetime.timezone.utc) + datetime.timedelta(
            seconds=float(Fraction(int(sample)) / Fraction(sr)))


class DrfProcessor(QRunnable):
    """Worker with the reference's constructor, attributes, slots and signals (drfProc.py:209-361).

    Differences, all opt-in: ``integrate=True`` selects Mode A, ``device`` the GPU, ``reader`` a
    Digital RF reader object.  The dB conversion is fused into the GPU epilogue.
    """

    def __init__(self, datasource, drfdir, tabID, fftbins, n_int, ntime, *args, integrate=False, device=0,
                 reader=None, raw_ingest=False, resident=False, resident_max_bytes=32 << 30, **kwargs):
        super(DrfProcessor, self).__init__()
        self.drfIn = DrfInput(drfdir, reader=reader)
        self.drf_path = Path(drfdir).expanduser()
        self.tabID = tabID
        self.fftbins = fftbins
        self.n_int = n_int
        self.ntime = ntime
        self.integrate = integrate
        self.device = device
        self.raw_ingest = raw_ingest  # ship the stored integer samples to the GPU, fold 1/ref into the kernel
        # keep the recording window on the GPU between iterations and address frames in place
        # (SURVEY.md section 8(f) N2); falls back to the reference's per-bin reads when the window is sparse
        self.resident = resident
        self.resident_max_bytes = int(resident_max_bytes)
        self._cache = {}
        self._last_key, self._last_out = None, None  # previous resident pass (incremental recompute)
        self.recomputes_skipped = 0
        self.bnds = self.drfIn.time_bnds
        self.chan_listing = list(self.drfIn.chan_2sub.keys())
        self.sub_chan_list = list(self.drfIn.chan_entries.keys())
        self.isrunning = False
        self.signals = ThreadProcessorSignals()
        self.reason = 0
        self.max_iterations = kwargs.get("max_iterations")  # None = run until abort(), like the reference
        if datasource.lower() == "streaming":
            self.streaming = True
            self.streamtime = 30
        else:
            self.streaming = False
            self.streamtime = None
        if reader is None and not self.drf_path.exists():
            self.terminate(1)
        self.isrunning = True
        self.curchan = list(self.drfIn.chan_2sub.keys())[0]

    def iterate_once(self, i=0):
        """One pass of the worker loop body (drfProc.py:279-314); returns what ``iterated`` carries."""
        ichan = self.curchan
        sr = self.drfIn.sr_dict[ichan]
        self.drfIn.bnds_update()
        self.updatesettings(self.fftbins, self.n_int, self.ntime, self.drfIn.time_bnds[0], self.drfIn.time_bnds[1])
        if self.streaming:
            end_time = self.drfIn.time_bnds[-1]
            st_time = end_time - self.streamtime
        else:
            st_time, end_time = self.bnds
        s_samp = _time_to_sample(st_time, sr)
        e_samp = _time_to_sample(end_time, sr)
        if self.resident:
            out = self._iterate_resident(i, ichan, sr, s_samp, e_samp)
            if out is not None:
                return out
        ref = 1.0
        if self.raw_ingest and hasattr(self.drfIn.drf_Obj, "read_vector_raw") and self.drfIn.ref_dict[ichan] != 1.0:
            n_st, d1, ref = self.drfIn.read_sti_raw(s_samp, ichan, e_samp, self.fftbins, self.n_int, self.ntime)
            if d1.ndim == 2 and d1.dtype.fields is not None:
                d1 = d1[:, :, np.newaxis]  # the viewer wants (nfft, ntime, nsub), drfview.py:1289
        else:
            n_st, d1 = self.drfIn.read_sti(s_samp, ichan, e_samp, self.fftbins, self.n_int, self.ntime)
        time_ar = np.array([_sample_to_datetime(istime, int(sr)) for istime in n_st])
        f, sxx_dbfs, sxx_med_dbfs = sti_proc_data_db(d1, sr, self.fftbins, integrate=self.integrate,
                                                     device=self.device, ref=ref)
        self.freqs_all = f
        self.signals.iterated.emit(i, self.tabID, time_ar, self.freqs_all, sxx_dbfs, sxx_med_dbfs)
        return time_ar, f, sxx_dbfs, sxx_med_dbfs

    def _iterate_resident(self, i, ichan, sr, s_samp, e_samp):
        """Loop body on a device-resident recording window: one contiguous read of what is not yet on
        the GPU, frames addressed through the start table (no gather, no re-upload).  Returns None
        when the window is too sparse or too large to keep resident (the caller then reads per bin)."""
        nfft, nint, ntime = int(self.fftbins), int(self.n_int), int(self.ntime)
        frames = nint if self.integrate else 1
        n_st = engine.frame_starts(s_samp, e_samp, nfft, nint, ntime)
        lo, hi = int(n_st.min()), int(n_st.max()) + frames * nfft
        nsub = len(self.drfIn.chan_2sub[ichan])
        raw = self.raw_ingest and hasattr(self.drfIn.drf_Obj, "read_vector_raw") and self.drfIn.ref_dict[ichan] != 1.0
        ebytes = (4 if raw else 8) * nsub
        used = ntime * frames * nfft
        if (hi - lo) * ebytes > self.resident_max_bytes or (hi - lo) > 8 * used:
            return None
        cache = self._cache.get(ichan)
        if cache is None:
            cache = self._cache[ichan] = engine.RecordingCache(self.device)
        ref = self.drfIn.ref_dict[ichan]
        if raw:
            read = lambda st, n: self.drfIn.read_raw(st, n, ichan)
            in_scale = 1.0 / ref
        else:
            read = lambda st, n: self.drfIn.read(st, n, ichan)  # already divided by ref (drfProc.py:129)
            in_scale = 1.0
        read_before = cache.samples_read
        buf, base = cache.ensure(read, lo, hi)
        # Incremental recompute: the reference's loop redoes the whole STI every pass even when neither
        # the settings nor the recording window moved (drfProc.py:275-321).  Same start table, same
        # settings, nothing new read -> the arrays of the previous pass are the answer.
        key = (ichan, nfft, frames, nsub, raw, float(sr), n_st.tobytes())
        if self._last_key == key and cache.samples_read == read_before and self._last_out is not None:
            time_ar, f, sxx_dbfs, sxx_med_dbfs = self._last_out
            self.recomputes_skipped += 1
            self.freqs_all = f
            self.signals.iterated.emit(i, self.tabID, time_ar, f, sxx_dbfs, sxx_med_dbfs)
            return time_ar, f, sxx_dbfs, sxx_med_dbfs
        torch = engine._torch()
        plan = engine.get_plan(nfft, self.device)
        offs = torch.from_numpy(((n_st - base) * nsub).astype(np.int64)).to(buf.device)
        lin, db = plan.run(buf, offs, frames, nfft, sample_stride=nsub, sub_stride=1, nsub=nsub, in_scale=in_scale,
                           eps=_EPS, want_lin=True, want_db=True, validate=True)
        _, mdb = plan.median(lin, eps=_EPS, want_lin=False, want_db=True)
        sxx_dbfs = db.permute(2, 1, 0).cpu().numpy()      # (nfft, ntime, nsub) as the viewer indexes it
        sxx_med_dbfs = mdb.t().cpu().numpy()               # (nfft, nsub)
        time_ar = np.array([_sample_to_datetime(istime, int(sr)) for istime in n_st])
        f = _freq_axis(nfft, sr)
        self._last_key, self._last_out = key, (time_ar, f, sxx_dbfs, sxx_med_dbfs)
        self.freqs_all = f
        self.signals.iterated.emit(i, self.tabID, time_ar, f, sxx_dbfs, sxx_med_dbfs)
        return time_ar, f, sxx_dbfs, sxx_med_dbfs

    @pyqtSlot()
    def run(self):
        counts = 0
        while not self.isrunning:
            counts += 1
            if self.reason:
                return
            elif counts > 100:
                self.terminate(3)
                return
            timemodule.sleep(0.1)
        if self.reason:
            return
        try:
            i = -1
            while self.isrunning:
                i += 1
                self.iterate_once(i)
                if self.max_iterations is not None and i + 1 >= self.max_iterations:
                    break
                timemodule.sleep(0.08 if self.streaming else 0.1)
        except Exception:
            self.isrunning = False
            self.terminate(4)
            trace_error()

    @pyqtSlot(float, float, float, float, float)
    def updatesettings_slot(self, fftbins, nint, ntime, bnd_beg, bnd_end):
        self.updatesettings(fftbins, nint, ntime, bnd_beg, bnd_end)

    def updatesettings(self, fftbins, nint, ntime, bnd_beg, bnd_end):
        self.fftbins = int(fftbins)
        self.n_int = int(nint)
        self.ntime = int(ntime)
        self.bnds = (bnd_beg, bnd_end)
        sr = self.drfIn.sr_dict[self.curchan]
        self.signals.statsupdated.emit(self.tabID, sr, self.fftbins, self.n_int, self.ntime, self.bnds)

    @pyqtSlot()
    def abort(self):
        self.terminate(0)
        return

    def terminate(self, reason):
        self.reason = reason
        self.isrunning = False
        self.signals.terminated.emit(self.tabID, reason)
