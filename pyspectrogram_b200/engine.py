"""Device-resident STI engine: torch tensors in, torch tensors out, compute in libpsgb200.so.

PyTorch is used for device memory and streams only; every kernel on this path is hand-written
sm_100a CUDA behind the C ABI (``include/psg_b200.h``).  Layout and mode definitions follow
SURVEY.md section 8(a):

* IQ element (sample n, sub-channel s) lives at ``iq[n*sample_stride + s*sub_stride]``
* column ``c`` averages ``frames_per_col`` frames starting at ``col_offsets[c] + k*hop*sample_stride``
* Mode R (``sti_proc_data`` as shipped, drfProc.py:364-403): ``frames_per_col=1``
* Mode A (per-bin averaging, read_sti's ``nint`` frames, drfProc.py:158): ``frames_per_col=nint, hop=nfft``
* Mode S (``proc_data``, drfProc.py:406-453): ``frames_per_col=n_int, hop=nfft-nfft//8``

Images are ``[nsub][ncol][nfft]`` float32, fftshifted (bin ``nfft/2`` is DC).
"""
from __future__ import annotations

import ctypes as C
import threading

import numpy as np

from . import _lib

DB_EPS = 1e-15  # drfProc.py:308
KAISER_BETA = 1.7  # drfProc.py:386, :435


def _torch():
    import torch
    return torch


class StiPlan:
    """One FFT length on one device (window + twiddle tables live on the GPU).

    A plan owns scratch and is used by one thread at a time (guarded by a lock so the viewer's
    worker threads, drfview.py:177-178, may share one safely).
    """

    def __init__(self, nfft: int, device: int = 0, window=("kaiser", KAISER_BETA)):
        self._lib = _lib.load()
        kind, beta = self._window_kind(window)
        handle = C.c_void_p()
        _lib.check(self._lib.psg_plan_create(C.byref(handle), int(nfft), kind, float(beta), int(device)))
        self._h = handle
        self.nfft = int(nfft)
        self.device = int(device)
        self._lock = threading.Lock()

    @staticmethod
    def _window_kind(window):
        if isinstance(window, str):
            window = (window, 0.0)
        name = window[0].lower()
        if name == "kaiser":
            return _lib.PSG_WINDOW_KAISER, float(window[1])
        if name in ("boxcar", "rect", "rectangular"):
            return _lib.PSG_WINDOW_BOXCAR, 0.0
        raise ValueError(f"window {window!r} is not available on the GPU path (kaiser, boxcar)")

    def close(self):
        if getattr(self, "_h", None):
            self._lib.psg_plan_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- introspection ---------------------------------------------------------------------
    def window_table(self) -> np.ndarray:
        """fp32 ``w/sum(w)`` as uploaded to the device."""
        out = np.empty(self.nfft, np.float32)
        _lib.check(self._lib.psg_plan_window(self._h, out.ctypes.data_as(C.c_void_p)))
        return out

    @property
    def variant(self) -> str:
        return self._lib.psg_plan_variant(self._h).decode()

    # ---- device path -----------------------------------------------------------------------
    def run(self, iq, col_offsets, frames_per_col=1, hop=None, *, sample_stride=1, sub_stride=0, nsub=1,
            in_scale=1.0, eps=DB_EPS, want_lin=True, want_db=False, out_lin=None, out_db=None, validate=False):
        """Fused frame->window->FFT->|X|^2->mean->fftshift->(dB) on device-resident IQ.

        ``iq``: CUDA tensor, complex64 (or float32 viewed as interleaved re/im), or raw integer IQ as
        int16 / int8 (re, im) pairs -- Digital RF's native sample formats; pass ``in_scale=1/ref``
        (``get_ref``, drfProc.py:182-201) instead of dividing on the host (drfProc.py:129).
        Strides and offsets are always in complex elements.
        ``col_offsets``: CUDA int64 tensor ``[ncol]`` of element offsets into ``iq``.
        Returns ``(lin, db)`` tensors ``[nsub][ncol][nfft]`` (``None`` for the one not requested).
        Work is enqueued on torch's current stream; nothing synchronises.

        ``iq`` must be contiguous (the kernels address its storage through the strides given here, not the
        tensor's own).  ``validate=True`` goes through ``psg_sti_run_checked``: the offset table is clamped on
        the device to what ``iq`` holds, and a table that reached outside (e.g. offsets computed against an
        old ``RecordingCache`` base) raises ``IndexError`` after the launch -- one 4-byte read back, so this
        form synchronises.
        """
        torch = _torch()
        if not iq.is_cuda or not col_offsets.is_cuda:
            raise ValueError("iq and col_offsets must be CUDA tensors (use StiPlan.host for host arrays)")
        iq_type = {torch.complex64: _lib.PSG_IQ_C64, torch.float32: _lib.PSG_IQ_C64, torch.int16: _lib.PSG_IQ_CI16,
                   torch.int8: _lib.PSG_IQ_CI8}.get(iq.dtype)
        if iq_type is None:
            raise TypeError(f"iq must be complex64 (or its float32 view), or raw int16 / int8 (re, im) pairs; got {iq.dtype}")
        if col_offsets.dtype != torch.int64 or not col_offsets.is_contiguous():
            raise TypeError("col_offsets must be a contiguous int64 tensor")
        if not iq.is_contiguous():
            raise ValueError("iq must be contiguous (pass strides through sample_stride / sub_stride)")
        if iq.device.index != self.device or col_offsets.device.index != self.device:
            raise ValueError("tensors are not on the plan's device")
        ncol = int(col_offsets.numel())
        hop = self.nfft if hop is None else int(hop)
        dev = iq.device
        shape = (int(nsub), ncol, self.nfft)
        if want_lin and out_lin is None:
            out_lin = torch.empty(shape, dtype=torch.float32, device=dev)
        if want_db and out_db is None:
            out_db = torch.empty(shape, dtype=torch.float32, device=dev)
        for o in (out_lin, out_db):
            if o is not None and (o.dtype != torch.float32 or not o.is_contiguous() or o.numel() != nsub * ncol * self.nfft):
                raise ValueError("output tensors must be contiguous float32 [nsub][ncol][nfft]")
        stream = torch.cuda.current_stream(dev).cuda_stream
        p_lin = C.c_void_p(out_lin.data_ptr() if out_lin is not None else None)
        p_db = C.c_void_p(out_db.data_ptr() if out_db is not None else None)
        if validate:
            iq_elems = int(iq.numel()) if iq.dtype == torch.complex64 else int(iq.numel()) // 2
            flag = torch.zeros(1, dtype=torch.int32, device=dev)
            with self._lock:
                _lib.check(self._lib.psg_sti_run_checked(
                    self._h, C.c_void_p(iq.data_ptr()), iq_type, iq_elems, int(sample_stride), int(sub_stride), int(nsub),
                    C.c_void_p(col_offsets.data_ptr()), ncol, int(frames_per_col), hop, float(in_scale), float(eps),
                    p_lin, p_db, C.c_void_p(flag.data_ptr()), C.c_void_p(stream)))
            if int(flag.item()):
                raise IndexError(f"col_offsets reach outside the {iq_elems}-element recording (columns were clamped)")
            return out_lin, out_db
        with self._lock:
            _lib.check(self._lib.psg_sti_run_typed(
                self._h, C.c_void_p(iq.data_ptr()), iq_type, int(sample_stride), int(sub_stride), int(nsub),
                C.c_void_p(col_offsets.data_ptr()), ncol, int(frames_per_col), hop, float(in_scale), float(eps),
                p_lin, p_db, C.c_void_p(stream)))
        return out_lin, out_db

    def median(self, img, *, eps=DB_EPS, want_lin=True, want_db=False, out_lin=None, out_db=None):
        """``np.median(sxx, axis=1)`` (drfProc.py:401) of a ``[nsub][ncol][nfft]`` linear image.  ``out_lin`` /
        ``out_db``: optional contiguous float32 ``[nsub][nfft]`` destinations (e.g. rows of a ``dist.PeerImage``)."""
        torch = _torch()
        if not img.is_cuda or img.dtype != torch.float32 or not img.is_contiguous() or img.dim() != 3:
            raise ValueError("img must be a contiguous CUDA float32 [nsub][ncol][nfft] tensor")
        nsub, ncol, nfft = (int(v) for v in img.shape)
        for o in (out_lin, out_db):
            if o is not None and (o.dtype != torch.float32 or not o.is_contiguous() or o.numel() != nsub * nfft):
                raise ValueError("median outputs must be contiguous float32 [nsub][nfft]")
        lin = out_lin if out_lin is not None else (torch.empty((nsub, nfft), dtype=torch.float32, device=img.device) if want_lin else None)
        db = out_db if out_db is not None else (torch.empty((nsub, nfft), dtype=torch.float32, device=img.device) if want_db else None)
        stream = torch.cuda.current_stream(img.device).cuda_stream
        with self._lock:
            _lib.check(self._lib.psg_median_time(
                self._h, C.c_void_p(img.data_ptr()), nsub, ncol, nfft, float(eps),
                C.c_void_p(lin.data_ptr() if lin is not None else None),
                C.c_void_p(db.data_ptr() if db is not None else None), C.c_void_p(stream)))
        return lin, db

    def minmax(self, img, *, eps=DB_EPS, want_lin=True, want_db=False):
        """Minimum and maximum over time of a ``[nsub][ncol][nfft]`` linear image (the "min" and "max"
        spectra of proc_data's docstring, drfProc.py:430-433).  Returns ``(min_lin, max_lin, min_db, max_db)``."""
        torch = _torch()
        if not img.is_cuda or img.dtype != torch.float32 or not img.is_contiguous() or img.dim() != 3:
            raise ValueError("img must be a contiguous CUDA float32 [nsub][ncol][nfft] tensor")
        nsub, ncol, nfft = (int(v) for v in img.shape)
        outs = [torch.empty((nsub, nfft), dtype=torch.float32, device=img.device) if w else None
                for w in (want_lin, want_lin, want_db, want_db)]
        stream = torch.cuda.current_stream(img.device).cuda_stream
        with self._lock:
            _lib.check(self._lib.psg_minmax_time(
                self._h, C.c_void_p(img.data_ptr()), nsub, ncol, nfft, float(eps),
                *[C.c_void_p(o.data_ptr() if o is not None else None) for o in outs], C.c_void_p(stream)))
        return tuple(outs)

    def gather_bins(self, img, indices, *, clamp=None):
        """``img[..., indices]`` (the viewer's ``plotindices``, drfview.py:1005-1023) on the device, with the
        optional colour-range clip of the PNG export (drfview.py:1515-1518).  ``img``: contiguous CUDA
        float32 ``[..., nfft]``; ``indices``: int sequence / array / CUDA int32 tensor."""
        torch = _torch()
        if not img.is_cuda or img.dtype != torch.float32 or not img.is_contiguous() or img.dim() < 1:
            raise ValueError("img must be a contiguous CUDA float32 tensor [..., nfft]")
        nfft = int(img.shape[-1])
        if not isinstance(indices, torch.Tensor):
            ind = np.asarray(indices)
            if ind.size and (ind.min() < -nfft or ind.max() >= nfft):
                raise IndexError(f"index out of bounds for axis of size {nfft}")
            indices = torch.from_numpy(np.ascontiguousarray(np.where(ind < 0, ind + nfft, ind).astype(np.int32))).to(img.device)
        if indices.dtype != torch.int32 or not indices.is_contiguous() or indices.dim() != 1:
            raise TypeError("indices must be a contiguous 1-D int32 tensor")
        count = int(indices.numel())
        rows = int(img.numel() // nfft)
        out = torch.empty(tuple(img.shape[:-1]) + (count,), dtype=torch.float32, device=img.device)
        if count == 0 or rows == 0:
            return out
        lo, hi = (1.0, 0.0) if clamp is None else (float(clamp[0]), float(clamp[1]))
        stream = torch.cuda.current_stream(img.device).cuda_stream
        with self._lock:
            _lib.check(self._lib.psg_gather_bins(
                self._h, C.c_void_p(img.data_ptr()), rows, nfft, C.c_void_p(indices.data_ptr()), count, lo, hi,
                C.c_void_p(out.data_ptr()), C.c_void_p(stream)))
        return out

    # ---- host path -------------------------------------------------------------------------
    def host(self, iq: np.ndarray, col_offsets, frames_per_col=1, hop=None, *, sample_stride=1, sub_stride=0,
             nsub=1, in_scale=1.0, eps=DB_EPS, want=("lin", "med")):
        """Host complex64 (or raw int16 / int8 pair) array in, host float32 arrays out (H2D, kernels,
        D2H inside the call).

        ``want``: any of ``"lin" "db" "med" "med_db"``.  Returns a dict of numpy arrays:
        images ``[nsub][ncol][nfft]``, medians ``[nsub][nfft]``.
        """
        iq, iq_type, nelem = _host_iq(iq)
        offs = np.ascontiguousarray(col_offsets, dtype=np.int64)
        ncol = int(offs.size)
        hop = self.nfft if hop is None else int(hop)
        out = {}
        ptr = {}
        for key, shape in (("lin", (nsub, ncol, self.nfft)), ("db", (nsub, ncol, self.nfft)),
                           ("med", (nsub, self.nfft)), ("med_db", (nsub, self.nfft))):
            if key in want:
                out[key] = np.empty(shape, np.float32)
                ptr[key] = out[key].ctypes.data_as(C.c_void_p)
            else:
                ptr[key] = C.c_void_p(None)
        with self._lock:
            _lib.check(self._lib.psg_sti_host_typed(
                self._h, iq.ctypes.data_as(C.c_void_p), iq_type, nelem, int(sample_stride), int(sub_stride),
                int(nsub), offs.ctypes.data_as(C.c_void_p), ncol, int(frames_per_col), hop, float(in_scale),
                float(eps), ptr["lin"], ptr["db"], ptr["med"], ptr["med_db"]))
        return out


def _host_iq(iq):
    """``(array, PSG_IQ_*, number of complex elements)`` for a host IQ array: complex64, or raw
    integer IQ as an int16 / int8 array whose last axis is (re, im), or Digital RF's structured
    dtype with two integer fields ('r', 'i')."""
    if not isinstance(iq, np.ndarray) or not iq.flags.c_contiguous:
        raise TypeError("iq must be a C-contiguous numpy array")
    if iq.dtype == np.complex64:
        return iq, _lib.PSG_IQ_C64, int(iq.size)
    if iq.dtype.fields is not None and len(iq.dtype.fields) == 2:
        base = {np.dtype(np.int16): _lib.PSG_IQ_CI16, np.dtype(np.int8): _lib.PSG_IQ_CI8}
        kinds = {v[0] for v in iq.dtype.fields.values()}
        if len(kinds) == 1 and next(iter(kinds)) in base and iq.dtype.itemsize == 2 * next(iter(kinds)).itemsize:
            return iq, base[next(iter(kinds))], int(iq.size)
    if iq.dtype in (np.int16, np.int8):
        if iq.size % 2:
            raise ValueError("raw integer IQ needs (re, im) pairs")
        return iq, (_lib.PSG_IQ_CI16 if iq.dtype == np.int16 else _lib.PSG_IQ_CI8), int(iq.size // 2)
    raise TypeError(f"iq dtype {iq.dtype} is not supported (complex64, int16 pairs, int8 pairs)")


class RecordingCache:
    """Device-resident window of one channel's recording.

    The reference re-reads ``ntime`` slices from disk and rebuilds the ``(nfft*nint, ntime, nsub)``
    array on every pass of the worker loop (drfProc.py:160-166, :275-321).  This cache keeps the
    samples ``[lo, hi)`` of a channel on the GPU as ``[sample][nsub]`` (complex64 or raw integer
    pairs); ``ensure`` reads only what is missing -- nothing when the same window is asked again, just
    the new tail when a streaming window slides forward -- and the STI kernel addresses frames in
    place through the start table (``col_offsets = (n_st - lo) * nsub``).
    """

    def __init__(self, device: int = 0, slack: float = 0.25):
        self.device = int(device)
        self.slack = float(slack)
        self.lo = self.hi = 0
        self.buf = None       # torch tensor [capacity, nsub(, 2)]
        self.samples_read = 0  # samples fetched from the reader so far (tests / stats)
        self.appended_in_place = 0  # slides that fitted the slack (no reallocation, no copy of the kept part)

    def _upload(self, arr):
        torch = _torch()
        arr = np.ascontiguousarray(arr)
        if arr.dtype.fields is not None:  # Digital RF's ('r','i') integer pairs
            base = next(iter(arr.dtype.fields.values()))[0]
            arr = arr.view(base).reshape(arr.shape + (2,))
        if arr.dtype == np.complex64:
            arr = arr.reshape(arr.shape[0], -1)
        elif arr.dtype in (np.int16, np.int8):
            arr = arr.reshape(arr.shape[0], -1, 2)
        else:
            arr = arr.astype(np.complex64).reshape(arr.shape[0], -1)
        return torch.from_numpy(arr).to(f"cuda:{self.device}", non_blocking=False)

    def ensure(self, read, lo: int, hi: int):
        """Make ``[lo, hi)`` resident; ``read(start, n)`` returns ``n`` samples from ``start`` as an
        ``(n,)`` / ``(n, nsub)`` array.  Returns the tensor and the absolute index of its row 0."""
        torch = _torch()
        lo, hi = int(lo), int(hi)
        if self.buf is not None and self.lo <= lo and hi <= self.hi:
            return self.buf, self.lo
        if self.buf is not None and self.lo <= lo < self.hi <= hi:
            # sliding window: keep [lo, self.hi), fetch only [self.hi, hi)
            new = self._upload(read(self.hi, hi - self.hi))
            self.samples_read += hi - self.hi
            if hi - self.lo <= int(self.buf.shape[0]):
                # the tail fits the slack behind the resident samples: appended in place, nothing moves and the
                # base stays (the rows before lo are simply no longer asked for)
                self.buf[self.hi - self.lo: hi - self.lo] = new
                self.hi = hi
                self.appended_in_place += 1
                return self.buf, self.lo
            keep = self.buf[lo - self.lo: self.hi - self.lo]
            cap = int((hi - lo) * (1.0 + self.slack)) + 1
            out = torch.empty((cap,) + tuple(keep.shape[1:]), dtype=keep.dtype, device=keep.device)
            out[: keep.shape[0]] = keep
            out[keep.shape[0]: keep.shape[0] + new.shape[0]] = new
            self.buf, self.lo, self.hi = out, lo, hi
            return self.buf, self.lo
        new = self._upload(read(lo, hi - lo))
        cap = int((hi - lo) * (1.0 + self.slack)) + 1
        self.buf = torch.empty((cap,) + tuple(new.shape[1:]), dtype=new.dtype, device=new.device)
        self.buf[: new.shape[0]] = new
        self.samples_read += hi - lo
        self.lo, self.hi = lo, hi
        return self.buf, self.lo

    def drop(self):
        self.buf = None
        self.lo = self.hi = 0


_plans = threading.local()


def get_plan(nfft: int, device: int = 0, window=("kaiser", KAISER_BETA)) -> StiPlan:
    """Plan cache keyed by (nfft, device, window), ONE CACHE PER THREAD.

    A plan owns device scratch (partial sums, the split path's buffers, staging of the host path) that the
    kernels of a call keep using after the call has returned to the host -- ``run`` / ``median`` / ``minmax``
    enqueue on the caller's stream and do not synchronise.  ``StiPlan._lock`` serialises the enqueueing only, so
    two of the viewer's worker threads (drfview.py:177-178) on different streams must not share a plan
    (``include/psg_b200.h``: one plan per worker thread).  The cache lives in thread-local storage: a thread's
    plans are destroyed with it.
    """
    key = (int(nfft), int(device), tuple(window) if not isinstance(window, str) else (window,))
    cache = getattr(_plans, "cache", None)
    if cache is None:
        cache = _plans.cache = {}
    plan = cache.get(key)
    if plan is None:
        plan = cache[key] = StiPlan(nfft, device, window)
    return plan


def frame_starts(st_sample, en_sample, nfft, nint, ntime) -> np.ndarray:
    """Frame index table of ``DrfInput.read_sti`` (drfProc.py:158-159).

    Evaluated by numpy's own ``linspace(..., dtype=int)`` on the host so that the float64
    quantisation of epoch-sized sample indices is reproduced bit for bit (SURVEY.md section 0,
    trap 2); the int64 table is what the kernel consumes.
    """
    n_sample = int(nint) * int(nfft)
    return np.linspace(st_sample, en_sample - n_sample, int(ntime), dtype=int)


def launch_count() -> int:
    return int(_lib.load().psg_launch_count())


def set_variant(name=None):
    _lib.check(_lib.load().psg_debug_set_variant(name.encode() if name else None))


def set_split_scratch(nbytes: int):
    """Scratch bytes per chunk of the large-nfft split path (default cap 2 GiB)."""
    _lib.check(_lib.load().psg_debug_set_split_scratch(int(nbytes)))


def set_host_chunk(nbytes: int):
    """Span above which ``StiPlan.host`` streams the recording in column chunks (default 1 GiB)."""
    _lib.check(_lib.load().psg_debug_set_host_chunk(int(nbytes)))


def set_force_generic(on: bool):
    _lib.check(_lib.load().psg_debug_set_force_generic(1 if on else 0))


def variants():
    lib = _lib.load()
    return [(lib.psg_variant_name(i).decode(), lib.psg_variant_logn(i)) for i in range(lib.psg_variant_count())]
