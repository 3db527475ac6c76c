"""CPU-side checks of the C ABI and the host logic around it (no compute calls: there is no GPU
here).  The library must load, export every symbol ``include/psg_b200.h`` declares, and fail loudly
-- never fall back -- when no sm_100 device is present."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from pyspectrogram_b200 import _lib, engine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "psg_b200.h")


def _declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(psg_[a-z0-9_]+)\s*\(", text)))


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_library_exports_every_declared_symbol():
    lib = C.CDLL(_lib.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} is declared in include/psg_b200.h but not exported"
    # and the ctypes binding describes exactly the declared surface
    assert sorted(_lib.exported_symbols()) == declared


def test_version_and_variant_table():
    lib = _lib.load()
    assert lib.psg_version() == 2  # PSG_ABI_VERSION: psg_sti_run_checked, psg_debug_* (thread-local) knobs
    names = [lib.psg_variant_name(i).decode() for i in range(lib.psg_variant_count())]
    assert len(names) == len(set(names)) and len(names) >= 20
    for i, name in enumerate(names):
        logn = lib.psg_variant_logn(i)
        radices = [int(r) for r in name.split("_")[1].split("x")]
        assert int(np.prod(radices)) == 1 << logn, name
    assert lib.psg_variant_name(10 ** 6) == b"" and lib.psg_variant_logn(-1) == -1
    with pytest.raises(ValueError):
        engine.set_variant("no_such_variant")
    assert b"no_such_variant" in lib.psg_last_error()
    engine.set_variant(None)


@pytest.mark.parametrize("nfft", [2, 16, 256, 1024, 4096, 65536])
def test_window_table_matches_scipy_kaiser(nfft):
    """w/sum(w) as the plan uploads it (fp64 math on the host, fp32 storage) against scipy's
    periodic Kaiser(1.7) (drfProc.py:386) and the golden table from the reference run."""
    import scipy.signal as sig
    lib = _lib.load()
    out = np.empty(nfft, np.float32)
    s = C.c_double()
    assert lib.psg_window_table(nfft, _lib.PSG_WINDOW_KAISER, 1.7, out.ctypes.data_as(C.c_void_p), C.byref(s)) == 0
    w = sig.get_window(("kaiser", 1.7), nfft)
    assert abs(s.value - w.sum()) <= 1e-12 * w.sum()
    assert np.array_equal(out, (w / w.sum()).astype(np.float32)) or np.abs(out / (w / w.sum()) - 1).max() <= 6e-8
    g = np.load(os.path.join(ROOT, "tests", "golden", "kaiser_tables.npz"))
    if f"w{nfft}" in g:
        assert np.abs(out.astype(np.float64) * s.value - g[f"w{nfft}"]).max() <= 1e-6
    box = np.empty(nfft, np.float32)
    assert lib.psg_window_table(nfft, _lib.PSG_WINDOW_BOXCAR, 0.0, box.ctypes.data_as(C.c_void_p), None) == 0
    assert np.all(box == np.float32(1.0 / nfft))
    assert lib.psg_window_table(nfft, 99, 0.0, box.ctypes.data_as(C.c_void_p), None) == _lib.PSG_ERR_ARG


@pytest.mark.skipif(_have_gpu(), reason="checks the no-device behaviour")
def test_no_device_fails_loudly_and_there_is_no_fallback():
    lib = _lib.load()
    h = C.c_void_p()
    rc = lib.psg_plan_create(C.byref(h), 1024, _lib.PSG_WINDOW_KAISER, 1.7, 0)
    assert rc == _lib.PSG_ERR_NODEVICE and not h.value
    assert len(lib.psg_last_error()) > 0
    with pytest.raises(_lib.PsgError):
        engine.StiPlan(1024)
    from pyspectrogram_b200 import drfProc as dp
    d1 = np.zeros((1024, 4), np.complex64)
    with pytest.raises(RuntimeError):
        dp.sti_proc_data(d1, 1.0e6, 1024)
    with pytest.raises(RuntimeError):
        dp.proc_data(np.zeros(8192, np.complex64), 1.0e6, 1024, 0.004)


def test_argument_errors_map_to_python_exceptions():
    lib = _lib.load()
    h = C.c_void_p()
    assert lib.psg_plan_create(C.byref(h), 1 << 21, 0, 1.7, 0) == _lib.PSG_ERR_UNSUPPORTED  # above PSG_MAX_NFFT
    assert b"outside" in lib.psg_last_error()
    assert lib.psg_plan_create(C.byref(h), 1, 0, 1.7, 0) == _lib.PSG_ERR_UNSUPPORTED
    assert lib.psg_plan_create(None, 1024, 0, 1.7, 0) == _lib.PSG_ERR_ARG
    assert lib.psg_sti_run(None, None, 1, 0, 1, None, 1, 1, 1, 1.0, 1e-15, None, None, None) == _lib.PSG_ERR_ARG
    assert lib.psg_plan_destroy(None) == 0
    with pytest.raises(NotImplementedError):
        _lib.check(_lib.PSG_ERR_UNSUPPORTED)
    with pytest.raises(ValueError):
        _lib.check(_lib.PSG_ERR_ARG)
    assert lib.psg_debug_set_split_scratch(1) == _lib.PSG_ERR_ARG
    assert lib.psg_sti_run_checked(None, None, 0, 0, 1, 0, 1, None, 1, 1, 1, 1.0, 1e-15, None, None, None, None) == _lib.PSG_ERR_ARG


def test_debug_knobs_are_per_thread():
    """psg_debug_set_variant in one thread does not change what another thread's calls pick (the viewer's seven
    workers share the library, drfview.py:177-178): a bad name fails only for its own call, and a knob set in a
    worker thread is gone with the thread."""
    import threading
    lib = _lib.load()
    seen = {}

    def worker():
        seen["set"] = lib.psg_debug_set_items_per_slot(7)
        seen["bad"] = lib.psg_debug_set_variant(b"no_such_variant")
        seen["msg"] = lib.psg_last_error()

    th = threading.Thread(target=worker)
    th.start()
    th.join()
    assert seen["set"] == 0 and seen["bad"] == _lib.PSG_ERR_ARG and b"no_such_variant" in seen["msg"]
    # this thread's state is untouched: its own override is still the automatic choice (setting "" succeeds and
    # the error message of the other thread is not visible here)
    assert lib.psg_debug_set_variant(None) == 0


def test_plan_cache_is_per_thread():
    """engine.get_plan hands every thread its own plans (a plan's scratch is in use by enqueued
    kernels after the call returns, so two threads on different streams must not share one)."""
    import threading
    got = {}

    class FakePlan:
        def __init__(self, nfft, device, window):
            self.key = (nfft, device)

    orig = engine.StiPlan
    engine.StiPlan = FakePlan
    try:
        a = engine.get_plan(1024)
        assert engine.get_plan(1024) is a and engine.get_plan(2048) is not a
        th = threading.Thread(target=lambda: got.update(p=engine.get_plan(1024), q=engine.get_plan(1024)))
        th.start()
        th.join()
        assert got["p"] is got["q"] and got["p"] is not a
    finally:
        engine.StiPlan = orig
        engine._plans.cache = {}


def test_host_side_input_checks_need_no_device():
    from pyspectrogram_b200 import drfProc as dp
    with pytest.raises(ValueError):
        dp.sti_proc_data(np.zeros((100, 4), np.complex64), 1.0, 128)  # fewer rows than nfft (scipy: ValueError)
    with pytest.raises(ValueError):
        dp.sti_proc_data(np.zeros(4096, np.complex64), 1.0, 128)  # 1-D input
    with pytest.raises(ValueError):
        dp.proc_data(np.zeros((4096, 2), np.complex64), 1.0, 128, 0.1)
    with pytest.raises(TypeError):
        engine._host_iq(np.zeros(8, np.float64))
    with pytest.raises(ValueError):
        engine._host_iq(np.zeros(7, np.int16))
    assert dp.get_ref(dict(H5Tget_class=1, H5Tget_precision=32, H5Tget_size=4)) == 1.0
    assert dp.get_ref(dict(H5Tget_class=0, H5Tget_precision=16, H5Tget_size=2)) == 2 ** 15.5
    f = dp._freq_axis(8, 8.0)
    assert np.array_equal(f, np.array([-4., -3., -2., -1., 0., 1., 2., 3.]))


@pytest.mark.skipif(_have_gpu(), reason="checks the no-device behaviour")
def test_plain_c_caller_builds_and_fails_loudly_without_a_device(tmp_path):
    """gcc + include/psg_b200.h + libpsgb200.so is all a C host needs; without a GPU the program
    gets PSG_ERR_NODEVICE from psg_plan_create (exit code 3), not a silent CPU result."""
    import subprocess
    exe = str(tmp_path / "abi_smoke")
    libdir = os.path.join(ROOT, "pyspectrogram_b200")
    subprocess.run(["gcc", "-O2", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "c", "abi_smoke.c"), "-o", exe, "-L", libdir, "-lpsgb200", "-lm",
                    f"-Wl,-rpath,{libdir}"], check=True)
    res = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert res.returncode == 3 and "psg_plan_create: -4" in res.stdout
