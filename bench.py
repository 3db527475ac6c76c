#!/usr/bin/env python
"""Benchmark of the fused STFT -> PSD -> STI hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W          # this repo's CUDA path
    python bench.py --impl reference --gpus N ...          # the reference's CPU path (scipy)

A step = one pass of the hot path over one batch of synthetic IQ: per GPU, BASELINE config 2
(1 channel, 25 MS/s, 60 s = 1.5e9 complex64 samples, nfft=4096, 1000 STI bins, every sample
read once: nint=366, Mode A) -> dB image + time-median.  N > 1 is weak scaling: one such channel
per GPU (the channel sharding of SURVEY.md section 8(e)), each rank computes its own columns and one
NCCL gather assembles the dB image (and the per-channel median rows) on rank 0 inside the timed region.

Prints ONE JSON line (rank 0).  ``value`` = Msamples/s with inputs resident in HBM; ``e2e`` = the same
metric through the host-buffer C-ABI call (pinned host IQ -> H2D -> kernels -> D2H of the image);
``roofline`` = algorithmic bytes of the fused kernel / its CUDA-event time vs the measured HBM
peak; ``cpu_baseline`` = the oracle port of the reference path timed on this box's host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "IQ Msamples/s and STI columns/s at 1/2/4/8 B200; % of HBM roofline"
FS = 25.0e6
NFFT = 4096
NTIME = 1000
SECONDS = 60
NSAMP = int(FS * SECONDS)            # 1.5e9 samples per channel
NINT = NSAMP // NTIME // NFFT        # 366: full coverage
HBM_FALLBACK_GBS = 6650.0


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.rows = []
        self._stop = threading.Event()
        self._th = None

    def _loop(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) >= 7:
                    self.rows.append(parts)
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        self._th = threading.Thread(target=self._loop, daemon=True)
        self._th.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._th.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(float(r[0]) for r in self.rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "samples": len(self.rows), "power_w_max": max(float(r[2]) for r in self.rows),
                "note": "sampled over the timed steps and an untimed continuation of the same steps (~1 s)"}


def synth_iq_device(torch, n, seed, device, chunk=1 << 26):
    """-40 dBFS complex noise + a -20 dBFS tone at 0.123 fs, made on device."""
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    iq = torch.empty(n, dtype=torch.complex64, device=device)
    view = torch.view_as_real(iq)
    sigma = 10 ** (-40 / 20) / np.sqrt(2)
    for lo in range(0, n, chunk):
        hi = min(n, lo + chunk)
        view[lo:hi].normal_(0.0, sigma, generator=gen)
        ph = (torch.arange(lo, hi, device=device, dtype=torch.float64) * 0.123) % 1.0
        ph = (ph * (2 * np.pi)).to(torch.float32)
        view[lo:hi, 0] += 0.1 * torch.cos(ph)
        view[lo:hi, 1] += 0.1 * torch.sin(ph)
        del ph
    return iq


_CPU_DATA = None


def _cpu_make(arg):
    """Untimed: every worker draws its own slice of the workload (ncols time bins x nint frames)."""
    global _CPU_DATA
    ncols, seed = arg
    rng = np.random.default_rng(os.getpid() if seed is None else seed)
    d1 = np.empty((NINT * NFFT, ncols), np.complex64)
    for c in range(ncols):
        d1[:, c] = ((rng.standard_normal(NINT * NFFT, dtype=np.float32)
                     + 1j * rng.standard_normal(NINT * NFFT, dtype=np.float32)) * np.float32(7e-3))
    _CPU_DATA = d1
    return ncols


def _cpu_touch(_):
    return _CPU_DATA.shape


def _cpu_compute(_):
    """Timed: the reference path on the worker's slice -- Mode A oracle (= the welch call the
    reference's periodogram forwards to, without the truncation) + fftshift + median + dB."""
    from oracle import ref_port
    f, sxx, med = ref_port.sti_mode_a(_CPU_DATA, FS, NFFT)
    db = ref_port.to_dbfs(sxx)
    mdb = ref_port.to_dbfs(med)
    return float(db[0, 0]) + float(mdb[0])


class CpuReference:
    """``cores`` worker processes over disjoint time bins of the cfg2 workload; data generation is
    outside the timed region, each ``step()`` is one timed pass over the resident sample."""

    def __init__(self, cores, cols_per_worker):
        import multiprocessing as mp
        self.cores, self.cols = cores, cols_per_worker
        self.pool = None
        if cores > 1:
            # the initializer runs exactly once in every worker: each draws its own resident slice
            self.pool = mp.get_context("fork").Pool(cores, initializer=_cpu_make, initargs=((cols_per_worker, None),))
            self.pool.map(_cpu_touch, range(cores), chunksize=1)
        else:
            _cpu_make((cols_per_worker, 1234))

    def step(self):
        t0 = time.perf_counter()
        if self.pool is None:
            _cpu_compute(0)
        else:
            self.pool.map(_cpu_compute, range(self.cores), chunksize=1)
        return time.perf_counter() - t0

    @property
    def nsamp(self):
        return self.cores * self.cols * NINT * NFFT

    def close(self):
        if self.pool is not None:
            self.pool.close()
            self.pool.join()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    cols = 8  # time bins per worker and step: 8 x 366 x 4096 = 12 Msamples (96 MB) per worker
    ref = CpuReference(cores, cols)
    times = []
    for i in range(args.warmup + args.steps):
        dt = ref.step()
        if i >= args.warmup:
            times.append(dt)
    ref.close()
    dt = float(np.mean(times))
    msps = ref.nsamp / dt / 1e6
    cps = cores * cols / dt
    sample = (f"{cores * cols} of {NTIME} time bins x nint={NINT} x nfft={NFFT} per step "
              f"({ref.nsamp / 1e6:.0f} Msamples), {cores} processes over disjoint bins, data resident before timing")
    line = {
        "impl": "reference", "metric": METRIC, "value": msps, "unit": "Msamples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64 (scipy 1.18 internals; complex64 in, float32 out)",
        "data": "synthetic", "columns_per_s": cps,
        "config": {"workload": "cfg2: 1 channel 25 MS/s 60 s, nfft=4096, 1000 STI bins, nint=366 (Mode A)",
                   "path": "oracle.ref_port.sti_mode_a (scipy.signal.welch noverlap=0 + fftshift + median + dB)"},
        "cpu_baseline": {"value": msps, "unit": "Msamples/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": msps, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--seconds", type=float, default=SECONDS, help="recording length per channel (default cfg2: 60)")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-raw", action="store_true", help="skip the extra raw-int16 ingest e2e measurement")
    ap.add_argument("--variant", default=None, help="force a kernel variant (tuning)")
    ap.add_argument("--items-per-slot", type=int, default=0, help="column split target (tuning)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from pyspectrogram_b200 import engine
    from pyspectrogram_b200 import dist as pdist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if args.variant:
        engine.set_variant(args.variant)
    if args.items_per_slot:
        from pyspectrogram_b200 import _lib
        _lib.check(_lib.load().psg_set_items_per_slot(args.items_per_slot))

    nsamp = int(FS * args.seconds)
    nint = nsamp // NTIME // NFFT
    plan = engine.StiPlan(NFFT, device=local_rank)
    iq = synth_iq_device(torch, nsamp, 20240112 + rank, dev)
    starts_np = engine.frame_starts(0, nsamp, NFFT, nint, NTIME)  # drfProc.py:158-159
    starts = torch.from_numpy(starts_np.astype(np.int64)).to(dev)
    out_db = torch.empty((1, NTIME, NFFT), dtype=torch.float32, device=dev)
    out_lin = torch.empty((1, NTIME, NFFT), dtype=torch.float32, device=dev)
    gathered = None

    ev = lambda: torch.cuda.Event(enable_timing=True)

    def step(evs=None):
        nonlocal gathered
        if evs:
            evs[0].record()
        plan.run(iq, starts, nint, NFFT, want_lin=True, want_db=True, out_lin=out_lin, out_db=out_db)
        if evs:
            evs[1].record()
        # the time-median is per channel: with one channel per rank it needs no other rank's columns
        _, med_db = plan.median(out_lin, want_lin=False, want_db=True)
        if world > 1:
            # one gather of the [ncol_local][nfft] dB slabs assembles the N-channel image on rank 0
            # (plus the N median rows)
            gathered = pdist.gather_columns(out_db[0], [NTIME] * world, dst=0)
            pdist.gather_columns(med_db, [1] * world, dst=0)

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches0 = engine.launch_count()
    kev = [(ev(), ev()) for _ in range(args.steps)]
    t_beg, t_end = ev(), ev()
    with ClockSampler(local_rank) as clk:
        torch.cuda.synchronize()
        t_beg.record()
        for i in range(args.steps):
            step(kev[i])
        t_end.record()
        torch.cuda.synchronize()
        launches_timed = engine.launch_count() - launches0
        # the timed region of a few steps is shorter than one nvidia-smi query: keep the same load
        # running (untimed) until the sampler has seen it for about a second
        t_load = time.perf_counter()
        extra_steps = 0
        while len(clk.rows) < 6 and time.perf_counter() - t_load < 1.5:
            for _ in range(8):
                step()
            extra_steps += 8
            torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches = launches_timed
    total_ms = t_beg.elapsed_time(t_end)
    kern_ms = float(np.mean([a.elapsed_time(b) for a, b in kev]))
    if world > 1:
        tt = torch.tensor([total_ms, kern_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        total_ms, kern_ms = float(tt[0]), float(tt[1])
    ms_per_step = total_ms / args.steps
    samples_per_step = nint * NFFT * NTIME * world
    value = samples_per_step / (ms_per_step * 1e-3) / 1e6
    cols_per_s = NTIME * world / (ms_per_step * 1e-3)

    # roofline of the dominant kernel (fused STFT->PSD->STI + its finalize), per rank
    peak, peak_src = measured_peak()
    alg_bytes = 8 * NFFT * nint * NTIME + 2 * 4 * NFFT * NTIME  # IQ in + dB image + linear image out
    achieved = alg_bytes / (kern_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "peak_source": peak_src, "kernel": plan.variant, "kernel_ms": kern_ms,
                "algorithmic_bytes": alg_bytes}
    prof = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(prof):
        try:
            roofline["traffic"] = json.load(open(prof)).get(plan.variant)
        except Exception:
            pass

    line = {
        "metric": METRIC, "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "columns_per_s": cols_per_s,
        "config": {"workload": f"cfg2 per GPU: 1 channel 25 MS/s {args.seconds:g} s ({nsamp} complex64 samples), "
                               f"nfft={NFFT}, {NTIME} STI bins, nint={nint} (Mode A, every sample read once); "
                               "outputs: linear + dB image, dB time-median",
                   "parallelism": f"channel-per-GPU x{world}" if world > 1 else "single GPU",
                   "l2": f"inputs ({8 * nsamp / 1e9:.1f} GB per step) exceed the 126 MB L2; no flush needed"},
        "roofline": roofline, "gpu_launches": int(launches), "clocks": clk.summary(),
    }

    if not args.no_e2e:
        # every rank pushes its own channel through the host-buffer entry point at the same time
        e2e = e2e_measure(torch, plan, iq, starts_np, nint, args, dist if world > 1 else None)
        if world > 1:
            tt = torch.tensor([e2e.get("ms_per_step") or 1e30], device=dev, dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            if float(tt[0]) < 1e29:
                e2e["ms_per_step"] = float(tt[0])
                e2e["value"] = nint * NFFT * NTIME * world / (float(tt[0]) * 1e-3) / 1e6
                e2e["h2d_bytes_per_step"] *= world
                e2e["d2h_bytes_per_step"] *= world
            else:
                e2e = {"value": None, "unit": "Msamples/s", "error": "pinned host allocation failed on a rank"}
        line["e2e"] = e2e
        if world == 1 and not args.no_raw:
            line["e2e_raw_int16"] = e2e_raw_int16(torch, plan, iq, starts_np, nint, args)
    if rank == 0 and world == 1 and not args.no_cpu:
        ref = CpuReference(1, 32)
        ref.step()
        dts = [ref.step() for _ in range(2)]
        dt = float(np.mean(dts))
        line["cpu_baseline"] = {"value": ref.nsamp / dt / 1e6, "unit": "Msamples/s", "cores": 1, "kind": "port",
                                "columns_per_s": 32 / dt,
                                "sample": f"32 of {NTIME} time bins x nint={NINT} x nfft={NFFT} "
                                          f"({ref.nsamp / 1e6:.0f} Msamples, {dt:.1f} s per pass, data resident before "
                                          "timing), oracle.ref_port.sti_mode_a + median + dB, 1 process "
                                          "(scipy.fft workers=1, the reference's default)"}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def e2e_measure(torch, plan, iq_dev, starts_np, nint, args, dist=None):
    """Same metric through the host-buffer C-ABI call: pinned host IQ -> H2D -> kernels -> D2H."""
    nsamp = iq_dev.numel()
    host, err = None, ""
    try:
        host = torch.empty(nsamp, dtype=torch.complex64, pin_memory=True)
    except Exception as exc:  # not enough lockable host memory on this box
        err = str(exc)
    ok = torch.tensor([1 if host is not None else 0], device=iq_dev.device)
    if dist is not None:
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)  # every rank takes the same branch (barriers below)
    if int(ok[0]) == 0:
        return {"value": None, "unit": "Msamples/s", "error": f"pinned host allocation failed: {err}"}
    host.copy_(iq_dev)
    torch.cuda.synchronize()
    h = host.numpy()
    times = []
    for i in range(1 + max(1, args.e2e_steps)):
        if dist is not None:
            dist.barrier()
        t0 = time.perf_counter()
        res = plan.host(h, starts_np, nint, NFFT, want=("db", "med_db"))
        dt = time.perf_counter() - t0
        if i:
            times.append(dt)
    dt = float(np.mean(times))
    d2h = res["db"].nbytes + res["med_db"].nbytes
    return {"value": nint * NFFT * NTIME / dt / 1e6, "unit": "Msamples/s", "ms_per_step": dt * 1e3,
            "h2d_bytes_per_step": int(8 * (starts_np[-1] + nint * NFFT - starts_np[0]) + 8 * NTIME),
            "d2h_bytes_per_step": int(d2h), "steps": len(times),
            "path": "psg_sti_host (pinned host complex64 recording + int64 start table -> dB image + dB median), "
                    "one channel per rank, all ranks at once"}


def e2e_raw_int16(torch, plan, iq_dev, starts_np, nint, args):
    """Extra (not the contract's ``e2e``): the same recording stored as Digital RF stores it --
    complex int16 -- pushed through the typed host entry point with 1/ref folded into the kernel
    (SURVEY.md section 8(f) N1).  Half the PCIe bytes of the complex64 path for the same samples."""
    ref = 2.0 ** 15.5  # get_ref for int16 (drfProc.py:199-201)
    nsamp = iq_dev.numel()
    try:
        host = torch.empty((nsamp, 2), dtype=torch.int16, pin_memory=True)
    except Exception as exc:
        return {"value": None, "unit": "Msamples/s", "error": f"pinned host allocation failed: {exc}"}
    view = torch.view_as_real(iq_dev)
    chunk = 1 << 26
    for lo in range(0, nsamp, chunk):
        hi = min(nsamp, lo + chunk)
        host[lo:hi].copy_((view[lo:hi] * ref).round_().clamp_(-32767, 32767).to(torch.int16))
    torch.cuda.synchronize()
    h = host.numpy()
    times = []
    for i in range(1 + max(1, args.e2e_steps)):
        t0 = time.perf_counter()
        res = plan.host(h, starts_np, nint, NFFT, in_scale=1.0 / ref, want=("db", "med_db"))
        dt = time.perf_counter() - t0
        if i:
            times.append(dt)
    dt = float(np.mean(times))
    return {"value": nint * NFFT * NTIME / dt / 1e6, "unit": "Msamples/s", "ms_per_step": dt * 1e3,
            "h2d_bytes_per_step": int(4 * (starts_np[-1] + nint * NFFT - starts_np[0]) + 8 * NTIME),
            "d2h_bytes_per_step": int(res["db"].nbytes + res["med_db"].nbytes), "steps": len(times),
            "path": "psg_sti_host_typed(PSG_IQ_CI16): pinned host complex-int16 recording -> dB image + dB median"}


if __name__ == "__main__":
    main()
