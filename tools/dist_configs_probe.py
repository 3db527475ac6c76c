#!/usr/bin/env python
"""BASELINE configs 3 and 4 on N GPUs of one box, over real NCCL (measurement aid; GPU box only).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/dist_configs_probe.py [--reps 5] [--out gpurun_out/dist_configs.json]

cfg3  nfft = 16384, 3600 time bins, nint = 64: one channel per rank (``shard_plan`` with nchan >= world);
      every rank computes its channel's image and time-median, one gather assembles the dB images on rank 0.
cfg4  nfft = 65536, 3600 time bins, nint = 16, ONE channel: contiguous time-bin ranges per rank
      (``shard_plan`` with nchan < world); every rank holds only the samples its bins touch, one gather of
      the linear slabs assembles the image on rank 0; the time-median (it needs every column of a row,
      drfProc.py:401) runs on frequency slabs after one pairwise exchange (dist.median_over_time_sharded),
      timed against the whole median on rank 0 and checked bit-identical to it.
Per config: per-rank kernel time (CUDA events, max over ranks), gather and whole-step time (barrier +
synchronize on both sides, max over ranks), Gsamples/s and columns/s for the whole job, and a check that
the slab rank 0 received from every rank is bit-identical to what that rank computed (fp64 checksums).
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--ntime", type=int, default=3600)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "dist_configs.json"))
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from bench import synth_iq_device
    from pyspectrogram_b200 import dist as pdist
    from pyspectrogram_b200 import engine

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ev = lambda: torch.cuda.Event(enable_timing=True)
    rows = []

    def run_config(name, nfft, nint, nchan):
        ntime = args.ntime
        plan_pieces = pdist.shard_plan(nchan, ntime, world)[rank]
        # this rank's columns: (channel, t_lo, t_hi) pieces; the samples it holds are exactly those of its bins
        ncol = sum(hi - lo for _, lo, hi in plan_pieces)
        n = ncol * nint * nfft
        iq = synth_iq_device(torch, n + 8, 20240112 + 1000 * rank, dev)
        starts = torch.from_numpy(engine.frame_starts(0, n, nfft, nint, ncol).astype(np.int64)).to(dev)
        plan = engine.StiPlan(nfft, device=local_rank)
        lin = torch.empty((1, ncol, nfft), dtype=torch.float32, device=dev)
        db = torch.empty((1, ncol, nfft), dtype=torch.float32, device=dev)
        by_channel = nchan >= world
        ncols_all = [sum(hi - lo for _, lo, hi in p) for p in pdist.shard_plan(nchan, ntime, world)]
        state = {}

        def step(kev=None, gev=None):
            if kev:
                kev[0].record()
            plan.run(iq, starts, nint, nfft, want_lin=True, want_db=True, out_lin=lin, out_db=db)
            if kev:
                kev[1].record()
            if by_channel:
                # the median is per channel: no other rank's columns are needed
                _, med_db = plan.median(lin, want_lin=False, want_db=True)
                if gev:
                    gev[0].record()
                state["img"] = pdist.gather_columns(db[0], ncols_all, dst=0) if world > 1 else db[0]
                if world > 1:
                    pdist.gather_columns(med_db, [1] * world, dst=0)
                if gev:
                    gev[1].record()
            else:
                if gev:
                    gev[0].record()
                img = pdist.gather_columns(lin[0], ncols_all, dst=0) if world > 1 else lin[0]
                if gev:
                    gev[1].record()
                state["img"] = img
                if state.get("root_median"):
                    if img is not None:  # rank 0 alone: median over ALL time bins, then dB of the median
                        state["med"] = plan.median(img.reshape(1, -1, nfft), want_lin=True, want_db=True)
                else:  # re-shard by frequency, median of the slabs on every rank, gather the rows
                    state["med"] = pdist.median_over_time_sharded(
                        lin[0], ncols_all, lambda x: plan.median(x, want_lin=True, want_db=True), dst=0)

        med_root = None
        if not by_channel:
            # reference point: gather, then the whole median on rank 0
            state["root_median"] = True
            for _ in range(2):
                step()
            torch.cuda.synchronize()
            troot = []
            for _ in range(args.reps):
                dist.barrier()
                torch.cuda.synchronize()
                t0, t1 = ev(), ev()
                t0.record()
                step()
                t1.record()
                torch.cuda.synchronize()
                v = torch.tensor([t0.elapsed_time(t1)], device=dev, dtype=torch.float64)
                dist.all_reduce(v, op=dist.ReduceOp.MAX)
                troot.append(float(v[0]))
            if rank == 0:
                med_root = [m.clone() for m in state["med"]]
            state["root_median"] = False
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        tot, kern, gat = [], [], []
        for _ in range(args.reps):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            t0, t1, k, g = ev(), ev(), (ev(), ev()), (ev(), ev())
            t0.record()
            step(k, g)
            t1.record()
            torch.cuda.synchronize()
            v = torch.tensor([t0.elapsed_time(t1), k[0].elapsed_time(k[1]), g[0].elapsed_time(g[1])], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(v, op=dist.ReduceOp.MAX)
            tot.append(float(v[0])); kern.append(float(v[1])); gat.append(float(v[2]))
        # the slab rank 0 holds for every rank is what that rank computed
        mine = (db if by_channel else lin)[0].double().sum().reshape(1)
        sums = [torch.zeros_like(mine) for _ in range(world)]
        if world > 1:
            dist.all_gather(sums, mine)
        else:
            sums = [mine]
        ok = True
        if rank == 0:
            img, off = state["img"], 0
            for r in range(world):
                got = img[off:off + ncols_all[r]].double().sum()
                ok = ok and bool(got == sums[r][0])
                off += ncols_all[r]
        ms, kms, gms = float(np.median(tot)), float(np.median(kern)), float(np.median(gat))
        samples = nint * nfft * ntime * nchan
        row = {"config": name, "n_gpus": world, "nfft": nfft, "ntime": ntime, "nint": nint, "channels": nchan,
               "sharding": "channel per rank" if by_channel else "contiguous time bins per rank",
               "columns_per_rank": ncols_all, "variant": plan.variant, "ms_step": ms, "ms_kernel_max_rank": kms,
               "ms_gather": gms, "gsamples_s": samples / ms / 1e6, "columns_s": ntime * nchan / (ms * 1e-3),
               "gathered_bytes": int(sum(ncols_all[1:]) * nfft * 4), "gather_matches_ranks": ok}
        if not by_channel:
            row["ms_step_median_on_root"] = float(np.median(troot))
            if rank == 0:
                row["sharded_median_bit_identical"] = bool(
                    torch.equal(state["med"][0], med_root[0][0]) and torch.equal(state["med"][1], med_root[1][0]))
        if rank == 0:
            rows.append(row)
            print(json.dumps(row), flush=True)
        del iq, lin, db, plan
        state.clear()
        torch.cuda.empty_cache()

    run_config("cfg3: 10 MS/s channels, nfft=16384, nint=64 (10.5 % duty), one channel per GPU", 16384, 64, world)
    run_config("cfg4: 1 channel 100 MS/s, nfft=65536, nint=16 (1.05 % duty), time bins sharded", 65536, 16, 1)
    if rank == 0:
        os.makedirs(os.path.dirname(args.out), exist_ok=True)
        json.dump(rows, open(args.out, "w"), indent=1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
