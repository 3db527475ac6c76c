"""Host-side sharding logic on CPU: shard plans, and the gather that
assembles the image, run on two gloo ranks."""
import os
import socket

import numpy as np
import pytest

from pyspectrogram_b200 import dist as pdist


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 100, 3600):
        for world in (1, 2, 3, 4, 8):
            pieces = [pdist.shard_range(n, r, world) for r in range(world)]
            assert pieces[0][0] == 0 and pieces[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(pieces, pieces[1:]))
            sizes = [hi - lo for lo, hi in pieces]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        pdist.shard_range(10, 2, 2)


def test_shard_plan_by_channel_and_by_time_bin():
    # cfg3: 8 channels over 8 ranks -> one whole channel each
    plan = pdist.shard_plan(8, 3600, 8)
    assert plan == [[(c, 0, 3600)] for c in range(8)]
    # cfg4: 1 channel, 3600 bins over 2/4/8 ranks -> contiguous time-bin ranges
    for world, per in ((2, 1800), (4, 900), (8, 450)):
        plan = pdist.shard_plan(1, 3600, world)
        assert plan == [[(0, r * per, (r + 1) * per)] for r in range(world)]
    # ragged: 3 channels x 10 bins over 4 ranks covers every column exactly once, in order
    plan = pdist.shard_plan(3, 10, 4)
    flat = [(c, t) for pieces in plan for (c, lo, hi) in pieces for t in range(lo, hi)]
    assert flat == [(c, t) for c in range(3) for t in range(10)]
    # more channels than ranks: whole channels, sizes differ by at most one
    plan = pdist.shard_plan(5, 7, 2)
    assert [len(p) for p in plan] == [3, 2]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ncols, nfft, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo = sum(ncols[:rank])
        full = torch.arange(sum(ncols) * nfft, dtype=torch.float32).reshape(sum(ncols), nfft)
        local = full[lo:lo + ncols[rank]].clone()
        img = pdist.gather_columns(local, ncols, dst=0)
        if rank == 0:
            q.put(bool(torch.equal(img, full)))
        else:
            q.put(img is None)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("ncols", [[5, 5], [4, 3], [1, 0]])
def test_gather_columns_two_gloo_ranks(ncols):
    """Equal and ragged slabs: rank 0 ends up with the columns of all ranks in rank order."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ncols, 16, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(results)


def test_sharded_starts_concatenate_to_the_single_gpu_table():
    """Time-bin sharding hands each rank a slice of numpy's own linspace table (row a1), so the
    union of the shards is bit-identical to the single-GPU table."""
    from pyspectrogram_b200 import engine
    st, en, nfft, nint, ntime = 170000000000000000, 170000000000000000 + 10**9, 65536, 16, 3600
    full = engine.frame_starts(st, en, nfft, nint, ntime)
    for world in (2, 4, 8):
        parts = [full[slice(*pdist.shard_range(ntime, r, world))] for r in range(world)]
        assert np.array_equal(np.concatenate(parts), full)


def _median_worker(rank, world, port, ncols, nfft, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(5)
        full = rng.random((sum(ncols), nfft)).astype(np.float32)
        lo = sum(ncols[:rank])
        local = torch.from_numpy(full[lo:lo + ncols[rank]].copy())

        def median_fn(img):  # [1][ntime][w] -> ([1][w] linear, None)
            return torch.from_numpy(np.median(img.numpy(), axis=1)), None

        # a FreqReshard without CUDA peers reports itself unavailable and the NCCL / gloo exchange runs
        rs = pdist.FreqReshard(ncols, nfft, device="cpu")
        assert not rs.available
        res = pdist.median_over_time_sharded(local, ncols, median_fn, dst=0, reshard=rs)
        if rank == 0:
            q.put(bool(np.array_equal(res[0].numpy(), np.median(full, axis=0)) and res[1] is None))
        else:
            q.put(res is None)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("ncols,nfft", [([5, 5], 16), ([4, 3], 10), ([6, 1], 7)])
def test_median_over_time_sharded_two_gloo_ranks(ncols, nfft):
    """Time-bin shards re-sharded by frequency (the one exchange of BASELINE config 4): the median of
    the slabs, gathered, is bit-identical to np.median of the assembled image -- equal and ragged
    shards, nfft not divisible by the number of ranks."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_median_worker, args=(r, 2, port, ncols, nfft, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(results)


def _peer_image_worker(rank, world, port, ncols, nfft, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = torch.arange(sum(ncols) * nfft, dtype=torch.float32).reshape(sum(ncols), nfft)
        lo = sum(ncols[:rank])
        img = pdist.PeerImage(ncols, nfft, device="cpu")  # no CUDA peers here: local slab + gather on publish()
        ok = img.mode == "gather" and tuple(img.rows.shape) == (ncols[rank], nfft)
        img.rows.copy_(full[lo:lo + ncols[rank]])  # stands for the kernel writing its columns
        out = img.publish()
        q.put(bool(ok and (torch.equal(out, full) if rank == 0 else out is None)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("ncols", [[5, 5], [4, 3]])
def test_peer_image_falls_back_to_the_gather(ncols):
    """dist.PeerImage without peer memory (gloo, CPU tensors): every rank writes its rows into a local slab and
    publish() assembles the image on rank 0 with the gather -- same result as the peer-memory form, which the
    multi-GPU bench exercises (rows aliasing rank 0's buffer over NVLink)."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_peer_image_worker, args=(r, 2, port, ncols, 8, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(results)


def test_peer_image_single_process():
    import torch
    img = pdist.PeerImage([7], 12, device="cpu")
    assert img.mode == "local" and img.publish() is img.rows and tuple(img.rows.shape) == (7, 12)
    assert img.rows.dtype == torch.float32
