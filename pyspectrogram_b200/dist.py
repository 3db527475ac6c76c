"""Sharding of STI columns over the GPUs of one box.

Every STI column depends only on its own samples (drfProc.py:161-166), so the flat list of
columns ``(channel, sub-channel, time bin)`` is block-partitioned across ranks -- whole channels
when there are at least as many channels as ranks, contiguous time-bin ranges otherwise -- and
each rank computes its slab with no data-path collective.  The only exchange is one gather of
``[ncol_local][nfft]`` float32 slabs when the image is returned to the host; the time-median
(drfProc.py:401) needs every column of a row and therefore runs after the gather.

One process per GPU; ``torch.distributed`` (NCCL over NVLink on the GPU box, gloo in CPU tests)
is plumbing only.
"""
from __future__ import annotations

from typing import List, Tuple


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block ``[lo, hi)`` of ``n`` items owned by ``rank``: the first ``n % world`` ranks
    own one extra item, so sizes differ by at most one and the concatenation over ranks is 0..n."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_plan(nchan: int, ntime: int, world: int) -> List[List[Tuple[int, int, int]]]:
    """Per rank, a list of ``(channel, t_lo, t_hi)`` pieces.

    ``nchan >= world``: whole channels per rank (BASELINE config 3).  Otherwise the flat column list
    ``channel-major x time`` is block-partitioned, which for one channel is a contiguous time-bin
    range per rank (BASELINE config 4)."""
    plan: List[List[Tuple[int, int, int]]] = [[] for _ in range(world)]
    if nchan >= world:
        for r in range(world):
            lo, hi = shard_range(nchan, r, world)
            plan[r] = [(c, 0, ntime) for c in range(lo, hi)]
        return plan
    total = nchan * ntime
    for r in range(world):
        lo, hi = shard_range(total, r, world)
        while lo < hi:
            c, t = divmod(lo, ntime)
            t_hi = min(ntime, t + (hi - lo))
            plan[r].append((c, t, t_hi))
            lo += t_hi - t
    return plan


def gather_columns(local, ncols_per_rank, dst=0, group=None):
    """Assemble the image on ``dst``: ``local`` is this rank's ``[ncol_local, nfft]`` slab (any
    backend's tensor); returns the ``[sum(ncols), nfft]`` image on ``dst`` and ``None`` elsewhere.

    Equal slabs use a single ``gather`` collective; ragged slabs are padded to the largest slab for
    the collective and trimmed afterwards (at most one padding row per rank with ``shard_range``)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    assert len(ncols_per_rank) == world and local.shape[0] == ncols_per_rank[rank]
    if world == 1:
        return local
    width = max(ncols_per_rank)
    send = local
    if local.shape[0] != width:
        send = torch.zeros((width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        send[: local.shape[0]] = local
    send = send.contiguous()
    if rank == dst:
        recv = torch.empty((world, width) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.gather(send, list(recv.unbind(0)), dst=dst, group=group)
        if all(n == width for n in ncols_per_rank):
            return recv.reshape((world * width,) + tuple(local.shape[1:]))
        return torch.cat([recv[r, : ncols_per_rank[r]] for r in range(world)], dim=0)
    dist.gather(send, None, dst=dst, group=group)
    return None
