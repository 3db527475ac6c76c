"""Sharding of STI columns over the GPUs of one box.

Every STI column depends only on its own samples (drfProc.py:161-166), so the flat list of
columns ``(channel, sub-channel, time bin)`` is block-partitioned across ranks -- whole channels
when there are at least as many channels as ranks, contiguous time-bin ranges otherwise -- and
each rank computes its slab with no data-path collective.  The image is assembled by one gather of
``[ncol_local][nfft]`` float32 slabs when it is returned to the host.  The time-median (drfProc.py:401)
needs every column of a row: per channel it stays on the channel's rank; with time-bin shards it is the
one real exchange of the path -- the image is re-sharded by frequency and every rank takes the median of
its slab (``median_over_time_sharded``).

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


class PeerImage:
    """The assembled ``[sum(ncols)][width]`` image on rank ``dst``, written in place by every rank.

    ``gather_columns`` moves finished slabs with NCCL kernels: a second pass over the image, and kernels that need SMs
    the persistent STI kernels occupy (a gather issued under a radix-32 kernel waits for it, or pushes its CTAs into a
    second wave: measured 10.2 against 8.5 ms per cfg3 step at two GPUs).  Over NVLink / NVSwitch every GPU can address
    its peers' memory, so the image is allocated once as symmetric memory and every rank gets ``rows`` -- a CUDA tensor
    that aliases ITS rows of rank ``dst``'s buffer.  Passed as ``out_db`` / ``out_lin`` (or the median's output) the
    kernels' own coalesced 128-bit stores land in the assembled image: the gather is fused into the epilogue, nothing
    else runs, nothing is copied twice.  ``publish()`` orders the writes before the root's reads (a barrier on the
    symmetric-memory signal pads, on the current stream).

    (Measured alternative for large slabs, removed: a local slab pushed into the root's image with one
    ``cudaMemcpyAsync`` per rank -- the copy engines take no SM from the kernels -- costs 0.33 ms per 236 MB slab at two
    GPUs where the fused stores cost 0.01 ms, and issued on a side stream under the next step's persistent radix-32
    kernel its barrier kernel waits for that kernel to retire: 9.85 against 8.20 ms per cfg3 step,
    ``profiles/r02_assembly_copy_engine_experiment.txt``.)

    Falls back to a local slab + ``gather_columns`` when symmetric memory is unavailable (gloo, one rank, no P2P):
    ``rows`` is then the local slab and ``publish()`` performs the gather.  ``image`` is the assembled tensor on
    ``dst`` (``None`` elsewhere; in the fallback it is valid after ``publish()``)."""

    def __init__(self, ncols_per_rank, width, dtype=None, device=None, dst=0, group=None, allow_peer=True):
        import torch
        import torch.distributed as dist

        self.group = group
        self.dst = dst
        self.ncols = list(ncols_per_rank)
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        assert len(self.ncols) == self.world
        dtype = dtype or torch.float32
        total = sum(self.ncols)
        off = sum(self.ncols[: self.rank])
        self.mode = "local"
        self._hdl = None
        self.image = None
        if self.world > 1 and allow_peer and device is not None and torch.device(device).type == "cuda":
            try:
                import torch.distributed._symmetric_memory as symm_mem
                buf = symm_mem.empty(total, width, dtype=dtype, device=device)
                self._hdl = symm_mem.rendezvous(buf, group if group is not None else dist.group.WORLD)
                root = self._hdl.get_buffer(dst, (total, width), dtype)
                self.rows = root[off: off + self.ncols[self.rank]]
                self.image = buf if self.rank == dst else None
                self._keep = buf
                self.mode = "peer"
            except Exception as exc:  # no symmetric memory on this build / fabric: NCCL gather
                self._why = str(exc)
            # every rank takes the same branch (publish() is collective either way)
            ok = torch.tensor([1 if self.mode == "peer" else 0], device=device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
            if int(ok[0]) == 0:
                self.mode = "local"
        if self.mode != "peer":
            self.rows = torch.empty((self.ncols[self.rank], width), dtype=dtype, device=device)
            self.mode = "gather" if self.world > 1 else "local"
            if self.world == 1:
                self.image = self.rows

    def publish(self):
        """Make every rank's rows visible in ``image`` on ``dst`` (stream-ordered on the current stream)."""
        if self.mode == "peer":
            self._hdl.barrier(channel=0)
        elif self.mode == "gather":
            self.image = gather_columns(self.rows, self.ncols, dst=self.dst, group=self.group)
        return self.image


class FreqReshard:
    """Re-sharding of a time-sharded image by frequency over peer memory (the exchange of ``median_over_time_sharded``).

    Every rank owns a symmetric ``[ntime][nfft / world]`` slab; ``exchange(local)`` writes this rank's columns' bins
    ``shard_range(nfft, s, world)`` straight into rows ``[off_r, off_r + ncol_r)`` of rank ``s``'s slab -- strided
    copy kernels whose stores cross NVLink, no staging copy, no NCCL send / receive pair per peer -- between two
    barriers on the signal pads ("the slabs of the previous call have been read" / "every block has landed").
    Needs ``nfft % world == 0`` and symmetric memory; ``available`` is False otherwise (the NCCL exchange is used)."""

    def __init__(self, ncols_per_rank, nfft, dtype=None, device=None, group=None):
        import torch
        import torch.distributed as dist

        self.ncols = list(ncols_per_rank)
        self.nfft = int(nfft)
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.available = False
        dtype = dtype or torch.float32
        if self.world < 2 or device is None or torch.device(device).type != "cuda":
            return
        ok_local = self.nfft % self.world == 0
        self.width = self.nfft // self.world
        ntime = sum(self.ncols)
        if ok_local:
            try:
                import torch.distributed._symmetric_memory as symm_mem
                buf = symm_mem.empty(ntime, self.width, dtype=dtype, device=device)
                self._hdl = symm_mem.rendezvous(buf, group if group is not None else dist.group.WORLD)
                self.slab = buf
                self._peers = [self._hdl.get_buffer(s, (ntime, self.width), dtype) for s in range(self.world)]
            except Exception as exc:
                self._why = str(exc)
                ok_local = False
        ok = torch.tensor([1 if ok_local else 0], device=device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        self.available = bool(int(ok[0]))

    def exchange(self, local):
        """``local``: this rank's ``[ncol_r][nfft]`` slab.  Returns this rank's ``[ntime][nfft / world]`` slab, complete
        when the current stream has passed this call."""
        off = sum(self.ncols[: self.rank])
        n = self.ncols[self.rank]
        self._hdl.barrier(channel=0)  # every rank is done reading the slabs of the previous exchange
        if n:
            for k in range(self.world):  # staggered destinations: no two ranks start on the same peer
                s = (self.rank + k) % self.world
                self._peers[s][off: off + n].copy_(local[:, s * self.width: (s + 1) * self.width])
        self._hdl.barrier(channel=1)  # every block has landed
        return self.slab


def median_over_time_sharded(local, ncols_per_rank, median_fn, dst=0, group=None, reshard=None):
    """Time-median of an image whose COLUMNS (time bins) are sharded over the ranks (BASELINE config 4).

    ``np.median(sxx, axis=1)`` (drfProc.py:401) needs every time bin of a frequency row.  Gathering the
    image on one rank and taking the median there leaves that rank with the whole ``[ntime][nfft]``
    selection while the others idle -- at nfft = 65536, ntime = 3600 that is as long as the transform
    of a two-GPU shard.  Here the image is re-sharded by FREQUENCY instead: rank ``s`` receives bins
    ``shard_range(nfft, s, world)`` of every rank's columns (the one exchange of the path: pairwise
    sends of ``[ncol_r][nfft / world]`` blocks over NVLink), stacks them in rank order -- which is time
    order -- runs ``median_fn`` on its ``[ntime][nfft / world]`` slab, and the median slabs are
    gathered on ``dst``.  A median is an order statistic of one frequency row, so the result is
    bit-identical to the median of the assembled image.

    ``reshard``: a ``FreqReshard`` built once for this shape does the exchange over peer memory (no staging copies, no
    NCCL kernels); without it, or when it is not ``available``, the blocks travel by NCCL ``batch_isend_irecv``.

    ``local``: this rank's ``[ncol_local][nfft]`` linear slab.  ``median_fn(img)`` maps a contiguous
    ``[1][ntime][w]`` tensor to a tuple of ``[1][w]`` tensors (e.g. ``StiPlan.median`` returning the
    linear and / or dB median; ``None`` entries are passed through).  Returns the tuple of ``[nfft]``
    tensors on ``dst``, ``None`` elsewhere."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    nfft = int(local.shape[1])
    assert len(ncols_per_rank) == world and local.shape[0] == ncols_per_rank[rank]
    if world == 1:
        return tuple(None if m is None else m[0] for m in median_fn(local.reshape(1, -1, nfft)))
    f_lo, f_hi = shard_range(nfft, rank, world)
    width = f_hi - f_lo
    ntime = sum(ncols_per_rank)
    if reshard is not None and reshard.available:
        slab = reshard.exchange(local)
        return _finish_sharded_median(slab, ntime, width, nfft, world, rank, median_fn, dst, group)
    slab = torch.empty((ntime, width), dtype=local.dtype, device=local.device)
    ops, keep = [], []
    off = 0
    for r in range(world):
        rows = slab[off:off + ncols_per_rank[r]]
        off += ncols_per_rank[r]
        if r == rank:
            rows.copy_(local[:, f_lo:f_hi])
            continue
        lo, hi = shard_range(nfft, r, world)
        if ncols_per_rank[rank] and hi > lo:
            send = local[:, lo:hi].contiguous()
            keep.append(send)
            ops.append(dist.P2POp(dist.isend, send, r, group))
        if ncols_per_rank[r] and width:
            ops.append(dist.P2POp(dist.irecv, rows, r, group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return _finish_sharded_median(slab, ntime, width, nfft, world, rank, median_fn, dst, group)


def _finish_sharded_median(slab, ntime, width, nfft, world, rank, median_fn, dst, group):
    meds = median_fn(slab.reshape(1, ntime, width))
    widths = [shard_range(nfft, r, world)[1] - shard_range(nfft, r, world)[0] for r in range(world)]
    out = []
    for m in meds:
        if m is None:
            out.append(None)
            continue
        full = gather_columns(m.reshape(width, 1), widths, dst=dst, group=group)
        out.append(full.reshape(nfft) if rank == dst else None)
    return tuple(out) if rank == dst else None
