"""Multi-GPU plumbing: one process per GPU, images sharded by rank, no collective in the compute path.

The reference is single-process; images are independent end to end, so a batch is
split into contiguous per-rank slices and each rank runs the whole hot path on its slice.  The only
exchange is the return of the fixed-capacity detection slabs ([B_local, max_det] yf_det + counts) to
rank 0 — one ``torch.distributed.gather`` (NCCL over NVLink for CUDA tensors, gloo for CPU tensors in
the CPU tests).  Order is restored by rank-major concatenation, which equals the original image order.
"""
import numpy as np
import torch
import torch.distributed as dist

from . import _lib


def shard_range(n_items, rank, world_size):
    """Contiguous slice [lo, hi) of rank `rank`; the first n_items % world_size ranks get one extra item."""
    base, extra = divmod(n_items, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


class _Buffers:
    cache = {}


def gather_detections(dets, counts, n_total, dst=0, group=None, to_host=True):
    """Gather per-rank detection slabs to `dst` with ONE collective.

    dets:   uint8 tensor [B_local, max_det, 56] (yf_det records), counts: int32 tensor [B_local];
            both on the same device (cuda -> NCCL, cpu -> gloo).  Ranks may hold different B_local
            (ragged shards from shard_range); slabs are padded to the largest shard and packed with their
            counts into one [max_local, max_det*56 + 4] byte tensor for the collective.
    Returns on `dst` (dets [n_total, max_det] structured numpy array, counts [n_total]) — or, with
    to_host=False, the packed device tensors per rank (no synchronisation) — and (None, None) elsewhere.
    """
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    max_local = -(-n_total // world)
    max_det = dets.shape[1]
    row = max_det * _lib.DET_DTYPE.itemsize
    key = (str(dets.device), max_local, max_det, world, rank == dst)
    buf = _Buffers.cache.get(key)
    if buf is None:
        packed = torch.zeros((max_local, row + 4), dtype=torch.uint8, device=dets.device)
        recv = [torch.empty_like(packed) for _ in range(world)] if rank == dst else None
        buf = _Buffers.cache[key] = (packed, recv)
    packed, recv = buf
    nl = dets.shape[0]
    packed[:nl, :row] = dets.reshape(nl, row)
    packed[:nl, row:] = counts.view(torch.uint8).reshape(nl, 4)
    dist.gather(packed, recv, dst=dst, group=group)
    if rank != dst:
        return None, None
    if not to_host:
        return recv, None
    out_d, out_c = [], []
    for r in range(world):
        lo, hi = shard_range(n_total, r, world)
        h = recv[r][:hi - lo].cpu().numpy()
        out_d.append(np.ascontiguousarray(h[:, :row]).view(_lib.DET_DTYPE).reshape(hi - lo, max_det))
        out_c.append(np.ascontiguousarray(h[:, row:]).view(np.int32).reshape(hi - lo))
    return np.concatenate(out_d, 0), np.concatenate(out_c, 0)
