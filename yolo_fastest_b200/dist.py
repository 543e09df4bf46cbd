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


def gather_detections(dets, counts, n_total, dst=0, group=None):
    """Gather per-rank detection slabs to `dst`.

    dets:   uint8 tensor [B_local, max_det, 56] (yf_det records), counts: int32 tensor [B_local];
            both on the same device (cuda -> NCCL, cpu -> gloo).  Ranks may hold different B_local
            (ragged shards from shard_range); slabs are padded to the largest shard for the collective.
    Returns (dets [n_total, max_det] structured numpy array, counts [n_total]) on `dst`, (None, None) elsewhere.
    """
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    max_local = -(-n_total // world)
    max_det = dets.shape[1]
    pad_d = torch.zeros((max_local, max_det, _lib.DET_DTYPE.itemsize), dtype=torch.uint8, device=dets.device)
    pad_c = torch.zeros((max_local,), dtype=torch.int32, device=dets.device)
    pad_d[:dets.shape[0]] = dets
    pad_c[:counts.shape[0]] = counts
    if rank == dst:
        all_d = [torch.empty_like(pad_d) for _ in range(world)]
        all_c = [torch.empty_like(pad_c) for _ in range(world)]
    else:
        all_d = all_c = None
    dist.gather(pad_d, all_d, dst=dst, group=group)
    dist.gather(pad_c, all_c, dst=dst, group=group)
    if rank != dst:
        return None, None
    out_d, out_c = [], []
    for r in range(world):
        lo, hi = shard_range(n_total, r, world)
        out_d.append(all_d[r][:hi - lo].cpu().numpy().view(_lib.DET_DTYPE).reshape(hi - lo, max_det))
        out_c.append(all_c[r][:hi - lo].cpu().numpy())
    return np.concatenate(out_d, 0), np.concatenate(out_c, 0)
