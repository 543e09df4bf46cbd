"""N > 1 path on CPU: two gloo ranks shard a batch, gather their detection slabs to rank 0, and rank 0 sees
exactly the single-process result in the original image order."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from yolo_fastest_b200 import _lib
from yolo_fastest_b200.dist import gather_compact, gather_detections, pack_host, shard_range, split_records


def test_shard_range_partitions():
    for n in (0, 1, 7, 8, 1024, 1025):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _fake_dets(n, max_det, seed):
    rng = np.random.default_rng(seed)
    d = np.zeros((n, max_det), dtype=_lib.DET_DTYPE)
    c = rng.integers(0, max_det + 1, size=n).astype(np.int32)
    for f in ("x1", "y1", "x2", "y2", "conf", "cls_score"):
        d[f] = rng.random((n, max_det))
    d["cls"] = rng.integers(0, 3, size=(n, max_det))
    d["src"] = rng.integers(0, 4800, size=(n, max_det))
    return d, c


def _worker(rank, world, port, n_total, max_det, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    d, c = _fake_dets(n_total, max_det, 11)
    lo, hi = shard_range(n_total, rank, world)
    dl = torch.from_numpy(d[lo:hi].copy().view(np.uint8).reshape(hi - lo, max_det, 56))
    cl = torch.from_numpy(c[lo:hi].copy())
    got = gather_detections(dl, cl, n_total, dst=0, rank_cap=(-(-n_total // world)) * max_det)
    if rank == 0:
        ok = len(got) == n_total and all(g.tobytes() == d[i, :c[i]].tobytes() for i, g in enumerate(got))
    else:
        ok = got is None
    # a message too small for a rank's detections must raise on rank 0, never truncate silently
    h = gather_compact(dl, cl, n_total, dst=0, rank_cap=1)
    if rank == 0:
        try:
            h.result()
            ok = False
        except _lib.YfError as e:
            ok = ok and "rank_cap" in str(e)
        q.put(ok)
    else:
        assert h.result() == (None, None) and ok
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gather_restores_order():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    n_total, max_det, world = 7, 5, 2          # ragged: 4 + 3 images
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, max_det, q)) for r in range(world)]
    for p in procs:
        p.start()
    ok = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert ok is True


def test_message_layout():
    """pack_host restates yf_compact_dets: int32 header [total, B, n_b ...] then the records back to back."""
    d, c = _fake_dets(5, 4, 3)
    c[1] = 9                                   # more than max_det: clipped to the slab capacity
    msg = pack_host(torch.from_numpy(d.view(np.uint8).reshape(5, 4, 56)), torch.from_numpy(c), 8, 32).numpy()
    hdr = msg[:32].view(np.int32)
    n = np.minimum(c, 4)
    assert hdr[0] == n.sum() and hdr[1] == 5 and list(hdr[2:7]) == list(n)
    recs = msg[32:32 + 56 * int(n.sum())].view(_lib.DET_DTYPE)
    parts = split_records(recs, n)
    assert all(p.tobytes() == d[i, :n[i]].tobytes() for i, p in enumerate(parts))
