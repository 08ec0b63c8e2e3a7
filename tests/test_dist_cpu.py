"""world_size-2 gloo test of the sharding + gather logic used by the multi-GPU path (runs on CPU)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_items, q):
    import sys
    sys.path.insert(0, ROOT)
    from e4s_b200.dist import gather_images, shard_range
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(n_items, rank, world)
    local = torch.stack([torch.full((3, 4, 4), float(i)) for i in range(lo, hi)]) if hi > lo else torch.zeros(0, 3, 4, 4)
    full = gather_images(local, n_items)
    ok = full.shape[0] == n_items and all(float(full[i, 0, 0, 0]) == i for i in range(n_items))
    auto = gather_images(local)            # sizes discovered by an extra all_gather
    ok = ok and torch.equal(auto, full)
    q.put((rank, ok))
    dist.destroy_process_group()


def test_shard_range_partitions():
    from e4s_b200.dist import shard_range
    for n in (0, 1, 7, 16, 64, 65):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_gather_two_ranks_gloo():
    ctx = mp.get_context("spawn")
    for n_items in (6, 5):          # even and ragged
        q = ctx.Queue()
        port = _free_port()
        procs = [ctx.Process(target=_worker, args=(r, 2, port, n_items, q)) for r in range(2)]
        for p in procs:
            p.start()
        results = [q.get(timeout=120) for _ in range(2)]
        for p in procs:
            p.join(timeout=60)
        assert all(ok for _, ok in results), results
