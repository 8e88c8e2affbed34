"""Host-side multi-process logic on CPU: world_size-2 gloo rendezvous on 127.0.0.1."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from plenoctree_b200.nerf.train import allreduce_gradients, shard_batch
        from plenoctree_b200.ops import grid_slab
        # [grads | stats] bucket: rank r contributes (r+1) everywhere
        gbuf = torch.full((1000 + 8,), float(rank + 1))
        w = allreduce_gradients(gbuf)
        ok = (w == world) and torch.allclose(gbuf, torch.full_like(gbuf, 3.0))
        mean = gbuf / w            # what Adam (grad_mult = 1/world) and the stats see
        ok = ok and torch.allclose(mean, torch.full_like(mean, 1.5))
        lo, hi = shard_batch(4096, rank, world)
        ok = ok and (hi - lo == 2048) and lo == rank * 2048
        x0, nx = grid_slab(512, rank, world)
        ok = ok and (x0, nx) == (rank * 256, 256)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_two_rank_gradient_bucket_and_sharding():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def test_single_process_is_identity():
    from plenoctree_b200.nerf.train import allreduce_gradients, shard_batch
    from plenoctree_b200.ops import grid_slab
    g = torch.arange(10.0)
    assert allreduce_gradients(g) == 1 and torch.equal(g, torch.arange(10.0))
    with pytest.raises(ValueError):
        shard_batch(1000, 0, 3)
    assert [grid_slab(10, r, 3) for r in range(3)] == [(0, 4), (4, 3), (7, 3)]
