"""CPU: the N>1 host logic (view sharding + gradient all-reduce) with two gloo processes."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge
    pkg = ge.load_package()
    from gsplat_b200 import multiview as mv  # registered by load_package
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        assert mv.views_for_rank(8, rank, world) == list(range(rank, 8, world))
        g = torch.Generator().manual_seed(1234 + rank)
        grads = [torch.randn(1000, 3, generator=g), torch.randn(1000, 4, generator=g), torch.randn(1000, generator=g),
                 torch.randn(1000, 16, 3, generator=g)]
        mine = [x.clone() for x in grads]
        mv.allreduce_gradients(grads, average_over=world)
        # recompute what the other rank drew
        expect = [torch.zeros_like(x) for x in mine]
        for r in range(world):
            gr = torch.Generator().manual_seed(1234 + r)
            for e, shape in zip(expect, [(1000, 3), (1000, 4), (1000,), (1000, 16, 3)]):
                e += torch.randn(*shape, generator=gr)
        ok = all(torch.allclose(a, e / world, atol=1e-6) for a, e in zip(grads, expect))
        tmax = mv.max_over_ranks(float(rank + 1), torch.device("cpu"))
        q.put((rank, ok, tmax))
    finally:
        dist.destroy_process_group()


def test_view_sharding_and_gradient_allreduce_gloo():
    world = 2
    port = 29500 + (os.getpid() % 2000)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res)
    assert all(abs(t - world) < 1e-9 for _, _, t in res)  # max over ranks of (rank + 1)


def test_views_for_rank_edge_cases(pkg):
    from gsplat_b200 import multiview as mv
    assert mv.views_for_rank(8, 0, 8) == [0] and mv.views_for_rank(8, 7, 8) == [7]
    assert mv.views_for_rank(3, 1, 2) == [1] and mv.views_for_rank(0, 0, 1) == []
    with pytest.raises(ValueError):
        mv.views_for_rank(8, 2, 2)
    # single-process: a no-op sum, only the averaging applies
    g = [torch.ones(4)]
    mv.allreduce_gradients(g, average_over=4)
    assert torch.allclose(g[0], torch.full((4,), 0.25))
