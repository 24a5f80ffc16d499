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


def _oracle_sh_views(degree, means, campos, coeffs, v_colors, v_means):
    """CPU stand-in for gsb_sh_bwd_views built from the oracle's single-view SH backward (test infrastructure)."""
    from oracle import oracle as orc
    orc.build()
    out = None
    for v in range(campos.shape[0]):
        dirs = (means - campos[v]).numpy()
        vco, vd = orc.sh_bwd(degree, dirs, coeffs.numpy(), None, v_colors[v].numpy(), True)
        out = torch.from_numpy(vco) if out is None else out + torch.from_numpy(vco)
        v_means += torch.from_numpy(vd)
    return out


def _compact_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge
    pkg = ge.load_package()
    from gsplat_b200 import multiview as mv
    from oracle import oracle as orc
    orc.build()
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        N, deg = 700, 3
        shared = torch.Generator().manual_seed(7)
        means = torch.randn(N, 3, generator=shared)
        coeffs = torch.randn(N, 16, 3, generator=shared)

        def rank_data(r):
            g = torch.Generator().manual_seed(100 + r)
            vc = torch.randn(N, 3, generator=g)
            vc[torch.rand(N, generator=g) < 0.4] = 0.0  # Gaussians this view did not blend
            campos = torch.randn(3, generator=g) * 5
            blend = {"means": torch.randn(N, 3, generator=g), "quats": torch.randn(N, 4, generator=g),
                     "scales": torch.randn(N, 3, generator=g), "opacities": torch.randn(N, generator=g)}
            return vc, campos, blend

        vc, campos, blend = rank_data(rank)
        params = {"means": means.clone().requires_grad_(True), "sh_coeffs": coeffs.clone().requires_grad_(True)}
        for k in ("quats", "scales", "opacities"):
            params[k] = torch.zeros_like(blend[k]).requires_grad_(True)
        for k in blend:
            params[k].grad = blend[k].clone()

        class Deferred:
            pass
        d = Deferred()
        d.v_colors, d.campos, d.sh_degree = vc, campos, deg
        mv.exchange_gradients_compact(params, d, sh_views_fn=_oracle_sh_views)

        # what the plain path gives: per-rank SH backward, then a sum of everything over the ranks
        want = {k: torch.zeros_like(params[k]) for k in params}
        for r in range(world):
            vcr, cpr, br = rank_data(r)
            vco, vd = orc.sh_bwd(deg, (means - cpr).numpy(), coeffs.numpy(), None, vcr.numpy(), True)
            want["sh_coeffs"] += torch.from_numpy(vco)
            want["means"] += br["means"] + torch.from_numpy(vd)
            for k in ("quats", "scales", "opacities"):
                want[k] += br[k]
        ok = all(torch.allclose(params[k].grad, want[k], rtol=1e-5, atol=1e-5) for k in want)
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_compact_gradient_exchange_equals_allreduce_gloo():
    """all-gather of the 12-byte colour gradients + all-reduce of the geometry gradients + local expansion of all
    views == all-reduce of the per-rank expanded gradients (the SH expansion stands in through the oracle)."""
    world = 2
    port = 31500 + (os.getpid() % 2000)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_compact_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok in res)


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
