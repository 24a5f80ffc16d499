"""Worker of tests/test_multigpu_nccl.py (one process per GPU, launched by torchrun): every rank renders its own view
of the same small scene and the ranks exchange gradients three ways -- plain all-reduce, compact exchange, compact
exchange with the geometry all-reduce overlapped on a second communicator, and the compact exchange with the colour
all-gather fused into the SH expansion kernel over NVLink peer memory (when symmetric memory can be mapped).  All must
give every rank the SUM of the per-view gradients, which rank 0 also computes alone by rendering all views itself."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import __graft_entry__ as ge
    import scenes
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    pg2 = dist.new_group()
    pkg = ge.load_package()
    pkg.load()
    from gsplat_b200 import multiview as mv
    names = ("means", "quats", "scales", "opacities", "sh_coeffs")

    def view_scene(v):
        return scenes.scene_b(N=30000, width=640, height=360, view=v, scale_mul=2.0)

    def render(P, sc, t, deferred=None):
        out = pkg.rasterize(P["means"], P["quats"], P["scales"], P["opacities"], P["sh_coeffs"], sc["sh_degree"],
                            t["viewmats"], t["Ks"], sc["width"], sc["height"], bg_color=t["background"],
                            sh_exchange=deferred)
        tgt = torch.full_like(out.render_colors, 0.4)
        ((out.render_colors - tgt) ** 2).mean().backward()

    def fresh(sc):
        t = {k: torch.from_numpy(v).to(dev) for k, v in sc.items() if isinstance(v, np.ndarray)}
        return t, {k: t[k].clone().requires_grad_(True) for k in names}

    sc = view_scene(rank)
    results = {}
    for mode in ("allreduce", "compact", "compact_overlap"):
        t, P = fresh(sc)
        if mode == "allreduce":
            render(P, sc, t)
            mv.allreduce_gradients([P[k].grad for k in names])
        else:
            d = pkg.DeferredSHBackward()
            render(P, sc, t, d)
            mv.exchange_gradients_compact(P, d, overlap_group=pg2 if mode == "compact_overlap" else None)
        torch.cuda.synchronize()
        results[mode] = {k: P[k].grad.clone() for k in names}

    # peer-memory exchange: construction is collective; all ranks agree on whether it is available
    peer, why = None, ""
    try:
        peer = mv.PeerColourExchange(sc["means"].shape[0], dev)
        okp = 1
    except Exception as e:  # no symmetric memory on this box: reported, the NCCL paths are still checked
        why, okp = repr(e)[:300], 0
    flagp = torch.tensor([okp], device=dev)
    dist.all_reduce(flagp, op=dist.ReduceOp.MIN)
    if int(flagp.item()) == 1:
        for mode in ("peer", "peer_overlap"):
            t, P = fresh(sc)
            d = pkg.DeferredSHBackward()
            render(P, sc, t, d)
            mv.exchange_gradients_peer(P, d, peer, overlap_group=pg2 if mode == "peer_overlap" else None)
            torch.cuda.synchronize()
            results[mode] = {k: P[k].grad.clone() for k in names}
    else:
        print(f"rank {rank}: peer-memory exchange unavailable: {why or 'another rank failed'}", flush=True)

    def rel(a, b):
        return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))

    ok = True
    msgs = []
    for k in names:
        for mode in [m for m in ("compact", "compact_overlap", "peer", "peer_overlap") if m in results]:
            e = rel(results[mode][k], results["allreduce"][k])
            msgs.append(f"rank {rank} {mode} vs allreduce {k}: {e:.2e}")
            ok = ok and e < 2e-5
    if "peer" in results:  # same sums in the same order as the NCCL all-gather path (each mode renders again, and the
        # blend's atomic adds retire in a different order every time: last-bit differences only)
        for k in names:
            e = rel(results["peer"][k], results["compact"][k])
            msgs.append(f"rank {rank} peer vs compact {k}: {e:.2e}")
            ok = ok and e < 5e-6
    if rank == 0:  # the sum of the views' gradients, computed without any collective
        total = None
        for v in range(world):
            scv = view_scene(v)
            t, P = fresh(scv)
            render(P, scv, t)
            g = {k: P[k].grad for k in names}
            total = g if total is None else {k: total[k] + g[k] for k in names}
        for k in names:
            e = rel(results["allreduce"][k], total[k])
            msgs.append(f"rank 0 allreduce vs local sum of {world} views {k}: {e:.2e}")
            ok = ok and e < 2e-5
    flag = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    print("\n".join(msgs), flush=True)
    dist.destroy_process_group()
    sys.exit(0 if float(flag.item()) == 1.0 else 1)


if __name__ == "__main__":
    main()
