"""Multi-GPU host logic of the hot path: the path shards by VIEW -- the kernels are
single-camera (RasterizeToPixelsFromWorld3DGSFwd.cu:197-200) and every rank keeps a full replica of the
parameters -- and has exactly one exchange step, the sum of the per-Gaussian gradients.

One process per GPU (torchrun); the collective is NCCL on GPUs and gloo in the CPU tests."""
from __future__ import annotations

from typing import Iterable, Sequence

import torch
import torch.distributed as dist


def views_for_rank(n_views: int, rank: int, world: int) -> list[int]:
    """Round-robin assignment of the step's views to ranks (config E: n_views == world -> [rank])."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    return list(range(rank, n_views, world))


def allreduce_gradients(grads: Sequence[torch.Tensor], average_over: int | None = None, group=None) -> None:
    """Sum the gradient tensors over all ranks IN PLACE with one coalesced collective launch
    (ncclGroupStart/End around five all-reduces: 236 B per Gaussian), optionally dividing by the number
    of views so the result is the gradient of the mean loss."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        if average_over and average_over != 1:
            for g in grads:
                g.div_(average_over)
        return
    grads = [g for g in grads if g is not None]
    dev = grads[0].device
    if dev.type == "cuda":
        with dist._coalescing_manager(group=group, device=dev, async_ops=False):
            for g in grads:
                dist.all_reduce(g, group=group)
    else:  # gloo: no coalescing manager needed
        for g in grads:
            dist.all_reduce(g, group=group)
    if average_over and average_over != 1:
        for g in grads:
            g.div_(average_over)


def exchange_gradients_compact(params: dict, deferred, sh_views_fn=None, average_over: int | None = None,
                               group=None) -> None:
    """The exchange step with 1/2 to 1/3 of the all-reduce's traffic.

    `params` maps 'means', 'quats', 'scales', 'opacities', 'sh_coeffs' to the leaf tensors; after
    backward() through rasterize(..., sh_exchange=deferred) their .grad hold this rank's BLEND gradients
    (sh_coeffs.grad is None) and `deferred` the view's 12-byte colour gradients.  Per view the SH gradient
    row is Y(dir) (x) v_color, so instead of all-reducing 192 B/Gaussian of expanded rows the ranks
      1. all-gather v_colors [V, N, 3] and the camera positions [V, 3],
      2. all-reduce the four geometry gradients (44 B/Gaussian, one coalesced launch),
      3. expand and sum all V views locally (gsb_sh_bwd_views): sh_coeffs.grad is written, the SH part of the
         position gradient is added to the (already reduced) means.grad.
    Result: every rank holds the same sums as allreduce_gradients() would give after per-rank SH backward.
    `sh_views_fn(degree, means, campos, coeffs, v_colors, v_means) -> v_coeffs` defaults to the product op;
    the gloo tests pass the oracle's."""
    if sh_views_fn is None:
        from . import default_backend
        sh_views_fn = default_backend().spherical_harmonics_bwd_views
    means, coeffs = params["means"], params["sh_coeffs"]
    vc, cp = deferred.v_colors, deferred.campos.reshape(1, 3).contiguous()
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    if world > 1:
        vc_all = torch.empty((world,) + tuple(vc.shape), dtype=vc.dtype, device=vc.device)
        cp_all = torch.empty((world, 3), dtype=cp.dtype, device=cp.device)
        dist.all_gather_into_tensor(vc_all.view(-1, vc.shape[-1]), vc, group=group)  # concatenation along dim 0
        dist.all_gather_into_tensor(cp_all, cp, group=group)
        allreduce_gradients([params[k].grad for k in ("means", "quats", "scales", "opacities")], group=group)
    else:
        vc_all, cp_all = vc.unsqueeze(0), cp
    with torch.no_grad():
        coeffs.grad = sh_views_fn(deferred.sh_degree, means.detach().contiguous(), cp_all.contiguous(),
                                  coeffs.detach().contiguous(), vc_all.contiguous(), means.grad)
    if average_over and average_over != 1:
        for k in ("means", "quats", "scales", "opacities", "sh_coeffs"):
            params[k].grad.div_(average_over)


def max_over_ranks(value: float, device, group=None) -> float:
    """Device timing of a multi-GPU step is the maximum over ranks."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return float(value)
    t = torch.tensor([value], device=device, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
