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


def max_over_ranks(value: float, device, group=None) -> float:
    """Device timing of a multi-GPU step is the maximum over ranks."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return float(value)
    t = torch.tensor([value], device=device, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
