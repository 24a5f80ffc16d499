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
    grads = [g for g in grads if g is not None]
    if not grads:
        return
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        if average_over and average_over != 1:
            for g in grads:
                g.div_(average_over)
        return
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


_side_streams: dict = {}


class _Marker:
    """Records named CUDA events into prof[name] (a list per name) -- only when a dict is passed; otherwise a no-op."""

    def __init__(self, prof, device):
        self.prof = prof if (prof is not None and device.type == "cuda") else None

    def __call__(self, name, stream=None):
        if self.prof is None:
            return
        ev = torch.cuda.Event(enable_timing=True)
        ev.record(stream) if stream is not None else ev.record()
        self.prof.setdefault(name, []).append(ev)


def phase_times(prof: dict) -> dict:
    """Average milliseconds between the markers of exchange_gradients_compact (call after a device synchronise)."""
    def avg(a, b):
        if a not in prof or b not in prof:
            return None
        return sum(x.elapsed_time(y) for x, y in zip(prof[a], prof[b])) / max(len(prof[a]), 1)
    return {"gather_ms": avg("start", "gathered"), "allreduce_after_gather_ms": avg("gathered", "reduced"),
            "expand_ms": avg("reduced" if "reduced" in prof else "gathered", "expanded"), "wait_and_add_ms": avg("expanded", "end"),
            "side_stream_allreduce_ms": avg("ar0", "ar1"), "total_ms": avg("start", "end")}


def _side_stream(device):
    key = (device.type, device.index)
    if key not in _side_streams:
        _side_streams[key] = torch.cuda.Stream(device)
    return _side_streams[key]


def exchange_gradients_compact(params: dict, deferred, sh_views_fn=None, average_over: int | None = None,
                               group=None, overlap_group=None, local_only: bool = False, prof: dict | None = None) -> None:
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
    the gloo tests pass the oracle's.

    `overlap_group`: a second process group over the same ranks (dist.new_group()).  The geometry all-reduce does not
    depend on the colour all-gather or on the SH expansion, so with its own communicator it runs on a side stream
    UNDER them (NVSwitch has the bandwidth for both); the expansion then writes the SH part of the position gradient
    into a scratch tensor that is added once the all-reduce has landed.  `deferred` may be a list (several views per
    rank): their colour gradients are stacked before the gather."""
    if sh_views_fn is None:
        from . import default_backend
        sh_views_fn = default_backend().spherical_harmonics_bwd_views
    means, coeffs = params["means"], params["sh_coeffs"]
    sinks = list(deferred) if isinstance(deferred, (list, tuple)) else [deferred]
    sh_degree = sinks[0].sh_degree
    if len(sinks) == 1:  # one view per rank (config E): views, no copy kernels
        vc = sinks[0].v_colors.reshape((1,) + tuple(sinks[0].v_colors.shape[-2:])).contiguous()  # [1, N, 3]
        cp = sinks[0].campos.reshape(1, 3).contiguous()
    else:
        vc = torch.stack([d.v_colors.reshape(d.v_colors.shape[-2:]) for d in sinks]).contiguous()  # [v_local, N, 3]
        cp = torch.stack([d.campos.reshape(3) for d in sinks]).contiguous()                        # [v_local, 3]
    geo = [params[k].grad for k in ("means", "quats", "scales", "opacities")]
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    if local_only:  # measurement aid: this rank's own views only, no collective (what the step costs without the exchange)
        world = 1
    pending, scratch = None, None
    mark = _Marker(prof, vc.device)  # measurement aid: CUDA events between the phases when `prof` is a dict
    mark("start")
    if world > 1:
        overlap = overlap_group is not None and vc.device.type == "cuda"
        if overlap:
            comp, side = torch.cuda.current_stream(vc.device), _side_stream(vc.device)
            side.wait_stream(comp)  # the blend gradients are complete
            with torch.cuda.stream(side):
                mark("ar0", side)
                allreduce_gradients(geo, group=overlap_group)
                mark("ar1", side)
            pending = side
            for g in geo:
                g.record_stream(side)
        vc_all = torch.empty((world * vc.shape[0],) + tuple(vc.shape[1:]), dtype=vc.dtype, device=vc.device)
        cp_all = torch.empty((world * cp.shape[0], 3), dtype=cp.dtype, device=cp.device)
        dist.all_gather_into_tensor(vc_all.view(-1, vc.shape[-1]), vc.view(-1, vc.shape[-1]), group=group)
        dist.all_gather_into_tensor(cp_all, cp, group=group)
        mark("gathered")
        if not overlap:
            allreduce_gradients(geo, group=group)
            mark("reduced")
    else:
        vc_all, cp_all = vc, cp
    with torch.no_grad():
        v_means = means.grad
        if pending is not None:  # means.grad is still being reduced: expand into a scratch tensor
            scratch = torch.zeros_like(means.grad)
            v_means = scratch
        coeffs.grad = sh_views_fn(sh_degree, means.detach().contiguous(), cp_all.contiguous(),
                                  coeffs.detach().contiguous(), vc_all.contiguous(), v_means)
        mark("expanded")
        if pending is not None:
            torch.cuda.current_stream(vc.device).wait_stream(pending)
            means.grad.add_(scratch)
        mark("end")
    if average_over and average_over != 1:
        for k in ("means", "quats", "scales", "opacities", "sh_coeffs"):
            params[k].grad.div_(average_over)


class PeerColourExchange:
    """This rank's block [N*3 colour gradients | camera position] in NVLink-mapped symmetric memory
    (torch.distributed._symmetric_memory: cuMem allocations exchanged between the ranks of one node, every rank holds
    the device address of every peer's block).  The SH expansion (gsb_sh_bwd_views_peer) reads its V views straight
    from the peers through NVSwitch: the all-gather of the exchange step runs INSIDE the kernel, overlapped with the
    expansion, and the gathered [V,N,3] tensor is never written or re-read.  Construction is collective; it raises
    when symmetric memory is not available (the caller then keeps the NCCL all-gather path)."""

    def __init__(self, n_gauss: int, device, group=None):
        import torch.distributed._symmetric_memory as symm
        grp = group if group is not None else dist.group.WORLD
        self.n = int(n_gauss)
        try:  # older releases want the group enabled first; newer ones deprecate the call
            symm.enable_symm_mem_for_group(grp.group_name)
        except Exception:
            pass
        self.buf = symm.empty(self.n * 3 + 4, dtype=torch.float32, device=device)
        self.hdl = symm.rendezvous(self.buf, grp)
        self.addrs = [int(a) for a in self.hdl.buffer_ptrs]
        self.world = int(self.hdl.world_size)
        if len(self.addrs) != self.world or any(a == 0 for a in self.addrs):
            raise RuntimeError("symmetric memory rendezvous returned no peer addresses")

    def publish(self, v_colors: torch.Tensor, campos: torch.Tensor) -> None:
        """Copy this rank's view into its block, then a device-side barrier: every rank's block is complete."""
        self.buf[: self.n * 3].copy_(v_colors.reshape(-1))
        self.buf[self.n * 3: self.n * 3 + 3].copy_(campos.reshape(3))
        self.hdl.barrier(channel=0)

    def release(self) -> None:
        """Device-side barrier after the expansion: nobody still reads a block when its owner overwrites it."""
        self.hdl.barrier(channel=1)


def exchange_gradients_peer(params: dict, deferred, peer: PeerColourExchange, sh_views_peer_fn=None,
                            average_over: int | None = None, group=None, overlap_group=None) -> None:
    """exchange_gradients_compact with the colour all-gather fused into the SH expansion kernel over peer memory
    (one view per rank).  Same sums in the same order as the compact path: the results are bit-identical."""
    if sh_views_peer_fn is None:
        from . import default_backend
        sh_views_peer_fn = default_backend().spherical_harmonics_bwd_views_peer
    if isinstance(deferred, (list, tuple)):
        if len(deferred) != 1:
            raise ValueError("the peer-memory exchange carries one view per rank")
        deferred = deferred[0]
    means, coeffs = params["means"], params["sh_coeffs"]
    geo = [params[k].grad for k in ("means", "quats", "scales", "opacities")]
    dev = means.device
    pending = None
    if overlap_group is not None:
        comp, side = torch.cuda.current_stream(dev), _side_stream(dev)
        side.wait_stream(comp)
        with torch.cuda.stream(side):
            allreduce_gradients(geo, group=overlap_group)
        pending = side
        for g in geo:
            g.record_stream(side)
    peer.publish(deferred.v_colors, deferred.campos)
    if pending is None:
        allreduce_gradients(geo, group=group)
    with torch.no_grad():
        v_means = torch.zeros_like(means.grad) if pending is not None else means.grad
        coeffs.grad = sh_views_peer_fn(deferred.sh_degree, means.detach().contiguous(), coeffs.detach().contiguous(),
                                       peer.addrs, v_means)
        peer.release()
        if pending is not None:
            torch.cuda.current_stream(dev).wait_stream(pending)
            means.grad.add_(v_means)
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
