"""Host-staged steps: the software pipeline behind the end-to-end number of bench.py.

When every input of a step lives in (pinned) host memory and every result has to return there, a
step is three transfers long: inputs host->device, the hot path, results device->host.  PCIe is full
duplex and the copy engines are independent of the SMs, so step k+1's inputs, step k's kernels and
step k-1's results can be in flight together on three streams.  Two device-side slots of inputs and
two host-side slots of results are enough; events order the slots, and the only host waits are the
hot path's own n_isects read-back and the (one step late) consumption of a result.

Nothing here is specific to the blend: `step_fn` is any callable that takes the dict of device
inputs, leaves gradients in `.grad` of the tensors named by `grad_names` and returns
(loss, image)."""
from __future__ import annotations

from typing import Callable, Dict, Sequence

import torch


class HostStagedSteps:
    def __init__(self, device: torch.device, host_inputs: Dict[str, torch.Tensor], grad_names: Sequence[str],
                 step_fn: Callable[[Dict[str, torch.Tensor]], tuple]):
        for k, t in host_inputs.items():
            if not t.is_pinned():
                raise ValueError(f"host input {k!r} must be in pinned memory (asynchronous copies need it)")
        self.dev = device
        self.host = host_inputs
        self.names = tuple(grad_names)
        self.step_fn = step_fn
        self.s_h2d = torch.cuda.Stream(device)
        self.s_d2h = torch.cuda.Stream(device)
        self.P = [{k: torch.empty_like(t, device=device).requires_grad_(k in self.names)
                   for k, t in host_inputs.items()} for _ in range(2)]
        self.host_grads = [{k: torch.empty_like(host_inputs[k]).pin_memory() for k in self.names} for _ in range(2)]
        self.host_img = [None, None]
        self.host_loss = [torch.empty((), dtype=torch.float32).pin_memory() for _ in range(2)]
        self.ev_h2d = [torch.cuda.Event() for _ in range(2)]
        self.ev_comp = [torch.cuda.Event() for _ in range(2)]
        self.ev_d2h = [torch.cuda.Event() for _ in range(2)]
        self._comp_recorded = [False, False]
        self.h2d_bytes = sum(t.numel() * t.element_size() for t in host_inputs.values())
        self.d2h_bytes = None

    def _prefetch(self, slot: int):
        if self._comp_recorded[slot]:  # the kernels that last read this slot must be done with it
            self.s_h2d.wait_event(self.ev_comp[slot])
        with torch.cuda.stream(self.s_h2d), torch.no_grad():
            for k, t in self.host.items():
                self.P[slot][k].copy_(t, non_blocking=True)
            self.ev_h2d[slot].record(self.s_h2d)

    def run(self, steps: int) -> list[float]:
        """`steps` pipelined steps; returns the loss of each (read back from the host copies)."""
        comp = torch.cuda.current_stream(self.dev)
        keep = [None, None]
        losses: list[float] = []
        self._comp_recorded = [False, False]
        self._prefetch(0)
        for k in range(steps):
            slot = k & 1
            if k + 1 < steps:
                self._prefetch(slot ^ 1)
            comp.wait_event(self.ev_h2d[slot])
            loss, image = self.step_fn(self.P[slot])
            self.ev_comp[slot].record(comp)
            self._comp_recorded[slot] = True
            grads = [self.P[slot][n].grad for n in self.names]
            if self.host_img[slot] is None:
                self.host_img[slot] = torch.empty(image.shape, dtype=image.dtype).pin_memory()
                self.d2h_bytes = (image.numel() * image.element_size() + 4 +
                                  sum(g.numel() * g.element_size() for g in grads))
            self.s_d2h.wait_event(self.ev_comp[slot])
            with torch.cuda.stream(self.s_d2h), torch.no_grad():
                self.host_img[slot].copy_(image.detach(), non_blocking=True)
                for n, g in zip(self.names, grads):
                    self.host_grads[slot][n].copy_(g, non_blocking=True)
                self.host_loss[slot].copy_(loss.detach(), non_blocking=True)
                self.ev_d2h[slot].record(self.s_d2h)
            keep[slot] = (loss, image, grads)  # alive until their copies have landed
            if k >= 1:  # consume the previous step's result while this one runs
                self.ev_d2h[slot ^ 1].synchronize()
                losses.append(float(self.host_loss[slot ^ 1]))
                keep[slot ^ 1] = None
        if steps:
            last = (steps - 1) & 1
            self.ev_d2h[last].synchronize()
            losses.append(float(self.host_loss[last]))
            keep[last] = None
        return losses
