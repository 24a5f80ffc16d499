"""Python mirror of the reference's caller of the fastgs rasterizer:
src/training/rasterization/fast_rasterizer_autograd.cpp:10-160 (FastGSRasterize) and fast_rasterizer.cpp:12-74
(fast_rasterize: parameters from SplatData, background composite), over the torch.ops binding of
fast_gs::rasterization::{forward,backward}_wrapper (shim/FastGs.cpp, shim/fastgs_binding.cpp).

`FastGsBackend` wraps one native library's binding; the product backend is `default_backend()`; tests and bench.py
build a second one over the reference's own fastgs kernels (oracle/ref_fastgs.py) and drive both through the same
call sites.  No CPU fallback: without the native library every call raises.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch


@dataclass
class FastGSSettings:
    """fast_gs::rasterization::FastGSSettings (rasterization_api.h:13-24)."""
    cam_position: torch.Tensor
    active_sh_bases: int
    width: int
    height: int
    focal_x: float
    focal_y: float
    center_x: float
    center_y: float
    near_plane: float = 0.01   # fast_rasterizer.cpp:36
    far_plane: float = 1e10    # fast_rasterizer.cpp:37


class FastGsBackend:
    def __init__(self, namespace_getter):
        self._get = namespace_getter

    @property
    def ns(self):
        return self._get()

    def forward(self, means, scales_raw, rotations_raw, opacities_raw, sh0, shN, w2c, s: FastGSSettings, capacity: int = 0):
        """-> image [3,H,W], alpha [1,H,W], ctx (what the reference's autograd context saves).
        capacity > 0 (this backend only, include/fastgs/rasterization_ext.h): the instance buffer holds `capacity` entries
        and the count stays on the device (ctx["n_instances_dev"]): no host read-back, capturable in a CUDA graph."""
        if capacity > 0:
            r = self.ns.fastgs_forward_capacity(means, scales_raw, rotations_raw, opacities_raw, sh0, shN, w2c, s.cam_position,
                                                s.active_sh_bases, s.width, s.height, s.focal_x, s.focal_y, s.center_x,
                                                s.center_y, s.near_plane, s.far_plane, int(capacity))
            ints = torch.tensor([-1, int(capacity), 0, 0, 0], dtype=torch.int64)
            return r[0], r[1], dict(buffers=(r[2], r[3], r[4], torch.empty(0, dtype=torch.uint8, device=means.device)),
                                    ints=ints, n_instances_dev=r[5])
        r = self.ns.fastgs_forward(means, scales_raw, rotations_raw, opacities_raw, sh0, shN, w2c, s.cam_position,
                                   s.active_sh_bases, s.width, s.height, s.focal_x, s.focal_y, s.center_x, s.center_y,
                                   s.near_plane, s.far_plane)
        return r[0], r[1], dict(buffers=(r[2], r[3], r[4], r[5]), ints=r[6])

    def backward(self, ctx, grad_image, grad_alpha, image, alpha, means, scales_raw, rotations_raw, shN, w2c,
                 s: FastGSSettings, densification_info=None, want_w2c_grad=False):
        dens = densification_info if densification_info is not None else torch.empty(0, device=means.device)
        b = ctx["buffers"]
        r = self.ns.fastgs_backward(dens, grad_image, grad_alpha, image, alpha, means, scales_raw, rotations_raw, shN, b[0],
                                    b[1], b[2], b[3], w2c, s.cam_position, s.active_sh_bases, s.width, s.height, s.focal_x,
                                    s.focal_y, s.center_x, s.center_y, s.near_plane, s.far_plane, ctx["ints"], want_w2c_grad)
        return r[0], r[1], r[2], r[3], r[4], r[5], (r[6] if r[6].numel() else None)


def default_backend() -> FastGsBackend:
    from . import load

    def ns():
        load()
        return torch.ops.gsplat_b200

    return FastGsBackend(ns)


class FastGSRasterize(torch.autograd.Function):
    """fast_rasterizer_autograd.cpp:10-160."""

    @staticmethod
    def forward(ctx, backend, means, scales_raw, rotations_raw, opacities_raw, sh0, shN, w2c, densification_info, settings):
        image, alpha, fctx = backend.forward(means, scales_raw, rotations_raw, opacities_raw, sh0, shN, w2c, settings)
        ctx.backend, ctx.fctx, ctx.settings = backend, fctx, settings
        ctx.densification_info = densification_info
        ctx.save_for_backward(image, alpha, means, scales_raw, rotations_raw, shN, w2c)
        return image, alpha

    @staticmethod
    def backward(ctx, grad_image, grad_alpha):
        image, alpha, means, scales_raw, rotations_raw, shN, w2c = ctx.saved_tensors
        g = ctx.backend.backward(ctx.fctx, grad_image.contiguous(), grad_alpha.contiguous(), image, alpha, means, scales_raw,
                                 rotations_raw, shN, w2c, ctx.settings, ctx.densification_info,
                                 want_w2c_grad=ctx.needs_input_grad[7])
        return (None, g[0], g[1], g[2], g[3], g[4], g[5], g[6], None, None)


def fast_rasterize(backend: FastGsBackend, means, scales_raw, rotations_raw, opacities_raw, sh0, shN, w2c,
                   settings: FastGSSettings, bg_color=None, densification_info=None):
    """fast_rasterizer.cpp:12-74: render + background composite.  Returns (image [3,H,W], alpha [1,H,W])."""
    image, alpha = FastGSRasterize.apply(backend, means, scales_raw, rotations_raw, opacities_raw, sh0, shN, w2c,
                                         densification_info, settings)
    if bg_color is not None:
        image = image + (1.0 - alpha) * bg_color.reshape(3, 1, 1)
    return image, alpha
