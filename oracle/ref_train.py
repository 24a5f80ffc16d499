"""The reference's training-step pieces around the rasterizer, driven through ITS OWN kernels (oracle/_ref):

  ref_photometric_loss   trainer.cpp:103-126 with fused_ssim(..., "valid") exactly as include/kernels/fused_ssim.cuh
                         wires the forward / backward kernels of src/training/kernels/ssim.cu into autograd
  RefFusedAdam           src/training/optimizers/fused_adam.cpp:22-95: one adam_step_cu launch per parameter tensor

TEST INFRASTRUCTURE ONLY (parity pin of SURVEY.md 8 f2 / f3, same-box baseline of the training iteration)."""
from __future__ import annotations

import math

import torch

from . import ref_ops

C1, C2 = 0.01 ** 2, 0.03 ** 2


class _RefFusedSSIM(torch.autograd.Function):
    """fs_internal::_FusedSSIM (fused_ssim.cuh:27-110), padding == "valid"."""

    @staticmethod
    def forward(ctx, img1, img2):
        ns = ref_ops._ns()
        img1, img2 = img1.contiguous(), img2.contiguous()
        m, dm1, ds1, ds12 = ns.fusedssim(C1, C2, img1, img2, True)
        h, w = m.shape[2], m.shape[3]
        ctx.crop = h > 10 and w > 10
        ctx.save_for_backward(img1.detach(), img2, dm1, ds1, ds12)
        return m[:, :, 5:h - 5, 5:w - 5] if ctx.crop else m

    @staticmethod
    def backward(ctx, g):
        img1, img2, dm1, ds1, ds12 = ctx.saved_tensors
        if ctx.crop:
            full = torch.zeros_like(img1)
            full[:, :, 5:img1.shape[2] - 5, 5:img1.shape[3] - 5] = g
            g = full
        return ref_ops._ns().fusedssim_backward(C1, C2, img1, img2, g.contiguous(), dm1, ds1, ds12), None


def ref_photometric_loss(image_chw, gt_chw, lambda_dssim=0.2):
    """image_chw = RenderOutput.image ([3,H,W], clamped), as Trainer::compute_photometric_loss receives it."""
    rendered = image_chw.unsqueeze(0) if image_chw.dim() == 3 else image_chw
    gt = gt_chw.unsqueeze(0) if gt_chw.dim() == 3 else gt_chw
    l1 = torch.nn.functional.l1_loss(rendered, gt)
    ssim = _RefFusedSSIM.apply(rendered, gt).mean()
    return (1.0 - lambda_dssim) * l1 + lambda_dssim * (1.0 - ssim)


class RefFusedAdam:
    def __init__(self, params: dict, lrs, iterations=30000, beta1=0.9, beta2=0.999, eps=1e-15):
        self.params, self.lr = params, list(lrs)
        self.b1, self.b2, self.eps = beta1, beta2, eps
        self.gamma = 0.01 ** (1.0 / iterations)
        self.m = {k: torch.zeros_like(v) for k, v in params.items()}
        self.v = {k: torch.zeros_like(v) for k, v in params.items()}
        self.steps = {k: 0 for k in params}

    def step(self, iteration, order):
        ns = ref_ops._ns()
        with torch.no_grad():
            for i, k in enumerate(order):
                p = self.params[k]
                if p.grad is None:
                    continue
                self.steps[k] += 1
                if i == 2 and iteration <= 1000:
                    continue
                t = self.steps[k]
                ns.adam_step(p, self.m[k], self.v[k], p.grad.contiguous(), self.lr[i], self.b1, self.b2, self.eps,
                             1.0 / (1.0 - self.b1 ** t), 1.0 / math.sqrt(1.0 - self.b2 ** t))
        self.lr[0] *= self.gamma

    def zero_grad(self):
        for p in self.params.values():
            p.grad = None
