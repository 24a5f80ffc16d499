"""Host-side mirror of the reference's training iteration around the hot path:

  photometric_loss   trainer.cpp:103-126 (L1 + fused SSIM "valid") as one kernel that also returns dLoss/d(render)
  FusedAdam          src/training/optimizers/fused_adam.cpp:22-117 + strategy_utils.cpp:27-55 (per-group learning rates,
                     eps 1e-15, exponential decay of the position learning rate, shN frozen for the first 1000
                     iterations), all groups stepped by ONE kernel launch (gsb_adam_step)
  TrainStep          render (extended operator on the raw tensors) -> loss -> backward -> Adam [-> MCMC noise], with
                     the intersection buffer sized from a capacity so that the iteration has no host read-back

Everything that computes lives behind the C ABI; this file only orders the calls, like the reference's trainer does."""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import torch

from . import _product_ns, load, rasterize_fused


class _PhotometricLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, renders, target, lambda_dssim):
        stats, v = _product_ns().photometric_loss_fused(renders.detach().contiguous(), target.contiguous(),
                                                        float(lambda_dssim), renders.requires_grad)
        ctx.save_for_backward(v)
        ctx.mark_non_differentiable(stats)
        return stats[0], stats

    @staticmethod
    def backward(ctx, g_loss, _g_stats):
        (v,) = ctx.saved_tensors
        return (v * g_loss if v.numel() else None), None, None


def photometric_loss(renders, target, lambda_dssim: float = 0.2):
    """(loss, stats) with stats = device tensor [loss, l1 mean, ssim mean].  renders [1,H,W,3] unclamped (the blend's
    output); target [3,H,W] (the reference's image layout) or [1,H,W,3]."""
    load()
    return _PhotometricLoss.apply(renders, target, lambda_dssim)


PARAM_GROUPS = ("means", "sh0", "shN", "scaling_raw", "rotation_raw", "opacity_raw")  # order of strategy_utils.cpp:27-55


@dataclass
class AdamConfig:
    """parameter/*_optimization_params.json defaults (include/core/parameters.hpp)."""
    means_lr: float = 0.00016
    shs_lr: float = 0.0025
    scaling_lr: float = 0.005
    rotation_lr: float = 0.001
    opacity_lr: float = 0.05
    scene_scale: float = 1.0
    iterations: int = 30000
    beta1: float = 0.9
    beta2: float = 0.999
    eps: float = 1e-15

    def lrs(self):
        return [self.means_lr * self.scene_scale, self.shs_lr, self.shs_lr / 20.0, self.scaling_lr, self.rotation_lr,
                self.opacity_lr]


class FusedAdam:
    """One launch per step for the six SplatData groups.  State tensors are allocated lazily like the reference's."""

    def __init__(self, params: dict, cfg: AdamConfig | None = None):
        self.cfg = cfg or AdamConfig()
        self.params = params
        self.lr = self.cfg.lrs()
        self.gamma = 0.01 ** (1.0 / self.cfg.iterations)  # ExponentialLR on group 0 (strategy_utils.cpp:52-55)
        self.exp_avg = {k: torch.zeros_like(v) for k, v in params.items()}
        self.exp_avg_sq = {k: torch.zeros_like(v) for k, v in params.items()}
        self.steps = {k: 0 for k in params}

    def step(self, iteration: int):
        names, counts = [], []
        for k in PARAM_GROUPS:
            p = self.params[k]
            if p.grad is None or p.numel() == 0:
                continue
            self.steps[k] += 1
            if k == "shN" and iteration <= 1000:  # fused_adam.cpp:68-70: the step count still advances
                continue
            names.append(k)
            counts.append(self.steps[k])
        if names:
            with torch.no_grad():
                _product_ns().fused_adam_step([self.params[k] for k in names], [self.params[k].grad for k in names],
                                              [self.exp_avg[k] for k in names], [self.exp_avg_sq[k] for k in names],
                                              [self.lr[PARAM_GROUPS.index(k)] for k in names], self.cfg.beta1,
                                              self.cfg.beta2, self.cfg.eps, counts)
        self.lr[0] *= self.gamma

    def zero_grad(self):
        for p in self.params.values():
            p.grad = None


@dataclass
class TrainStep:
    """One training iteration on resident raw parameters: camera + target in, loss (device scalar) out."""
    params: dict
    sh_degree: int
    width: int
    height: int
    lambda_dssim: float = 0.2
    optimizer: FusedAdam | None = None
    mcmc_noise_lr: float = 0.0          # > 0: gsplat::add_noise every iteration (mcmc.cpp:349-367)
    capacity: int = 0                    # intersections the buffers are sized for; 0 = size exactly (one read-back)
    iteration: int = 0
    last: dict = field(default_factory=dict)

    def __call__(self, viewmat, K, target, background=None):
        P = self.params
        self.iteration += 1
        out = rasterize_fused(P["means"], P["sh0"], P["shN"], P["scaling_raw"], P["rotation_raw"], P["opacity_raw"],
                              self.sh_degree, viewmat, K, self.width, self.height, bg_color=background,
                              isect_capacity=self.capacity)
        loss, stats = photometric_loss(out.render_colors, target, self.lambda_dssim)
        loss.backward()
        if self.optimizer is not None:
            self.optimizer.step(self.iteration)
            if self.mcmc_noise_lr > 0.0:
                with torch.no_grad():
                    noise = torch.randn_like(P["means"])
                    _product_ns().add_noise(P["opacity_raw"].detach().reshape(-1), P["scaling_raw"].detach(),
                                            P["rotation_raw"].detach(), noise, P["means"].detach(),
                                            self.mcmc_noise_lr * self.optimizer.lr[0])
            self.optimizer.zero_grad()
        self.last = {"n_isects": out.n_isects, "stats": stats}
        return loss

    def size_capacity(self, viewmat, K, target, background=None, slack: float = 1.25):
        """An exact (synchronising) render sizes the intersection capacity for the following steps."""
        with torch.no_grad():
            P = self.params
            out = rasterize_fused(P["means"], P["sh0"], P["shN"], P["scaling_raw"], P["rotation_raw"], P["opacity_raw"],
                                  self.sh_degree, viewmat, K, self.width, self.height, bg_color=background)
            n = int(out.n_isects.item())
        self.capacity = int(n * slack) + 1024
        return n
