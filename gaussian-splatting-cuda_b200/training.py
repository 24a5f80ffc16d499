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


class GraphedTrainStep:
    """The whole iteration -- extended render operator, loss-and-gradient kernel, blend backward, fused back kernel, Adam --
    captured ONCE in a CUDA graph and replayed: one graph launch per iteration instead of ~25 kernel launches and their
    torch bookkeeping.  Possible because nothing inside the iteration needs a host-side value: the intersection buffers
    are sized from a capacity (the count stays on the device, `overflowed()` checks it afterwards) and Adam's
    step-dependent scalars (learning rates, bias corrections, the shN freeze of the first 1000 iterations) are read from a
    small device tensor that is refreshed before every replay.

    Per iteration the host copies the camera (viewmat, K), the background and the target image into static device buffers
    (pinned -> device, on the replay stream) and reads the loss back when it wants to."""

    def __init__(self, params: dict, sh_degree: int, width: int, height: int, cfg: AdamConfig | None = None,
                 lambda_dssim: float = 0.2, capacity_slack: float = 1.25):
        load()
        self.params, self.sh_degree, self.W, self.H = params, int(sh_degree), int(width), int(height)
        self.lambda_dssim = float(lambda_dssim)
        self.opt = FusedAdam(params, cfg)
        self.slack = capacity_slack
        dev = params["means"].device
        self.dev = dev
        self.static = {"viewmat": torch.zeros((1, 4, 4), device=dev), "K": torch.zeros((1, 3, 3), device=dev),
                       "background": torch.zeros((1, 3), device=dev),
                       "target": torch.zeros((3, self.H, self.W), device=dev)}
        self.dyn = torch.zeros((len(PARAM_GROUPS), 4), device=dev)
        self.dyn_host = torch.zeros((len(PARAM_GROUPS), 4)).pin_memory()
        self.graph = None
        self.loss = None
        self.stats = None
        self.n_isects = None
        self.capacity = 0
        self.iteration = 0

    def _body(self):
        P, S = self.params, self.static
        out = rasterize_fused(P["means"], P["sh0"], P["shN"], P["scaling_raw"], P["rotation_raw"], P["opacity_raw"],
                              self.sh_degree, S["viewmat"], S["K"], self.W, self.H, bg_color=S["background"],
                              isect_capacity=self.capacity)
        loss, stats = photometric_loss(out.render_colors, S["target"], self.lambda_dssim)
        loss.backward()
        names = [k for k in PARAM_GROUPS if P[k].numel() > 0]
        with torch.no_grad():
            _product_ns().fused_adam_step_dynamic([P[k] for k in names], [P[k].grad for k in names],
                                                  [self.opt.exp_avg[k] for k in names],
                                                  [self.opt.exp_avg_sq[k] for k in names],
                                                  self.dyn[[PARAM_GROUPS.index(k) for k in names]].contiguous()
                                                  if len(names) != len(PARAM_GROUPS) else self.dyn,
                                                  self.opt.cfg.beta1, self.opt.cfg.beta2, self.opt.cfg.eps)
        return loss, stats, out.n_isects

    def _set_inputs(self, viewmat, K, target, background):
        S = self.static
        S["viewmat"].copy_(viewmat.reshape(1, 4, 4), non_blocking=True)
        S["K"].copy_(K.reshape(1, 3, 3), non_blocking=True)
        S["target"].copy_(target.reshape(3, self.H, self.W), non_blocking=True)
        if background is not None:
            S["background"].copy_(background.reshape(1, 3), non_blocking=True)

    def _set_scalars(self):
        """fused_adam.cpp:61-81 on the host: step counts, bias corrections, the shN freeze; strategy_utils.cpp:52-55 lr."""
        o = self.opt
        for i, k in enumerate(PARAM_GROUPS):
            o.steps[k] += 1
            t = o.steps[k]
            enabled = 0.0 if (k == "shN" and self.iteration <= 1000) else 1.0
            self.dyn_host[i, 0] = o.lr[i]
            self.dyn_host[i, 1] = 1.0 / (1.0 - o.cfg.beta1 ** t)
            self.dyn_host[i, 2] = 1.0 / math.sqrt(1.0 - o.cfg.beta2 ** t)
            self.dyn_host[i, 3] = enabled
        self.dyn.copy_(self.dyn_host, non_blocking=True)
        o.lr[0] *= o.gamma

    def capture(self, viewmat, K, target, background=None):
        """Sizes the intersection capacity with one exact render, warms up and captures the graph."""
        self._set_inputs(viewmat, K, target, background)
        with torch.no_grad():
            P, S = self.params, self.static
            out = rasterize_fused(P["means"], P["sh0"], P["shN"], P["scaling_raw"], P["rotation_raw"], P["opacity_raw"],
                                  self.sh_degree, S["viewmat"], S["K"], self.W, self.H, bg_color=S["background"])
            n = int(out.n_isects.item())
        self.capacity = int(n * self.slack) + 1024
        self.dyn.zero_()  # warm-up iterations must not move the parameters: every group disabled
        side = torch.cuda.Stream(self.dev)
        side.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(side):
            for _ in range(3):
                for p in self.params.values():
                    p.grad = None
                self._body()
        torch.cuda.current_stream(self.dev).wait_stream(side)
        for p in self.params.values():
            p.grad = None
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss, self.stats, self.n_isects = self._body()
        return n

    def __call__(self, viewmat, K, target, background=None):
        if self.graph is None:
            self.capture(viewmat, K, target, background)
        self.iteration += 1
        self._set_inputs(viewmat, K, target, background)
        self._set_scalars()
        self.graph.replay()
        return self.loss

    def overflowed(self) -> bool:
        """True when the last iteration had more intersections than the captured buffers hold (re-capture then)."""
        return int(self.n_isects.item()) > self.capacity


@dataclass
class FastGsTrainStep:
    """One training iteration on the reference's DEFAULT rasterizer path (SURVEY.md 8 f4 + f2 + f3), ordered like
    the reference's trainer (trainer.cpp: fast_rasterize -> photometric loss -> backward -> optimizer step):
    fastgs forward (raw parameters in) -> background composite (fast_rasterizer.cpp:71) -> fused SSIM + L1 loss and
    gradient on the [3,H,W] image -> fastgs backward -> one-launch Adam.  No autograd graph is built: the loss kernel
    returns dLoss/d(image) and the backward takes it directly."""
    params: dict                          # PARAM_GROUPS -> leaf tensors (opacity_raw [N,1] or [N])
    sh_degree: int
    width: int
    height: int
    lambda_dssim: float = 0.2
    optimizer: FusedAdam | None = None
    backend: object = None                # fastgs.FastGsBackend; default: this library
    capacity: int = 0                     # > 0: instance buffer sized from it, no host read-back (rasterization_ext.h)
    iteration: int = 0
    last: dict = field(default_factory=dict)

    def size_capacity(self, w2c, settings, slack: float = 1.25):
        """One exact (synchronising) forward sizes the instance capacity for the following iterations."""
        from . import fastgs as fg
        be = self.backend or fg.default_backend()
        P = self.params
        with torch.no_grad():
            _, _, ctx = be.forward(P["means"], P["scaling_raw"], P["rotation_raw"], P["opacity_raw"].reshape(-1, 1), P["sh0"],
                                   P["shN"], w2c, settings)
        n = int(ctx["ints"][1])
        self.capacity = int(n * slack) + 1024
        return n

    def __call__(self, w2c, settings, target, background=None, densification_info=None):
        from . import fastgs as fg
        be = self.backend or fg.default_backend()
        P = self.params
        self.iteration += 1
        with torch.no_grad():
            opac = P["opacity_raw"].reshape(-1, 1)
            image, alpha, ctx = be.forward(P["means"], P["scaling_raw"], P["rotation_raw"], opac, P["sh0"], P["shN"], w2c, settings,
                                           capacity=self.capacity)
            final = image + (1.0 - alpha) * background.reshape(3, 1, 1) if background is not None else image
            stats, v_final = _product_ns().photometric_loss_fused(final.contiguous(), target.contiguous(),
                                                                  float(self.lambda_dssim), True)
            v_alpha = (-(v_final * background.reshape(3, 1, 1)).sum(0, keepdim=True) if background is not None
                       else torch.zeros_like(alpha))
            g = be.backward(ctx, v_final, v_alpha.contiguous(), image, alpha, P["means"], P["scaling_raw"], P["rotation_raw"],
                            P["shN"], w2c, settings, densification_info)
            for k, gk in zip(("means", "scaling_raw", "rotation_raw", "opacity_raw", "sh0", "shN"), g[:6]):
                P[k].grad = gk.reshape(P[k].shape)
        if self.optimizer is not None:
            self.optimizer.step(self.iteration)
            self.optimizer.zero_grad()
        # capacity mode: the count is a device tensor (compare it with the capacity when convenient)
        self.last = {"n_instances": ctx.get("n_instances_dev", int(ctx["ints"][1])), "stats": stats}
        return stats[0]


class GraphedFastGsTrainStep(GraphedTrainStep):
    """GraphedTrainStep on the fastgs path: fastgs forward (capacity mode) -> background composite -> loss-and-gradient
    kernel -> fastgs backward -> Adam with device-side scalars, captured once and replayed.  The camera pose, camera
    position, background and target are static device buffers refreshed before each replay; the intrinsics and the image
    size are launch arguments and therefore fixed per captured graph (one graph per camera resolution)."""

    def __init__(self, params: dict, sh_degree: int, settings, cfg: AdamConfig | None = None, lambda_dssim: float = 0.2,
                 capacity_slack: float = 1.25):
        super().__init__(params, sh_degree, settings.width, settings.height, cfg, lambda_dssim, capacity_slack)
        from . import fastgs as fg
        dev = self.dev
        self.static = {"w2c": torch.zeros((4, 4), device=dev), "campos": torch.zeros((3,), device=dev),
                       "background": torch.zeros((3,), device=dev), "target": torch.zeros((3, self.H, self.W), device=dev)}
        self.settings = fg.FastGSSettings(cam_position=self.static["campos"], active_sh_bases=settings.active_sh_bases,
                                          width=settings.width, height=settings.height, focal_x=settings.focal_x,
                                          focal_y=settings.focal_y, center_x=settings.center_x, center_y=settings.center_y,
                                          near_plane=settings.near_plane, far_plane=settings.far_plane)
        self.be = fg.default_backend()
        self.step = FastGsTrainStep(params, sh_degree, settings.width, settings.height, lambda_dssim, optimizer=None,
                                    backend=self.be)

    def _body(self):
        P, S = self.params, self.static
        self.step.capacity = self.capacity
        loss = self.step(S["w2c"], self.settings, S["target"], S["background"])
        names = [k for k in PARAM_GROUPS if P[k].numel() > 0]
        with torch.no_grad():
            _product_ns().fused_adam_step_dynamic([P[k] for k in names], [P[k].grad for k in names],
                                                  [self.opt.exp_avg[k] for k in names],
                                                  [self.opt.exp_avg_sq[k] for k in names],
                                                  self.dyn[[PARAM_GROUPS.index(k) for k in names]].contiguous()
                                                  if len(names) != len(PARAM_GROUPS) else self.dyn,
                                                  self.opt.cfg.beta1, self.opt.cfg.beta2, self.opt.cfg.eps)
        return loss, self.step.last["stats"], self.step.last["n_instances"]

    def _set_inputs(self, w2c, cam_position, target, background):
        S = self.static
        S["w2c"].copy_(w2c.reshape(4, 4), non_blocking=True)
        S["campos"].copy_(cam_position.reshape(3), non_blocking=True)
        S["target"].copy_(target.reshape(3, self.H, self.W), non_blocking=True)
        if background is not None:
            S["background"].copy_(background.reshape(3), non_blocking=True)

    def capture(self, w2c, cam_position, target, background=None):
        self._set_inputs(w2c, cam_position, target, background)
        self.step.capacity = 0
        n = self.step.size_capacity(self.static["w2c"], self.settings, self.slack)
        self.capacity = self.step.capacity
        self.dyn.zero_()  # warm-up iterations must not move the parameters: every group disabled
        side = torch.cuda.Stream(self.dev)
        side.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(side):
            for _ in range(3):
                self._body()
        torch.cuda.current_stream(self.dev).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss, self.stats, self.n_isects = self._body()
        return n

    def __call__(self, w2c, cam_position, target, background=None):
        if self.graph is None:
            self.capture(w2c, cam_position, target, background)
        self.iteration += 1
        self._set_inputs(w2c, cam_position, target, background)
        self._set_scalars()
        self.graph.replay()
        return self.loss
