#!/usr/bin/env python
"""bench.py -- Gaussians rasterized/sec (fwd+bwd) on BASELINE.json configs[1] (and configs[4] for N>1).

One "step" = one pass of the hot path over one view of the synthetic scene:
  projection_ut -> SH fwd -> intersect_tile/offset -> blend fwd -> L1 loss grad -> blend bwd -> SH bwd
through the reference-facing API (the gsplat:: shim, driven by the L3 mirror in the package),
exactly the call sequence of gs::training::rasterize + loss.backward().

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

N>1 (launched by torchrun, one rank per GPU): every rank renders its own view of the same 1 M
Gaussians (config E) and the per-Gaussian gradients are summed with one NCCL all-reduce group per
step; weak scaling, value = N_gaussians x n_gpus / max-over-ranks time.

--impl reference times the reference's CPU path of the same hot path: the oracle port
(oracle/gut_oracle.c, OpenMP over all host cores) on a bounded sample (a sub-frustum crop of the
same scene, same Gaussians-per-pixel density); rank 0 only.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "gaussians_rasterized_per_sec_fwd_bwd"
UNIT = "Gaussians/s"


# ---------------------------------------------------------------------------------------------
# helpers
# ---------------------------------------------------------------------------------------------
def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def crop_scene(sc, x0, y0, w, h, margin=0.12):
    """Sub-frustum sample of a scene: the Gaussians whose centres project into the window
    (plus a margin) and a camera whose principal point is shifted to the window."""
    V = sc["viewmats"][0].astype(np.float64)
    K = sc["Ks"][0].astype(np.float64)
    pc = sc["means"].astype(np.float64) @ V[:3, :3].T + V[:3, 3]
    u = pc[:, 0] / pc[:, 2] * K[0, 0] + K[0, 2]
    v = pc[:, 1] / pc[:, 2] * K[1, 1] + K[1, 2]
    keep = (pc[:, 2] > 0.01) & (u > x0 - margin * w) & (u < x0 + w * (1 + margin)) & \
           (v > y0 - margin * h) & (v < y0 + h * (1 + margin))
    out = dict(sc)
    for k in ("means", "quats", "scales", "opacities", "sh_coeffs"):
        out[k] = np.ascontiguousarray(sc[k][keep])
    Kc = sc["Ks"].copy()
    Kc[0, 0, 2] -= x0
    Kc[0, 1, 2] -= y0
    out["Ks"] = Kc
    out["width"], out["height"] = w, h
    return out


def oracle_step(orc, sc, target):
    """One fwd+bwd pass of the hot path on the CPU oracle with an L1 loss against `target`."""
    o = orc.render_pipeline(sc, precision="f32")
    diff = o["renders"] - target
    v_rc = (np.sign(diff) / diff.size).astype(np.float32)
    v_ra = np.zeros(o["alphas"].shape, np.float32)
    orc.backward_pipeline(sc, o, v_rc, v_ra, precision="f32")
    return float(np.abs(diff).mean()), o


def cpu_sample_scene(n_gauss):
    import scenes
    sc = scenes.scene_b(N=n_gauss)
    # central 480x272 window of the 1920x1080 frame: 1/16 of the pixels and ~1/16 of the Gaussians
    w, h = 480, 272
    x0, y0 = (sc["width"] - w) // 2, (sc["height"] - h) // 2
    return crop_scene(sc, x0, y0, w, h), f"central {w}x{h} window of the {sc['width']}x{sc['height']} frame"


def time_oracle(steps, warmup, n_gauss, threads=None):
    from oracle import oracle as orc
    orc.build()
    cores = threads or orc.max_threads()
    orc.set_threads(cores)
    sc, what = cpu_sample_scene(n_gauss)
    target = np.full((1, sc["height"], sc["width"], 3), 0.5, np.float32)
    for _ in range(warmup):
        oracle_step(orc, sc, target)
    t0 = time.perf_counter()
    for _ in range(steps):
        oracle_step(orc, sc, target)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    orc.set_threads(1)
    ns = sc["means"].shape[0]
    return {"value": ns / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{what}: {ns} of the {n_gauss} Gaussians (sub-frustum, same per-pixel density), "
                      f"fwd+bwd, {dt * 1e3:.1f} ms/step, oracle/gut_oracle.c with OpenMP"}, dt


# ---------------------------------------------------------------------------------------------
# reference arm
# ---------------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cb, dt = time_oracle(args.steps, args.warmup, args.gaussians)
    line = {
        "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "1M synthetic Gaussians, 1920x1080, SH deg 3, 16x16 tiles, fwd+bwd (configs[1]); "
                               "CPU sample: " + cb["sample"]},
        "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------
# (kernel, gaussians, width, height) -> DRAM bytes per launch measured by ncu (profiles/r1c_ncu_summary.md)
NCU_DRAM_BYTES = {("raster_bwd", 1_000_000, 1920, 1080): 89_998_592 + 4_392_448,
                  ("raster_fwd", 1_000_000, 1920, 1080): 50_757_632 + 9_520_640}


def run_b200(args):
    import torch
    import torch.distributed as dist

    import __graft_entry__ as ge
    import scenes

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the B200 backend has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    pkg = ge.load_package()
    pkg.load()
    from gsplat_b200 import multiview
    cabi = ctypes.CDLL(pkg.CABI_PATH)
    cabi.gsb_launch_count.restype = ctypes.c_uint64
    cabi.gsb_profile_read.argtypes = [ctypes.c_char_p, ctypes.POINTER(ctypes.c_double)]

    N = args.gaussians
    view = rank if world > 1 else None  # config E: one camera of the ring per rank
    sc = scenes.scene_b(N=N, view=view)
    W, H, deg = sc["width"], sc["height"], sc["sh_degree"]
    host = {k: torch.from_numpy(sc[k]).pin_memory() for k in
            ("means", "quats", "scales", "opacities", "sh_coeffs", "viewmats", "Ks", "background")}
    rng = np.random.default_rng(123 + rank)
    host["target"] = torch.from_numpy(rng.random((1, H, W, 3), dtype=np.float32)).pin_memory()
    names = ("means", "quats", "scales", "opacities", "sh_coeffs")
    P = {k: host[k].to(dev).requires_grad_(k in names) for k in host}
    host_out = {k: torch.empty_like(host[k]).pin_memory() for k in names}
    host_img = torch.empty((1, H, W, 3), dtype=torch.float32).pin_memory()

    stats = {}

    compact = world > 1 and args.exchange == "compact"

    def step(Pd, backend=None):
        for k in names:
            Pd[k].grad = None
        deferred = pkg.DeferredSHBackward() if compact else None
        out = pkg.rasterize(Pd["means"], Pd["quats"], Pd["scales"], Pd["opacities"], Pd["sh_coeffs"], deg,
                            Pd["viewmats"], Pd["Ks"], W, H, bg_color=Pd["background"], backend=backend,
                            sh_exchange=deferred)
        loss = (out.render_colors - Pd["target"]).abs().mean()
        loss.backward()
        if compact:
            # all-gather of the 12-byte colour gradients + all-reduce of the 44 B of geometry gradients + local
            # expansion of all views (gsb_sh_bwd_views): the same sums as the all-reduce below, 1/2-1/3 of the bytes
            multiview.exchange_gradients_compact(Pd, deferred)
        elif world > 1:
            # one fused NCCL launch for the five gradient tensors (236 B/Gaussian)
            multiview.allreduce_gradients([Pd[k].grad for k in names])
        stats["n_isects"], stats["vis"] = out.n_isects, out.visibility
        return loss, out

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        sync_all()
        return multiview.max_over_ranks(e0.elapsed_time(e1), dev)

    # ---- warm-up + device-resident timing -------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        step(P)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = cabi.gsb_launch_count()
    ms_total = timed(lambda: step(P), args.steps)
    launches = int(cabi.gsb_launch_count() - launches0)
    # per-kernel CUDA-event timing in a SEPARATE pass (its event records are not in the timed region)
    cabi.gsb_profile_enable(1)
    timed(lambda: step(P), args.steps)
    prof = {}
    for kname in ("projection_ut", "sh_fwd", "isect_count", "isect_depth_sort", "isect_tile_hist", "isect_emit", "isect_sort",
                  "isect_offsets", "raster_prep",
                  "raster_fwd", "raster_bwd", "raster_finalize", "sh_bwd", "sh_bwd_views"):
        tot = ctypes.c_double(0.0)
        n = cabi.gsb_profile_read(kname.encode(), ctypes.byref(tot))
        if n:
            prof[kname] = {"launches": n, "avg_ms": tot.value / n}
    cabi.gsb_profile_enable(0)
    ms_step = ms_total / args.steps
    value = N * world / (ms_step * 1e-3)

    # ---- end-to-end: every op input from pinned host memory, image + gradients back to the host ----
    def step_e2e():
        Pd = {k: host[k].to(dev, non_blocking=True).requires_grad_(k in names) for k in host}
        loss, out = step(Pd)
        host_img.copy_(out.render_colors.detach(), non_blocking=True)
        for k in names:
            host_out[k].copy_(Pd[k].grad, non_blocking=True)
        return float(loss.item())  # D2H read of the step's result (also drains the copies)

    for _ in range(2):
        step_e2e()
    ms_e2e_seq = timed(step_e2e, args.steps) / args.steps
    h2d = sum(host[k].numel() * host[k].element_size() for k in host)
    d2h = host_img.numel() * 4 + sum(host_out[k].numel() * 4 for k in names) + 4

    # The same traffic through the package's host-staged pipeline (hoststream.py): H2D of step k+1, kernels
    # of step k and D2H of step k-1 overlap on three streams.  Every step still moves all of its inputs from
    # pinned host memory and all of its results back, inside the timed region.
    from gsplat_b200 import hoststream

    def staged_step(Pd):
        loss, out = step(Pd)
        return loss, out.render_colors
    staged = hoststream.HostStagedSteps(dev, host, names, staged_step)
    staged.run(3)
    e2e_losses = []
    ms_e2e = timed(lambda: e2e_losses.extend(staged.run(args.steps)), 1) / args.steps
    assert staged.h2d_bytes == h2d and staged.d2h_bytes == d2h
    clocks = sampler.stop() if rank == 0 else None

    # ---- resident-parameter end-to-end (what a training step moves: camera + target in, loss out) ----
    def step_e2e_resident():
        for k in ("viewmats", "Ks", "background", "target"):
            P[k] = host[k].to(dev, non_blocking=True)
        loss, _ = step(P)
        return float(loss.item())

    step_e2e_resident()
    ms_e2e_res = timed(step_e2e_resident, args.steps) / args.steps

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel (blend backward), SURVEY.md 8(d) algorithmic bytes ---------
    peak, peak_src = load_peaks()
    I, Pn = stats["n_isects"], W * H
    dom = max(prof, key=lambda k: prof[k]["avg_ms"]) if prof else None
    algo = {"raster_bwd": 60 * I + 24 * Pn + 112 * N, "raster_fwd": 48 * I + 20 * Pn}
    roof = None
    if dom in algo:
        achieved = algo[dom] / (prof[dom]["avg_ms"] * 1e-3) / 1e9
        # dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of this kernel at this
        # workload (profiles/r1c_ncu_summary.md); other sizes have no capture
        traffic = NCU_DRAM_BYTES.get((dom, N, W, H))
        roof = {"kernel": dom, "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": algo[dom], "avg_kernel_ms": prof[dom]["avg_ms"],
                "note": "the blend is FP32-issue/LSU/atomic bound by construction (SURVEY.md 8d): the HBM fraction is "
                        "the figure north_star asks for, the pipe utilisations are in profiles/"}

    cb = None
    if world == 1 and not args.no_cpu_baseline:
        cb, _ = time_oracle(1, 1, N)

    # ---- same-box GPU baseline: the reference's own gsplat CUDA kernels (oracle/_ref), same call sites ----
    ref_cuda = None
    if world == 1 and not args.no_ref_cuda:
        try:
            from oracle import ref_ops
            if ref_ops.available():
                rb = ref_ops.backend(pkg)
                for _ in range(3):
                    step(P, rb)
                ms_ref = timed(lambda: step(P, rb), args.steps) / args.steps
                ref_cuda = {"ms_per_step": ms_ref, "value": N / (ms_ref * 1e-3), "unit": UNIT,
                            "what": "reference gsplat/*.cu compiled unmodified (-O3 --use_fast_math, sm_100) by "
                                    "oracle/build_ref.py, driven through the same L3 call sequence on the same inputs",
                            "speedup_device_resident": ms_ref / ms_step}
        except Exception as e:  # the baseline is optional evidence, never a failure of the bench
            ref_cuda = {"unavailable": repr(e)[:200]}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "iters_per_sec": 1e3 / ms_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": ("1M synthetic Gaussians, 1920x1080, SH deg 3, 16x16 tiles, fwd+bwd on 1xB200 "
                                "(BASELINE.json configs[1])" if world == 1 else
                                f"1M Gaussians x {world} synthetic cameras/step, per-view shard + NCCL gradient "
                                f"exchange ({args.exchange}) (BASELINE.json configs[4])"),
                   "gaussians": N, "visible": int(stats["vis"].sum().item()), "intersections": I, "image": [W, H],
                   "loss": "L1 vs synthetic target", "parallelism": f"view-dp{world}",
                   "l2": "per-step working set (236 MB parameters + 64 MB records + ~0.35 GB intersection "
                         "buffers + 236 MB gradients) exceeds the 126 MB L2; no explicit flush"},
        "e2e": {"value": N * world / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "what": "all op inputs (parameters, camera, target) from pinned host memory each step; image + all "
                        "gradients + loss back to pinned host; copies of neighbouring steps overlap the kernels "
                        "(gsplat_b200.hoststream.HostStagedSteps)",
                "ms_per_step_unpipelined": ms_e2e_seq, "loss": e2e_losses[-1] if e2e_losses else None},
        "e2e_resident": {"value": N * world / (ms_e2e_res * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e_res,
                         "what": "parameters resident (as in training); camera + target image H2D, loss D2H per step"},
        "gpu_launches": launches,
        # the quantity the blend's cost is proportional to; rank 0's count x ranks
        "intersections_per_sec": I * world / (ms_step * 1e-3),
        "kernels_ms": {k: round(v["avg_ms"], 4) for k, v in prof.items()},
        "clocks": clocks,
    }
    if roof:
        line["roofline"] = roof
    if cb:
        line["cpu_baseline"] = cb
    if ref_cuda:
        line["reference_cuda"] = ref_cuda
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--exchange", choices=("compact", "allreduce"), default="compact",
                    help="N>1 gradient exchange: compact (all-gather colour gradients + multi-view SH backward) or "
                         "plain all-reduce of the five gradient tensors")
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--gaussians", type=int, default=1_000_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ref-cuda", action="store_true")
    ap.add_argument("--quick", action="store_true", help="device-resident timing only (profiling runs)")
    args = ap.parse_args()
    if args.quick:
        args.no_cpu_baseline = args.no_ref_cuda = True
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
