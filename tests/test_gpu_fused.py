"""GPU: the extended operator over the raw SplatData tensors (include/gsplat/FusedOps.h) against
the unfused sequence the reference runs -- torch activations (splat_data.cpp:267-286) + the eleven gsplat::
operators + autograd -- on the same raw tensors, and against the CPU oracle."""
import numpy as np
import pytest
import torch

import scenes
from parity import assert_grad_close, rel, to_dev

pytestmark = pytest.mark.gpu
RAW = ("means", "sh0", "shN", "scaling_raw", "rotation_raw", "opacity_raw")


def raw_leaves(native, t, quat_scale=None, requires_grad=True):
    raw = native.raw_from_activated(t["means"], t["quats"], t["scales"], t["opacities"], t["sh_coeffs"])
    if quat_scale is not None:  # the stored quaternions are NOT unit length in training
        raw["rotation_raw"] = raw["rotation_raw"] * quat_scale
    return {k: v.detach().clone().requires_grad_(requires_grad) for k, v in raw.items()}


def run_pair(native, sc, dev, deg, seed=0, quat_scale=True, capacity=0, **camera):
    t = to_dev(sc, dev)
    W, H = sc["width"], sc["height"]
    g = torch.Generator(device=dev).manual_seed(seed)
    qs = (torch.rand((t["means"].shape[0], 1), device=dev, generator=g) * 1.5 + 0.5) if quat_scale else None
    vr = torch.randn((1, H, W, 3), device=dev, generator=g)
    va = torch.randn((1, H, W, 1), device=dev, generator=g)
    out = {}
    for tag in ("unfused", "fused"):
        P = raw_leaves(native, t, qs)
        if tag == "fused":
            o = native.rasterize_fused(P["means"], P["sh0"], P["shN"], P["scaling_raw"], P["rotation_raw"], P["opacity_raw"],
                                       deg, t["viewmats"], t["Ks"], W, H, bg_color=t.get("background"),
                                       isect_capacity=capacity, **camera)
            img, alpha, radii, n = o.render_colors, o.alpha, o.radii, int(o.n_isects.item())
        else:
            o = native.rasterize_from_raw(P, deg, t["viewmats"], t["Ks"], W, H, bg_color=t.get("background"), **camera)
            img, alpha, n = o.render_colors, o.alpha.permute(1, 2, 0)[None], o.n_isects
            radii = None
        ((img * vr).sum() + (alpha * va).sum()).backward()
        out[tag] = dict(img=img.detach(), alpha=alpha.detach(), n=n, radii=radii, out=o,
                        grads={k: (P[k].grad if P[k].grad is not None else torch.zeros_like(P[k])) for k in RAW})
    return out


def compare(out, n_gauss, tag, rel_tol=2e-4):
    # an activation that differs from torch's in the last bit can flip a radius at a ceil() boundary: a handful of
    # intersections out of millions (the alpha there is below 1/255, so images and gradients do not notice)
    nf, nu = out["fused"]["n"], out["unfused"]["n"]
    assert abs(nf - nu) <= max(8, int(2e-5 * nu)), (nf, nu)
    e = rel(out["fused"]["img"], out["unfused"]["img"])
    ea = rel(out["fused"]["alpha"], out["unfused"]["alpha"])
    print(f"[{tag}] fused vs unfused: image rel_l2 {e:.2e}, alpha {ea:.2e}, intersections {out['fused']['n']}")
    assert e < 1e-5 and ea < 1e-5
    for k in RAW:
        assert_grad_close(out["fused"]["grads"][k], out["unfused"]["grads"][k], k, n_gauss, rel_tol=rel_tol,
                          tag=f"{tag} fused vs unfused")


@pytest.mark.parametrize("name", ["small_rot", "b30k", "a_deg0"])
def test_fused_equals_unfused_sequence(native, cuda_device, name):
    sc = {"small_rot": lambda: scenes.scene_small(N=3000, width=200, height=120, view=1),
          "b30k": lambda: scenes.scene_b(N=30000, width=640, height=360, view=3, scale_mul=2.0),
          "a_deg0": lambda: scenes.scene_a(background=False)}[name]()
    out = run_pair(native, sc, cuda_device, sc["sh_degree"])
    # same projection arithmetic as the a1 operator: the radii agree exactly unless an activation differs from torch's
    # in the last bit at a ceil() boundary
    o = out["unfused"]["out"]
    mism = int((out["fused"]["radii"][0].max(-1).values != o.radii).sum())
    assert mism <= 2, mism
    compare(out, sc["means"].shape[0], name)


@pytest.mark.parametrize("K,deg", [(16, 3), (16, 2), (16, 1), (16, 0), (4, 1), (9, 2), (25, 4)])
def test_fused_sh_layouts_and_degrees(native, cuda_device, K, deg):
    sc = scenes.scene_small(N=1500, width=160, height=96, view=2)
    rng = np.random.default_rng(K * 10 + deg)
    sc["sh_coeffs"] = ((rng.random((1500, K, 3), dtype=np.float32) - 0.5) * 0.6).astype(np.float32)
    out = run_pair(native, sc, cuda_device, deg)
    compare(out, 1500, f"K={K} deg={deg}")
    if deg < 3 and K > (deg + 1) ** 2:  # inactive degrees get exact zeros
        g = out["fused"]["grads"]["shN"]
        assert float(g[:, (deg + 1) ** 2 - 1:].abs().max()) == 0.0


def test_fused_k1_no_shn_and_ragged_count(native, cuda_device):
    sc = scenes.scene_small(N=1237, width=150, height=70, sh_degree=0, view=5)  # N not a multiple of the slab, K == 1
    out = run_pair(native, sc, cuda_device, 0)
    compare(out, 1237, "K=1")


def test_fused_capacity_mode_needs_no_readback(native, cuda_device):
    sc = scenes.scene_b(N=20000, width=640, height=360, view=2, scale_mul=2.0)
    exact = run_pair(native, sc, cuda_device, 3)
    n = exact["fused"]["n"]
    roomy = run_pair(native, sc, cuda_device, 3, capacity=int(n * 1.3) + 7)
    assert roomy["fused"]["n"] == n
    assert torch.equal(roomy["fused"]["img"], exact["fused"]["img"])
    for k in RAW:
        assert rel(roomy["fused"]["grads"][k], exact["fused"]["grads"][k]) < 1e-5
    # too small: nothing faults, the count tells the caller to retry
    t = to_dev(sc, cuda_device)
    P = raw_leaves(native, t, None, requires_grad=False)
    o = native.rasterize_fused(P["means"], P["sh0"], P["shN"], P["scaling_raw"], P["rotation_raw"], P["opacity_raw"], 3,
                               t["viewmats"], t["Ks"], sc["width"], sc["height"], bg_color=t["background"],
                               isect_capacity=n // 2)
    torch.cuda.synchronize()
    assert int(o.n_isects.item()) == n and o.capacity == n // 2
    assert torch.isfinite(o.render_colors).all()


@pytest.mark.parametrize("model", ["opencv", "fisheye"])
def test_fused_distorted_cameras(native, cuda_device, model):
    from test_gpu_parity import DISTORTED
    cfg = DISTORTED[model]
    sc = scenes.scene_small(N=2500, width=208, height=128, view=2)
    dev = lambda a: None if a is None else torch.from_numpy(a).to(cuda_device)
    out = run_pair(native, sc, cuda_device, 3, camera_model=cfg["camera_model"], radial_coeffs=dev(cfg["radial"]),
                   tangential_coeffs=dev(cfg["tangential"]), thin_prism_coeffs=dev(cfg["thin_prism"]))
    # distorted cameras: the per-pixel Newton undistortion amplifies last-bit differences of the activations, and the
    # gradient sums are order-dependent (REDs): the standard gradient tolerance applies, not the tight fused one
    compare(out, 2500, model, rel_tol=1e-3)


def test_fused_vs_cpu_oracle(native, orc, cuda_device):
    """Independent of the unfused CUDA path: the CPU oracle pipeline on the activated parameters, chained to the raw
    parameters by hand (exp / sigmoid / normalise derivatives in float64)."""
    sc = scenes.scene_small(N=3000, width=200, height=120, view=1)
    W, H = sc["width"], sc["height"]
    rng = np.random.default_rng(4)
    vrc = rng.standard_normal((1, H, W, 3)).astype(np.float32)
    vra = rng.standard_normal((1, H, W, 1)).astype(np.float32)
    ref = orc.render_pipeline(sc, "f32", True, vrc, vra)
    t = to_dev(sc, cuda_device)
    P = raw_leaves(native, t, None)
    o = native.rasterize_fused(P["means"], P["sh0"], P["shN"], P["scaling_raw"], P["rotation_raw"], P["opacity_raw"], 3,
                               t["viewmats"], t["Ks"], W, H, bg_color=t["background"])
    assert abs(int(o.n_isects.item()) - len(ref["flatten_ids"])) <= 8
    ((o.render_colors * torch.from_numpy(vrc).to(cuda_device)).sum() +
     (o.alpha * torch.from_numpy(vra).to(cuda_device)).sum()).backward()
    assert rel(o.render_colors.detach().cpu().numpy(), ref["renders"]) < 1e-4
    op = sc["opacities"].astype(np.float64)
    want = {"means": ref["v_means"] + ref["v_dirs"][0], "sh0": ref["v_sh_coeffs"][:, :1], "shN": ref["v_sh_coeffs"][:, 1:],
            "scaling_raw": ref["v_scales"] * sc["scales"], "rotation_raw": ref["v_quats"],
            "opacity_raw": (ref["v_opacities"][0] * op * (1 - op))[:, None]}
    for k, w in want.items():
        assert_grad_close(P[k].grad, w, k, 3000, tag="fused vs CPU oracle")
