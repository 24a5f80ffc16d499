"""A/B of the blend-forward loop (gsb_debug_fwd_variant: 0 = two records per trip, 1 = one) on config B / D, GUT and EWA."""
import ctypes, os, sys, json, importlib
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np, torch
import __graft_entry__ as ge
import scenes
pkg = ge.load_package(); pkg.load()
fg = importlib.import_module(pkg.__name__ + ".fastgs")
cabi = ctypes.CDLL(pkg.CABI_PATH)
dev = torch.device("cuda:0")
names = ("means", "quats", "scales", "opacities", "sh_coeffs")
def prof(fn, scopes, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); cabi.gsb_profile_enable(1)
    for _ in range(iters): fn()
    torch.cuda.synchronize()
    res = {}
    for kn in scopes:
        tot = ctypes.c_double(0.0); n = cabi.gsb_profile_read(kn.encode(), ctypes.byref(tot)); res[kn] = round(tot.value / max(n, 1), 4)
    cabi.gsb_profile_enable(0)
    return res
for cfg, N in (("B", 1_000_000), ("D", 6_000_000)):
    sc = scenes.scene_b(N=N)
    P = {k: torch.from_numpy(sc[k]).to(dev).requires_grad_(k in names) for k in names + ("viewmats", "Ks", "background")}
    tgt = torch.rand((1, 1080, 1920, 3), device=dev)
    img = {}
    def step():
        for k in names: P[k].grad = None
        o = pkg.rasterize(P["means"], P["quats"], P["scales"], P["opacities"], P["sh_coeffs"], 3, P["viewmats"], P["Ks"], 1920, 1080, bg_color=P["background"])
        ((o.render_colors - tgt).abs().mean()).backward(); img["x"] = o.render_colors.detach()
    inp = scenes.fastgs_inputs(sc)
    F = {k: torch.from_numpy(inp[k]).to(dev).requires_grad_(True) for k in ("means", "scales_raw", "rotations_raw", "opacities_raw", "sh0", "shN")}
    w2c = torch.from_numpy(inp["w2c"]).to(dev)
    st = fg.FastGSSettings(cam_position=torch.from_numpy(inp["cam_position"]).to(dev), active_sh_bases=16, width=1920, height=1080, focal_x=inp["fx"], focal_y=inp["fy"], center_x=inp["cx"], center_y=inp["cy"])
    be = fg.default_backend()
    def fstep():
        for k in F: F[k].grad = None
        im, _ = fg.fast_rasterize(be, F["means"], F["scales_raw"], F["rotations_raw"], F["opacities_raw"], F["sh0"], F["shN"], w2c, st)
        (im - tgt[0].permute(2, 0, 1)).abs().mean().backward(); img["f"] = im.detach()
    ref = {}
    for v in (0, 1, 0, 1):
        cabi.gsb_debug_fwd_variant(ctypes.c_int(v))
        r = prof(step, ("raster_fwd", "raster_bwd")); r.update(prof(fstep, ("ewa_blend_fwd", "ewa_blend_bwd")))
        if v not in ref: ref[v] = (img["x"].clone(), img["f"].clone())
        print(cfg, "variant", v, r, "gut equal:", bool(torch.equal(img["x"], ref[0][0])), "ewa equal:", bool(torch.equal(img["f"], ref[0][1])), flush=True)
    del P, F
    torch.cuda.empty_cache()
