"""A/B of the blend-backward experiment variants (gsb_debug_bwd_variant) on config B / D: kernel time from the library's
event profile, gradients compared against variant 0."""
import ctypes, os, sys, json
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np, torch
import __graft_entry__ as ge
import scenes
pkg = ge.load_package(); pkg.load()
cabi = ctypes.CDLL(pkg.CABI_PATH)
dev = torch.device("cuda:0")
names = ("means", "quats", "scales", "opacities", "sh_coeffs")
out = {}
for cfg, N in (("B", 1_000_000), ("D", 6_000_000)):
    sc = scenes.scene_b(N=N)
    P = {k: torch.from_numpy(sc[k]).to(dev).requires_grad_(k in names) for k in names + ("viewmats", "Ks", "background")}
    tgt = torch.rand((1, 1080, 1920, 3), device=dev)
    def step():
        for k in names: P[k].grad = None
        o = pkg.rasterize(P["means"], P["quats"], P["scales"], P["opacities"], P["sh_coeffs"], 3, P["viewmats"], P["Ks"], 1920, 1080, bg_color=P["background"])
        ((o.render_colors - tgt).abs().mean()).backward()
    ref = None
    for v in (0, 1, 2, 3, 4, 0):
        cabi.gsb_debug_bwd_variant(ctypes.c_int(v))
        for _ in range(3): step()
        torch.cuda.synchronize()
        cabi.gsb_profile_enable(1)
        for _ in range(20): step()
        torch.cuda.synchronize()
        res = {}
        for kn in ("raster_bwd", "raster_fwd"):
            tot = ctypes.c_double(0.0)
            n = cabi.gsb_profile_read(kn.encode(), ctypes.byref(tot))
            res[kn] = tot.value / max(n, 1)
        cabi.gsb_profile_enable(0)
        g = {k: P[k].grad.detach().clone() for k in names}
        if ref is None: ref = g
        err = max(float((g[k] - ref[k]).norm() / ref[k].norm()) for k in names)
        print(cfg, "variant", v, {k: round(x, 4) for k, x in res.items()}, "max rel diff vs variant 0: %.2e" % err, flush=True)
        out[f"{cfg}_{v}"] = res
    del P
    torch.cuda.empty_cache()
json.dump(out, open("gpurun_out/ab_bwd.json", "w"))
