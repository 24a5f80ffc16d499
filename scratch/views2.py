import os, sys, json, ctypes
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np, torch
import __graft_entry__ as ge
import scenes, bench
pkg = ge.load_package(); pkg.load()
cabi = ctypes.CDLL(pkg.CABI_PATH)
cabi.gsb_profile_read.argtypes = [ctypes.c_char_p, ctypes.POINTER(ctypes.c_double)]
dev = torch.device("cuda:0")
names = ("means", "quats", "scales", "opacities", "sh_coeffs")
for view in (0, 1):
    sc = scenes.scene_b(N=1_000_000, view=view)
    P = {k: torch.from_numpy(sc[k]).to(dev).requires_grad_(k in names) for k in names + ("viewmats", "Ks", "background")}
    tgt = torch.rand((1, 1080, 1920, 3), device=dev)
    def step():
        for k in names: P[k].grad = None
        o = pkg.rasterize(P["means"], P["quats"], P["scales"], P["opacities"], P["sh_coeffs"], 3, P["viewmats"], P["Ks"], 1920, 1080, bg_color=P["background"])
        ((o.render_colors - tgt).abs().mean()).backward()
        return o
    for _ in range(3): o = step()
    torch.cuda.synchronize(); cabi.gsb_profile_enable(1)
    for _ in range(10): o = step()
    torch.cuda.synchronize()
    prof = {}
    for k in bench.KERNEL_NAMES:
        tot = ctypes.c_double(0.0); n = cabi.gsb_profile_read(k.encode(), ctypes.byref(tot))
        if n: prof[k] = round(tot.value / n, 4)
    cabi.gsb_profile_enable(0)
    # tile list statistics
    radii, means2d, depths, _, _ = pkg.projection_ut_3dgs_fused(P["means"].detach(), P["quats"].detach(), P["scales"].detach(), P["opacities"].detach(), P["viewmats"], P["Ks"], 1920, 1080, 0.3, 0.01, 1e4, 0.0)
    tpg, ids, flat = pkg.intersect_tile(means2d, radii, depths, 1, 16, 120, 68, True)
    off = pkg.intersect_offset(ids, 1, 120, 68).reshape(-1).long()
    cnt = torch.diff(torch.cat([off, torch.tensor([ids.shape[0]], device=dev)]))
    r = radii[0].float()
    print("view", view, prof, "tile list max/mean %d/%.0f" % (int(cnt.max()), float(cnt.float().mean())), "radius mean/max %.1f/%d" % (float(r[r>0].mean()), int(r.max())), "tiles per gauss max", int(tpg.max()), flush=True)
