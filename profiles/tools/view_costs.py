import os, sys, json
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np, torch
import __graft_entry__ as ge
import scenes
pkg = ge.load_package(); pkg.load()
dev = torch.device("cuda:0")
names = ("means", "quats", "scales", "opacities", "sh_coeffs")
for view in range(8):
    sc = scenes.scene_b(N=1_000_000, view=view)
    P = {k: torch.from_numpy(sc[k]).to(dev).requires_grad_(k in names) for k in names + ("viewmats", "Ks", "background")}
    tgt = torch.rand((1, 1080, 1920, 3), device=dev)
    def step():
        for k in names: P[k].grad = None
        o = pkg.rasterize(P["means"], P["quats"], P["scales"], P["opacities"], P["sh_coeffs"], 3, P["viewmats"], P["Ks"], 1920, 1080, bg_color=P["background"])
        ((o.render_colors - tgt).abs().mean()).backward()
        return o
    for _ in range(3): o = step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): o = step()
    e1.record(); torch.cuda.synchronize()
    print("view", view, "ms/step %.3f" % (e0.elapsed_time(e1) / 20), "isects", o.n_isects, "visible", int(o.visibility.sum()), flush=True)
