import os, sys
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np, torch
import __graft_entry__ as ge
import scenes
pkg = ge.load_package(); pkg.load()
dev = torch.device("cuda:0")
view = int(sys.argv[1]) if len(sys.argv) > 1 else 1
sc = scenes.scene_b(N=1_000_000, view=view)
P = {k: torch.from_numpy(sc[k]).to(dev) for k in ("means", "quats", "scales", "opacities", "viewmats", "Ks")}
radii, means2d, depths, _, _ = pkg.projection_ut_3dgs_fused(P["means"], P["quats"], P["scales"], P["opacities"], P["viewmats"], P["Ks"], 1920, 1080, 0.3, 0.01, 1e4, 0.0)
for _ in range(4):
    tpg, ids, flat = pkg.intersect_tile(means2d, radii, depths, 1, 16, 120, 68, True)
torch.cuda.synchronize()
# group-size statistics of the placement: chunks of ~I/888 intersections in depth order
I = ids.shape[0]; Pn = 888
tile = (ids >> 32).long()
vis = (radii[0] > 0).all(-1)
order = torch.argsort(torch.where(vis, depths[0], torch.full_like(depths[0], 1e30)), stable=True)
rank = torch.empty_like(order); rank[order] = torch.arange(order.numel(), device=dev)
cum = torch.cumsum(tpg.reshape(-1)[order], 0)
chunk_of_gauss = torch.empty_like(order); chunk_of_gauss[order] = torch.clamp((cum - 1) * Pn // I, 0, Pn - 1)
grp = chunk_of_gauss[flat.long()] * 8160 + tile
_, counts = torch.unique(grp, return_counts=True)
q = torch.tensor([0.5, 0.9, 0.99, 0.999], device=dev)
print("view", view, "isects", I, "groups", counts.numel(), "mean %.2f" % float(counts.float().mean()), "quantiles", torch.quantile(counts.float(), q).tolist(), "max", int(counts.max()),
      "frac of isects in groups >8: %.3f, >16: %.3f, >128: %.3f" % tuple(float(counts[counts > t].sum()) / I for t in (8, 16, 128)),
      "runs with > 32 tiles: %.4f" % float((tpg > 32).float().mean()))
