import sys, os, importlib
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np, torch
import __graft_entry__ as ge
pkg = ge.load_package(); pkg.load()
fg = importlib.import_module(pkg.__name__ + ".fastgs")
from oracle import ref_fastgs
import scenes
dev = torch.device("cuda:0")
rb = ref_fastgs.backend(fg); mb = fg.default_backend()
for N in (100_000, 300_000, 1_000_000):
    sc = scenes.scene_b(N=N)
    inp = scenes.fastgs_inputs(sc)
    P = {k: torch.from_numpy(inp[k]).to(dev) for k in ("means", "scales_raw", "rotations_raw", "opacities_raw", "sh0", "shN")}
    w2c = torch.from_numpy(inp["w2c"]).to(dev)
    s = fg.FastGSSettings(cam_position=torch.from_numpy(inp["cam_position"]).to(dev), active_sh_bases=16, width=1920, height=1080,
                          focal_x=inp["fx"], focal_y=inp["fy"], center_x=inp["cx"], center_y=inp["cy"])
    for name, be in (("mine", mb), ("ref", rb)):
        try:
            img, al, c = be.forward(P["means"], P["scales_raw"], P["rotations_raw"], P["opacities_raw"], P["sh0"], P["shN"], w2c, s)
            torch.cuda.synchronize()
            print(N, name, "ints", c["ints"].tolist(), "img mean", float(img.mean()), flush=True)
        except Exception as e:
            print(N, name, "FAILED", repr(e)[:300], flush=True)
