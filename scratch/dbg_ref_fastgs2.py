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
which = sys.argv[1] if len(sys.argv) > 1 else "sweep"
sizes = [(640, 368), (1280, 720), (1920, 1080)] if which == "sweep" else [(1920, 1080)]
for (W, H) in sizes:
    N = 100_000
    sc = scenes.scene_b(N=N, width=W, height=H)
    sc["Ks"][0, 0, 0] = sc["Ks"][0, 1, 1] = 1600.0 * W / 1920
    inp = scenes.fastgs_inputs(sc)
    P = {k: torch.from_numpy(inp[k]).to(dev) for k in ("means", "scales_raw", "rotations_raw", "opacities_raw", "sh0", "shN")}
    w2c = torch.from_numpy(inp["w2c"]).to(dev)
    s = fg.FastGSSettings(cam_position=torch.from_numpy(inp["cam_position"]).to(dev), active_sh_bases=16, width=W, height=H,
                          focal_x=inp["fx"], focal_y=inp["fy"], center_x=inp["cx"], center_y=inp["cy"])
    for name, be in ((("mine", mb),) if which == "sweep" else ()) + (("ref", rb),):
        try:
            img, al, c = be.forward(P["means"], P["scales_raw"], P["rotations_raw"], P["opacities_raw"], P["sh0"], P["shN"], w2c, s)
            torch.cuda.synchronize()
            print(W, H, name, "ints", c["ints"].tolist(), "img mean", float(img.mean()), flush=True)
        except Exception as e:
            print(W, H, name, "FAILED", repr(e)[:200], flush=True)
