"""Profiling driver: a few training iterations (extended rasterizer + fused loss + fused Adam) on config B or D, so
that ncu can capture ssim_l1_kernel / fused_front_kernel / fused_back_kernel / adam kernels in their real context.
    ncu --set full -k regex:ssim_l1 -s 2 -c 1 -o gpurun_out/prof_ssim python profiles/tools/train_prof.py
Prints the per-kernel averages of the library's own event profile when run without a profiler."""
import argparse
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="B")
    ap.add_argument("--iters", type=int, default=4)
    args = ap.parse_args()
    import numpy as np
    import torch

    import __graft_entry__ as ge
    import bench
    pkg = ge.load_package()
    pkg.load()
    from gsplat_b200 import training
    dev = torch.device("cuda", 0)
    sc = bench.make_scene(args.config)
    W, H, deg = sc["width"], sc["height"], sc["sh_degree"]
    act = {k: torch.from_numpy(sc[k]).to(dev) for k in ("means", "quats", "scales", "opacities", "sh_coeffs")}
    raw = pkg.raw_from_activated(act["means"], act["quats"], act["scales"], act["opacities"], act["sh_coeffs"])
    P = {k: v.detach().clone().requires_grad_(True) for k, v in raw.items()}
    viewmats, Ks = torch.from_numpy(sc["viewmats"]).to(dev), torch.from_numpy(sc["Ks"]).to(dev)
    tgt = torch.from_numpy(np.random.default_rng(5).random((3, H, W), dtype=np.float32)).to(dev)
    bg = torch.zeros(1, 3, device=dev)
    ts = training.TrainStep(P, deg, W, H, optimizer=training.FusedAdam(P))
    ts.size_capacity(viewmats, Ks, tgt, bg)
    cabi = ctypes.CDLL(pkg.CABI_PATH)
    cabi.gsb_profile_read.argtypes = [ctypes.c_char_p, ctypes.POINTER(ctypes.c_double)]
    for _ in range(2):
        ts(viewmats, Ks, tgt, bg)
    torch.cuda.synchronize()
    cabi.gsb_profile_enable(1)
    for _ in range(args.iters):
        loss = ts(viewmats, Ks, tgt, bg)
    torch.cuda.synchronize()
    prof = {}
    for k in bench.KERNEL_NAMES:
        tot = ctypes.c_double(0.0)
        n = cabi.gsb_profile_read(k.encode(), ctypes.byref(tot))
        if n:
            prof[k] = round(tot.value / n, 4)
    print(json.dumps({"loss": float(loss.item()), "kernels_ms": prof}))


if __name__ == "__main__":
    main()
