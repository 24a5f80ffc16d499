"""Profiling driver: a few forward + backward passes of the fastgs (EWA) path on config B or D, so
that ncu can capture fgs_front_kernel / fgs_back_kernel / the EWA instantiations of the blend kernels / the filtered
intersect walks in their real context.
    ncu --set full -k regex:fgs_front -s 2 -c 1 -o gpurun_out/prof_fgs_front python profiles/tools/fastgs_prof.py
Prints the per-kernel averages of the library's own event profile when run without a profiler; --reference runs the
reference's own fastgs kernels (oracle/_ref/libfastgs_ref.so) instead."""
import argparse
import ctypes
import importlib
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
    ap.add_argument("--reference", action="store_true")
    args = ap.parse_args()
    import torch

    import __graft_entry__ as ge
    import bench
    import scenes
    pkg = ge.load_package()
    pkg.load()
    fg = importlib.import_module(pkg.__name__ + ".fastgs")
    dev = torch.device("cuda", 0)
    inp = scenes.fastgs_inputs(bench.make_scene(args.config))
    names = ("means", "scales_raw", "rotations_raw", "opacities_raw", "sh0", "shN")
    P = {k: torch.from_numpy(inp[k]).to(dev).requires_grad_(True) for k in names}
    w2c = torch.from_numpy(inp["w2c"]).to(dev)
    st = fg.FastGSSettings(cam_position=torch.from_numpy(inp["cam_position"]).to(dev), active_sh_bases=inp["active_sh_bases"],
                           width=inp["width"], height=inp["height"], focal_x=inp["fx"], focal_y=inp["fy"], center_x=inp["cx"],
                           center_y=inp["cy"])
    tgt = torch.rand((3, inp["height"], inp["width"]), device=dev)
    if args.reference:
        from oracle import ref_fastgs
        be = ref_fastgs.backend(fg)
    else:
        be = fg.default_backend()

    def step():
        for k in names:
            P[k].grad = None
        img, _ = fg.fast_rasterize(be, P["means"], P["scales_raw"], P["rotations_raw"], P["opacities_raw"], P["sh0"], P["shN"],
                                   w2c, st)
        loss = (img - tgt).abs().mean()
        loss.backward()
        return loss
    cabi = ctypes.CDLL(pkg.CABI_PATH)
    cabi.gsb_profile_read.argtypes = [ctypes.c_char_p, ctypes.POINTER(ctypes.c_double)]
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    cabi.gsb_profile_enable(1)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.iters):
        loss = step()
    ev1.record()
    torch.cuda.synchronize()
    prof = {}
    for k in bench.KERNEL_NAMES:
        tot = ctypes.c_double(0.0)
        n = cabi.gsb_profile_read(k.encode(), ctypes.byref(tot))
        if n:
            prof[k] = round(tot.value / n, 4)
    print(json.dumps({"loss": float(loss.item()), "ms_per_step": ev0.elapsed_time(ev1) / args.iters, "kernels_ms": prof}))


if __name__ == "__main__":
    main()
