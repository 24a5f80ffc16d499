"""Generate tests/golden/torch_impl_*.npz from the reference's own torch oracle.

Runs in the build container only (needs /root/reference).  Compiles
/root/reference/tests/torch_impl.cpp (where it lies) together with torch_impl_binding.cpp into a
scratch .so under /tmp, evaluates it on seeded inputs following the reference's tests
(tests/test_garden_data.cpp:531-569 TileIntersectionTest; tests/test_numerical_gradients.cpp:158-229
SH forward/backward vs autograd of the torch oracle), and stores inputs + outputs as small fixtures.
"""
import os
import subprocess
import sys

import numpy as np
import torch
from torch.utils import cpp_extension as ce

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/tests"
SO = "/tmp/ref_torch_impl.so"


def build():
    if os.path.exists(SO):
        return
    inc = []
    for i in ce.include_paths() + [REF]:
        inc += ["-I", i]
    tl = ce.library_paths()[0]
    cmd = ["/usr/bin/g++", "-O1", "-std=c++20", "-fPIC", "-shared", "-D_GLIBCXX_USE_CXX11_ABI=1",
           os.path.join(REF, "torch_impl.cpp"), os.path.join(HERE, "torch_impl_binding.cpp"), "-o", SO] + inc + \
          ["-L", tl, "-ltorch", "-ltorch_cpu", "-lc10", "-Wl,-rpath," + tl]
    subprocess.run(cmd, check=True)


def main():
    build()
    torch.ops.load_library(SO)
    R = torch.ops.ref_torch_impl

    # ---- TileIntersectionTest (test_garden_data.cpp:531-569): C=3, N=1000, 40x60, tile 16 -------
    torch.manual_seed(42)
    C, N, W, H, ts = 3, 1000, 40, 60, 16
    tw, th = (W + ts - 1) // ts, (H + ts - 1) // ts
    means2d = torch.randn(C, N, 2) * W
    radii = torch.randint(0, W, (C, N, 2), dtype=torch.int32)
    depths = torch.rand(C, N)
    tpg, ids, flat = R.isect_tiles(means2d, radii, depths, ts, tw, th, True)
    np.savez_compressed(os.path.join(HERE, "torch_impl_isect_c3.npz"), means2d=means2d.numpy(), radii=radii.numpy(),
                        depths=depths.numpy(), tile_size=ts, tile_width=tw, tile_height=th,
                        tiles_per_gauss=tpg.numpy().astype(np.int32), isect_ids=ids.numpy(), flatten_ids=flat.numpy())
    # a C=1 case on a 1080p-like tile grid (n_tiles not a power of two) with some negative centres
    torch.manual_seed(7)
    C, N, W, H = 1, 3000, 300, 200
    tw, th = (W + ts - 1) // ts, (H + ts - 1) // ts
    means2d = torch.rand(C, N, 2) * torch.tensor([W * 1.4, H * 1.4]) - torch.tensor([W * 0.2, H * 0.2])
    radii = torch.randint(0, 40, (C, N, 2), dtype=torch.int32)
    depths = torch.rand(C, N) * 5 + 0.1
    depths[0, 10] = depths[0, 11]  # an exact depth tie
    tpg, ids, flat = R.isect_tiles(means2d, radii, depths, ts, tw, th, True)
    np.savez_compressed(os.path.join(HERE, "torch_impl_isect_c1.npz"), means2d=means2d.numpy(), radii=radii.numpy(),
                        depths=depths.numpy(), tile_size=ts, tile_width=tw, tile_height=th,
                        tiles_per_gauss=tpg.numpy().astype(np.int32), isect_ids=ids.numpy(), flatten_ids=flat.numpy())

    # ---- SH forward + autograd backward of the torch oracle, degrees 0..4, K = 25 ----------------
    torch.manual_seed(42)
    N, K = 1000, 25
    out = {}
    dirs = torch.randn(N, 3)
    coeffs = torch.randn(N, K, 3)
    v_colors = torch.randn(N, 3)
    out.update(dirs=dirs.numpy(), coeffs=coeffs.numpy(), v_colors=v_colors.numpy())
    for deg in range(5):
        d = dirs.clone().requires_grad_(True)
        c = coeffs.clone().requires_grad_(True)
        col = R.spherical_harmonics(deg, d, c)
        (col * v_colors).sum().backward()
        out[f"colors_deg{deg}"] = col.detach().numpy()
        out[f"v_coeffs_deg{deg}"] = c.grad.numpy()
        out[f"v_dirs_deg{deg}"] = (d.grad if d.grad is not None else torch.zeros_like(d)).numpy()
    np.savez_compressed(os.path.join(HERE, "torch_impl_sh.npz"), **out)

    # ---- quat -> rotmat ------------------------------------------------------------------------------
    torch.manual_seed(3)
    q = torch.randn(256, 4)
    np.savez_compressed(os.path.join(HERE, "torch_impl_quat.npz"), quats=q.numpy(), rotmats=R.quat_to_rotmat(q).numpy())
    print("wrote golden fixtures to", HERE)


if __name__ == "__main__":
    sys.exit(main())
