"""CPU (build container only): the reference's own L3 caller compiles, unmodified, against this repo's
drop-in headers and links against libgsplat_b200.so -- i.e. the gsplat:: API surface is complete for
src/training/rasterization/rasterizer_autograd.cpp.  core/camera.hpp and core/splat_data.hpp
are replaced by empty stubs: that translation unit does not use anything from them."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SRC = "/root/reference/src/training/rasterization/rasterizer_autograd.cpp"


@pytest.mark.skipif(not os.path.exists(REF_SRC), reason="reference tree not present (GPU box)")
def test_reference_autograd_caller_links_against_b200_backend(pkg, tmp_path):
    from torch.utils import cpp_extension as ce

    pkg.build()
    inc = []
    for i in [os.path.join(ROOT, "tests", "link_stubs"), os.path.join(ROOT, "include", "gsplat"),
              os.path.join(ROOT, "include"), "/root/reference/src/training"] + ce.include_paths() + \
             ["/usr/local/cuda/include"]:
        inc += ["-I", i]
    obj = str(tmp_path / "rasterizer_autograd.o")
    subprocess.run(["/usr/bin/g++", "-std=c++20", "-O0", "-fPIC", "-DGSB_NO_GLM", "-c", REF_SRC, "-o", obj] + inc,
                   check=True)
    lib_dir = os.path.join(ROOT, "gaussian-splatting-cuda_b200", "lib")
    tl = ce.library_paths()[0]
    so = str(tmp_path / "libcaller.so")
    # -Wl,--no-undefined: every gsplat:: symbol the caller references must be provided by the backend
    subprocess.run(["/usr/bin/g++", "-shared", "-o", so, obj, "-L", lib_dir, "-lgsplat_b200", "-lgsb200", "-L", tl,
                    "-ltorch", "-ltorch_cpu", "-lc10", "-ltorch_cuda", "-lc10_cuda", "-Wl,--no-undefined",
                    "-Wl,-rpath," + lib_dir, "-Wl,-rpath," + tl], check=True)
    out = subprocess.run(["nm", "-D", "--undefined-only", "-C", so], capture_output=True, text=True).stdout
    for fn in ("gsplat::spherical_harmonics_fwd", "gsplat::spherical_harmonics_bwd", "gsplat::projection_ut_3dgs_fused",
               "gsplat::rasterize_to_pixels_from_world_3dgs_fwd", "gsplat::rasterize_to_pixels_from_world_3dgs_bwd"):
        assert fn in out, f"{fn} not referenced by the caller object?"


FAST_SRC = "/root/reference/src/training/rasterization/fast_rasterizer_autograd.cpp"


@pytest.mark.skipif(not os.path.exists(FAST_SRC), reason="reference tree not present (GPU box)")
def test_reference_fastgs_caller_links_against_b200_backend(pkg, tmp_path):
    """the reference's caller of its default rasterizer (FastGSRasterize, fast_rasterizer_autograd.cpp)
    compiles unmodified against include/fastgs/rasterization_api.h and links against libgsplat_b200.so."""
    from torch.utils import cpp_extension as ce

    pkg.build()
    inc = []
    for i in [os.path.join(ROOT, "include", "fastgs"), "/root/reference/src/training"] + ce.include_paths() + \
             ["/usr/local/cuda/include"]:
        inc += ["-I", i]
    obj = str(tmp_path / "fast_rasterizer_autograd.o")
    subprocess.run(["/usr/bin/g++", "-std=c++20", "-O0", "-fPIC", "-c", FAST_SRC, "-o", obj] + inc, check=True)
    lib_dir = os.path.join(ROOT, "gaussian-splatting-cuda_b200", "lib")
    tl = ce.library_paths()[0]
    so = str(tmp_path / "libfastcaller.so")
    subprocess.run(["/usr/bin/g++", "-shared", "-o", so, obj, "-L", lib_dir, "-lgsplat_b200", "-lgsb200", "-L", tl,
                    "-ltorch", "-ltorch_cpu", "-lc10", "-ltorch_cuda", "-lc10_cuda", "-Wl,--no-undefined",
                    "-Wl,-rpath," + lib_dir, "-Wl,-rpath," + tl], check=True)
    out = subprocess.run(["nm", "-D", "--undefined-only", "-C", so], capture_output=True, text=True).stdout
    for fn in ("fast_gs::rasterization::forward_wrapper", "fast_gs::rasterization::backward_wrapper"):
        assert fn in out, f"{fn} not referenced by the caller object?"
