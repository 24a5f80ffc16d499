"""ctypes front-end of oracle/fastgs_oracle.c -- the CPU restatement of the reference's fastgs (EWA) rasterizer
(SURVEY.md 8 f4).  TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__ and bench.py's CPU legs.

render(...) mirrors fast_gs::rasterization::forward_wrapper / backward_wrapper
(/root/reference/fastgs/rasterization/include/rasterization_api.h:25-75): raw parameters in, image [3,H,W] and alpha
[1,H,W] out; with grad_image / grad_alpha the gradients of the raw parameters as well (float64).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS: dict[str, C.CDLL] = {}


def build(force: bool = False) -> None:
    src = os.path.join(_HERE, "fastgs_oracle.c")
    libs = [os.path.join(_HERE, f"libfastgs_oracle_{p}.so") for p in ("f32", "f64", "f64s")]
    need = force or not all(os.path.exists(l) for l in libs) or any(os.path.getmtime(l) < os.path.getmtime(src) for l in libs)
    if need:
        subprocess.run(["make", "-C", _HERE, "-s", "-B", "libfastgs_oracle_f32.so", "libfastgs_oracle_f64.so",
                        "libfastgs_oracle_f64s.so"], check=True)


def lib(precision: str = "f32") -> C.CDLL:
    if precision not in _LIBS:
        path = os.path.join(_HERE, f"libfastgs_oracle_{precision}.so")
        if not os.path.exists(path):
            build()
        _LIBS[precision] = C.CDLL(path)
    return _LIBS[precision]


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a, ty=C.c_float):
    return None if a is None else a.ctypes.data_as(C.POINTER(ty))


def render(means, scales_raw, rotations_raw, opacities_raw, sh0, shN, w2c, cam_position, active_sh_bases, width, height,
           fx, fy, cx, cy, near_plane=0.01, far_plane=1e10, grad_image=None, grad_alpha=None, want_w2c_grad=False,
           densification_info=None, precision="f32"):
    n = int(means.shape[0])
    means, scales_raw, rotations_raw = _f(means), _f(scales_raw), _f(rotations_raw)
    opacities_raw, sh0 = _f(opacities_raw).reshape(-1), _f(sh0).reshape(n, 3)
    shN = _f(shN).reshape(n, -1, 3)
    total_rest = int(shN.shape[1])
    w2c, cam_position = _f(w2c).reshape(4, 4), _f(cam_position).reshape(3)
    image = np.zeros((3, height, width), np.float32)
    alpha = np.zeros((1, height, width), np.float32)
    n_touched = np.zeros(n, np.int32)
    n_inst = C.c_int64(0)
    out = {}
    bwd = grad_image is not None
    gi = ga = gm = gs = gq = go = g0 = gN = gw = None
    if bwd:
        gi = _f(grad_image).reshape(3, height, width)
        ga = _f(grad_alpha).reshape(height, width) if grad_alpha is not None else np.zeros((height, width), np.float32)
        gm, gs, gq = np.zeros((n, 3)), np.zeros((n, 3)), np.zeros((n, 4))
        go, g0, gN = np.zeros((n, 1)), np.zeros((n, 1, 3)), np.zeros((n, total_rest, 3))
        gw = np.zeros((4, 4)) if want_w2c_grad else None
    dens = None
    if densification_info is not None:
        dens = np.ascontiguousarray(densification_info, dtype=np.float32)
    rc = lib(precision).fgo_render(
        C.c_uint32(n), _p(means), _p(scales_raw), _p(rotations_raw), _p(opacities_raw), _p(sh0), _p(shN),
        C.c_uint32(total_rest), C.c_uint32(active_sh_bases), _p(w2c), _p(cam_position), C.c_uint32(width),
        C.c_uint32(height), C.c_float(fx), C.c_float(fy), C.c_float(cx), C.c_float(cy), C.c_float(near_plane),
        C.c_float(far_plane), _p(image), _p(alpha), _p(n_touched, C.c_int32), C.byref(n_inst), _p(gi), _p(ga),
        _p(gm, C.c_double), _p(gs, C.c_double), _p(gq, C.c_double), _p(go, C.c_double), _p(g0, C.c_double),
        _p(gN, C.c_double), _p(gw, C.c_double), _p(dens))
    if rc != 0:
        raise RuntimeError(f"fastgs oracle failed ({rc})")
    out.update(image=image, alpha=alpha, n_touched=n_touched, n_instances=int(n_inst.value))
    if bwd:
        out.update(grad_means=gm, grad_scales_raw=gs, grad_rotations_raw=gq, grad_opacities_raw=go, grad_sh0=g0,
                   grad_shN=gN, grad_w2c=gw)
    if dens is not None:
        out["densification_info"] = dens
    return out
