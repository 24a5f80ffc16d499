"""CPU: the error convention of the C ABI (include/gsb200.h).  Every entry point validates its arguments
before it touches the device, so these paths run without a GPU: 0 for the empty cases the reference
treats as no-ops, GSB_E_INVALID (-1) for null / inconsistent arguments, GSB_E_UNSUPPORTED (-2) for what
the reference's API allows but this backend does not implement, GSB_E_WORKSPACE (-3) for a short or
misaligned workspace.  No compute call is made."""
import ctypes as C
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OK, E_INVALID, E_UNSUPPORTED, E_WORKSPACE = 0, -1, -2, -3
PINHOLE, ORTHO, FISHEYE = 0, 1, 2
SHUTTER_ROLLING, SHUTTER_GLOBAL = 0, 4  # ShutterType values of the reference (Cameras.h)


class GsbUTParams(C.Structure):
    _fields_ = [("alpha", C.c_float), ("beta", C.c_float), ("kappa", C.c_float),
                ("in_image_margin_factor", C.c_float), ("require_all_sigma_points_valid", C.c_int32)]


class GsbCamera(C.Structure):
    _fields_ = [("viewmats0", C.c_void_p), ("viewmats1", C.c_void_p), ("Ks", C.c_void_p),
                ("camera_model", C.c_int32), ("shutter_type", C.c_int32),
                ("radial_coeffs", C.c_void_p), ("radial_count", C.c_int32),
                ("tangential_coeffs", C.c_void_p), ("tangential_count", C.c_int32),
                ("thin_prism_coeffs", C.c_void_p), ("thin_prism_count", C.c_int32), ("ut", GsbUTParams)]


@pytest.fixture(scope="module")
def lib(pkg):
    path = os.path.join(ROOT, "gaussian-splatting-cuda_b200", "lib", "libgsb200.so")
    if not os.path.exists(path):
        pkg.build()
    L = C.CDLL(path)
    L.gsb_error_string.restype = C.c_char_p
    for n in ("gsb_raster_fwd_workspace", "gsb_raster_bwd_workspace", "gsb_isect_count_workspace",
              "gsb_isect_plan_workspace", "gsb_isect_sort_workspace"):
        getattr(L, n).restype = C.c_size_t
    return L


# any non-null address will do: the calls below return before dereferencing it on the device
_buf = np.zeros(8192, np.uint8)
PTR = C.c_void_p(_buf.ctypes.data)
NULL = C.c_void_p(None)
u32, u64, f32, i32, sz = C.c_uint32, C.c_uint64, C.c_float, C.c_int32, C.c_size_t


def camera(model=PINHOLE, shutter=SHUTTER_GLOBAL, viewmats1=None):
    cam = GsbCamera()
    cam.viewmats0, cam.Ks = PTR.value, PTR.value
    cam.viewmats1 = viewmats1
    cam.camera_model, cam.shutter_type = model, shutter
    cam.ut = GsbUTParams(0.1, 2.0, 0.0, 0.1, 1)
    return cam


def test_error_strings(lib):
    assert lib.gsb_error_string(OK) == b"ok"
    for code in (E_INVALID, E_UNSUPPORTED, E_WORKSPACE):
        assert len(lib.gsb_error_string(code)) > 3


def test_sh_argument_validation(lib):
    assert lib.gsb_sh_fwd(u32(0), u32(16), u32(3), NULL, NULL, NULL, NULL, NULL) == OK  # empty: no-op
    assert lib.gsb_sh_fwd(u32(5), u32(16), u32(3), NULL, PTR, NULL, PTR, NULL) == E_INVALID
    assert lib.gsb_sh_fwd(u32(5), u32(4), u32(3), PTR, PTR, NULL, PTR, NULL) == E_INVALID  # 16 bases > K = 4
    assert lib.gsb_sh_fwd(u32(5), u32(36), u32(5), PTR, PTR, NULL, PTR, NULL) == E_INVALID  # degree > 4
    assert lib.gsb_sh_bwd(u32(0), u32(16), u32(3), NULL, NULL, NULL, NULL, NULL, NULL, NULL) == OK
    assert lib.gsb_sh_bwd(u32(5), u32(16), u32(3), PTR, PTR, NULL, NULL, PTR, PTR, NULL) == E_INVALID
    assert lib.gsb_sh_bwd_views(u32(0), u32(16), u32(3), u32(2), NULL, NULL, NULL, NULL, NULL, NULL, NULL) == OK
    assert lib.gsb_sh_bwd_views(u32(5), u32(16), u32(3), u32(2), PTR, NULL, PTR, PTR, PTR, PTR, NULL) == E_INVALID
    assert lib.gsb_sh_bwd_views(u32(5), u32(9), u32(3), u32(2), PTR, PTR, PTR, PTR, PTR, PTR, NULL) == E_INVALID


def test_projection_argument_validation(lib):
    args = (f32(0.3), f32(0.01), f32(1e4), f32(0.0), PTR, PTR, PTR, PTR, NULL, NULL)
    assert lib.gsb_projection_ut(u32(1), u32(8), PTR, PTR, PTR, PTR, NULL, u32(64), u32(64), *args) == E_INVALID
    cam = camera()
    assert lib.gsb_projection_ut(u32(1), u32(0), NULL, NULL, NULL, NULL, C.byref(cam), u32(64), u32(64), *args) == OK
    assert lib.gsb_projection_ut(u32(1), u32(8), NULL, PTR, PTR, PTR, C.byref(cam), u32(64), u32(64), *args) == E_INVALID
    for bad in (camera(model=ORTHO),):  # rolling shutter is implemented; only the orthographic model is not
        assert lib.gsb_projection_ut(u32(1), u32(8), PTR, PTR, PTR, PTR, C.byref(bad), u32(64), u32(64),
                                     *args) == E_UNSUPPORTED


def test_intersect_argument_validation(lib):
    assert lib.gsb_isect_count(u32(1), u32(0), NULL, NULL, u32(16), u32(4), u32(4), NULL, NULL, NULL, sz(0), NULL) == OK
    assert lib.gsb_isect_count(u32(1), u32(8), NULL, PTR, u32(16), u32(4), u32(4), PTR, PTR, PTR, sz(4096),
                               NULL) == E_INVALID
    assert lib.gsb_isect_count(u32(1), u32(8), PTR, PTR, u32(16), u32(4), u32(4), PTR, PTR, PTR, sz(8),
                               NULL) == E_WORKSPACE
    assert lib.gsb_isect_count_workspace(u64(1000)) >= 256
    # plan: the host needs somewhere to read n_isects from; the key must fit 64 bits (Intersect.cpp:50)
    need = lib.gsb_isect_plan_workspace(u32(1), u32(8), u32(4), u32(4))
    assert need >= 8 * (4 + 8)
    aligned = C.c_void_p((PTR.value + 255) & ~255)
    assert lib.gsb_isect_plan(u32(1), u32(8), PTR, PTR, PTR, u32(16), u32(4), u32(4), PTR, NULL, NULL, i32(0), aligned, sz(need),
                              NULL) == E_INVALID
    assert lib.gsb_isect_plan(u32(1), u32(8), NULL, PTR, PTR, u32(16), u32(4), u32(4), PTR, PTR, NULL, i32(0), aligned, sz(need),
                              NULL) == E_INVALID
    assert lib.gsb_isect_plan(u32(4), u32(8), PTR, PTR, PTR, u32(16), u32(1 << 16), u32(1 << 15), PTR, PTR, NULL, i32(0), aligned,
                              sz(1 << 20), NULL) == E_INVALID  # 32 tile bits + 3 camera bits > 32
    assert lib.gsb_isect_plan(u32(1), u32(8), PTR, PTR, PTR, u32(16), u32(1 << 16), u32(1), PTR, PTR, NULL, i32(0), aligned,
                              sz(1 << 20), NULL) == E_INVALID  # tile grid side > 65535 (box packing)
    assert lib.gsb_isect_plan(u32(1), u32(8), PTR, PTR, PTR, u32(16), u32(4), u32(4), PTR, PTR, NULL, i32(0), aligned, sz(16),
                              NULL) == E_WORKSPACE
    misaligned = C.c_void_p(aligned.value + 8)
    assert lib.gsb_isect_plan(u32(1), u32(8), PTR, PTR, PTR, u32(16), u32(4), u32(4), PTR, PTR, NULL, i32(0), misaligned, sz(need),
                              NULL) == E_WORKSPACE
    # emit: nothing to do without intersections; a short or misaligned plan workspace is refused
    assert lib.gsb_isect_emit_planned(u32(1), u32(8), PTR, u32(4), u32(4), u64(0), aligned, sz(need), PTR, PTR, NULL) == OK
    assert lib.gsb_isect_emit_planned(u32(1), u32(8), PTR, u32(4), u32(4), u64(100), aligned, sz(64), PTR, PTR,
                                      NULL) == E_WORKSPACE
    assert lib.gsb_isect_emit_planned(u32(1), u32(8), PTR, u32(4), u32(4), u64(100), aligned, sz(need), PTR, NULL,
                                      NULL) == E_INVALID
    assert lib.gsb_isect_sort(u64(0), u32(1), u32(4), u32(4), NULL, NULL, NULL, NULL, NULL, sz(0), NULL) == OK
    assert lib.gsb_isect_sort(u64(10), u32(1), u32(4), u32(4), PTR, PTR, PTR, PTR, PTR, sz(8), NULL) == E_WORKSPACE


def test_blend_argument_validation(lib):
    fwd_tail = (PTR, PTR, PTR, PTR, PTR, PTR, sz(1 << 20), NULL)

    def fwd(C_=1, N=8, n_isects=10, W=64, H=64, tile=16, cam=None, ws_bytes=None, means=PTR):
        cam = cam if cam is not None else camera()
        ws = sz(ws_bytes if ws_bytes is not None else lib.gsb_raster_fwd_workspace(u32(N)))
        return lib.gsb_raster_fwd(u32(C_), u32(N), u64(n_isects), means, PTR, PTR, PTR, PTR, NULL, NULL, u32(W), u32(H),
                                  u32(tile), C.byref(cam), PTR, PTR, PTR, PTR, PTR,
                                  C.c_void_p((_buf.ctypes.data + 255) & ~255), ws, NULL)

    assert lib.gsb_raster_fwd(u32(1), u32(8), u64(10), PTR, PTR, PTR, PTR, PTR, NULL, NULL, u32(64), u32(64), u32(16), NULL,
                              *fwd_tail) == E_INVALID  # no camera
    assert fwd(W=0) == OK                              # empty image: no-op
    assert fwd(C_=2) == E_UNSUPPORTED                  # the reference's kernels are single-camera too
    assert fwd(tile=8) == E_UNSUPPORTED
    assert fwd(cam=camera(model=ORTHO)) == E_UNSUPPORTED
    assert fwd(cam=camera(shutter=7)) == E_INVALID       # not a ShutterType
    assert fwd(means=NULL) == E_INVALID
    assert fwd(n_isects=1 << 31) == E_INVALID          # int32 tile offsets
    assert fwd(ws_bytes=64) == E_WORKSPACE
    assert lib.gsb_raster_fwd_workspace(u32(1000)) >= 64 * 1000
    assert lib.gsb_raster_bwd_workspace(u32(1000)) >= 2 * 64 * 1000
    # backward: same camera / shape rules
    cam = camera()
    bwd = lambda C_, tile, cm: lib.gsb_raster_bwd(
        u32(C_), u32(8), u64(10), PTR, PTR, PTR, PTR, PTR, NULL, NULL, u32(64), u32(64), u32(tile), C.byref(cm), PTR, PTR,
        PTR, PTR, PTR, PTR, PTR, PTR, PTR, PTR, PTR, PTR, sz(64), NULL)
    assert bwd(2, 16, cam) == E_UNSUPPORTED
    assert bwd(1, 32, cam) == E_UNSUPPORTED
    assert bwd(1, 16, camera(model=ORTHO)) == E_UNSUPPORTED
    assert bwd(1, 16, cam) == E_WORKSPACE              # 64-byte workspace for 8 Gaussians


def test_strategy_helpers_argument_validation(lib):
    assert lib.gsb_quat_to_rotmat(u32(0), NULL, NULL, NULL) == OK
    assert lib.gsb_quat_to_rotmat(u32(4), NULL, PTR, NULL) == E_INVALID
    assert lib.gsb_relocation(u32(0), NULL, NULL, NULL, NULL, i32(51), NULL, NULL, NULL) == OK
    assert lib.gsb_relocation(u32(4), PTR, PTR, NULL, PTR, i32(51), PTR, PTR, NULL) == E_INVALID
    assert lib.gsb_add_noise(u32(0), NULL, NULL, NULL, NULL, NULL, f32(1e-3), NULL) == OK
    assert lib.gsb_add_noise(u32(4), PTR, PTR, PTR, NULL, PTR, f32(1e-3), NULL) == E_INVALID


def test_shim_rejects_cpu_tensors_like_the_reference(pkg):
    """CHECK_INPUT of the reference (Common.h:12-17): every operator insists on CUDA, contiguous tensors and throws
    c10::Error (RuntimeError in Python) otherwise -- checked here with CPU tensors, no device needed."""
    import torch
    pkg.load()
    N = 8
    means, quats, scales = torch.randn(N, 3), torch.randn(N, 4), torch.rand(N, 3)
    with pytest.raises(RuntimeError, match="CUDA"):
        pkg.spherical_harmonics_fwd(3, torch.randn(N, 3), torch.randn(N, 16, 3), None)
    with pytest.raises(RuntimeError, match="CUDA"):
        pkg.spherical_harmonics_bwd(16, 3, torch.randn(N, 3), torch.randn(N, 16, 3), None, torch.randn(N, 3), True)
    with pytest.raises(RuntimeError, match="CUDA"):
        pkg.intersect_tile(torch.randn(1, N, 2), torch.ones(1, N, 2, dtype=torch.int32), torch.rand(1, N), 1, 16, 4, 4)
    with pytest.raises(RuntimeError, match="CUDA"):
        pkg.intersect_offset(torch.zeros(4, dtype=torch.int64), 1, 4, 4)
    with pytest.raises(RuntimeError, match="CUDA"):
        pkg.projection_ut_3dgs_fused(means, quats, scales, torch.rand(N), torch.eye(4)[None], torch.eye(3)[None], 64, 64)
    with pytest.raises(RuntimeError, match="CUDA"):
        pkg.quats_to_rotmats(quats)
    with pytest.raises(RuntimeError, match="CUDA"):
        pkg.default_backend().spherical_harmonics_bwd_views(3, means, torch.zeros(1, 3), torch.randn(N, 16, 3),
                                                           torch.randn(1, N, 3), torch.zeros(N, 3))


class GsbFastgsView(C.Structure):
    _fields_ = [("w2c", C.c_void_p), ("cam_position", C.c_void_p), ("width", C.c_uint32), ("height", C.c_uint32),
                ("focal_x", C.c_float), ("focal_y", C.c_float), ("center_x", C.c_float), ("center_y", C.c_float),
                ("near_plane", C.c_float), ("far_plane", C.c_float), ("active_sh_bases", C.c_uint32),
                ("total_bases_sh_rest", C.c_uint32)]


def fastgs_view(active=16, rest=15, w=64, h=48):
    v = GsbFastgsView()
    v.w2c, v.cam_position = PTR.value, PTR.value
    v.width, v.height = w, h
    v.focal_x = v.focal_y = 50.0
    v.center_x, v.center_y = w / 2, h / 2
    v.near_plane, v.far_plane = 0.01, 1e10
    v.active_sh_bases, v.total_bases_sh_rest = active, rest
    return v


def test_fastgs_argument_validation(lib):
    """SURVEY.md 8 f4 entry points (gsb_fastgs_*): view / buffer checks run before anything touches the device."""
    lib.gsb_fastgs_primitive_bytes.restype = C.c_size_t
    lib.gsb_fastgs_tile_bytes.restype = C.c_size_t
    pb = lib.gsb_fastgs_primitive_bytes(u32(100), u32(64), u32(48))
    tb = lib.gsb_fastgs_tile_bytes(u32(64), u32(48))
    assert pb >= 100 * (64 + 64 + 4 + 8 + 4 + 4 + 16 + 8) and pb % 256 == 0
    assert tb >= (4 * 3 + 1) * 4 + 64 * 48 * 4
    assert lib.gsb_fastgs_primitive_bytes(u32(200), u32(64), u32(48)) > pb
    n_out = np.zeros(1, np.int64)
    NOUT = C.c_void_p(n_out.ctypes.data)
    v = fastgs_view()

    def plan(view, prim=PTR, prim_bytes=pb, tile=PTR, tile_bytes=tb, nout=NOUT):
        return lib.gsb_fastgs_forward_plan(u32(100), PTR, PTR, PTR, PTR, PTR, PTR, view, prim, sz(prim_bytes), tile,
                                           sz(tile_bytes), nout, NULL)

    assert plan(None) == E_INVALID
    assert plan(C.byref(fastgs_view(active=5))) == E_INVALID            # not a full SH degree
    assert plan(C.byref(fastgs_view(active=16, rest=3))) == E_INVALID   # more active bases than stored
    assert plan(C.byref(fastgs_view(w=0))) == E_INVALID
    assert plan(C.byref(v), nout=NULL) == E_INVALID
    assert plan(C.byref(v), prim_bytes=1024) == E_WORKSPACE
    assert plan(C.byref(v), tile_bytes=16) == E_WORKSPACE
    assert plan(C.byref(v), prim=NULL) == E_WORKSPACE
    big = np.zeros(pb + 512, np.uint8)
    aligned = (big.ctypes.data + 255) & ~255
    assert plan(C.byref(v), prim=C.c_void_p(aligned + 4)) == E_WORKSPACE  # misaligned
    assert lib.gsb_fastgs_forward_blend(u32(100), C.byref(v), PTR, sz(pb), PTR, sz(tb), NULL, u64(10), PTR, PTR,
                                        NULL) == E_INVALID              # capacity without an instance buffer
    assert lib.gsb_fastgs_forward_blend(u32(100), C.byref(v), PTR, sz(pb), PTR, sz(tb), PTR, u64(10), NULL, PTR,
                                        NULL) == E_INVALID
    # backward with N == 0 is a no-op; missing gradients are rejected
    assert lib.gsb_fastgs_backward(u32(0), NULL, NULL, NULL, NULL, C.byref(v), NULL, sz(0), NULL, sz(0), NULL, u64(0), NULL,
                                   NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL) == OK
    assert lib.gsb_fastgs_backward(u32(100), PTR, PTR, PTR, PTR, C.byref(v), PTR, sz(pb), PTR, sz(tb), PTR, u64(10), PTR, PTR,
                                   PTR, NULL, PTR, PTR, PTR, PTR, PTR, NULL, NULL, NULL) == E_INVALID


def test_fastgs_shim_rejects_cpu_tensors(pkg):
    import torch
    pkg.load()
    N = 8
    with pytest.raises(RuntimeError, match="CUDA"):
        torch.ops.gsplat_b200.fastgs_forward(torch.randn(N, 3), torch.randn(N, 3), torch.randn(N, 4), torch.randn(N, 1),
                                             torch.randn(N, 1, 3), torch.randn(N, 15, 3), torch.eye(4), torch.zeros(3), 16,
                                             64, 48, 50.0, 50.0, 32.0, 24.0, 0.01, 1e10)
