// fastgs_binding.cpp -- torch.ops binding of fast_gs::rasterization::forward_wrapper / backward_wrapper, used by the
// tests, the Python mirror and bench.py.  The SAME file is compiled twice:
//   * into libgsplat_b200.so against include/fastgs/rasterization_api.h (this backend)   -> torch.ops.gsplat_b200.fastgs_*
//   * by oracle/build_ref.py against the reference's own rasterization_api.h and sources -> torch.ops.fastgs_ref.fastgs_*
// so both libraries are driven through identical call sites.  The five ints of the forward travel as a CPU int64 tensor.
#include <torch/library.h>

#include "rasterization_api.h"
#ifndef FGS_REFERENCE_LIBRARY
#include "rasterization_ext.h"
#endif

namespace {
using at::Tensor;

std::tuple<Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor> fastgs_forward(
    const Tensor &means, const Tensor &scales_raw, const Tensor &rotations_raw, const Tensor &opacities_raw,
    const Tensor &sh0, const Tensor &shN, const Tensor &w2c, const Tensor &cam_position, int64_t active_sh_bases,
    int64_t width, int64_t height, double fx, double fy, double cx, double cy, double near_plane, double far_plane) {
    auto r = fast_gs::rasterization::forward_wrapper(means, scales_raw, rotations_raw, opacities_raw, sh0, shN, w2c,
                                                     cam_position, (int)active_sh_bases, (int)width, (int)height, (float)fx,
                                                     (float)fy, (float)cx, (float)cy, (float)near_plane, (float)far_plane);
    Tensor ints = at::empty({5}, at::TensorOptions().dtype(at::kLong));
    int64_t *p = ints.data_ptr<int64_t>();
    p[0] = std::get<6>(r); p[1] = std::get<7>(r); p[2] = std::get<8>(r); p[3] = std::get<9>(r); p[4] = std::get<10>(r);
    return std::make_tuple(std::get<0>(r), std::get<1>(r), std::get<2>(r), std::get<3>(r), std::get<4>(r), std::get<5>(r), ints);
}

std::tuple<Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor> fastgs_backward(
    Tensor densification_info, const Tensor &grad_image, const Tensor &grad_alpha, const Tensor &image, const Tensor &alpha,
    const Tensor &means, const Tensor &scales_raw, const Tensor &rotations_raw, const Tensor &shN, const Tensor &per_primitive,
    const Tensor &per_tile, const Tensor &per_instance, const Tensor &per_bucket, const Tensor &w2c, const Tensor &cam_position,
    int64_t active_sh_bases, int64_t width, int64_t height, double fx, double fy, double cx, double cy, double near_plane,
    double far_plane, const Tensor &ints, bool want_w2c_grad) {
    const int64_t *p = ints.data_ptr<int64_t>();
    Tensor w = w2c.detach();
    if (want_w2c_grad) w = w.clone().requires_grad_(true);
    auto r = fast_gs::rasterization::backward_wrapper(
        densification_info, grad_image, grad_alpha, image, alpha, means, scales_raw, rotations_raw, shN, per_primitive, per_tile,
        per_instance, per_bucket, w, cam_position, (int)active_sh_bases, (int)width, (int)height, (float)fx, (float)fy, (float)cx,
        (float)cy, (float)near_plane, (float)far_plane, (int)p[0], (int)p[1], (int)p[2], (int)p[3], (int)p[4]);
    Tensor gw = std::get<6>(r);
    if (!gw.defined()) gw = at::empty({0}, means.options());
    return std::make_tuple(std::get<0>(r), std::get<1>(r), std::get<2>(r), std::get<3>(r), std::get<4>(r), std::get<5>(r), gw);
}
#ifndef FGS_REFERENCE_LIBRARY
std::tuple<Tensor, Tensor, Tensor, Tensor, Tensor, Tensor> fastgs_forward_capacity(
    const Tensor &means, const Tensor &scales_raw, const Tensor &rotations_raw, const Tensor &opacities_raw,
    const Tensor &sh0, const Tensor &shN, const Tensor &w2c, const Tensor &cam_position, int64_t active_sh_bases,
    int64_t width, int64_t height, double fx, double fy, double cx, double cy, double near_plane, double far_plane,
    int64_t instance_capacity) {
    return fast_gs::rasterization::forward_capacity(means, scales_raw, rotations_raw, opacities_raw, sh0, shN, w2c, cam_position,
                                                    (int)active_sh_bases, (int)width, (int)height, (float)fx, (float)fy,
                                                    (float)cx, (float)cy, (float)near_plane, (float)far_plane, instance_capacity);
}
#endif
} // namespace

#ifdef FGS_REFERENCE_LIBRARY
TORCH_LIBRARY(fastgs_ref, m) {
#else
TORCH_LIBRARY_FRAGMENT(gsplat_b200, m) {
#endif
    m.def("fastgs_forward", &fastgs_forward);
#ifndef FGS_REFERENCE_LIBRARY
    m.def("fastgs_forward_capacity", &fastgs_forward_capacity);
#endif
    m.def("fastgs_backward(Tensor(a!) densification_info, Tensor grad_image, Tensor grad_alpha, Tensor image, Tensor alpha, "
          "Tensor means, Tensor scales_raw, Tensor rotations_raw, Tensor shN, Tensor per_primitive, Tensor per_tile, "
          "Tensor per_instance, Tensor per_bucket, Tensor w2c, Tensor cam_position, int active_sh_bases, int width, "
          "int height, float fx, float fy, float cx, float cy, float near_plane, float far_plane, Tensor ints, "
          "bool want_w2c_grad) -> (Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor)",
          &fastgs_backward);
}
