// ref_fastgs_l3_harness.cpp -- runs the REFERENCE's own fastgs caller on the GPU through this repository's drop-in library.
//
// /root/reference/src/training/rasterization/fast_rasterizer_autograd.cpp (FastGSRasterize: forward AND backward, the
// saved-tensor / saved_data bookkeeping around forward_wrapper / backward_wrapper) is compiled unmodified against
// include/fastgs/rasterization_api.h and linked against libgsplat_b200.so.  This file is the part of
// fast_rasterizer.cpp:12-74 that does not need the product's Camera / SplatData classes, statement for statement:
// settings from plain arguments, FastGSRasterize::apply, background composite.
// TEST INFRASTRUCTURE ONLY (tests/test_gpu_fastgs.py): turns "the reference's caller links" into "it runs".
#include <torch/library.h>
#include <torch/torch.h>

#include "rasterization/fast_rasterizer_autograd.hpp"

namespace {
using torch::Tensor;

std::tuple<Tensor, Tensor> fast_render(const Tensor &means, const Tensor &raw_scales, const Tensor &raw_rotations,
                                       const Tensor &raw_opacities, const Tensor &sh0, const Tensor &shN, const Tensor &w2c,
                                       const Tensor &cam_position, int64_t sh_degree, int64_t width, int64_t height, double fx,
                                       double fy, double cx, double cy, const Tensor &bg_color, Tensor densification_info) {
    const int active_sh_bases = (int)((sh_degree + 1) * (sh_degree + 1)); // fast_rasterizer.cpp:33-34
    constexpr float near_plane = 0.01f;                                    // :36-37
    constexpr float far_plane = 1e10f;
    fast_gs::rasterization::FastGSSettings settings;
    settings.cam_position = cam_position;
    settings.active_sh_bases = active_sh_bases;
    settings.width = (int)width;
    settings.height = (int)height;
    settings.focal_x = (float)fx;
    settings.focal_y = (float)fy;
    settings.center_x = (float)cx;
    settings.center_y = (float)cy;
    settings.near_plane = near_plane;
    settings.far_plane = far_plane;
    auto raster_outputs = gs::training::FastGSRasterize::apply(means, raw_scales, raw_rotations, raw_opacities, sh0, shN, w2c,
                                                               densification_info, settings); // :53-62
    Tensor image = raster_outputs[0], alpha = raster_outputs[1];
    image = image + (1.0f - alpha) * bg_color.unsqueeze(-1).unsqueeze(-1); // :71
    return std::make_tuple(image, alpha);
}
} // namespace

TORCH_LIBRARY(ref_fastgs_l3_b200, m) {
    m.def("fast_render(Tensor means, Tensor raw_scales, Tensor raw_rotations, Tensor raw_opacities, Tensor sh0, Tensor shN, "
          "Tensor w2c, Tensor cam_position, int sh_degree, int width, int height, float fx, float fy, float cx, float cy, "
          "Tensor bg_color, Tensor(a!) densification_info) -> (Tensor, Tensor)",
          &fast_render);
}
