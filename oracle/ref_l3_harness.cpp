// ref_l3_harness.cpp -- runs the REFERENCE's own L3 autograd code on the GPU through this repository's drop-in library.
//
// /root/reference/src/training/rasterization/rasterizer_autograd.cpp (SphericalHarmonicsFunction,
// fully_fused_projection_with_ut, GUTRasterizationFunction: forward AND backward) is compiled unmodified and linked
// against libgsplat_b200.so; this file only strings them together the way gs::training::rasterize does
// (rasterizer.cpp:226-361, RenderMode::RGB): it is the part of rasterizer.cpp that does not need the product's
// Camera / SplatData classes, statement for statement (activated parameters in, image out).
// TEST INFRASTRUCTURE ONLY (tests/test_gpu_reference_l3.py): turns "the reference's caller links" into "it runs".
#include <torch/library.h>
#include <torch/torch.h>

#include "Ops.h"
#include "rasterization/rasterizer_autograd.hpp"

namespace {
using torch::Tensor;
using torch::indexing::None;
using torch::indexing::Slice;

// rasterizer.cpp:14-43
Tensor spherical_harmonics(int sh_degree, const Tensor &dirs, const Tensor &coeffs, const Tensor &masks) {
    auto sh_degree_tensor = torch::tensor({sh_degree}, torch::TensorOptions().dtype(torch::kInt32).device(dirs.device()));
    return gs::training::SphericalHarmonicsFunction::apply(sh_degree_tensor, dirs.contiguous(), coeffs.contiguous(),
                                                            masks.defined() ? masks.contiguous() : masks)[0];
}

std::tuple<Tensor, Tensor, Tensor, Tensor> render(const Tensor &means3D, const Tensor &rotations, const Tensor &scales,
                                                  const Tensor &opacities, const Tensor &sh_coeffs, int64_t sh_degree,
                                                  const Tensor &viewmat, const Tensor &K, int64_t image_width,
                                                  int64_t image_height, const Tensor &prepared_bg_color) {
    using namespace gs::training;
    const int tile_size = 16;                                                  // rasterizer.cpp:176-180
    const float eps2d = 0.3f, near_plane = 0.01f, far_plane = 10000.0f, radius_clip = 0.0f, scaling_modifier = 1.0f;
    auto ut_params = UnscentedTransformParameters{};
    GUTProjectionSettings proj_settings{(int)image_width, (int)image_height, eps2d, near_plane, far_plane, radius_clip,
                                        scaling_modifier, gsplat::CameraModelType::PINHOLE};
    auto proj_outputs = fully_fused_projection_with_ut(means3D, rotations, scales, opacities, viewmat, K, std::nullopt,
                                                       std::nullopt, std::nullopt, proj_settings, ut_params);
    auto radii = proj_outputs[0], means2d = proj_outputs[1], depths = proj_outputs[2];
    auto means2d_with_grad = means2d.contiguous();
    // rasterizer.cpp:249-266
    auto viewmat_inv = torch::inverse(viewmat);
    auto campos = viewmat_inv.index({Slice(), Slice(None, 3), 3});
    auto dirs = means3D.unsqueeze(0) - campos.unsqueeze(1);
    auto masks = (radii > 0).all(-1);
    auto shs = sh_coeffs.unsqueeze(0);
    auto colors = spherical_harmonics((int)sh_degree, dirs, shs, masks);
    colors = torch::clamp_min(colors + 0.5f, 0.0f);
    Tensor final_bg = prepared_bg_color.defined() ? prepared_bg_color : at::empty({0}, colors.options());
    auto final_opacities = opacities.unsqueeze(0);
    // rasterizer.cpp:313-361
    const int tile_width = ((int)image_width + tile_size - 1) / tile_size;
    const int tile_height = ((int)image_height + tile_size - 1) / tile_size;
    const auto isect_results = gsplat::intersect_tile(means2d_with_grad, radii, depths, {}, {}, 1, tile_size, tile_width,
                                                      tile_height, true);
    const auto isect_ids = std::get<1>(isect_results);
    const auto flatten_ids = std::get<2>(isect_results);
    auto isect_offsets = gsplat::intersect_offset(isect_ids, 1, tile_width, tile_height);
    isect_offsets = isect_offsets.reshape({1, tile_height, tile_width});
    auto raster_settings = GUTRasterizationSettings{(int)image_width, (int)image_height, tile_size, scaling_modifier,
                                                    gsplat::CameraModelType::PINHOLE};
    auto raster_outputs = GUTRasterizationFunction::apply(means3D, rotations, scales, colors, final_opacities, final_bg,
                                                          std::nullopt, viewmat, K, std::nullopt, std::nullopt, std::nullopt,
                                                          isect_offsets, flatten_ids, raster_settings, ut_params);
    return std::make_tuple(raster_outputs[0], raster_outputs[1], radii, flatten_ids);
}
} // namespace

TORCH_LIBRARY(ref_l3_b200, m) { m.def("render", &render); }
