// FusedOps.h -- EXTENDED operators of the B200 backend: the hot path over the RAW SplatData
// tensors, with the L3 glue of the reference's caller folded in.  Not part of the reference's gsplat/Ops.h; the
// eleven operators of Ops.h stay available and unchanged.  What a maintainer replaces with one call each:
//
//   forward   src/core/splat_data.cpp:267-286 (get_means / get_opacity / get_rotation / get_scaling / get_shs:
//             exp, sigmoid, normalize, cat) + src/training/rasterization/rasterizer.cpp:250-266 (inverse(viewmat),
//             dirs, masks, spherical harmonics, clamp_min(+0.5)) + projection_ut_3dgs_fused + intersect_tile +
//             intersect_offset + rasterize_to_pixels_from_world_3dgs_fwd
//   backward  rasterize_to_pixels_from_world_3dgs_bwd + spherical_harmonics_bwd + the autograd backward of all of
//             the above, down to the gradients of the raw parameter tensors
//
// C == 1 (one camera per call, like the reference's kernels), RGB, tile size 16, global shutter.
#pragma once

#include <ATen/core/Tensor.h>
#include <c10/util/Optional.h>

#include <tuple>
#include <vector>

#include "Cameras.h"
#include "Common.h"

namespace gsplat {

struct FusedForwardResult {
    at::Tensor renders;      // [1,H,W,3]  (unclamped, background composited)
    at::Tensor alphas;       // [1,H,W,1]
    at::Tensor radii;        // [1,N,2] int32, 0 = culled
    at::Tensor means2d;      // [1,N,2]
    at::Tensor depths;       // [1,N]
    // context for the backward (opaque to callers)
    at::Tensor last_ids;     // [1,H,W] int32
    at::Tensor tile_offsets; // [th*tw + 1] int32, closed: the last entry is n_isects
    at::Tensor flatten_ids;  // [capacity] int32
    at::Tensor workspace;    // bytes: blend records + gradient moments
    at::Tensor n_isects;     // [1] int64 on the device (compare with flatten_ids.size(0) when a capacity was given)
};

// isect_capacity > 0: flatten_ids is allocated with that many entries and NOTHING in the call synchronises with
// the host (the whole step can be captured in a CUDA graph); intersections beyond the capacity are dropped, so the
// caller compares n_isects with the capacity afterwards.  isect_capacity <= 0: the call reads the count back once
// and allocates exactly, like gsplat::intersect_tile.
FusedForwardResult rasterize_from_world_fused_fwd(
    const at::Tensor means,        // [N,3]
    const at::Tensor sh0,          // [N,1,3]
    const at::Tensor shN,          // [N,K-1,3]
    const at::Tensor scaling_raw,  // [N,3]
    const at::Tensor rotation_raw, // [N,4]
    const at::Tensor opacity_raw,  // [N,1] or [N]
    const uint32_t sh_degree, const float scaling_modifier,
    const at::Tensor viewmat,      // [1,4,4]
    const at::Tensor K,            // [1,3,3]
    const uint32_t image_width, const uint32_t image_height, const float eps2d, const float near_plane,
    const float far_plane, const float radius_clip, const at::optional<at::Tensor> backgrounds, // [1,3]
    const CameraModelType camera_model, const UnscentedTransformParameters ut_params,
    const at::optional<at::Tensor> radial_coeffs, const at::optional<at::Tensor> tangential_coeffs,
    const at::optional<at::Tensor> thin_prism_coeffs, const int64_t isect_capacity, const bool prepare_backward);

// gradients of (means, sh0, shN, scaling_raw, rotation_raw, opacity_raw), shaped like the inputs
std::tuple<at::Tensor, at::Tensor, at::Tensor, at::Tensor, at::Tensor, at::Tensor> rasterize_from_world_fused_bwd(
    const at::Tensor means, const at::Tensor sh0, const at::Tensor shN, const at::Tensor scaling_raw,
    const at::Tensor rotation_raw, const at::Tensor opacity_raw, const uint32_t sh_degree, const float scaling_modifier,
    const at::Tensor viewmat, const at::Tensor K, const uint32_t image_width, const uint32_t image_height,
    const at::optional<at::Tensor> backgrounds, const CameraModelType camera_model,
    const UnscentedTransformParameters ut_params, const at::optional<at::Tensor> radial_coeffs,
    const at::optional<at::Tensor> tangential_coeffs, const at::optional<at::Tensor> thin_prism_coeffs,
    const at::Tensor radii, const at::Tensor tile_offsets, const at::Tensor flatten_ids, at::Tensor workspace,
    const at::Tensor render_alphas, const at::Tensor last_ids, const at::Tensor v_render_colors,
    const at::Tensor v_render_alphas);

// loss = (1 - lambda) * l1(clamp(render, 0, 1), target) + lambda * (1 - fused_ssim(.., "valid")) of
// src/training/trainer.cpp:103-126 and, when `compute_grad`, dLoss/d(renders) in the same kernel.
// renders [1,H,W,3] (the from-world blend's output) or [3,H,W] (the fastgs image), target [3,H,W] / [1,3,H,W] (the
// reference's layout) or [1,H,W,3].
// Returns (stats = device float[3]: loss, l1 mean, ssim mean;  v_renders shaped like renders, or an undefined tensor).
std::tuple<at::Tensor, at::Tensor> photometric_loss_fused(const at::Tensor renders, const at::Tensor target,
                                                          const float lambda_dssim, const bool compute_grad);

// one Adam step over all parameter groups in one launch (fastgs/optimizer/include/adam_kernels.cuh:13-36,
// src/training/optimizers/fused_adam.cpp:22-95).  step_counts are the per-group counts AFTER the increment.
void fused_adam_step(const std::vector<at::Tensor> &params, const std::vector<at::Tensor> &grads,
                     const std::vector<at::Tensor> &exp_avg, const std::vector<at::Tensor> &exp_avg_sq,
                     const std::vector<double> &lr, const double beta1, const double beta2, const double eps,
                     const std::vector<int64_t> &step_counts);

// The same step with (lr, 1/(1-beta1^t), 1/sqrt(1-beta2^t), enabled) per group in a DEVICE tensor [n_groups,4]: nothing
// step-dependent is baked into the launch, so the iteration can be replayed from a CUDA graph.
void fused_adam_step_dynamic(const std::vector<at::Tensor> &params, const std::vector<at::Tensor> &grads,
                             const std::vector<at::Tensor> &exp_avg, const std::vector<at::Tensor> &exp_avg_sq,
                             const at::Tensor dynamic_scalars, const double beta1, const double beta2, const double eps);

} // namespace gsplat
