/*
 * fastgs/rasterization_api.h -- drop-in for the reference's fastgs rasterizer API, SURVEY.md 8 f4.
 *
 * Same namespace, struct, functions, argument order and return tuples as
 * /root/reference/fastgs/rasterization/include/rasterization_api.h:11-75, so that the reference's caller
 * (src/training/rasterization/fast_rasterizer_autograd.cpp:10-160) compiles and links against
 * libgsplat_b200.so unchanged.  Implemented in gaussian-splatting-cuda_b200/shim/FastGs.cpp on top
 * of the C ABI (include/gsb200.h: gsb_fastgs_*).
 *
 * What the opaque return values hold here (the caller only stores them between forward and backward):
 *   per_primitive_buffers   blend records, gradient moments, tile boxes / counts, the intersect plan
 *   per_tile_buffers        closed tile offsets, per-pixel last contributor
 *   per_instance_buffers    n_instances int32 primitive indices in (tile, depth) order
 *   per_bucket_buffers      empty: the backward needs no stored blend state
 *   n_visible_primitives    -1 (not needed; counting them would cost a second read-back)
 *   n_instances             number of (primitive, tile) instances
 *   n_buckets, *_selector   0
 * Launches go to the CURRENT torch CUDA stream (the reference uses the legacy default stream plus three blocking
 * cudaMemcpy read-backs, forward.cu:99-102,177); one stream synchronisation remains, to size per_instance_buffers.
 */
#pragma once

#include <torch/torch.h>
#include <tuple>

namespace fast_gs::rasterization {

    struct FastGSSettings {
        torch::Tensor cam_position;
        int active_sh_bases;
        int width;
        int height;
        float focal_x;
        float focal_y;
        float center_x;
        float center_y;
        float near_plane;
        float far_plane;
    };

    // image [3,H,W], alpha [1,H,W], per_primitive / per_tile / per_instance / per_bucket buffers,
    // n_visible_primitives, n_instances, n_buckets, primitive selector, instance selector
    std::tuple<torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, int, int, int, int, int>
    forward_wrapper(
        const torch::Tensor& means,                // [N,3]
        const torch::Tensor& scales_raw,           // [N,3] log-scales
        const torch::Tensor& rotations_raw,        // [N,4] (w,x,y,z), un-normalised
        const torch::Tensor& opacities_raw,        // [N,1] logits
        const torch::Tensor& sh_coefficients_0,    // [N,1,3]
        const torch::Tensor& sh_coefficients_rest, // [N,B-1,3]
        const torch::Tensor& w2c,                  // [4,4] (or [1,4,4])
        const torch::Tensor& cam_position,         // [3]
        const int active_sh_bases,
        const int width,
        const int height,
        const float focal_x,
        const float focal_y,
        const float center_x,
        const float center_y,
        const float near_plane,
        const float far_plane);

    // grad_means, grad_scales_raw, grad_rotations_raw, grad_opacities_raw, grad_sh_coefficients_0,
    // grad_sh_coefficients_rest, grad_w2c (undefined unless w2c.requires_grad())
    std::tuple<torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor>
    backward_wrapper(
        torch::Tensor& densification_info, // [2,N] accumulated in place, or empty
        const torch::Tensor& grad_image,
        const torch::Tensor& grad_alpha,
        const torch::Tensor& image,
        const torch::Tensor& alpha,
        const torch::Tensor& means,
        const torch::Tensor& scales_raw,
        const torch::Tensor& rotations_raw,
        const torch::Tensor& sh_coefficients_rest,
        const torch::Tensor& per_primitive_buffers,
        const torch::Tensor& per_tile_buffers,
        const torch::Tensor& per_instance_buffers,
        const torch::Tensor& per_bucket_buffers,
        const torch::Tensor& w2c,
        const torch::Tensor& cam_position,
        const int active_sh_bases,
        const int width,
        const int height,
        const float focal_x,
        const float focal_y,
        const float center_x,
        const float center_y,
        const float near_plane,
        const float far_plane,
        const int n_visible_primitives,
        const int n_instances,
        const int n_buckets,
        const int primitive_primitive_indices_selector,
        const int instance_primitive_indices_selector);

} // namespace fast_gs::rasterization
