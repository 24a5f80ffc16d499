/*
 * fastgs/rasterization_api.h -- drop-in for the reference's fastgs rasterizer API, SURVEY.md 8 f4.
 *
 * Declares, with identical types, order and return tuples, the struct and the two functions of
 * /root/reference/fastgs/rasterization/include/rasterization_api.h:11-75, so that the reference's caller
 * (src/training/rasterization/fast_rasterizer_autograd.cpp:10-160) compiles and links against libgsplat_b200.so
 * unchanged.  Implemented in gaussian-splatting-cuda_b200/shim/FastGs.cpp on top of the C ABI
 * (include/gsb200.h: gsb_fastgs_*).  tests/test_link_reference_caller.py and tests/test_gpu_fastgs.py build and run that
 * caller against this header.
 *
 * What the opaque return values hold here (the caller only stores them between forward and backward):
 *   buffer 1 ("per primitive")  blend records, gradient moments, tile boxes / exact counts / tile-test masks, the plan
 *   buffer 2 ("per tile")       closed tile offsets, per-pixel last contributor
 *   buffer 3 ("per instance")   n_instances int32 primitive indices in (tile, depth) order
 *   buffer 4 ("per bucket")     empty: the backward needs no stored blend state
 *   the five ints               -1 (visible primitives are not counted: it would cost a second read-back),
 *                               n_instances, 0, 0, 0
 * Launches go to the CURRENT torch CUDA stream (the reference uses the legacy default stream plus three blocking
 * cudaMemcpy read-backs, forward.cu:99-102,177); one stream synchronisation remains, to size buffer 3
 * (rasterization_ext.h has the variant without it).
 */
#pragma once

#include <torch/torch.h>
#include <tuple>

namespace fast_gs::rasterization {

    using Tensor = torch::Tensor;

    // Camera / raster settings as the caller's autograd node carries them (same members, same order as the reference).
    struct FastGSSettings {
        Tensor cam_position; // [3]
        int active_sh_bases; // 1, 4, 9 or 16
        int width, height;
        float focal_x, focal_y;
        float center_x, center_y;
        float near_plane, far_plane;
    };

    // image [3,H,W], alpha [1,H,W], the four opaque buffers, the five ints.
    using ForwardResult = std::tuple<Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, int, int, int, int, int>;
    // gradients of positions, log-scales, quaternions, opacity logits, sh0, shN, and of w2c (undefined unless it requires one).
    using BackwardResult = std::tuple<Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor>;

    // RAW parameters in: positions [N,3], log-scales [N,3], un-normalised (w,x,y,z) quaternions [N,4], opacity logits
    // [N,1], SH band 0 [N,1,3], remaining bands [N,B-1,3]; w2c [4,4] (or [1,4,4]); camera position [3].
    ForwardResult forward_wrapper(const Tensor& positions, const Tensor& log_scales, const Tensor& quaternions,
                                  const Tensor& opacity_logits, const Tensor& sh_band0, const Tensor& sh_rest,
                                  const Tensor& world_to_camera, const Tensor& camera_position, const int active_sh_bases,
                                  const int width, const int height, const float focal_x, const float focal_y,
                                  const float center_x, const float center_y, const float near_plane, const float far_plane);

    // densification_info: [2,N] accumulated in place, or empty.  The buffers and ints are forward_wrapper's.
    BackwardResult backward_wrapper(Tensor& densification_info, const Tensor& grad_image, const Tensor& grad_alpha,
                                    const Tensor& image, const Tensor& alpha, const Tensor& positions, const Tensor& log_scales,
                                    const Tensor& quaternions, const Tensor& sh_rest, const Tensor& buffer_primitives,
                                    const Tensor& buffer_tiles, const Tensor& buffer_instances, const Tensor& buffer_buckets,
                                    const Tensor& world_to_camera, const Tensor& camera_position, const int active_sh_bases,
                                    const int width, const int height, const float focal_x, const float focal_y,
                                    const float center_x, const float center_y, const float near_plane, const float far_plane,
                                    const int n_visible_primitives, const int n_instances, const int n_buckets,
                                    const int primitive_selector, const int instance_selector);

} // namespace fast_gs::rasterization
