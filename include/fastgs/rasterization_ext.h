/*
 * fastgs/rasterization_ext.h -- one addition to the reference's fastgs API (include/fastgs/rasterization_api.h):
 * a forward that never reads a device value on the host.
 *
 * forward_wrapper must return an exactly-sized per_instance_buffers tensor and `n_instances` as an int
 * (rasterization_api.h:25-44), which forces a stream synchronisation per frame.  forward_capacity takes the size of the
 * instance buffer from the caller instead (keep ~1.25 x the last count), leaves the count on the device and can therefore
 * be captured in a CUDA graph together with the loss, the backward and the optimizer step
 * (gaussian-splatting-cuda_b200/training.py: GraphedFastGsTrainStep).  Instances beyond the capacity are dropped (every
 * tile range is clipped); compare n_instances_dev with the capacity afterwards and re-run on overflow.
 * backward_wrapper is used unchanged: pass instance_capacity as its n_instances argument.
 */
#pragma once

#include <torch/torch.h>
#include <tuple>

namespace fast_gs::rasterization {

    // image [3,H,W], alpha [1,H,W], per_primitive / per_tile / per_instance buffers, n_instances_dev (int64 [1], device)
    std::tuple<torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor>
    forward_capacity(
        const torch::Tensor& means,
        const torch::Tensor& scales_raw,
        const torch::Tensor& rotations_raw,
        const torch::Tensor& opacities_raw,
        const torch::Tensor& sh_coefficients_0,
        const torch::Tensor& sh_coefficients_rest,
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
        const int64_t instance_capacity);

} // namespace fast_gs::rasterization
