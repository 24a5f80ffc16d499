// Cameras.h -- drop-in replacement for the reference's gsplat/Cameras.h (public value types).
//
// Same names, enumerator order, field order and defaults as /root/reference/gsplat/Cameras.h:16-61:
// both types cross the gsplat:: API by value, so their layout is ABI.  GLM is not needed here.
#pragma once

#include <cstdint>
#include <torch/torch.h>

// Sensor read-out direction (Cameras.h:16-22).  GLOBAL is what every caller in src/ passes
// (rasterizer_autograd.cpp:234,311,369).  Numeric values equal GSB_SHUTTER_*.
enum class ShutterType {
    ROLLING_TOP_TO_BOTTOM,
    ROLLING_LEFT_TO_RIGHT,
    ROLLING_BOTTOM_TO_TOP,
    ROLLING_RIGHT_TO_LEFT,
    GLOBAL
};

// Sigma-point parameters of the unscented transform (Cameras.h:27-44; Wan & van der Merwe 2000).
struct UnscentedTransformParameters {
    float alpha = 0.1;
    float beta = 2.f;
    float kappa = 0.f;
    // a projected sigma point may fall this fraction of the image size outside the image
    float in_image_margin_factor = 0.1f;
    // true: every sigma point must project validly; false: one is enough
    bool require_all_sigma_points_valid = true;

    // The autograd node stores the parameters as a 5-float CPU tensor (rasterizer_autograd.cpp:290,364).
    torch::Tensor to_tensor() const {
        return torch::tensor({alpha, beta, kappa, in_image_margin_factor,
                              static_cast<float>(require_all_sigma_points_valid)},
                             torch::TensorOptions().dtype(torch::kFloat32));
    }

    static UnscentedTransformParameters from_tensor(const torch::Tensor& tensor) {
        TORCH_CHECK(tensor.dim() == 1 && tensor.size(0) == 5,
                    "UnscentedTransformParameters must be a 1D tensor of size 5");
        return UnscentedTransformParameters{
            tensor[0].item<float>(), tensor[1].item<float>(),
            tensor[2].item<float>(), tensor[3].item<float>(),
            tensor[4].item<bool>()};
    }
};
