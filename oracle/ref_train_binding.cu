// ref_train_binding.cu -- exposes the REFERENCE's own training-step kernels next to its gsplat operators
// (torch.ops.gsplat_ref.*): the fused SSIM forward / backward of /root/reference/src/training/kernels/ssim.cu (compiled
// unmodified by oracle/build_ref.py) and the Adam kernel of /root/reference/fastgs/optimizer/include/adam_kernels.cuh
// (header-only; launched here with the reference's block size, fastgs/optimizer/include/optimizer_config.h).
// TEST INFRASTRUCTURE ONLY: the parity pin of SURVEY.md 8 f2 / f3 and the same-box baseline of the training step.
#include <torch/library.h>
#include <torch/torch.h>

#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>

#include "adam_kernels.cuh"

// defined in the reference's ssim.cu
std::tuple<torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor> fusedssim(float C1, float C2, torch::Tensor &img1,
                                                                                 torch::Tensor &img2, bool train);
torch::Tensor fusedssim_backward(float C1, float C2, torch::Tensor &img1, torch::Tensor &img2, torch::Tensor &dL_dmap,
                                 torch::Tensor &dm_dmu1, torch::Tensor &dm_dsigma1_sq, torch::Tensor &dm_dsigma12);

namespace {
using at::Tensor;

std::tuple<Tensor, Tensor, Tensor, Tensor> ssim_fwd(double C1, double C2, Tensor img1, Tensor img2, bool train) {
    return fusedssim((float)C1, (float)C2, img1, img2, train);
}
Tensor ssim_bwd(double C1, double C2, Tensor img1, Tensor img2, Tensor dL_dmap, Tensor dm_dmu1, Tensor dm_dsigma1_sq,
                Tensor dm_dsigma12) {
    return fusedssim_backward((float)C1, (float)C2, img1, img2, dL_dmap, dm_dmu1, dm_dsigma1_sq, dm_dsigma12);
}
// fastgs/optimizer/src/adam.cu:10-36 (block_size_adam_step = 256, default stream like the reference)
void adam_step(Tensor param, Tensor exp_avg, Tensor exp_avg_sq, const Tensor &grad, double lr, double beta1, double beta2,
               double eps, double bc1_rcp, double bc2_sqrt_rcp) {
    const c10::cuda::CUDAGuard guard(param.device());
    const int n = (int)param.numel();
    if (n == 0) return;
    fast_gs::optimizer::kernels::adam::adam_step_cu<<<(n + 255) / 256, 256, 0, c10::cuda::getCurrentCUDAStream()>>>(
        param.data_ptr<float>(), exp_avg.data_ptr<float>(), exp_avg_sq.data_ptr<float>(), grad.data_ptr<float>(), n,
        (float)lr, (float)beta1, (float)beta2, (float)eps, (float)bc1_rcp, (float)bc2_sqrt_rcp);
}
} // namespace

#ifndef REF_NS
#define REF_NS gsplat_ref
#endif
#define REF_TORCH_LIBRARY_FRAGMENT(ns, m) TORCH_LIBRARY_FRAGMENT(ns, m)

REF_TORCH_LIBRARY_FRAGMENT(REF_NS, m) {
    m.def("fusedssim", &ssim_fwd);
    m.def("fusedssim_backward", &ssim_bwd);
    m.def("adam_step(Tensor(a!) param, Tensor(b!) exp_avg, Tensor(c!) exp_avg_sq, Tensor grad, float lr, float beta1, "
          "float beta2, float eps, float bc1_rcp, float bc2_sqrt_rcp) -> ()",
          &adam_step);
}
