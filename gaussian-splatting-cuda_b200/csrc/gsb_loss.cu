// gsb_loss.cu -- SURVEY.md 8(f2): the photometric loss of the training step and its gradient in ONE kernel.
//
// The reference computes (src/training/trainer.cpp:103-126)
//     image = clamp(render, 0, 1) permuted to CHW                      (rasterizer.cpp:401)
//     loss  = (1 - lambda) * l1_loss(image, gt) + lambda * (1 - fused_ssim(image, gt, "valid"))
// with fused_ssim = mean over the map cropped by 5 pixels per side of the 11x11 Gaussian-window SSIM evaluated with
// zero padding (src/training/kernels/ssim.cu:64-283, include/kernels/fused_ssim.cuh:27-117), and gets the gradient
// from a second stencil kernel over three stored derivative maps plus the autograd of clamp / permute / l1 / mean
// (ssim.cu:284-420).  Here one kernel reads the blend's [H,W,3] output and the target once and writes the loss and
// dLoss/d(render) -- exactly what gsb_raster_bwd consumes: per 16x16 tile (the blend's tiling) it stages a 36x36
// neighbourhood (tile + two halos), runs the separable 11-tap window twice (statistics on 26x26, then the three
// derivative maps back onto the tile), and folds in the L1 term, the crop, the means and the clamp mask.  No
// derivative maps, no permuted copies and no 2-million-element reductions go through HBM.
#include "gsb_common.cuh"

namespace gsb {

constexpr int kLossThreads = 128;
constexpr int kT = 16;           // tile side
constexpr int kHalo = 5;         // window radius
constexpr int kR1 = kT + 2 * kHalo;      // 26: pixels whose SSIM statistics the tile's gradient needs
constexpr int kR2 = kT + 4 * kHalo;      // 36: input neighbourhood

__constant__ float c_win[11] = {0.001028380084f, 0.007598758135f, 0.03600077213f, 0.1093606895f, 0.2130055377f,
                                0.2660117249f,   0.2130055377f,   0.1093606895f,  0.03600077213f, 0.007598758135f,
                                0.001028380084f}; // exp(-x^2 / (2 * 1.5^2)) normalised, x = -5..5

struct LossParams {
    uint32_t H, W;
    const float *renders;   // [H,W,3], unclamped
    const float *target;    // [3,H,W] (target_chw) or [H,W,3]
    int target_chw;
    float lambda;           // weight of the SSIM term (lambda_dssim)
    float grad_scale;       // upstream gradient of the scalar loss
    float *v_renders;       // [H,W,3] or null (evaluation)
    double *acc;            // [2] running sums (|x - y|, ssim), zeroed before the launch
    unsigned int *done;     // [1] CTA counter, zeroed before the launch
    float *loss_out;        // [3] loss, l1 mean, ssim mean
    uint32_t y_lo, y_hi, x_lo, x_hi; // valid (cropped) region of the SSIM map
};

__global__ void __launch_bounds__(kLossThreads) ssim_l1_kernel(const LossParams p) {
    __shared__ float sX[kR2][kR2 + 1], sY[kR2][kR2 + 1];  // clamped render / target, zero outside the image
    __shared__ float sH[5][kR2][kR1 + 1];                  // horizontal pass: X, X^2, Y, Y^2, XY
    __shared__ float sM[3][kR1][kR1 + 1];                  // chain * (dm/dmu1, dm/dsigma1^2, dm/dsigma12)
    __shared__ float sG[3][kR1][kT + 1];                   // horizontal pass of sM
    __shared__ float s_red[2][kLossThreads / 32];
    const int tid = threadIdx.x;
    const int tx0 = blockIdx.x * kT, ty0 = blockIdx.y * kT;
    const int H = (int)p.H, W = (int)p.W;
    const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
    const float n_valid = (float)(p.y_hi - p.y_lo) * (float)(p.x_hi - p.x_lo) * 3.0f;
    const float chain = -p.lambda * p.grad_scale / n_valid;
    const float l1_w = (1.0f - p.lambda) * p.grad_scale / ((float)H * (float)W * 3.0f);
    float l1_sum = 0.f, ssim_sum = 0.f;

    for (int c = 0; c < 3; ++c) {
        // (0) neighbourhood of the tile
        for (int i = tid; i < kR2 * kR2; i += kLossThreads) {
            const int r = i / kR2, q = i - r * kR2;
            const int gy = ty0 + r - 2 * kHalo, gx = tx0 + q - 2 * kHalo;
            float X = 0.f, Y = 0.f;
            if (gy >= 0 && gy < H && gx >= 0 && gx < W) {
                X = fminf(fmaxf(p.renders[((size_t)gy * W + gx) * 3 + c], 0.f), 1.f);
                Y = p.target_chw ? p.target[((size_t)c * H + gy) * W + gx] : p.target[((size_t)gy * W + gx) * 3 + c];
            }
            sX[r][q] = X;
            sY[r][q] = Y;
        }
        __syncthreads();
        // (1) horizontal window on 36 rows x 26 columns, a strip of 13 outputs per work item.  The five windowed
        //     quantities ride on packed fp32 pairs: (X, Y), (X^2, Y^2) and a scalar XY -- three FMA issues per tap
        for (int it = tid; it < kR2 * 2; it += kLossThreads) {
            const int r = it >> 1, q0 = (it & 1) * 13;
            f2 xy[23], sq[23];
            float pr[23];
#pragma unroll
            for (int k = 0; k < 23; ++k) {
                const float xv = sX[r][q0 + k], yv = sY[r][q0 + k];
                xy[k] = f2_make(xv, yv);
                sq[k] = f2_mul(xy[k], xy[k]);
                pr[k] = xv * yv;
            }
#pragma unroll
            for (int o = 0; o < 13; ++o) {
                f2 a02 = f2_bc(0.f), a13 = f2_bc(0.f);
                float a4 = 0.f;
#pragma unroll
                for (int k = 0; k < 11; ++k) {
                    const float w = c_win[k];
                    a02 = f2_fma(f2_bc(w), xy[o + k], a02);
                    a13 = f2_fma(f2_bc(w), sq[o + k], a13);
                    a4 += w * pr[o + k];
                }
                sH[0][r][q0 + o] = f2_lo(a02); sH[2][r][q0 + o] = f2_hi(a02);
                sH[1][r][q0 + o] = f2_lo(a13); sH[3][r][q0 + o] = f2_hi(a13);
                sH[4][r][q0 + o] = a4;
            }
        }
        __syncthreads();
        // (2) vertical window -> statistics of the 26x26 pixels -> SSIM value and the three derivative maps
        for (int it = tid; it < kR1 * 2; it += kLossThreads) {
            const int q = it >> 1, r0 = (it & 1) * 13;
            float acc5[5][13];
            {
                f2 v02[23], v13[23];
                float v4[23];
#pragma unroll
                for (int k = 0; k < 23; ++k) {
                    v02[k] = f2_make(sH[0][r0 + k][q], sH[2][r0 + k][q]);
                    v13[k] = f2_make(sH[1][r0 + k][q], sH[3][r0 + k][q]);
                    v4[k] = sH[4][r0 + k][q];
                }
#pragma unroll
                for (int o = 0; o < 13; ++o) {
                    f2 a02 = f2_bc(0.f), a13 = f2_bc(0.f);
                    float a4 = 0.f;
#pragma unroll
                    for (int k = 0; k < 11; ++k) {
                        const float w = c_win[k];
                        a02 = f2_fma(f2_bc(w), v02[o + k], a02);
                        a13 = f2_fma(f2_bc(w), v13[o + k], a13);
                        a4 += w * v4[o + k];
                    }
                    acc5[0][o] = f2_lo(a02); acc5[2][o] = f2_hi(a02);
                    acc5[1][o] = f2_lo(a13); acc5[3][o] = f2_hi(a13);
                    acc5[4][o] = a4;
                }
            }
#pragma unroll
            for (int o = 0; o < 13; ++o) {
                const int r = r0 + o;
                const int gy = ty0 + r - kHalo, gx = tx0 + q - kHalo;
                float m0 = 0.f, m1 = 0.f, m2 = 0.f;
                const bool in_img = gy >= 0 && gy < H && gx >= 0 && gx < W;
                const bool in_valid = gy >= (int)p.y_lo && gy < (int)p.y_hi && gx >= (int)p.x_lo && gx < (int)p.x_hi;
                if (in_img && in_valid) {
                    const float mu1 = acc5[0][o], mu2 = acc5[2][o];
                    const float mu1_sq = mu1 * mu1, mu2_sq = mu2 * mu2;
                    const float sigma1_sq = acc5[1][o] - mu1_sq, sigma2_sq = acc5[3][o] - mu2_sq;
                    const float sigma12 = acc5[4][o] - mu1 * mu2;
                    const float A = mu1_sq + mu2_sq + C1, B = sigma1_sq + sigma2_sq + C2;
                    const float Cc = 2.f * mu1 * mu2 + C1, D = 2.f * sigma12 + C2;
                    const float iAB = 1.0f / (A * B);
                    // ssim.cu:253-263
                    m0 = chain * ((mu2 * 2.f * D) * iAB - (mu2 * 2.f * Cc) * iAB - (mu1 * 2.f * Cc * D) * iAB / A +
                                  (mu1 * 2.f * Cc * D) * iAB / B);
                    m1 = chain * ((-Cc * D) * iAB / B);
                    m2 = chain * ((2.f * Cc) * iAB);
                    // the tile's own pixels add their SSIM value to the mean
                    if (r >= kHalo && r < kHalo + kT && q >= kHalo && q < kHalo + kT) ssim_sum += (Cc * D) * iAB;
                }
                sM[0][r][q] = m0; sM[1][r][q] = m1; sM[2][r][q] = m2;
            }
        }
        __syncthreads();
        // (3) horizontal window of the derivative maps: 26 rows x 16 columns, strips of 8; maps 0 and 1 packed
        for (int it = tid; it < kR1 * 2; it += kLossThreads) {
            const int r = it >> 1, q0 = (it & 1) * 8;
            f2 v01[18];
            float v2[18];
#pragma unroll
            for (int k = 0; k < 18; ++k) {
                v01[k] = f2_make(sM[0][r][q0 + k], sM[1][r][q0 + k]);
                v2[k] = sM[2][r][q0 + k];
            }
#pragma unroll
            for (int o = 0; o < 8; ++o) {
                f2 a01 = f2_bc(0.f);
                float a2 = 0.f;
#pragma unroll
                for (int k = 0; k < 11; ++k) {
                    a01 = f2_fma(f2_bc(c_win[k]), v01[o + k], a01);
                    a2 += c_win[k] * v2[o + k];
                }
                sG[0][r][q0 + o] = f2_lo(a01); sG[1][r][q0 + o] = f2_hi(a01); sG[2][r][q0 + o] = a2;
            }
        }
        __syncthreads();
        // (4) vertical window + L1 term + clamp mask -> dLoss / d(render): 16 columns x 2 strips of 8 rows
        for (int it = tid; it < kT * 2; it += kLossThreads) {
            const int q = it >> 1, r0 = (it & 1) * 8;
            float s3[3][8];
            {
                f2 v01[18];
                float v2[18];
#pragma unroll
                for (int k = 0; k < 18; ++k) {
                    v01[k] = f2_make(sG[0][r0 + k][q], sG[1][r0 + k][q]);
                    v2[k] = sG[2][r0 + k][q];
                }
#pragma unroll
                for (int o = 0; o < 8; ++o) {
                    f2 a01 = f2_bc(0.f);
                    float a2 = 0.f;
#pragma unroll
                    for (int k = 0; k < 11; ++k) {
                        a01 = f2_fma(f2_bc(c_win[k]), v01[o + k], a01);
                        a2 += c_win[k] * v2[o + k];
                    }
                    s3[0][o] = f2_lo(a01); s3[1][o] = f2_hi(a01); s3[2][o] = a2;
                }
            }
#pragma unroll
            for (int o = 0; o < 8; ++o) {
                const int gy = ty0 + r0 + o, gx = tx0 + q;
                if (gy < H && gx < W) {
                    const float p1 = sX[r0 + o + 2 * kHalo][q + 2 * kHalo], p2 = sY[r0 + o + 2 * kHalo][q + 2 * kHalo];
                    const float d = p1 - p2;
                    l1_sum += fabsf(d);
                    if (p.v_renders) {
                        float g = s3[0][o] + (2.f * p1) * s3[1][o] + p2 * s3[2][o]; // ssim.cu:411
                        g += l1_w * (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f));
                        const float raw = p.renders[((size_t)gy * W + gx) * 3 + c];
                        p.v_renders[((size_t)gy * W + gx) * 3 + c] = (raw >= 0.f && raw <= 1.f) ? g : 0.f;
                    }
                }
            }
        }
        __syncthreads();
    }

    // loss: CTA sums -> two double atomics; the last CTA turns the sums into the scalar
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        l1_sum += __shfl_xor_sync(0xffffffffu, l1_sum, o);
        ssim_sum += __shfl_xor_sync(0xffffffffu, ssim_sum, o);
    }
    if ((tid & 31) == 0) { s_red[0][tid >> 5] = l1_sum; s_red[1][tid >> 5] = ssim_sum; }
    __syncthreads();
    if (tid == 0) {
        double a = 0.0, b = 0.0;
        for (int w = 0; w < kLossThreads / 32; ++w) { a += (double)s_red[0][w]; b += (double)s_red[1][w]; }
        atomicAdd(&p.acc[0], a);
        atomicAdd(&p.acc[1], b);
        __threadfence();
        const unsigned int ticket = atomicAdd(p.done, 1u);
        if (ticket == gridDim.x * gridDim.y - 1) {
            __threadfence();
            const double l1 = *((volatile double *)&p.acc[0]) / ((double)H * (double)W * 3.0);
            const double ss = *((volatile double *)&p.acc[1]) / (double)n_valid;
            p.loss_out[0] = (float)((1.0 - (double)p.lambda) * l1 + (double)p.lambda * (1.0 - ss));
            p.loss_out[1] = (float)l1;
            p.loss_out[2] = (float)ss;
        }
    }
}

} // namespace gsb

extern "C" size_t gsb_ssim_l1_workspace(void) { return 256; }

extern "C" int gsb_ssim_l1(uint32_t image_width, uint32_t image_height, const float *renders, const float *target,
                           int target_chw, float lambda_dssim, float grad_scale, float *v_renders, float *loss_out,
                           void *workspace, size_t workspace_bytes, gsb_stream_t stream) {
    using namespace gsb;
    if (image_width == 0 || image_height == 0) return GSB_E_INVALID;
    if (!renders || !target || !loss_out) return GSB_E_INVALID;
    if (!workspace || (reinterpret_cast<uintptr_t>(workspace) & 255) || workspace_bytes < gsb_ssim_l1_workspace())
        return GSB_E_WORKSPACE;
    cudaStream_t s = as_stream(stream);
    GSB_CUDA_TRY(cudaMemsetAsync(workspace, 0, 64, s));
    LossParams p;
    p.H = image_height; p.W = image_width;
    p.renders = renders; p.target = target; p.target_chw = target_chw;
    p.lambda = lambda_dssim; p.grad_scale = grad_scale; p.v_renders = v_renders;
    p.acc = reinterpret_cast<double *>(workspace);
    p.done = reinterpret_cast<unsigned int *>(reinterpret_cast<char *>(workspace) + 32);
    p.loss_out = loss_out;
    // fused_ssim(..., "valid"): the map is cropped by 5 per side only when both sides exceed 10 (fused_ssim.cuh:63-67)
    const bool crop = image_height > 10 && image_width > 10;
    p.y_lo = crop ? 5 : 0; p.y_hi = crop ? image_height - 5 : image_height;
    p.x_lo = crop ? 5 : 0; p.x_hi = crop ? image_width - 5 : image_width;
    const dim3 grid((image_width + kT - 1) / kT, (image_height + kT - 1) / kT);
    {
        ProfScope ps(v_renders ? "ssim_l1_bwd" : "ssim_l1_fwd", s);
        ssim_l1_kernel<<<grid, kLossThreads, 0, s>>>(p);
    }
    GSB_LAUNCH_CHECK();
    return GSB_OK;
}
