// gsb_sh.cu -- a3/a4: spherical-harmonics colour evaluation and its gradient.
//
// Implements gsplat::spherical_harmonics_fwd / _bwd (reference:
// gsplat/SphericalHarmonicsCUDA.cu:21-110 fwd, :113-371 vjp, :374-481 kernels).
// Real SH basis in Sloan's ordering, degrees 0..4, constants as in the reference.
//
// Design differences from the reference (same results):
//  * one thread per element evaluates the basis ONCE and produces all three channels (the
//    reference runs one thread per channel and re-evaluates the basis three times, reading the
//    coefficient rows at a 12-byte stride);
//  * the backward writes every v_coeffs / v_dirs element itself (zeros where the reference
//    relies on a 192 B/Gaussian memset) and needs no atomics for v_dirs because the three
//    channels live in one thread.
#include "gsb_common.cuh"

namespace gsb {

constexpr int kShThreads = 128;

// Basis values b[0..nb) for the unit direction (x,y,z).  MAXB is the compile-time bound.
template <int DEG>
__device__ __forceinline__ void sh_basis(float x, float y, float z, float *b) {
    b[0] = 0.2820947917738781f;
    if constexpr (DEG >= 1) {
        b[1] = -0.48860251190292f * y;
        b[2] = 0.48860251190292f * z;
        b[3] = -0.48860251190292f * x;
    }
    if constexpr (DEG >= 2) {
        const float z2 = z * z;
        const float fTmp0B = -1.092548430592079f * z;
        const float fC1 = x * x - y * y;
        const float fS1 = 2.f * x * y;
        b[6] = 0.9461746957575601f * z2 - 0.3153915652525201f;
        b[7] = fTmp0B * x;
        b[5] = fTmp0B * y;
        b[8] = 0.5462742152960395f * fC1;
        b[4] = 0.5462742152960395f * fS1;
        if constexpr (DEG >= 3) {
            const float fTmp0C = -2.285228997322329f * z2 + 0.4570457994644658f;
            const float fTmp1B = 1.445305721320277f * z;
            const float fC2 = x * fC1 - y * fS1;
            const float fS2 = x * fS1 + y * fC1;
            b[12] = z * (1.865881662950577f * z2 - 1.119528997770346f);
            b[13] = fTmp0C * x;
            b[11] = fTmp0C * y;
            b[14] = fTmp1B * fC1;
            b[10] = fTmp1B * fS1;
            b[15] = -0.5900435899266435f * fC2;
            b[9] = -0.5900435899266435f * fS2;
            if constexpr (DEG >= 4) {
                const float fTmp0D = z * (-4.683325804901025f * z2 + 2.007139630671868f);
                const float fTmp1C = 3.31161143515146f * z2 - 0.47308734787878f;
                const float fTmp2B = -1.770130769779931f * z;
                const float fC3 = x * fC2 - y * fS2;
                const float fS3 = x * fS2 + y * fC2;
                b[20] = 1.984313483298443f * z * b[12] - 1.006230589874905f * b[6];
                b[21] = fTmp0D * x;
                b[19] = fTmp0D * y;
                b[22] = fTmp1C * fC1;
                b[18] = fTmp1C * fS1;
                b[23] = fTmp2B * fC2;
                b[17] = fTmp2B * fS2;
                b[24] = 0.6258357354491763f * fC3;
                b[16] = 0.6258357354491763f * fS3;
            }
        }
    }
}

// Gradient of sum_k b_k(x,y,z) * w_k with respect to the unit direction, where w_k is the
// per-coefficient weight sum_c v_colour_c * coeff[k][c] (SphericalHarmonicsCUDA.cu:137-352).
template <int DEG>
__device__ __forceinline__ void sh_basis_vjp(float x, float y, float z, const float *w, float &vx, float &vy,
                                             float &vz) {
    vx = vy = vz = 0.f;
    if constexpr (DEG >= 1) {
        vx += -0.48860251190292f * w[3];
        vy += -0.48860251190292f * w[1];
        vz += 0.48860251190292f * w[2];
    }
    if constexpr (DEG >= 2) {
        const float z2 = z * z;
        const float fTmp0B = -1.092548430592079f * z;
        const float fC1 = x * x - y * y;
        const float fS1 = 2.f * x * y;
        const float fTmp0B_z = -1.092548430592079f;
        const float fC1_x = 2.f * x, fC1_y = -2.f * y, fS1_x = 2.f * y, fS1_y = 2.f * x;
        const float pSH6_z = 2.f * 0.9461746957575601f * z;
        vx += 0.5462742152960395f * fS1_x * w[4] + 0.5462742152960395f * fC1_x * w[8] + fTmp0B * w[7];
        vy += 0.5462742152960395f * fS1_y * w[4] + 0.5462742152960395f * fC1_y * w[8] + fTmp0B * w[5];
        vz += pSH6_z * w[6] + fTmp0B_z * x * w[7] + fTmp0B_z * y * w[5];
        if constexpr (DEG >= 3) {
            const float fTmp0C = -2.285228997322329f * z2 + 0.4570457994644658f;
            const float fTmp1B = 1.445305721320277f * z;
            const float fC2 = x * fC1 - y * fS1;
            const float fS2 = x * fS1 + y * fC1;
            const float fTmp0C_z = -2.285228997322329f * 2.f * z;
            const float fTmp1B_z = 1.445305721320277f;
            const float fC2_x = fC1 + x * fC1_x - y * fS1_x;
            const float fC2_y = x * fC1_y - fS1 - y * fS1_y;
            const float fS2_x = fS1 + x * fS1_x + y * fC1_x;
            const float fS2_y = x * fS1_y + fC1 + y * fC1_y;
            const float pSH12 = z * (1.865881662950577f * z2 - 1.119528997770346f);
            const float pSH12_z = 3.f * 1.865881662950577f * z2 - 1.119528997770346f;
            vx += -0.5900435899266435f * fS2_x * w[9] + -0.5900435899266435f * fC2_x * w[15] +
                  fTmp1B * fS1_x * w[10] + fTmp1B * fC1_x * w[14] + fTmp0C * w[13];
            vy += -0.5900435899266435f * fS2_y * w[9] + -0.5900435899266435f * fC2_y * w[15] +
                  fTmp1B * fS1_y * w[10] + fTmp1B * fC1_y * w[14] + fTmp0C * w[11];
            vz += pSH12_z * w[12] + fTmp0C_z * x * w[13] + fTmp0C_z * y * w[11] + fTmp1B_z * fC1 * w[14] +
                  fTmp1B_z * fS1 * w[10];
            if constexpr (DEG >= 4) {
                const float fTmp0D = z * (-4.683325804901025f * z2 + 2.007139630671868f);
                const float fTmp1C = 3.31161143515146f * z2 - 0.47308734787878f;
                const float fTmp2B = -1.770130769779931f * z;
                const float fTmp0D_z = 3.f * -4.683325804901025f * z2 + 2.007139630671868f;
                const float fTmp1C_z = 2.f * 3.31161143515146f * z;
                const float fTmp2B_z = -1.770130769779931f;
                const float fC3_x = fC2 + x * fC2_x - y * fS2_x;
                const float fC3_y = x * fC2_y - fS2 - y * fS2_y;
                const float fS3_x = fS2 + y * fC2_x + x * fS2_x;
                const float fS3_y = x * fS2_y + fC2 + y * fC2_y;
                const float pSH20_z = 1.984313483298443f * (pSH12 + z * pSH12_z) + -1.006230589874905f * pSH6_z;
                vx += 0.6258357354491763f * fS3_x * w[16] + 0.6258357354491763f * fC3_x * w[24] +
                      fTmp2B * fS2_x * w[17] + fTmp2B * fC2_x * w[23] + fTmp1C * fS1_x * w[18] +
                      fTmp1C * fC1_x * w[22] + fTmp0D * w[21];
                vy += 0.6258357354491763f * fS3_y * w[16] + 0.6258357354491763f * fC3_y * w[24] +
                      fTmp2B * fS2_y * w[17] + fTmp2B * fC2_y * w[23] + fTmp1C * fS1_y * w[18] +
                      fTmp1C * fC1_y * w[22] + fTmp0D * w[19];
                vz += pSH20_z * w[20] + fTmp0D_z * x * w[21] + fTmp0D_z * y * w[19] + fTmp1C_z * fC1 * w[22] +
                      fTmp1C_z * fS1 * w[18] + fTmp2B_z * fC2 * w[23] + fTmp2B_z * fS2 * w[17];
            }
        }
    }
}

template <int DEG>
__global__ void __launch_bounds__(kShThreads) sh_fwd_kernel(uint32_t M, uint32_t K, const float *__restrict__ dirs,
                                                            const float *__restrict__ coeffs,
                                                            const uint8_t *__restrict__ masks,
                                                            float *__restrict__ colors) {
    constexpr int NB = (DEG + 1) * (DEG + 1);
    const uint32_t e = blockIdx.x * kShThreads + threadIdx.x;
    if (e >= M) return;
    if (masks != nullptr && !masks[e]) return; // masked rows stay untouched (:391-393)
    float b[NB];
    float x = dirs[(size_t)e * 3], y = dirs[(size_t)e * 3 + 1], z = dirs[(size_t)e * 3 + 2];
    if constexpr (DEG >= 1) {
        const float inorm = rsqrtf(x * x + y * y + z * z);
        x *= inorm; y *= inorm; z *= inorm;
    }
    sh_basis<DEG>(x, y, z, b);
    const float *c = coeffs + (size_t)e * K * 3;
    float r0 = 0.f, r1 = 0.f, r2 = 0.f;
    if ((K * 3) % 4 == 0 && NB * 3 % 4 == 0) {
        // 16-byte aligned rows: 128-bit loads (K = 16, deg 3: twelve LDG.128 per element)
        const float4 *c4 = reinterpret_cast<const float4 *>(c);
        float f[NB * 3];
#pragma unroll
        for (int i = 0; i < NB * 3 / 4; ++i) {
            const float4 v = __ldg(c4 + i);
            f[i * 4] = v.x; f[i * 4 + 1] = v.y; f[i * 4 + 2] = v.z; f[i * 4 + 3] = v.w;
        }
#pragma unroll
        for (int k = 0; k < NB; ++k) {
            r0 += b[k] * f[k * 3]; r1 += b[k] * f[k * 3 + 1]; r2 += b[k] * f[k * 3 + 2];
        }
    } else {
#pragma unroll
        for (int k = 0; k < NB; ++k) {
            r0 += b[k] * __ldg(c + k * 3);
            r1 += b[k] * __ldg(c + k * 3 + 1);
            r2 += b[k] * __ldg(c + k * 3 + 2);
        }
    }
    colors[(size_t)e * 3] = r0;
    colors[(size_t)e * 3 + 1] = r1;
    colors[(size_t)e * 3 + 2] = r2;
}

template <int DEG>
__global__ void __launch_bounds__(kShThreads) sh_bwd_kernel(uint32_t M, uint32_t K, const float *__restrict__ dirs,
                                                            const float *__restrict__ coeffs,
                                                            const uint8_t *__restrict__ masks,
                                                            const float *__restrict__ v_colors,
                                                            float *__restrict__ v_coeffs,
                                                            float *__restrict__ v_dirs) {
    constexpr int NB = (DEG + 1) * (DEG + 1);
    const uint32_t e = blockIdx.x * kShThreads + threadIdx.x;
    if (e >= M) return;
    float *vc = v_coeffs + (size_t)e * K * 3;
    const bool active = (masks == nullptr) || masks[e];
    if (!active) {
        for (uint32_t i = 0; i < K * 3; ++i) vc[i] = 0.f;
        if (v_dirs) { v_dirs[(size_t)e * 3] = 0.f; v_dirs[(size_t)e * 3 + 1] = 0.f; v_dirs[(size_t)e * 3 + 2] = 0.f; }
        return;
    }
    float b[NB];
    const float dx = dirs[(size_t)e * 3], dy = dirs[(size_t)e * 3 + 1], dz = dirs[(size_t)e * 3 + 2];
    float x = dx, y = dy, z = dz, inorm = 1.f;
    if constexpr (DEG >= 1) {
        inorm = rsqrtf(dx * dx + dy * dy + dz * dz);
        x *= inorm; y *= inorm; z *= inorm;
    }
    sh_basis<DEG>(x, y, z, b);
    const float g0 = v_colors[(size_t)e * 3], g1 = v_colors[(size_t)e * 3 + 1], g2 = v_colors[(size_t)e * 3 + 2];
    const float *c = coeffs + (size_t)e * K * 3;
    float w[NB];
    if ((K * 3) % 4 == 0 && NB * 3 % 4 == 0) {
        float4 *vc4 = reinterpret_cast<float4 *>(vc);
        const float4 *c4 = reinterpret_cast<const float4 *>(c);
        float f[NB * 3], o[NB * 3];
        if (v_dirs) {
#pragma unroll
            for (int i = 0; i < NB * 3 / 4; ++i) {
                const float4 v = __ldg(c4 + i);
                f[i * 4] = v.x; f[i * 4 + 1] = v.y; f[i * 4 + 2] = v.z; f[i * 4 + 3] = v.w;
            }
        }
#pragma unroll
        for (int k = 0; k < NB; ++k) {
            o[k * 3] = b[k] * g0; o[k * 3 + 1] = b[k] * g1; o[k * 3 + 2] = b[k] * g2;
            if (v_dirs) w[k] = g0 * f[k * 3] + g1 * f[k * 3 + 1] + g2 * f[k * 3 + 2];
        }
#pragma unroll
        for (int i = 0; i < NB * 3 / 4; ++i) vc4[i] = make_float4(o[i * 4], o[i * 4 + 1], o[i * 4 + 2], o[i * 4 + 3]);
    } else {
#pragma unroll
        for (int k = 0; k < NB; ++k) {
            vc[k * 3] = b[k] * g0; vc[k * 3 + 1] = b[k] * g1; vc[k * 3 + 2] = b[k] * g2;
            if (v_dirs) w[k] = g0 * __ldg(c + k * 3) + g1 * __ldg(c + k * 3 + 1) + g2 * __ldg(c + k * 3 + 2);
        }
    }
    for (uint32_t i = NB * 3; i < K * 3; ++i) vc[i] = 0.f; // inactive degrees: zero (reference: memset)
    if (v_dirs) {
        float vx = 0.f, vy = 0.f, vz = 0.f;
        if constexpr (DEG >= 1) {
            sh_basis_vjp<DEG>(x, y, z, w, vx, vy, vz);
            const float d = vx * x + vy * y + vz * z;
            vx = (vx - d * x) * inorm; vy = (vy - d * y) * inorm; vz = (vz - d * z) * inorm;
        }
        v_dirs[(size_t)e * 3] = vx; v_dirs[(size_t)e * 3 + 1] = vy; v_dirs[(size_t)e * 3 + 2] = vz;
    }
}

// ------------------------------------------------------------------------------------------
// K == 16 fast path (SH degree <= 3 with the reference's [N,16,3] layout, 192-byte rows).
// Every thread moves ITS OWN 192-byte coefficient row with one TMA bulk copy (global -> shared
// through an mbarrier, shared -> global through a bulk group), so all global traffic is full,
// aligned 32-byte sectors regardless of the 192-byte element stride, and the per-thread shared rows
// are padded to 208 bytes: float4 slot (13 t + i) mod 8 is conflict-free for 8 consecutive lanes.
// ------------------------------------------------------------------------------------------
constexpr int kRow16 = 48;        // floats per row
constexpr int kRowStride16 = 52;  // padded row stride in floats (208 B)

template <int DEG>
__global__ void __launch_bounds__(kShThreads) sh_fwd_k16_kernel(uint32_t M, const float *__restrict__ dirs,
                                                                const float *__restrict__ coeffs,
                                                                const uint8_t *__restrict__ masks,
                                                                float *__restrict__ colors) {
    constexpr int NB = (DEG + 1) * (DEG + 1);
    extern __shared__ __align__(128) float s_rows[]; // [kShThreads][kRowStride16]
    __shared__ __align__(8) uint64_t s_bar;
    const uint32_t tid = threadIdx.x;
    const uint32_t e = blockIdx.x * kShThreads + tid;
    const bool active = (e < M) && (masks == nullptr || masks[e]);
    if (tid == 0) { mbar_init(&s_bar, kShThreads); mbar_fence_init(); }
    __syncthreads();
    float *row = s_rows + tid * kRowStride16;
    constexpr uint32_t kInBytes = (NB * 12 + 15) / 16 * 16; // only the active degrees' bytes
    if (active) {
        mbar_arrive_expect_tx(&s_bar, kInBytes);
        bulk_g2s(row, coeffs + (size_t)e * kRow16, kInBytes, &s_bar);
    } else {
        mbar_arrive(&s_bar);
    }
    float b[NB];
    float x = 0.f, y = 0.f, z = 1.f;
    if (active) { x = dirs[(size_t)e * 3]; y = dirs[(size_t)e * 3 + 1]; z = dirs[(size_t)e * 3 + 2]; }
    if constexpr (DEG >= 1) {
        const float inorm = rsqrtf(x * x + y * y + z * z);
        x *= inorm; y *= inorm; z *= inorm;
    }
    sh_basis<DEG>(x, y, z, b);
    mbar_wait(&s_bar, 0);
    if (!active) return; // masked rows stay untouched
    float r0 = 0.f, r1 = 0.f, r2 = 0.f;
    const float4 *r4 = reinterpret_cast<const float4 *>(row);
    float f[(NB * 3 + 3) / 4 * 4];
#pragma unroll
    for (int i = 0; i < (NB * 3 + 3) / 4; ++i) {
        const float4 v = r4[i];
        f[i * 4] = v.x; f[i * 4 + 1] = v.y; f[i * 4 + 2] = v.z; f[i * 4 + 3] = v.w;
    }
#pragma unroll
    for (int k = 0; k < NB; ++k) { r0 += b[k] * f[k * 3]; r1 += b[k] * f[k * 3 + 1]; r2 += b[k] * f[k * 3 + 2]; }
    colors[(size_t)e * 3] = r0; colors[(size_t)e * 3 + 1] = r1; colors[(size_t)e * 3 + 2] = r2;
}

template <int DEG>
__global__ void __launch_bounds__(kShThreads) sh_bwd_k16_kernel(uint32_t M, const float *__restrict__ dirs,
                                                                const float *__restrict__ coeffs,
                                                                const uint8_t *__restrict__ masks,
                                                                const float *__restrict__ v_colors,
                                                                float *__restrict__ v_coeffs,
                                                                float *__restrict__ v_dirs) {
    constexpr int NB = (DEG + 1) * (DEG + 1);
    extern __shared__ __align__(128) float s_rows[]; // in rows [T][52] then out rows [T][52]
    __shared__ __align__(8) uint64_t s_bar;
    const uint32_t tid = threadIdx.x;
    const uint32_t e = blockIdx.x * kShThreads + tid;
    const bool inside = e < M;
    const bool active = inside && (masks == nullptr || masks[e]);
    const bool need_in = active && (v_dirs != nullptr) && (DEG >= 1);
    if (tid == 0) { mbar_init(&s_bar, kShThreads); mbar_fence_init(); }
    __syncthreads();
    float *row_in = s_rows + tid * kRowStride16;
    float *row_out = s_rows + (kShThreads + tid) * kRowStride16;
    constexpr uint32_t kInBytes = (NB * 12 + 15) / 16 * 16;
    if (need_in) {
        mbar_arrive_expect_tx(&s_bar, kInBytes);
        bulk_g2s(row_in, coeffs + (size_t)e * kRow16, kInBytes, &s_bar);
    } else {
        mbar_arrive(&s_bar);
    }
    float b[NB];
    float x = 0.f, y = 0.f, z = 1.f, inorm = 1.f, g0 = 0.f, g1 = 0.f, g2 = 0.f;
    if (active) {
        x = dirs[(size_t)e * 3]; y = dirs[(size_t)e * 3 + 1]; z = dirs[(size_t)e * 3 + 2];
        g0 = v_colors[(size_t)e * 3]; g1 = v_colors[(size_t)e * 3 + 1]; g2 = v_colors[(size_t)e * 3 + 2];
    }
    if constexpr (DEG >= 1) {
        inorm = rsqrtf(x * x + y * y + z * z);
        x *= inorm; y *= inorm; z *= inorm;
    }
    sh_basis<DEG>(x, y, z, b);
    // v_coeffs row: b_k * g_c for the active degrees, zeros elsewhere (and for masked elements)
    float4 *o4 = reinterpret_cast<float4 *>(row_out);
    {
        float o[kRow16];
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const float bk = (k < NB) ? b[k < NB ? k : 0] : 0.f;
            o[k * 3] = bk * g0; o[k * 3 + 1] = bk * g1; o[k * 3 + 2] = bk * g2;
        }
#pragma unroll
        for (int i = 0; i < kRow16 / 4; ++i) o4[i] = make_float4(o[i * 4], o[i * 4 + 1], o[i * 4 + 2], o[i * 4 + 3]);
    }
    if (inside) {
        fence_proxy_async(); // this thread's generic-proxy writes -> visible to its bulk store
        bulk_s2g(v_coeffs + (size_t)e * kRow16, row_out, kRow16 * 4);
        bulk_commit();
    }
    mbar_wait(&s_bar, 0);
    if (inside && v_dirs != nullptr) {
        float vx = 0.f, vy = 0.f, vz = 0.f;
        if constexpr (DEG >= 1) {
            if (active) {
                float w[NB];
                const float4 *r4 = reinterpret_cast<const float4 *>(row_in);
                float f[(NB * 3 + 3) / 4 * 4];
#pragma unroll
                for (int i = 0; i < (NB * 3 + 3) / 4; ++i) {
                    const float4 v = r4[i];
                    f[i * 4] = v.x; f[i * 4 + 1] = v.y; f[i * 4 + 2] = v.z; f[i * 4 + 3] = v.w;
                }
#pragma unroll
                for (int k = 0; k < NB; ++k) w[k] = g0 * f[k * 3] + g1 * f[k * 3 + 1] + g2 * f[k * 3 + 2];
                sh_basis_vjp<DEG>(x, y, z, w, vx, vy, vz);
                const float d = vx * x + vy * y + vz * z;
                vx = (vx - d * x) * inorm; vy = (vy - d * y) * inorm; vz = (vz - d * z) * inorm;
            }
        }
        v_dirs[(size_t)e * 3] = vx; v_dirs[(size_t)e * 3 + 1] = vy; v_dirs[(size_t)e * 3 + 2] = vz;
    }
    if (inside) bulk_wait_read_all(); // the row must stay in shared memory until the bulk store has read it
}

// ------------------------------------------------------------------------------------------
// Multi-view SH backward (the multi-GPU exchange step).  The SH gradient of one view is the
// rank-1 row  v_coeffs[n] = Y(dir_v(n)) (x) v_color_v[n]  -- 192 B per Gaussian made from 12 B.  Ranks
// therefore exchange the 12-byte colour gradients (all-gather) instead of all-reducing the 192-byte rows,
// and every rank expands and sums all V views here: one thread per Gaussian, the coefficient row comes in
// and the summed gradient row goes out with one TMA bulk copy each (K = 16), views whose colour gradient is
// exactly zero (Gaussian not blended in that view) cost one 12-byte load.  The gradient w.r.t. the view
// directions is chained to the Gaussian's position (dir = mean - campos) and ACCUMULATED into v_means.
// ------------------------------------------------------------------------------------------
template <int DEG, bool kRow16Path>
__global__ void __launch_bounds__(kShThreads) sh_bwd_views_kernel(uint32_t M, uint32_t K, uint32_t V,
                                                                  const float *__restrict__ means,
                                                                  const float *__restrict__ campos,
                                                                  const float *__restrict__ coeffs,
                                                                  const float *__restrict__ v_colors,
                                                                  float *__restrict__ v_coeffs,
                                                                  float *__restrict__ v_means) {
    constexpr int NB = (DEG + 1) * (DEG + 1);
    extern __shared__ __align__(128) float s_rows[]; // row path: in rows [T][52] then out rows [T][52]
    __shared__ __align__(8) uint64_t s_bar;
    const uint32_t tid = threadIdx.x;
    const uint32_t e = blockIdx.x * kShThreads + tid;
    const bool inside = e < M;
    float *row_in = nullptr, *row_out = nullptr;
    if constexpr (kRow16Path) {
        if (tid == 0) { mbar_init(&s_bar, kShThreads); mbar_fence_init(); }
        __syncthreads();
        row_in = s_rows + tid * kRowStride16;
        row_out = s_rows + (kShThreads + tid) * kRowStride16;
        constexpr uint32_t kInBytes = (NB * 12 + 15) / 16 * 16;
        if (inside && DEG >= 1) {
            mbar_arrive_expect_tx(&s_bar, kInBytes);
            bulk_g2s(row_in, coeffs + (size_t)e * kRow16, kInBytes, &s_bar);
        } else {
            mbar_arrive(&s_bar);
        }
    }
    float acc[NB * 3];
#pragma unroll
    for (int i = 0; i < NB * 3; ++i) acc[i] = 0.f;
    float mx = 0.f, my = 0.f, mz = 0.f, vmx = 0.f, vmy = 0.f, vmz = 0.f;
    if (inside) { mx = means[(size_t)e * 3]; my = means[(size_t)e * 3 + 1]; mz = means[(size_t)e * 3 + 2]; }
    if constexpr (kRow16Path) mbar_wait(&s_bar, 0);
    if (inside) {
        for (uint32_t v = 0; v < V; ++v) {
            const float *gc = v_colors + ((size_t)v * M + e) * 3;
            const float g0 = gc[0], g1 = gc[1], g2 = gc[2];
            if (g0 == 0.f && g1 == 0.f && g2 == 0.f) continue;
            float x = mx - campos[v * 3], y = my - campos[v * 3 + 1], z = mz - campos[v * 3 + 2], inorm = 1.f;
            if constexpr (DEG >= 1) {
                inorm = rsqrtf(x * x + y * y + z * z);
                x *= inorm; y *= inorm; z *= inorm;
            }
            float b[NB];
            sh_basis<DEG>(x, y, z, b);
#pragma unroll
            for (int k = 0; k < NB; ++k) {
                acc[k * 3] = fmaf(b[k], g0, acc[k * 3]);
                acc[k * 3 + 1] = fmaf(b[k], g1, acc[k * 3 + 1]);
                acc[k * 3 + 2] = fmaf(b[k], g2, acc[k * 3 + 2]);
            }
            if constexpr (DEG >= 1) {
                float w[NB];
#pragma unroll
                for (int k = 0; k < NB; ++k) {
                    float f0, f1, f2;
                    if constexpr (kRow16Path) { f0 = row_in[k * 3]; f1 = row_in[k * 3 + 1]; f2 = row_in[k * 3 + 2]; }
                    else {
                        const float *f = coeffs + ((size_t)e * K + k) * 3;
                        f0 = f[0]; f1 = f[1]; f2 = f[2];
                    }
                    w[k] = g0 * f0 + g1 * f1 + g2 * f2;
                }
                float vx = 0.f, vy = 0.f, vz = 0.f;
                sh_basis_vjp<DEG>(x, y, z, w, vx, vy, vz);
                const float d = vx * x + vy * y + vz * z;
                vmx += (vx - d * x) * inorm; vmy += (vy - d * y) * inorm; vmz += (vz - d * z) * inorm;
            }
        }
    }
    if constexpr (kRow16Path) {
        float4 *o4 = reinterpret_cast<float4 *>(row_out);
#pragma unroll
        for (int i = 0; i < kRow16 / 4; ++i) {
            float t[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) t[j] = (i * 4 + j < NB * 3) ? acc[(i * 4 + j < NB * 3) ? i * 4 + j : 0] : 0.f;
            o4[i] = make_float4(t[0], t[1], t[2], t[3]);
        }
        if (inside) {
            fence_proxy_async();
            bulk_s2g(v_coeffs + (size_t)e * kRow16, row_out, kRow16 * 4);
            bulk_commit();
        }
    } else if (inside) {
        float *o = v_coeffs + (size_t)e * K * 3;
        for (uint32_t k = 0; k < K; ++k)
            for (int c = 0; c < 3; ++c) o[k * 3 + c] = (k < (uint32_t)NB) ? acc[(k < (uint32_t)NB ? k : 0) * 3 + c] : 0.f;
    }
    if (inside && DEG >= 1) {
        v_means[(size_t)e * 3] += vmx; v_means[(size_t)e * 3 + 1] += vmy; v_means[(size_t)e * 3 + 2] += vmz;
    }
    if constexpr (kRow16Path) {
        if (inside) bulk_wait_read_all();
    }
}

static inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

} // namespace gsb

extern "C" int gsb_sh_fwd(uint32_t M, uint32_t K, uint32_t degree, const float *dirs, const float *coeffs,
                          const uint8_t *masks, float *colors, gsb_stream_t stream) {
    if (M == 0) return GSB_OK;
    if (!dirs || !coeffs || !colors) return GSB_E_INVALID;
    if (degree > 4 || (degree + 1) * (degree + 1) > K) return GSB_E_INVALID;
    const dim3 grid((M + gsb::kShThreads - 1) / gsb::kShThreads);
    cudaStream_t s = gsb::as_stream(stream);
    gsb::ProfScope ps("sh_fwd", s);
    if (K == 16 && degree >= 1 && degree <= 3 && gsb::aligned16(coeffs)) {
        const size_t smem = (size_t)gsb::kShThreads * gsb::kRowStride16 * 4;
        switch (degree) {
        case 1: gsb::sh_fwd_k16_kernel<1><<<grid, gsb::kShThreads, smem, s>>>(M, dirs, coeffs, masks, colors); break;
        case 2: gsb::sh_fwd_k16_kernel<2><<<grid, gsb::kShThreads, smem, s>>>(M, dirs, coeffs, masks, colors); break;
        default: gsb::sh_fwd_k16_kernel<3><<<grid, gsb::kShThreads, smem, s>>>(M, dirs, coeffs, masks, colors); break;
        }
        GSB_LAUNCH_CHECK();
        return GSB_OK;
    }
    switch (degree) {
    case 0: gsb::sh_fwd_kernel<0><<<grid, gsb::kShThreads, 0, s>>>(M, K, dirs, coeffs, masks, colors); break;
    case 1: gsb::sh_fwd_kernel<1><<<grid, gsb::kShThreads, 0, s>>>(M, K, dirs, coeffs, masks, colors); break;
    case 2: gsb::sh_fwd_kernel<2><<<grid, gsb::kShThreads, 0, s>>>(M, K, dirs, coeffs, masks, colors); break;
    case 3: gsb::sh_fwd_kernel<3><<<grid, gsb::kShThreads, 0, s>>>(M, K, dirs, coeffs, masks, colors); break;
    default: gsb::sh_fwd_kernel<4><<<grid, gsb::kShThreads, 0, s>>>(M, K, dirs, coeffs, masks, colors); break;
    }
    GSB_LAUNCH_CHECK();
    return GSB_OK;
}

extern "C" int gsb_sh_bwd(uint32_t M, uint32_t K, uint32_t degree, const float *dirs, const float *coeffs,
                          const uint8_t *masks, const float *v_colors, float *v_coeffs, float *v_dirs,
                          gsb_stream_t stream) {
    if (M == 0) return GSB_OK;
    if (!dirs || !coeffs || !v_colors || !v_coeffs) return GSB_E_INVALID;
    if (degree > 4 || (degree + 1) * (degree + 1) > K) return GSB_E_INVALID;
    const dim3 grid((M + gsb::kShThreads - 1) / gsb::kShThreads);
    cudaStream_t s = gsb::as_stream(stream);
    gsb::ProfScope ps("sh_bwd", s);
    if (K == 16 && degree <= 3 && gsb::aligned16(coeffs) && gsb::aligned16(v_coeffs)) {
        const size_t smem = (size_t)2 * gsb::kShThreads * gsb::kRowStride16 * 4; // 53 248 B
        #define GSB_SH_BWD16(D)                                                                                        \
            cudaFuncSetAttribute(gsb::sh_bwd_k16_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);   \
            gsb::sh_bwd_k16_kernel<D><<<grid, gsb::kShThreads, smem, s>>>(M, dirs, coeffs, masks, v_colors, v_coeffs, v_dirs)
        switch (degree) {
        case 0: GSB_SH_BWD16(0); break;
        case 1: GSB_SH_BWD16(1); break;
        case 2: GSB_SH_BWD16(2); break;
        default: GSB_SH_BWD16(3); break;
        }
        #undef GSB_SH_BWD16
        GSB_LAUNCH_CHECK();
        return GSB_OK;
    }
    switch (degree) {
    case 0: gsb::sh_bwd_kernel<0><<<grid, gsb::kShThreads, 0, s>>>(M, K, dirs, coeffs, masks, v_colors, v_coeffs, v_dirs); break;
    case 1: gsb::sh_bwd_kernel<1><<<grid, gsb::kShThreads, 0, s>>>(M, K, dirs, coeffs, masks, v_colors, v_coeffs, v_dirs); break;
    case 2: gsb::sh_bwd_kernel<2><<<grid, gsb::kShThreads, 0, s>>>(M, K, dirs, coeffs, masks, v_colors, v_coeffs, v_dirs); break;
    case 3: gsb::sh_bwd_kernel<3><<<grid, gsb::kShThreads, 0, s>>>(M, K, dirs, coeffs, masks, v_colors, v_coeffs, v_dirs); break;
    default: gsb::sh_bwd_kernel<4><<<grid, gsb::kShThreads, 0, s>>>(M, K, dirs, coeffs, masks, v_colors, v_coeffs, v_dirs); break;
    }
    GSB_LAUNCH_CHECK();
    return GSB_OK;
}

extern "C" int gsb_sh_bwd_views(uint32_t M, uint32_t K, uint32_t degree, uint32_t V, const float *means,
                                const float *campos, const float *coeffs, const float *v_colors, float *v_coeffs,
                                float *v_means, gsb_stream_t stream) {
    if (M == 0) return GSB_OK;
    if (!means || !coeffs || !v_coeffs || !v_means || (V > 0 && (!campos || !v_colors))) return GSB_E_INVALID;
    if (degree > 4 || (degree + 1) * (degree + 1) > K) return GSB_E_INVALID;
    const dim3 grid((M + gsb::kShThreads - 1) / gsb::kShThreads);
    cudaStream_t s = gsb::as_stream(stream);
    gsb::ProfScope ps("sh_bwd_views", s);
    if (K == 16 && degree <= 3 && gsb::aligned16(coeffs) && gsb::aligned16(v_coeffs)) {
        const size_t smem = (size_t)2 * gsb::kShThreads * gsb::kRowStride16 * 4;
        #define GSB_SH_VIEWS16(D)                                                                                      \
            cudaFuncSetAttribute(gsb::sh_bwd_views_kernel<D, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,       \
                                 (int)smem);                                                                           \
            gsb::sh_bwd_views_kernel<D, true><<<grid, gsb::kShThreads, smem, s>>>(M, K, V, means, campos, coeffs,      \
                                                                                  v_colors, v_coeffs, v_means)
        switch (degree) {
        case 0: GSB_SH_VIEWS16(0); break;
        case 1: GSB_SH_VIEWS16(1); break;
        case 2: GSB_SH_VIEWS16(2); break;
        default: GSB_SH_VIEWS16(3); break;
        }
        #undef GSB_SH_VIEWS16
        GSB_LAUNCH_CHECK();
        return GSB_OK;
    }
    #define GSB_SH_VIEWS(D)                                                                                            \
        gsb::sh_bwd_views_kernel<D, false><<<grid, gsb::kShThreads, 0, s>>>(M, K, V, means, campos, coeffs, v_colors,  \
                                                                            v_coeffs, v_means)
    switch (degree) {
    case 0: GSB_SH_VIEWS(0); break;
    case 1: GSB_SH_VIEWS(1); break;
    case 2: GSB_SH_VIEWS(2); break;
    case 3: GSB_SH_VIEWS(3); break;
    default: GSB_SH_VIEWS(4); break;
    }
    #undef GSB_SH_VIEWS
    GSB_LAUNCH_CHECK();
    return GSB_OK;
}
