// gsb_sh.cuh -- real spherical-harmonics basis (Sloan's ordering, degrees 0..4) and its VJP, shared by the
// stand-alone SH operators (gsb_sh.cu: a3/a4) and the fused front / back kernels (gsb_fused.cu: f1).
// Constants as in the reference (gsplat/SphericalHarmonicsCUDA.cu:21-110 forward, :113-371 VJP), which are
// Sloan's published generated code ("Efficient Spherical Harmonic Evaluation", JCGT 2013).
#pragma once

#include "gsb_common.cuh"

namespace gsb {

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

} // namespace gsb
