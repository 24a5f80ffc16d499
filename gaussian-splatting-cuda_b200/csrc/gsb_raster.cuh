// gsb_raster.cuh -- shared definitions of the from-world blend kernels (a7/a8).
//
// ---- the reformulation --------------------------------------------------------------------
// The reference evaluates, per (pixel, Gaussian) pair (RasterizeToPixelsFromWorld3DGSFwd.cu:228-248),
//     gro = M (o - mu),  grd = normalize(M d),  power = -1/2 |grd x gro|^2,   M = S^-1 R_g^T
// with o the camera centre and d the pixel's world ray: two 3x3 mat-vecs, a normalisation, a
// cross product, ~60 FLOP and 13 shared-memory floats per pair.  With a global shutter, o is the
// same for every pixel and d = B c(p) / |c(p)| with B = R_cam^-1 and c(p) = (u, v, 1) the
// pinhole direction of the pixel.  Writing A = M B (3x3 per Gaussian) and p_c = (u_c, v_c, 1)
// the projection of the Gaussian centre (mu_c = B^-1 mu + t = z_c p_c), one gets exactly
//     grd_un = A c(p) = G + du A0 + dv A1,          G = A p_c,  du = u - u_c, dv = v - v_c
//     grd_un x gro = du (A0 x gro) + dv (A1 x gro)  because gro = -z_c G is parallel to G
//     power = -1/2 |du E0 + dv E1|^2 / |G + du A0 + dv A1|^2,   E_i = A_i x gro
// i.e. a ratio of two quadratic forms in the pixel offset from the projected centre:
//     power = -1/2 (n0 du^2 + n1 du dv + n2 dv^2) / (d0 + d1 du + d2 dv + d3 du^2 + d4 du dv + d5 dv^2)
// The numerator is a positive semi-definite form in (du, dv) -- no cancellation, unlike the
// reference's cross product of a ~1000-long vector with a unit vector -- so float32 is enough
// for the per-pair arithmetic while the nine coefficients are computed once per Gaussian in
// float64 (gsb_prep_records_kernel).  Per pair this is ~16 FLOP and 12 shared-memory floats, and
// the alpha < 1/255 rejection needs no MUFU at all (N >= tau * D test).  The backward pass
// accumulates 15 moments  sum w {x, y, x^2, xy, y^2, ...}  per Gaussian instead of running the
// reference's 3x3/quaternion VJP per pair; the chain rule back to (mean, quat, scale) runs once
// per Gaussian in float64 (gsb_finalize_grads_kernel).  Derivation and numerical validation:
// DESIGN.md "Blend kernels".
#pragma once

#include "gsb_camera.cuh"

namespace gsb {

// 64-byte per-Gaussian record consumed by the blend kernels.  Pixel units: x = px - pcx.
//   e(x, y) = Ns / Ds + lop,  alpha_raw = 2^e
//   Ns = n0 x^2 + n1 x y + n2 y^2          (already scaled by -0.5 log2(e) / d0)
//   Ds = 1 + d1 x + d2 y + d3 x^2 + d4 x y + d5 y^2
//   tau: conservative MUFU-free rejection threshold: alpha_raw < 1/255  <=  Ns < tau * Ds
struct __align__(16) GaussRec {
    float pcx, pcy, n0, n1;
    float n2, d1, d2, d3;
    float d4, d5, lop, tau;
    float r, g, b;
    int32_t gid; // index of this Gaussian (row of the gradient-moment buffer)
};
static_assert(sizeof(GaussRec) == 64, "GaussRec must be 64 bytes");

constexpr int kRecFloats = 16;
constexpr int kMomFloats = 16; // 15 moments + pad, one 64-byte row per Gaussian
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr float kLog2AlphaThr = -7.994353436858858f; // log2(1/255)
constexpr float kTauMargin = 2e-3f;                   // in log2 units; >> float rounding of the test

// Camera constants shared by prep / blend / finalize, computed on the device from the same
// float view matrix the reference reads (Cameras.cuh:33-71,261-265): B = mat3_cast(inverse(q)).
struct CamConst {
    double B[3][3];    // camera->world rotation used for rays (R_inv of the reference)
    double Binv[3][3]; // its exact inverse (NOT the transpose: q is only unit to ~1e-7)
    double t[3];
    double fx, fy, cx, cy;
};

__device__ inline void cam_const_from(const float *viewmat, const float *K, CamConst &c) {
    const CamPose p = cam_pose_from_viewmat(viewmat);
    // glm::inverse(quat) = conjugate / dot, evaluated in float like the reference
    const float d = p.qw * p.qw + p.qx * p.qx + p.qy * p.qy + p.qz * p.qz;
    const float iw = p.qw / d, ix = -p.qx / d, iy = -p.qy / d, iz = -p.qz / d;
    const M3<float> Bf = rotmat_raw(iw, ix, iy, iz);
    M3<double> B;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) B.m[i][j] = (double)Bf.m[i][j];
    const M3<double> Bi = inverse3(B);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) { c.B[i][j] = B.m[i][j]; c.Binv[i][j] = Bi.m[i][j]; }
    c.t[0] = p.tx; c.t[1] = p.ty; c.t[2] = p.tz;
    c.fx = K[0]; c.fy = K[4]; c.cx = K[2]; c.cy = K[5];
}

// Geometry of one Gaussian in the camera-relative parametrisation.  T = double in prep (the projected
// centre is cancellation-sensitive), T = float in finalize (the chain rule is not: see DESIGN.md 4.1).
template <typename T>
struct GaussGeom {
    V3<T> A0, A1, A2, G, gro, E0, E1;
    M3<T> Rg;       // rotation of the normalised quaternion (row-major math matrix)
    T inv_s[3];
    T qn[4], inv_qnorm;
    T zc, uc, vc;
    T n[3], d0, d[5];
    bool degenerate;
};

template <typename T>
__device__ inline void gauss_geom(const CamConst &c, const float *mean, const float *quat, const float *scale,
                                  GaussGeom<T> &g) {
    const T qw = quat[0], qx = quat[1], qy = quat[2], qz = quat[3];
    const T nn = qw * qw + qx * qx + qy * qy + qz * qz;
    g.inv_qnorm = T(1) / sqrt(nn);
    g.qn[0] = qw * g.inv_qnorm; g.qn[1] = qx * g.inv_qnorm; g.qn[2] = qy * g.inv_qnorm; g.qn[3] = qz * g.inv_qnorm;
    g.Rg = rotmat_raw<T>(g.qn[0], g.qn[1], g.qn[2], g.qn[3]);
    g.inv_s[0] = T(1) / (T)scale[0]; g.inv_s[1] = T(1) / (T)scale[1]; g.inv_s[2] = T(1) / (T)scale[2];
    // M = diag(1/s) Rg^T ; A = M B
    M3<T> A;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            T s = T(0);
            for (int k = 0; k < 3; ++k) s += g.Rg.m[k][i] * (T)c.B[k][j];
            A.m[i][j] = s * g.inv_s[i];
        }
    g.A0 = col(A, 0); g.A1 = col(A, 1); g.A2 = col(A, 2);
    const T mx = mean[0], my = mean[1], mz = mean[2];
    const T xc = (T)c.Binv[0][0] * mx + (T)c.Binv[0][1] * my + (T)c.Binv[0][2] * mz + (T)c.t[0];
    const T yc = (T)c.Binv[1][0] * mx + (T)c.Binv[1][1] * my + (T)c.Binv[1][2] * mz + (T)c.t[1];
    const T zc = (T)c.Binv[2][0] * mx + (T)c.Binv[2][1] * my + (T)c.Binv[2][2] * mz + (T)c.t[2];
    g.zc = zc;
    g.degenerate = !(fabs((double)zc) > 1e-12 * (fabs((double)xc) + fabs((double)yc) + 1e-300)) || !isfinite((double)zc);
    const T iz = g.degenerate ? T(0) : T(1) / zc;
    g.uc = xc * iz; g.vc = yc * iz;
    g.G = g.A0 * g.uc + g.A1 * g.vc + g.A2;
    g.gro = g.G * (-zc);
    g.E0 = cross(g.A0, g.gro);
    g.E1 = cross(g.A1, g.gro);
    g.n[0] = dot(g.E0, g.E0); g.n[1] = T(2) * dot(g.E0, g.E1); g.n[2] = dot(g.E1, g.E1);
    g.d0 = dot(g.G, g.G);
    g.d[0] = T(2) * dot(g.G, g.A0); g.d[1] = T(2) * dot(g.G, g.A1);
    g.d[2] = dot(g.A0, g.A0); g.d[3] = T(2) * dot(g.A0, g.A1); g.d[4] = dot(g.A1, g.A1);
    if (!(g.d0 > T(0)) || !isfinite((double)g.d0)) g.degenerate = true;
}

// Per-pair evaluation, shared VERBATIM by the forward and backward kernels (explicitly rounded
// operations: the two kernels must agree bit-for-bit on which pairs pass the alpha test), for a
// thread's two pixels at once on packed fp32 pairs (lo = pixel 0, hi = pixel 1):
//   Ns = n0 x^2 + n1 xy + n2 y^2          (pre-scaled by -1/2 log2 e / d0)
//   Ds = 1 + d1 x + d2 y + d3 x^2 + d4 xy + d5 y^2
//   pass <=> Ns >= tau Ds                  (MUFU-free rejection, tau carries a safety margin)
struct PairEval2 {
    f2 Ns, Ds;
    f2 xx, xy, yy;
    bool pass0, pass1;
};
__device__ __forceinline__ PairEval2 pair_eval2(const float4 q0, const float4 q1, const float4 q2, f2 x, f2 y) {
    PairEval2 r;
    r.xx = f2_mul(x, x); r.xy = f2_mul(x, y); r.yy = f2_mul(y, y);
    r.Ns = f2_fma(f2_bc(q1.x), r.yy, f2_fma(f2_bc(q0.w), r.xy, f2_mul(f2_bc(q0.z), r.xx)));
    f2 D = f2_fma(f2_bc(q1.y), x, f2_bc(1.0f));
    D = f2_fma(f2_bc(q1.z), y, D);
    D = f2_fma(f2_bc(q1.w), r.xx, D);
    D = f2_fma(f2_bc(q2.x), r.xy, D);
    r.Ds = f2_fma(f2_bc(q2.y), r.yy, D);
    const f2 t = f2_fma(f2_bc(-q2.w), r.Ds, r.Ns); // Ns >= tau * Ds
    r.pass0 = f2_lo(t) >= 0.0f;
    r.pass1 = f2_hi(t) >= 0.0f;
    return r;
}
// Warp-level culling: can ANY pixel of the warp's block, whose (undistorted) pixel coordinates lie
// in the box [bx0, bx1] x [by0, by1], pass the rejection test of this record?  The pass region {Ns - tau Ds >= 0} is the interior of an
// ellipse (the form F = A x^2 + B xy + C y^2 + D x + E y + F0 is concave for every realistic
// Gaussian), so the exact answer is max_box F >= 0: the unconstrained maximiser if it lies in the
// box, else the best point on the (at most two) box edges facing it.  One lane tests one record,
// so a warp classifies 32 records per pass; `slack` absorbs the float rounding of both this test
// and pair_eval's, keeping the cull strictly conservative.  Non-concave forms are never culled.
__device__ __forceinline__ bool block_may_pass(const float4 q0, const float4 q1, const float4 q2, float bx0,
                                               float bx1, float by0, float by1) {
    const float tau = q2.w;
    if (!(tau < 3.0e38f)) return false; // dead record (tau = +inf)
    const float A = __fmaf_rn(-tau, q1.w, q0.z), B = __fmaf_rn(-tau, q2.x, q0.w), C = __fmaf_rn(-tau, q2.y, q1.x);
    const float D = -tau * q1.y, E = -tau * q1.z, F0 = -tau;
    const float x0 = bx0 - q0.x, x1 = bx1 - q0.x, y0 = by0 - q0.y, y1 = by1 - q0.y;
    const float det = 4.0f * A * C - B * B;
    if (!(A < 0.0f && C < 0.0f && det > 0.0f)) return true;
    const float idet = 1.0f / det;
    const float xs = (B * E - 2.0f * C * D) * idet, ys = (B * D - 2.0f * A * E) * idet;
    const bool xin = (xs >= x0) && (xs <= x1), yin = (ys >= y0) && (ys <= y1);
    auto Fv = [&](float x, float y) { return (A * x + B * y + D) * x + (C * y + E) * y + F0; };
    float fmax;
    if (xin && yin) {
        fmax = Fv(xs, ys);
    } else {
        fmax = -3.0e38f;
        if (!xin) {
            const float xc = fminf(fmaxf(xs, x0), x1);
            const float yc = fminf(fmaxf(-(E + B * xc) / (2.0f * C), y0), y1);
            fmax = fmaxf(fmax, Fv(xc, yc));
        }
        if (!yin) {
            const float yc = fminf(fmaxf(ys, y0), y1);
            const float xc = fminf(fmaxf(-(D + B * yc) / (2.0f * A), x0), x1);
            fmax = fmaxf(fmax, Fv(xc, yc));
        }
    }
    return fmax >= -0.02f;
}

// alpha_raw = opac * exp(power) = 2^(Ns/Ds + lop)
__device__ __forceinline__ float pair_alpha_raw(float Ns, float Ds, float lop) {
    return fast_ex2(__fmaf_rn(Ns, fast_rcp(Ds), lop));
}

} // namespace gsb
