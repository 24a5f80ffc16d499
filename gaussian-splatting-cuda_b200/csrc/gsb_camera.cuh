// gsb_camera.cuh -- a2: the reference's camera models for a GLOBAL shutter
// (gsplat/Cameras.cuh:416-471 perfect pinhole, :473-755 OpenCV pinhole with radial / tangential /
// thin-prism distortion, :817-1001 OpenCV fisheye), restated without GLM / CRTP.
//
//   cam_project    camera-space point -> image point + validity   (camera_ray_to_image_point)
//   cam_unproject  image point -> camera-space ray direction + validity (image_point_to_camera_ray)
//
// Rolling shutter (Cameras.cuh:346-413, iterative pose refinement per point) is not implemented:
// no caller of the reference passes anything but ShutterType::GLOBAL (rasterizer_autograd.cpp:234,
// 311, 369) and the quadratic-form blend relies on a single camera centre per image.
#pragma once

#include "gsb_common.cuh"

namespace gsb {

enum : int { kCamPerfectPinhole = 0, kCamOpenCVPinhole = 1, kCamOpenCVFisheye = 2 };

struct CamModel {
    int kind;
    float W, H;
    float fx, fy, cx, cy;
    float k[6], p[2], s[4];  // OpenCV pinhole (Cameras.cuh:481-487)
    float fwd[5], dfwd[5];   // fisheye forward polynomial (odd) / derivative (even) (:837-842)
    float bwd1;              // fisheye linear guess of the inverse (:881-884)
    float max_angle;
};

__device__ __forceinline__ float poly_horner5(const float *c, float x) {
    float y = c[4];
    y = x * y + c[3]; y = x * y + c[2]; y = x * y + c[1]; y = x * y + c[0];
    return y;
}

// Smallest positive real root of 1 + a x + b x^2 + c x^3 (+inf when there is none): where the derivative of the
// fisheye polynomial first vanishes, i.e. the largest angle up to which the model is monotonic (what
// Cameras.cuh:759-815 computes).  Linear / quadratic cases directly; the cubic through its depressed form
// t^3 + P t + Q (x = t - B/3 for the monic x^3 + B x^2 + A x + C0), Cardano for one real root, the
// trigonometric form for three.
__device__ inline float smallest_positive_root_cubic(float a, float b, float c) {
    const float INF = 3.402823466e+38f;
    if (c == 0.0f) {
        if (b == 0.0f) return a < 0.0f ? -1.0f / a : INF;                 // 1 + a x
        const float disc = a * a - 4.0f * b;                              // 1 + a x + b x^2
        if (disc < 0.0f) return INF;
        // the root with the smaller magnitude in its cancellation-free form: 2 / (-a + sqrt(disc)); the other root is
        // either negative or larger
        const float q = sqrtf(disc) - a;
        return q > 0.0f ? 2.0f / q : INF;
    }
    const float B = b / c, A = a / c, C0 = 1.0f / c;                      // monic coefficients
    const float P = A - B * B / 3.0f;
    const float Q = 2.0f * B * B * B / 27.0f - A * B / 3.0f + C0;
    const float shift = -B / 3.0f;
    const float D = 0.25f * Q * Q + P * P * P / 27.0f;
    if (D >= 0.0f) {                                                      // one real root
        const float sD = sqrtf(D);
        const float u = cbrtf(-0.5f * Q + sD);
        // v = cbrt(-Q/2 - sqrt(D)) = -P / (3 u): avoids the second cube root and its cancellation
        if (u == 0.0f) return INF;
        const float x = u - P / (3.0f * u) + shift;
        return x > 0.0f ? x : INF;
    }
    const float m = 2.0f * sqrtf(-P / 3.0f);                              // three real roots
    const float phi = atan2f(sqrtf(-D), -0.5f * Q) / 3.0f;
    const float third = 2.0943951023931953f;                              // 2 pi / 3
    float best = INF;
#pragma unroll
    for (int k = -1; k <= 1; ++k) {
        const float x = m * cosf(phi + (float)k * third) + shift;
        if (x > 0.0f) best = fminf(best, x);
    }
    return best;
}

// Built once per CTA from the DEVICE camera arrays (the reference rebuilds it in every thread).
// n_* = floats per camera behind each pointer; missing coefficients are zero.
__device__ inline void cam_model_build(CamModel &m, int camera_model, uint32_t W, uint32_t H, const float *K,
                                       const float *radial, int nr, const float *tangential, int nt,
                                       const float *thin_prism, int ns) {
    m.W = (float)W; m.H = (float)H;
    m.fx = K[0]; m.fy = K[4]; m.cx = K[2]; m.cy = K[5];
    for (int i = 0; i < 6; ++i) m.k[i] = 0.f;
    for (int i = 0; i < 2; ++i) m.p[i] = 0.f;
    for (int i = 0; i < 4; ++i) m.s[i] = 0.f;
    for (int i = 0; i < 5; ++i) { m.fwd[i] = 0.f; m.dfwd[i] = 0.f; }
    m.bwd1 = 0.f; m.max_angle = 0.f;
    if (camera_model == GSB_CAMERA_FISHEYE) { // Cameras.cuh:830-885
        m.kind = kCamOpenCVFisheye;
        float fk[4] = {0.f, 0.f, 0.f, 0.f};
        for (int i = 0; i < 4 && radial && i < nr; ++i) fk[i] = radial[i];
        m.fwd[0] = 1.f; m.fwd[1] = fk[0]; m.fwd[2] = fk[1]; m.fwd[3] = fk[2]; m.fwd[4] = fk[3];
        m.dfwd[0] = 1.f; m.dfwd[1] = 3.f * fk[0]; m.dfwd[2] = 5.f * fk[1]; m.dfwd[3] = 7.f * fk[2]; m.dfwd[4] = 9.f * fk[3];
        const float mdx = fmaxf(m.W - m.cx, m.cx), mdy = fmaxf(m.H - m.cy, m.cy);
        const float max_radius_pixels = sqrtf(mdx * mdx + mdy * mdy);
        if (fk[3] == 0.f) {
            m.max_angle = sqrtf(smallest_positive_root_cubic(3.f * fk[0], 5.f * fk[1], 7.f * fk[2]));
        } else {
            const float dd[4] = {6.f * fk[0], 20.f * fk[1], 42.f * fk[2], 72.f * fk[3]};
            float x = 1.57f;
            bool converged = false;
            for (int j = 0; j < 20; ++j) {
                const float x2 = x * x;
                const float dfdx = x * (((dd[3] * x2 + dd[2]) * x2 + dd[1]) * x2 + dd[0]);
                const float residual = poly_horner5(m.dfwd, x2);
                const float dx = residual / dfdx;
                x -= dx;
                if (fabsf(dx) < 1e-6f) { converged = true; break; }
            }
            m.max_angle = (!converged || x <= 0.f) ? 3.402823466e+38f : x;
        }
        m.max_angle = fminf(m.max_angle, fmaxf(max_radius_pixels / m.fx, max_radius_pixels / m.fy));
        const float max_normalized_dist = fmaxf(m.W / 2.f / m.fx, m.H / 2.f / m.fy);
        m.bwd1 = m.max_angle / max_normalized_dist;
    } else if (radial || tangential || thin_prism) { // ProjectionUT3DGSFused.cu:98-115
        m.kind = kCamOpenCVPinhole;
        for (int i = 0; i < 6 && radial && i < nr; ++i) m.k[i] = radial[i];
        for (int i = 0; i < 2 && tangential && i < nt; ++i) m.p[i] = tangential[i];
        for (int i = 0; i < 4 && thin_prism && i < ns; ++i) m.s[i] = thin_prism[i];
    } else {
        m.kind = kCamPerfectPinhole;
    }
}

// OpenCV pinhole distortion (rational radial g = N(r)/D(r), r = x^2 + y^2; tangential p1, p2; thin prism s1..s4):
//   u = g x + 2 p1 x y + p2 (r + 2 x^2) + s1 r + s2 r^2,   v = g y + p1 (r + 2 y^2) + 2 p2 x y + s3 r + s4 r^2
// and, when asked, the Jacobian d(u, v)/d(x, y) (with dr/dx = 2x, dr/dy = 2y and g' = (N' D - N D') / D^2).
struct OpenCVDistortion {
    float u, v, radial;
    float ux, uy, vx, vy;
};
template <bool kJacobian>
__device__ __forceinline__ OpenCVDistortion opencv_distort(const CamModel &m, float x, float y) {
    OpenCVDistortion o;
    const float r = x * x + y * y;
    const float N = 1.0f + r * (m.k[0] + r * (m.k[1] + r * m.k[2]));
    const float D = 1.0f + r * (m.k[3] + r * (m.k[4] + r * m.k[5]));
    const float g = N / D;
    o.radial = g;
    const float xy2 = 2.0f * x * y;
    o.u = g * x + m.p[0] * xy2 + m.p[1] * (r + 2.0f * x * x) + r * (m.s[0] + r * m.s[1]);
    o.v = g * y + m.p[0] * (r + 2.0f * y * y) + m.p[1] * xy2 + r * (m.s[2] + r * m.s[3]);
    o.ux = o.uy = o.vx = o.vy = 0.f;
    if (kJacobian) {
        const float Nr = m.k[0] + r * (2.0f * m.k[1] + r * (3.0f * m.k[2]));
        const float Dr = m.k[3] + r * (2.0f * m.k[4] + r * (3.0f * m.k[5]));
        const float gr = (Nr * D - N * Dr) / (D * D);
        const float tu = m.s[0] + 2.0f * m.s[1] * r, tv = m.s[2] + 2.0f * m.s[3] * r; // d(thin prism)/dr
        const float rx = 2.0f * x, ry = 2.0f * y;
        o.ux = g + x * gr * rx + 2.0f * m.p[0] * y + 6.0f * m.p[1] * x + tu * rx;
        o.uy = x * gr * ry + 2.0f * m.p[0] * x + 2.0f * m.p[1] * y + tu * ry;
        o.vx = y * gr * rx + 2.0f * m.p[0] * x + 2.0f * m.p[1] * y + tv * rx;
        o.vy = g + y * gr * ry + 6.0f * m.p[0] * y + 2.0f * m.p[1] * x + tv * ry;
    }
    return o;
}

// Cameras.cuh:228-240
__device__ __forceinline__ bool in_bounds_margin(float px, float py, float W, float H, float margin) {
    const float MX = W * margin, MY = H * margin;
    return (-MX <= px) && (px < W + MX) && (-MY <= py) && (py < H + MY);
}

// camera_ray_to_image_point (Cameras.cuh:431-455, 535-597, 894-959)
__device__ inline bool cam_project(const CamModel &m, V3<float> cam, float margin, float &ox, float &oy) {
    ox = 0.f; oy = 0.f;
    if (cam.z <= 0.f) return false;
    if (m.kind == kCamPerfectPinhole) {
        ox = (cam.x / cam.z) * m.fx + m.cx;
        oy = (cam.y / cam.z) * m.fy + m.cy;
        return in_bounds_margin(ox, oy, m.W, m.H, margin);
    }
    if (m.kind == kCamOpenCVPinhole) {
        const OpenCVDistortion d = opencv_distort<false>(m, cam.x / cam.z, cam.y / cam.z);
        ox = d.u * m.fx + m.cx;
        oy = d.v * m.fy + m.cy;
        return (d.radial > 0.8f) && in_bounds_margin(ox, oy, m.W, m.H, margin); // Cameras.cuh:590-596
    }
    // fisheye
    const float ax = fabsf(cam.x), ay = fabsf(cam.y);
    const float mn = fminf(ax, ay), mx = fmaxf(ax, ay);
    float xy_norm = 0.f;
    if (mx > 0.f) { const float r = mn / mx; xy_norm = mx * sqrtf(1.f + r * r); }
    if (xy_norm <= 0.f) xy_norm = 1.1920929e-07f;
    const float theta_full = atan2f(xy_norm, cam.z);
    const float theta = theta_full < m.max_angle ? theta_full : m.max_angle;
    const float delta = theta * poly_horner5(m.fwd, theta * theta) / xy_norm;
    if (delta <= 0.f) return false;
    ox = m.fx * delta * cam.x + m.cx;
    oy = m.fy * delta * cam.y + m.cy;
    return in_bounds_margin(ox, oy, m.W, m.H, margin) && (theta <= m.max_angle);
}

// image_point_to_camera_ray (Cameras.cuh:457-470, 698-754, 961-1000).  Returns the ray as
// (x/z, y/z) -- the blend works with the undistorted normalised image coordinates -- and whether
// it is valid.  Rays with z <= 0 (a fisheye beyond 90 degrees off-axis) are reported invalid: see
// DESIGN.md, deviations.
__device__ inline bool cam_unproject_normalized(const CamModel &m, float px, float py, float &xn, float &yn) {
    const float u0 = (px - m.cx) / m.fx, v0 = (py - m.cy) / m.fy;
    xn = u0; yn = v0;
    if (m.kind == kCamPerfectPinhole) return true;
    if (m.kind == kCamOpenCVPinhole) {
        // Newton on distort(x, y) = (u0, v0) from the distorted point itself; the reference's iteration budget and
        // stopping rules (Cameras.cuh:698-740): at most 5 steps, stop on a non-positive radial factor, a Jacobian
        // determinant below 1e-6, or steps below 1e-6
        float x = u0, y = v0;
        bool converged = false;
        for (int iter = 0; iter < 5; ++iter) {
            const OpenCVDistortion d = opencv_distort<true>(m, x, y);
            if (d.radial <= 0.f) break;
            const float ru = u0 - d.u, rv = v0 - d.v;          // residual of the target
            const float det = d.ux * d.vy - d.uy * d.vx;
            if (fabsf(det) < 1e-6f) break;
            const float sx = (ru * d.vy - d.uy * rv) / det;     // Cramer: J (sx, sy) = (ru, rv)
            const float sy = (d.ux * rv - ru * d.vx) / det;
            x += sx; y += sy;
            if (fabsf(sx) < 1e-6f && fabsf(sy) < 1e-6f) { converged = true; break; }
        }
        xn = x; yn = y;
        return converged;
    }
    // fisheye: invert theta * P(theta^2) = |uv| by Newton from the linear guess
    const float delta = sqrtf(u0 * u0 + v0 * v0);
    float th = m.bwd1 * delta;
    bool converged = false;
    for (int j = 0; j < 20; ++j) {
        const float t2 = th * th;
        const float dfdx = poly_horner5(m.dfwd, t2);
        const float residual = th * poly_horner5(m.fwd, t2) - delta;
        const float dx = residual / dfdx;
        th -= dx;
        if (fabsf(dx) < 1e-6f) { converged = true; break; }
    }
    if (th < 0.f || th >= m.max_angle || !converged) return false;
    if (delta >= 1e-6f) {
        const float c = cosf(th);
        if (!(c > 0.f)) return false; // ray at/behind the image plane: not representable as (x/z, y/z)
        const float sf = sinf(th) / (delta * c);
        xn = sf * u0; yn = sf * v0;
    } else {
        xn = 0.f; yn = 0.f;
    }
    return true;
}

// ---- rolling shutter (Cameras.cuh:268-320) ------------------------------------------------------------------------------
struct Quat {
    float w, x, y, z;
};
// glm::slerp: shortest arc, linear interpolation when the endpoints (almost) coincide (not renormalised, like GLM)
__device__ __forceinline__ Quat quat_slerp(Quat a, Quat b, float t) {
    float c = a.w * b.w + a.x * b.x + a.y * b.y + a.z * b.z;
    if (c < 0.0f) { b.w = -b.w; b.x = -b.x; b.y = -b.y; b.z = -b.z; c = -c; }
    Quat r;
    if (c > 1.0f - 1.1920929e-07f) {
        r.w = a.w * (1.0f - t) + b.w * t; r.x = a.x * (1.0f - t) + b.x * t;
        r.y = a.y * (1.0f - t) + b.y * t; r.z = a.z * (1.0f - t) + b.z * t;
        return r;
    }
    const float ang = acosf(c);
    const float s0 = sinf((1.0f - t) * ang), s1 = sinf(t * ang), sd = sinf(ang);
    r.w = (s0 * a.w + s1 * b.w) / sd; r.x = (s0 * a.x + s1 * b.x) / sd;
    r.y = (s0 * a.y + s1 * b.y) / sd; r.z = (s0 * a.z + s1 * b.z) / sd;
    return r;
}
// Relative frame time in [0, 1] of an image point for the four rolling directions (0 for a global shutter)
__device__ __forceinline__ float shutter_relative_time(int shutter_type, float px, float py, float W, float H) {
    switch (shutter_type) {
        case GSB_SHUTTER_ROLLING_TOP_TO_BOTTOM: return floorf(py) / (H - 1.0f);
        case GSB_SHUTTER_ROLLING_LEFT_TO_RIGHT: return floorf(px) / (W - 1.0f);
        case GSB_SHUTTER_ROLLING_BOTTOM_TO_TOP: return (H - ceilf(py)) / (H - 1.0f);
        case GSB_SHUTTER_ROLLING_RIGHT_TO_LEFT: return (W - ceilf(px)) / (W - 1.0f);
        default: return 0.0f;
    }
}

} // namespace gsb
