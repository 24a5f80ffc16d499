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

// Cameras.cuh:759-815: smallest positive root of 1 + a x + b x^2 + c x^3
__device__ inline float fisheye_max_angle_cubic(float a, float b, float c) {
    const float INF = 3.402823466e+38f;
    if (c == 0.0f) {
        if (b == 0.0f) return a >= 0.0f ? INF : -1.0f / a;
        float delta = a * a - 4.0f * b;
        if (delta >= 0.0f) {
            delta = sqrtf(delta) - a;
            if (delta > 0.0f) return 2.0f / delta;
        }
    } else {
        const float boc = b / c, boc2 = boc * boc;
        const float t1 = (9.0f * a * boc - 2.0f * b * boc2 - 27.0f) / c;
        const float t2 = 3.0f * a / c - boc2;
        const float delta = t1 * t1 + 4.0f * t2 * t2 * t2;
        if (delta >= 0.0f) {
            const float d2 = sqrtf(delta);
            const float cube_root = cbrtf((d2 + t1) / 2.0f);
            if (cube_root != 0.0f) {
                const float soln = (cube_root - (t2 / cube_root) - boc) / 3.0f;
                if (soln > 0.0f) return soln;
            }
        } else {
            const float theta = atan2f(sqrtf(-delta), t1) / 3.0f;
            const float two_third_pi = 2.0f * 3.14159265358979323846f / 3.0f;
            const float t3 = 2.0f * sqrtf(-t2);
            float soln = INF;
            for (int i = -1; i <= 1; ++i) {
                const float sv = (t3 * cosf(theta + (float)i * two_third_pi) - boc) / 3.0f;
                if (sv > 0.0f) soln = fminf(soln, sv);
            }
            return soln;
        }
    }
    return INF;
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
            m.max_angle = sqrtf(fisheye_max_angle_cubic(3.f * fk[0], 5.f * fk[1], 7.f * fk[2]));
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
        const float u = cam.x / cam.z, v = cam.y / cam.z;
        const float u2 = u * u, v2 = v * v, r2 = u2 + v2;
        const float a1 = 2.f * u * v, a2 = r2 + 2.f * u2, a3 = r2 + 2.f * v2;
        const float num = 1.f + r2 * (m.k[0] + r2 * (m.k[1] + r2 * m.k[2]));
        const float den = 1.f + r2 * (m.k[3] + r2 * (m.k[4] + r2 * m.k[5]));
        const float icD = num / den;
        const float dx = m.p[0] * a1 + m.p[1] * a2 + r2 * (m.s[0] + r2 * m.s[1]);
        const float dy = m.p[0] * a3 + m.p[1] * a1 + r2 * (m.s[2] + r2 * m.s[3]);
        ox = (icD * u + dx) * m.fx + m.cx;
        oy = (icD * v + dy) * m.fy + m.cy;
        return (icD > 0.8f) && in_bounds_margin(ox, oy, m.W, m.H, margin);
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
    if (m.kind == kCamOpenCVPinhole) { // Newton, at most 5 iterations
        float x = u0, y = v0;
        bool converged = false;
        for (int iter = 0; iter < 5; ++iter) {
            const float r = x * x + y * y, r2 = r * r;
            const float alpha = 1.0f + r * (m.k[0] + r * (m.k[1] + r * m.k[2]));
            const float beta = 1.0f + r * (m.k[3] + r * (m.k[4] + r * m.k[5]));
            const float d = alpha / beta;
            if (d <= 0.f) break;
            const float p1 = m.p[0], p2 = m.p[1], s1 = m.s[0], s2 = m.s[1], s3 = m.s[2], s4 = m.s[3];
            const float fx_ = d * x + 2.f * p1 * x * y + p2 * (r + 2.f * x * x) + s1 * r + s2 * r2 - u0;
            const float fy_ = d * y + 2.f * p2 * x * y + p1 * (r + 2.f * y * y) + s3 * r + s4 * r2 - v0;
            const float alpha_r = m.k[0] + r * (2.0f * m.k[1] + r * (3.0f * m.k[2]));
            const float beta_r = m.k[3] + r * (2.0f * m.k[4] + r * (3.0f * m.k[5]));
            const float d_r = (alpha_r * beta - alpha * beta_r) / (beta * beta);
            const float d_x = 2.0f * x * d_r, d_y = 2.0f * y * d_r;
            float fx_x = d + d_x * x + 2.0f * p1 * y + 6.0f * p2 * x;
            fx_x += 2.0f * x * (s1 + 2.0f * s2 * r);
            float fx_y = d_y * x + 2.0f * p1 * x + 2.0f * p2 * y;
            fx_y += 2.0f * y * (s1 + 2.0f * s2 * r);
            float fy_x = d_x * y + 2.0f * p2 * y + 2.0f * p1 * x;
            fy_x += 2.0f * x * (s3 + 2.0f * s4 * r);
            float fy_y = d + d_y * y + 2.0f * p2 * x + 6.0f * p1 * y;
            fy_y += 2.0f * y * (s3 + 2.0f * s4 * r);
            const float det = fx_y * fy_x - fx_x * fy_y;
            if (fabsf(det) < 1e-6f) break;
            const float dx = (fx_ * fy_y - fy_ * fx_y) / det;
            const float dy = (fy_ * fx_x - fx_ * fy_x) / det;
            x += dx; y += dy;
            if (fabsf(dx) < 1e-6f && fabsf(dy) < 1e-6f) { converged = true; break; }
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

} // namespace gsb
