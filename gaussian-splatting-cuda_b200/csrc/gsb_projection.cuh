// gsb_projection.cuh -- the per-Gaussian arithmetic of the unscented-transform projection (a1), shared by the
// stand-alone operator (gsb_projection.cu) and the fused front kernel (gsb_fused.cu).
//
// Restates gsplat/ProjectionUT3DGSFused.cu:47-202 + gsplat/Cameras.cuh:1034-1150 in the reference's operation
// order.  Every translation unit that includes this header is compiled with -fmad=false: the radii come out of
// ceilf() and feed the bit-exact tile intersection, so the arithmetic must not be FMA-contracted (the CPU oracle
// is built with -ffp-contract=off for the same reason).
#pragma once

#include "gsb_camera.cuh"

namespace gsb {

// GLM operator*(quat, vec3): v + 2 (w (u x v) + u x (u x v))   [glm/detail/type_quat.inl]
__device__ __forceinline__ V3<float> quat_rotate(float qw, float qx, float qy, float qz, V3<float> v) {
    const V3<float> u = {qx, qy, qz};
    const V3<float> uv = cross(u, v);
    const V3<float> uuv = cross(u, uv);
    return v + ((uv * qw) + uuv) * 2.0f;
}

// Scalars of the projection call.
struct ProjConsts {
    uint32_t W, H;
    float eps2d, near_plane, far_plane, radius_clip;
    GsbUTParams ut;
};

// Camera pose terms: the start pose (global shutter: also the end pose) and the centre-of-shutter pose
// slerp(q, q, 0.5), 0.5 t + 0.5 t of Cameras.cuh:268-280, evaluated like GLM does.
struct ProjPose {
    CamPose pose;       // start of frame
    CamPose pose1;      // end of frame (== pose for a global shutter)
    int shutter;        // GSB_SHUTTER_*
    float mw, mx, my, mz;
    V3<float> t_mid, t0;
};
__device__ __forceinline__ ProjPose proj_pose_from_viewmat(const float *viewmat, const float *viewmat1 = nullptr,
                                                           int shutter = GSB_SHUTTER_GLOBAL) {
    ProjPose r;
    const CamPose pose = cam_pose_from_viewmat(viewmat);
    if (viewmat1 != nullptr && shutter != GSB_SHUTTER_GLOBAL) { // centre-of-shutter pose between two different poses
        const CamPose p1 = cam_pose_from_viewmat(viewmat1);
        const Quat qm = quat_slerp(Quat{pose.qw, pose.qx, pose.qy, pose.qz}, Quat{p1.qw, p1.qx, p1.qy, p1.qz}, 0.5f);
        r.pose = pose; r.pose1 = p1; r.shutter = shutter;
        r.mw = qm.w; r.mx = qm.x; r.my = qm.y; r.mz = qm.z;
        r.t_mid = {0.5f * pose.tx + 0.5f * p1.tx, 0.5f * pose.ty + 0.5f * p1.ty, 0.5f * pose.tz + 0.5f * p1.tz};
        r.t0 = {pose.tx, pose.ty, pose.tz};
        return r;
    }
    float mw, mx, my, mz;
    {
        const float cosT = pose.qw * pose.qw + pose.qx * pose.qx + pose.qy * pose.qy + pose.qz * pose.qz;
        if (cosT > 1.0f - 1.1920929e-07f) {
            mw = pose.qw * 0.5f + pose.qw * 0.5f; mx = pose.qx * 0.5f + pose.qx * 0.5f;
            my = pose.qy * 0.5f + pose.qy * 0.5f; mz = pose.qz * 0.5f + pose.qz * 0.5f;
        } else {
            const float ang = acosf(cosT);
            const float s0 = sinf(0.5f * ang), sd = sinf(ang);
            mw = (s0 * pose.qw + s0 * pose.qw) / sd; mx = (s0 * pose.qx + s0 * pose.qx) / sd;
            my = (s0 * pose.qy + s0 * pose.qy) / sd; mz = (s0 * pose.qz + s0 * pose.qz) / sd;
        }
    }
    r.pose = pose; r.pose1 = pose; r.shutter = GSB_SHUTTER_GLOBAL;
    r.mw = mw; r.mx = mx; r.my = my; r.mz = mz;
    r.t_mid = {0.5f * pose.tx + 0.5f * pose.tx, 0.5f * pose.ty + 0.5f * pose.ty, 0.5f * pose.tz + 0.5f * pose.tz};
    r.t0 = {pose.tx, pose.ty, pose.tz};
    return r;
}

// world_point_to_image_point_shutter_pose for a rolling shutter (Cameras.cuh:371-413): the projection with the
// start-of-frame pose (or, if that one is invalid, the end-of-frame pose) seeds ten fixed-point iterations
// "time of the row / column the point falls on -> interpolated pose -> projection".  (px, py, valid) enter with the
// start-pose projection and leave with the result.
__device__ __forceinline__ bool project_rolling(const ProjConsts &p, const CamModel &cm, const ProjPose &pp, V3<float> pt,
                                                bool valid_start, float &px, float &py) {
    const CamPose &a = pp.pose, &b = pp.pose1;
    const V3<float> t1 = {b.tx, b.ty, b.tz};
    float ex, ey;
    const bool valid_end = cam_project(cm, quat_rotate(b.qw, b.qx, b.qy, b.qz, pt) + t1, p.ut.in_image_margin_factor, ex, ey);
    if (!valid_start) {
        px = ex; py = ey;
        if (!valid_end) return false;
    }
    const Quat qa{a.qw, a.qx, a.qy, a.qz}, qb{b.qw, b.qx, b.qy, b.qz};
    for (int it = 0; it < 10; ++it) {
        const float tau = shutter_relative_time(pp.shutter, px, py, cm.W, cm.H);
        const Quat q = quat_slerp(qa, qb, tau);
        const V3<float> t = {(1.0f - tau) * a.tx + tau * b.tx, (1.0f - tau) * a.ty + tau * b.ty,
                             (1.0f - tau) * a.tz + tau * b.tz};
        float nx, ny;
        cam_project(cm, quat_rotate(q.w, q.x, q.y, q.z, pt) + t, p.ut.in_image_margin_factor, nx, ny);
        px = nx; py = ny;
    }
    return true;
}

struct ProjResult {
    bool keep;          // false: culled (radii 0, other outputs untouched)
    int32_t rx, ry;
    float mx, my, depth;
    float c0, c1, c2;   // conic
    float comp;
};

// One Gaussian: activated scale, quaternion as stored (normalised here like the reference, :63), optional opacity.
// kRolling = false compiles the rolling-shutter iteration out (global-shutter callers keep their register budget).
template <bool kRolling>
__device__ __forceinline__ ProjResult project_gaussian(const ProjConsts &p, const CamModel &s_cm, const ProjPose &pp,
                                                       V3<float> mean, const float (&sc)[3], float qw, float qx,
                                                       float qy, float qz, bool has_opacity, float opacity_in) {
    ProjResult out;
    out.keep = false;
    out.rx = out.ry = 0;
    out.mx = out.my = out.depth = out.c0 = out.c1 = out.c2 = out.comp = 0.f;
    const CamPose &pose = pp.pose;
    do {
        { // glm::normalize(quat)
            const float len = sqrtf(qw * qw + qx * qx + qy * qy + qz * qz);
            if (len <= 0.f) { qw = 1.f; qx = qy = qz = 0.f; }
            else { const float ool = 1.0f / len; qw *= ool; qx *= ool; qy *= ool; qz *= ool; }
        }
        const V3<float> mean_c = quat_rotate(pp.mw, pp.mx, pp.my, pp.mz, mean) + pp.t_mid;
        if (mean_c.z < p.near_plane || mean_c.z > p.far_plane) break;

        // sigma points (Cameras.cuh:1034-1083)
        const float alpha = p.ut.alpha, beta = p.ut.beta, kappa = p.ut.kappa;
        const float D = 3.0f;
        const float lambda = alpha * alpha * (D + kappa) - D;
        const M3<float> R = rotmat_raw(qw, qx, qy, qz);
        const float sq = sqrtf(D + lambda);
        const float w0m = lambda / (D + lambda);
        const float w0c = lambda / (D + lambda) + (1.0f - alpha * alpha + beta);
        const float wi = 1.0f / (2.0f * (D + lambda));

        float ipx[7], ipy[7];
        float mx2 = 0.f, my2 = 0.f;
        bool valid = p.ut.require_all_sigma_points_valid != 0;
        bool early = false;
#pragma unroll
        for (int i = 0; i < 7; ++i) {
            V3<float> pt = mean;
            if (i > 0) {
                const int a = (i - 1) % 3;
                const V3<float> delta = col(R, a) * (sq * sc[a]);
                pt = (i <= 3) ? (mean + delta) : (mean - delta);
            }
            const V3<float> cam = quat_rotate(pose.qw, pose.qx, pose.qy, pose.qz, pt) + pp.t0;
            float px, py;
            bool pv = cam_project(s_cm, cam, p.ut.in_image_margin_factor, px, py);
            if constexpr (kRolling) {
                if (pp.shutter != GSB_SHUTTER_GLOBAL) pv = project_rolling(p, s_cm, pp, pt, pv, px, py);
            }
            if (p.ut.require_all_sigma_points_valid) {
                valid = valid && pv;
                if (!pv) { early = true; break; }
            } else {
                valid = valid || pv;
            }
            ipx[i] = px; ipy[i] = py;
            const float w = (i == 0) ? w0m : wi;
            mx2 += w * px;
            my2 += w * py;
        }
        if (early || !valid) break;
        float cxx = 0.f, cxy = 0.f, cyy = 0.f;
#pragma unroll
        for (int i = 0; i < 7; ++i) {
            const float w = (i == 0) ? w0c : wi;
            const float dx = ipx[i] - mx2, dy = ipy[i] - my2;
            cxx += w * (dx * dx);
            cxy += w * (dx * dy);
            cyy += w * (dy * dy);
        }
        // add_blur (Utils.cuh:171-179)
        const float det_orig = cxx * cyy - cxy * cxy;
        cxx += p.eps2d;
        cyy += p.eps2d;
        const float det = cxx * cyy - cxy * cxy;
        const float compensation = sqrtf(fmaxf(0.f, det_orig / det));
        if (det <= 0.f) break;
        const float ood = 1.0f / (cxx * cyy - cxy * cxy); // glm::inverse(mat2)

        float extend = 3.33f;
        if (has_opacity) {
            float opacity = opacity_in;
            opacity *= compensation; // multiplied even when compensations are not returned (:156-157)
            if (opacity < kAlphaThreshold) break;
            extend = fminf(extend, sqrtf(2.0f * logf(opacity / kAlphaThreshold)));
        }
        const float b = 0.5f * (cxx + cyy);
        const float tmp = sqrtf(fmaxf(0.01f, b * b - det));
        const float v1 = b + tmp;
        const float r1 = extend * sqrtf(v1);
        const float radius_x = ceilf(fminf(extend * sqrtf(cxx), r1));
        const float radius_y = ceilf(fminf(extend * sqrtf(cyy), r1));
        if (radius_x <= p.radius_clip && radius_y <= p.radius_clip) break;
        if (mx2 + radius_x <= 0 || mx2 - radius_x >= (float)p.W || my2 + radius_y <= 0 ||
            my2 - radius_y >= (float)p.H)
            break;
        out.keep = true;
        out.rx = (int32_t)radius_x; out.ry = (int32_t)radius_y;
        out.mx = mx2; out.my = my2; out.depth = mean_c.z;
        out.c0 = cyy * ood; out.c1 = -cxy * ood; out.c2 = cxx * ood;
        out.comp = compensation;
    } while (false);
    return out;
}

} // namespace gsb
