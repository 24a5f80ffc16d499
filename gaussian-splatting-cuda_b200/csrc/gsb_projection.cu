// gsb_projection.cu -- a1: unscented-transform projection of world Gaussians to image Gaussians.
//
// Implements gsplat::projection_ut_3dgs_fused (reference: gsplat/ProjectionUT3DGSFused.cu:17-203,
// gsplat/Cameras.cuh:1034-1150, gsplat/Utils.cuh:171-179) for the perfect-pinhole / global-shutter
// camera, the only configuration the reference's callers reach (rasterizer_autograd.cpp:217-237).
//
// B200 notes: the kernel is a pure stream (44 B in, <= 36 B out per Gaussian, ~500 FLOP), so the
// design goal is full-sector HBM traffic: a CTA stages its 256-Gaussian slab of the AoS inputs
// (means/scales 12 B rows, quats 16 B rows, opacities) into shared memory with TMA bulk copies
// (cp.async.bulk, one per array) and writes the 12-byte conic rows back through shared memory
// with a bulk store, so every global transaction is a full 128-byte line.
//
// This file is compiled with -fmad=false: radii come out of ceilf() and feed the bit-exact tile
// intersection, so the arithmetic deliberately follows the reference's operation order without
// FMA contraction (the CPU oracle is built with -ffp-contract=off for the same reason).
#include "gsb_camera.cuh"

namespace gsb {

constexpr int kProjThreads = 256;

struct ProjParams {
    uint32_t C, N;
    const float *means, *quats, *scales, *opacities;
    const float *viewmats0, *Ks;
    uint32_t W, H;
    float eps2d, near_plane, far_plane, radius_clip;
    GsbUTParams ut;
    int32_t camera_model;
    const float *radial, *tangential, *thin_prism; // per-camera coefficient rows, nullable
    int32_t n_radial, n_tangential, n_thin_prism;
    int32_t *radii;
    float *means2d, *depths, *conics, *compensations;
    uint32_t inputs_aligned16; // every input base pointer is 16-byte aligned: full slabs may use bulk copies
};

// GLM operator*(quat, vec3): v + 2 (w (u x v) + u x (u x v))   [glm/detail/type_quat.inl]
__device__ __forceinline__ V3<float> quat_rotate(float qw, float qx, float qy, float qz, V3<float> v) {
    const V3<float> u = {qx, qy, qz};
    const V3<float> uv = cross(u, v);
    const V3<float> uuv = cross(u, uv);
    return v + ((uv * qw) + uuv) * 2.0f;
}

__global__ void __launch_bounds__(kProjThreads) projection_ut_kernel(const ProjParams p) {
    __shared__ __align__(128) float s_means[kProjThreads * 3];
    __shared__ __align__(128) float s_scales[kProjThreads * 3];
    __shared__ __align__(128) float s_quats[kProjThreads * 4];
    __shared__ __align__(128) float s_opac[kProjThreads];
    __shared__ __align__(128) float s_conics[kProjThreads * 3];
    __shared__ __align__(8) uint64_t s_bar;

    const uint32_t cid = blockIdx.y;
    const uint32_t g0 = blockIdx.x * kProjThreads;
    const uint32_t cnt = min((uint32_t)kProjThreads, p.N - g0);
    const uint32_t tid = threadIdx.x;
    // cp.async.bulk needs 16-byte aligned global addresses: true for every full slab when the base pointers
    // are (256 rows x 12 / 16 / 4 bytes), not for a contiguous view such as means[1:] -- the reference accepts
    // those, so they take the plain staging loop instead of faulting
    const bool full = (cnt == kProjThreads);
    const bool bulk_in = full && p.inputs_aligned16;

    // ---- stage the slab ------------------------------------------------------------------
    if (bulk_in) {
        if (tid == 0) {
            mbar_init(&s_bar, 1);
            mbar_fence_init();
        }
        __syncthreads();
        if (tid == 0) {
            const uint32_t bytes = kProjThreads * (12 + 12 + 16) + (p.opacities ? kProjThreads * 4 : 0);
            mbar_arrive_expect_tx(&s_bar, bytes);
            bulk_g2s(s_means, p.means + (size_t)g0 * 3, kProjThreads * 12, &s_bar);
            bulk_g2s(s_scales, p.scales + (size_t)g0 * 3, kProjThreads * 12, &s_bar);
            bulk_g2s(s_quats, p.quats + (size_t)g0 * 4, kProjThreads * 16, &s_bar);
            if (p.opacities) bulk_g2s(s_opac, p.opacities + g0, kProjThreads * 4, &s_bar);
        }
        mbar_wait(&s_bar, 0);
    } else {
        for (uint32_t i = tid; i < cnt * 3; i += kProjThreads) {
            s_means[i] = p.means[(size_t)g0 * 3 + i];
            s_scales[i] = p.scales[(size_t)g0 * 3 + i];
        }
        for (uint32_t i = tid; i < cnt * 4; i += kProjThreads) s_quats[i] = p.quats[(size_t)g0 * 4 + i];
        if (p.opacities)
            for (uint32_t i = tid; i < cnt; i += kProjThreads) s_opac[i] = p.opacities[g0 + i];
        __syncthreads();
    }

    // ---- camera (Cameras.cuh:33-71, 268-280) -----------------------------------------------
    __shared__ CamModel s_cm;
    if (tid == 0)
        cam_model_build(s_cm, p.camera_model, p.W, p.H, p.Ks + cid * 9,
                        p.radial ? p.radial + (size_t)cid * p.n_radial : nullptr, p.n_radial,
                        p.tangential ? p.tangential + (size_t)cid * p.n_tangential : nullptr, p.n_tangential,
                        p.thin_prism ? p.thin_prism + (size_t)cid * p.n_thin_prism : nullptr, p.n_thin_prism);
    __syncthreads();
    const CamPose pose = cam_pose_from_viewmat(p.viewmats0 + cid * 16);
    // centre-of-shutter pose = slerp(q, q, 0.5), 0.5 t + 0.5 t (global shutter: start == end)
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
    const V3<float> t_mid = {0.5f * pose.tx + 0.5f * pose.tx, 0.5f * pose.ty + 0.5f * pose.ty,
                             0.5f * pose.tz + 0.5f * pose.tz};
    const V3<float> t0 = {pose.tx, pose.ty, pose.tz};

    // ---- per-Gaussian work -----------------------------------------------------------------
    int32_t rx_i = 0, ry_i = 0;
    float o_mx = 0.f, o_my = 0.f, o_depth = 0.f, o_c0 = 0.f, o_c1 = 0.f, o_c2 = 0.f, o_comp = 0.f;
    bool keep = false;
    if (tid < cnt) {
        do {
            const V3<float> mean = {s_means[tid * 3], s_means[tid * 3 + 1], s_means[tid * 3 + 2]};
            const float sc[3] = {s_scales[tid * 3], s_scales[tid * 3 + 1], s_scales[tid * 3 + 2]};
            float qw = s_quats[tid * 4], qx = s_quats[tid * 4 + 1], qy = s_quats[tid * 4 + 2], qz = s_quats[tid * 4 + 3];
            { // glm::normalize(quat)
                const float len = sqrtf(qw * qw + qx * qx + qy * qy + qz * qz);
                if (len <= 0.f) { qw = 1.f; qx = qy = qz = 0.f; }
                else { const float ool = 1.0f / len; qw *= ool; qx *= ool; qy *= ool; qz *= ool; }
            }
            const V3<float> mean_c = quat_rotate(mw, mx, my, mz, mean) + t_mid;
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
                const V3<float> cam = quat_rotate(pose.qw, pose.qx, pose.qy, pose.qz, pt) + t0;
                float px, py;
                const bool pv = cam_project(s_cm, cam, p.ut.in_image_margin_factor, px, py);
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
            if (p.opacities) {
                float opacity = s_opac[tid];
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
            keep = true;
            rx_i = (int32_t)radius_x; ry_i = (int32_t)radius_y;
            o_mx = mx2; o_my = my2; o_depth = mean_c.z;
            o_c0 = cyy * ood; o_c1 = -cxy * ood; o_c2 = cxx * ood;
            o_comp = compensation;
        } while (false);
    }

    // ---- outputs ---------------------------------------------------------------------------
    const size_t idx = (size_t)cid * p.N + g0 + tid;
    if (tid < cnt) {
        reinterpret_cast<int2 *>(p.radii)[idx] = make_int2(rx_i, ry_i);
        if (keep) {
            reinterpret_cast<float2 *>(p.means2d)[idx] = make_float2(o_mx, o_my);
            p.depths[idx] = o_depth;
            if (p.compensations) p.compensations[idx] = o_comp;
        }
    }
    // conics: culled rows must stay untouched (Projection.cpp:70-73), so the full-line bulk store
    // is only used when every Gaussian of the slab survived; otherwise scalar stores.
    const bool all_keep = __syncthreads_and(keep || tid >= cnt) != 0;
    float *cdst = p.conics + ((size_t)cid * p.N + g0) * 3;
    if (all_keep && full && ((reinterpret_cast<uintptr_t>(cdst) & 15) == 0)) {
        s_conics[tid * 3] = o_c0; s_conics[tid * 3 + 1] = o_c1; s_conics[tid * 3 + 2] = o_c2;
        fence_proxy_async();
        __syncthreads();
        if (tid == 0) {
            bulk_s2g(cdst, s_conics, kProjThreads * 12);
            bulk_commit();
            bulk_wait_read_all();
        }
    } else if (keep) {
        cdst[tid * 3] = o_c0; cdst[tid * 3 + 1] = o_c1; cdst[tid * 3 + 2] = o_c2;
    }
}

} // namespace gsb

extern "C" int gsb_projection_ut(
    uint32_t C, uint32_t N, const float *means, const float *quats, const float *scales,
    const float *opacities, const GsbCamera *cam, uint32_t image_width, uint32_t image_height,
    float eps2d, float near_plane, float far_plane, float radius_clip, int32_t *radii, float *means2d,
    float *depths, float *conics, float *compensations, gsb_stream_t stream) {
    if (!cam || !cam->viewmats0 || !cam->Ks) return GSB_E_INVALID;
    if (C == 0 || N == 0) return GSB_OK; // ProjectionUT3DGSFused.cu:242-245
    if (!means || !quats || !scales || !radii || !means2d || !depths || !conics) return GSB_E_INVALID;
    if ((cam->camera_model != GSB_CAMERA_PINHOLE && cam->camera_model != GSB_CAMERA_FISHEYE) || cam->viewmats1 ||
        cam->shutter_type != GSB_SHUTTER_GLOBAL)
        return GSB_E_UNSUPPORTED; // orthographic / rolling shutter: no caller of the reference uses them
    gsb::ProjParams p;
    p.C = C; p.N = N;
    p.means = means; p.quats = quats; p.scales = scales; p.opacities = opacities;
    p.viewmats0 = cam->viewmats0; p.Ks = cam->Ks;
    p.W = image_width; p.H = image_height;
    p.eps2d = eps2d; p.near_plane = near_plane; p.far_plane = far_plane; p.radius_clip = radius_clip;
    p.ut = cam->ut;
    p.camera_model = cam->camera_model;
    p.radial = cam->radial_coeffs; p.tangential = cam->tangential_coeffs; p.thin_prism = cam->thin_prism_coeffs;
    p.n_radial = cam->radial_count; p.n_tangential = cam->tangential_count; p.n_thin_prism = cam->thin_prism_count;
    p.radii = radii; p.means2d = means2d; p.depths = depths; p.conics = conics; p.compensations = compensations;
    const uintptr_t bits = reinterpret_cast<uintptr_t>(means) | reinterpret_cast<uintptr_t>(quats) |
                           reinterpret_cast<uintptr_t>(scales) | reinterpret_cast<uintptr_t>(opacities);
    p.inputs_aligned16 = (bits & 15) == 0 ? 1u : 0u;
    dim3 grid((N + gsb::kProjThreads - 1) / gsb::kProjThreads, C);
    {
        gsb::ProfScope ps("projection_ut", gsb::as_stream(stream));
        gsb::projection_ut_kernel<<<grid, gsb::kProjThreads, 0, gsb::as_stream(stream)>>>(p);
    }
    GSB_LAUNCH_CHECK();
    return GSB_OK;
}
