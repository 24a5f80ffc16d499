// gsb_raster_rs.cu -- a7/a8 for ROLLING-SHUTTER cameras (gsplat/Cameras.cuh:268-339): every pixel has its own camera
// pose, interpolated between the start- and end-of-frame poses by the time its row / column is read out.
//
// The quadratic-form blend of gsb_raster.cu relies on one camera centre per image (the ray origin is common to all
// pixels, so the response of a Gaussian is a ratio of two quadratic forms in the pixel position).  With a rolling
// shutter the origin and the rotation change from row to row, and the response has to be evaluated per pair in the
// reference's general form (RasterizeToPixelsFromWorld3DGSFwd.cu:226-258):
//     M = S^-1 R_g^T,  gro = M (o - mu),  grd = normalize(M d),  power = -1/2 |grd x gro|^2
// No caller of the reference passes a rolling shutter (rasterizer_autograd.cpp:234,311,369); this path exists so
// that the operator API is complete, and is laid out for clarity rather than for the last cycle:
//   rs_prep      one thread per Gaussian: 64-byte record (M, mu, opacity, colour), moment row zeroed
//   rs_fwd       one CTA per 16x16 tile, one thread per pixel; records staged 256 at a time in shared memory
//   rs_bwd       same tiling, back to front; per pair the VJP of the formulas above, 16 values per pair
//                (dL/dM 9, dL/dmu 3, dL/dopacity 1, dL/dcolour 3) reduced over the warp with the 16-value butterfly
//                of gsb_raster.cuh and added to the Gaussian's moment row with one 16-lane RED
//   rs_finalize  one thread per Gaussian: dL/dM -> dL/dquat, dL/dscale (M_ij = R_ji / s_i), everything written
#include "gsb_raster.cuh"

namespace gsb {

constexpr int kRsThreads = 256;

struct __align__(16) RsRec {
    float M[9];   // row-major S^-1 R^T
    float mu[3];
    float opac;
    float r, g, b;
};
static_assert(sizeof(RsRec) == 64, "RsRec must be 64 bytes");

struct RsParams {
    uint32_t N, n_isects, W, H, tile_w, tile_h;
    const RsRec *recs;
    const float *backgrounds;
    const uint8_t *masks;
    const int32_t *tile_offsets, *flatten_ids;
    const float *viewmat0, *viewmat1, *Ks;
    int32_t camera_model, shutter;
    const float *radial, *tangential, *thin_prism;
    int32_t n_radial, n_tangential, n_thin_prism;
};

// rotation of the NORMALISED quaternion (Utils.cuh:80-102)
__device__ __forceinline__ M3<float> rotmat_normalized(const float *q, float &inv_norm, float (&qn)[4]) {
    inv_norm = rsqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    qn[0] = q[0] * inv_norm; qn[1] = q[1] * inv_norm; qn[2] = q[2] * inv_norm; qn[3] = q[3] * inv_norm;
    return rotmat_raw(qn[0], qn[1], qn[2], qn[3]);
}

__global__ void __launch_bounds__(kRsThreads) rs_prep_kernel(uint32_t N, const float *__restrict__ means,
                                                             const float *__restrict__ quats,
                                                             const float *__restrict__ scales,
                                                             const float *__restrict__ colors,
                                                             const float *__restrict__ opacities, RsRec *__restrict__ recs,
                                                             float *__restrict__ moments /*nullable*/) {
    const uint32_t g = blockIdx.x * kRsThreads + threadIdx.x;
    if (g >= N) return;
    float inv_norm, qn[4];
    const M3<float> R = rotmat_normalized(quats + (size_t)g * 4, inv_norm, qn);
    RsRec rec;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const float is = 1.0f / scales[(size_t)g * 3 + i];
#pragma unroll
        for (int j = 0; j < 3; ++j) rec.M[i * 3 + j] = R.m[j][i] * is;
    }
    rec.mu[0] = means[(size_t)g * 3]; rec.mu[1] = means[(size_t)g * 3 + 1]; rec.mu[2] = means[(size_t)g * 3 + 2];
    rec.opac = opacities[g];
    rec.r = colors[(size_t)g * 3]; rec.g = colors[(size_t)g * 3 + 1]; rec.b = colors[(size_t)g * 3 + 2];
    recs[g] = rec;
    if (moments) {
        float4 *m4 = reinterpret_cast<float4 *>(moments + (size_t)g * kMomFloats);
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        m4[0] = z; m4[1] = z; m4[2] = z; m4[3] = z;
    }
}

// World ray of a pixel under the pose of its read-out time (image_point_to_world_ray_shutter_pose, Cameras.cuh:322-339)
struct RsRay {
    float ox, oy, oz, dx, dy, dz;
    bool valid;
};
__device__ __forceinline__ RsRay rs_pixel_ray(const RsParams &p, const CamModel &cm, const CamPose &a, const CamPose &b,
                                              float px, float py) {
    RsRay r;
    r.ox = r.oy = r.oz = r.dx = r.dy = r.dz = 0.f;
    float xn, yn;
    r.valid = cam_unproject_normalized(cm, px, py, xn, yn);
    if (!r.valid) return r;
    const float inv = rsqrtf(xn * xn + yn * yn + 1.0f);
    const V3<float> dc = {xn * inv, yn * inv, inv};
    const float tau = shutter_relative_time(p.shutter, px, py, cm.W, cm.H);
    const Quat q = quat_slerp(Quat{a.qw, a.qx, a.qy, a.qz}, Quat{b.qw, b.qx, b.qy, b.qz}, tau);
    const V3<float> t = {(1.0f - tau) * a.tx + tau * b.tx, (1.0f - tau) * a.ty + tau * b.ty,
                         (1.0f - tau) * a.tz + tau * b.tz};
    // camera_ray_to_world_ray (Cameras.cuh:261-265): R_inv = mat3_cast(inverse(q))
    const float d = q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z;
    const M3<float> Ri = rotmat_raw(q.w / d, -q.x / d, -q.y / d, -q.z / d);
    const V3<float> o = mulv(Ri, t), dw = mulv(Ri, dc);
    r.ox = -o.x; r.oy = -o.y; r.oz = -o.z;
    r.dx = dw.x; r.dy = dw.y; r.dz = dw.z;
    return r;
}

struct RsPair {
    float gro[3], grd[3], grdn[3], k[3];
    float vis, alpha;
    bool ok;
};
__device__ __forceinline__ RsPair rs_pair(const RsRec &g, const RsRay &ray) {
    RsPair e;
    const float ex = ray.ox - g.mu[0], ey = ray.oy - g.mu[1], ez = ray.oz - g.mu[2];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        e.gro[i] = g.M[i * 3] * ex + g.M[i * 3 + 1] * ey + g.M[i * 3 + 2] * ez;
        e.grd[i] = g.M[i * 3] * ray.dx + g.M[i * 3 + 1] * ray.dy + g.M[i * 3 + 2] * ray.dz;
    }
    const float l2 = e.grd[0] * e.grd[0] + e.grd[1] * e.grd[1] + e.grd[2] * e.grd[2];
    const float il = l2 > 0.f ? rsqrtf(l2) : 1.0f; // safe_normalize (Utils.cuh:181-184)
    e.grdn[0] = e.grd[0] * il; e.grdn[1] = e.grd[1] * il; e.grdn[2] = e.grd[2] * il;
    e.k[0] = e.grdn[1] * e.gro[2] - e.grdn[2] * e.gro[1];
    e.k[1] = e.grdn[2] * e.gro[0] - e.grdn[0] * e.gro[2];
    e.k[2] = e.grdn[0] * e.gro[1] - e.grdn[1] * e.gro[0];
    const float power = -0.5f * (e.k[0] * e.k[0] + e.k[1] * e.k[1] + e.k[2] * e.k[2]);
    e.vis = __expf(power);
    e.alpha = fminf(kMaxAlpha, g.opac * e.vis);
    e.ok = !(power > 0.f) && e.alpha >= kAlphaThreshold;
    return e;
}

__global__ void __launch_bounds__(kRsThreads) rs_fwd_kernel(const RsParams p, float *__restrict__ renders,
                                                            float *__restrict__ alphas, int32_t *__restrict__ last_ids) {
    __shared__ RsRec s_rec[kRsThreads];
    __shared__ CamModel s_cm;
    const uint32_t tile_id = blockIdx.x, tile_y = tile_id / p.tile_w, tile_x = tile_id % p.tile_w;
    const uint32_t tid = threadIdx.x;
    const uint32_t x = tile_x * 16 + (tid & 15), y = tile_y * 16 + (tid >> 4);
    const bool inside = x < p.W && y < p.H;
    const size_t pix = (size_t)y * p.W + x;
    const bool has_bg = p.backgrounds != nullptr;
    const float bg0 = has_bg ? p.backgrounds[0] : 0.f, bg1 = has_bg ? p.backgrounds[1] : 0.f;
    const float bg2 = has_bg ? p.backgrounds[2] : 0.f;
    if (p.masks != nullptr && !p.masks[tile_id]) {
        if (inside) { renders[pix * 3] = bg0; renders[pix * 3 + 1] = bg1; renders[pix * 3 + 2] = bg2; }
        return;
    }
    if (tid == 0)
        cam_model_build(s_cm, p.camera_model, p.W, p.H, p.Ks, p.radial, p.n_radial, p.tangential, p.n_tangential,
                        p.thin_prism, p.n_thin_prism);
    __syncthreads();
    const CamPose a = cam_pose_from_viewmat(p.viewmat0), b = cam_pose_from_viewmat(p.viewmat1);
    const RsRay ray = rs_pixel_ray(p, s_cm, a, b, (float)x + 0.5f, (float)y + 0.5f);
    bool done = !inside || !ray.valid;
    const int32_t range_start = p.tile_offsets[tile_id];
    const int32_t range_end = (tile_id == p.tile_w * p.tile_h - 1) ? (int32_t)p.n_isects : p.tile_offsets[tile_id + 1];
    float T = 1.f, cr = 0.f, cg = 0.f, cb = 0.f;
    int32_t last = 0;
    for (int32_t b0 = range_start; b0 < range_end; b0 += kRsThreads) {
        if (__syncthreads_and(done)) break; // also keeps the previous batch alive until everyone is through
        const int32_t i = b0 + (int32_t)tid;
        if (i < range_end) s_rec[tid] = p.recs[p.flatten_ids[i]];
        __syncthreads();
        const int32_t cnt = min((int32_t)kRsThreads, range_end - b0);
        for (int32_t t = 0; t < cnt && !done; ++t) {
            const RsPair e = rs_pair(s_rec[t], ray);
            if (!e.ok) continue;
            const float nT = T * (1.0f - e.alpha);
            if (nT <= kMinTransmittance) { done = true; break; }
            const float w = e.alpha * T;
            cr += s_rec[t].r * w; cg += s_rec[t].g * w; cb += s_rec[t].b * w;
            last = b0 + t;
            T = nT;
        }
    }
    if (inside) {
        alphas[pix] = 1.0f - T;
        renders[pix * 3] = has_bg ? cr + T * bg0 : cr;
        renders[pix * 3 + 1] = has_bg ? cg + T * bg1 : cg;
        renders[pix * 3 + 2] = has_bg ? cb + T * bg2 : cb;
        last_ids[pix] = last;
    }
}

// slots of the moment row on this path
enum : int { kR_M = 0, kR_MU = 9, kR_OP = 12, kR_C = 13 };

__global__ void __launch_bounds__(kRsThreads) rs_bwd_kernel(const RsParams p, const float *__restrict__ render_alphas,
                                                            const int32_t *__restrict__ last_ids,
                                                            const float *__restrict__ v_render_colors,
                                                            const float *__restrict__ v_render_alphas,
                                                            float *__restrict__ moments) {
    __shared__ RsRec s_rec[kRsThreads];
    __shared__ int32_t s_gid[kRsThreads];
    __shared__ CamModel s_cm;
    const uint32_t tile_id = blockIdx.x, tile_y = tile_id / p.tile_w, tile_x = tile_id % p.tile_w;
    if (p.masks != nullptr && !p.masks[tile_id]) return;
    const uint32_t tid = threadIdx.x;
    const uint32_t x = tile_x * 16 + (tid & 15), y = tile_y * 16 + (tid >> 4);
    const bool hi16 = (tid & 16) != 0;
    if (tid == 0)
        cam_model_build(s_cm, p.camera_model, p.W, p.H, p.Ks, p.radial, p.n_radial, p.tangential, p.n_tangential,
                        p.thin_prism, p.n_thin_prism);
    __syncthreads();
    const CamPose a = cam_pose_from_viewmat(p.viewmat0), b = cam_pose_from_viewmat(p.viewmat1);
    const RsRay ray = rs_pixel_ray(p, s_cm, a, b, (float)x + 0.5f, (float)y + 0.5f);
    const bool inside = x < p.W && y < p.H && ray.valid;
    const size_t pix = (size_t)y * p.W + x;
    float T_final = 1.f, T = 1.f, vr = 0.f, vg = 0.f, vb = 0.f, va = 0.f;
    int32_t bin_final = -1;
    if (inside) {
        T_final = 1.0f - render_alphas[pix];
        T = T_final;
        bin_final = last_ids[pix];
        vr = v_render_colors[pix * 3]; vg = v_render_colors[pix * 3 + 1]; vb = v_render_colors[pix * 3 + 2];
        va = v_render_alphas[pix];
    }
    float bgdot = 0.f;
    if (p.backgrounds) bgdot = p.backgrounds[0] * vr + p.backgrounds[1] * vg + p.backgrounds[2] * vb;
    float buf_r = 0.f, buf_g = 0.f, buf_b = 0.f;
    const int32_t range_start = p.tile_offsets[tile_id];
    const int32_t range_end = (tile_id == p.tile_w * p.tile_h - 1) ? (int32_t)p.n_isects : p.tile_offsets[tile_id + 1];
    for (int32_t top = range_end - 1; top >= range_start; top -= kRsThreads) { // slot t of a batch holds index top - t
        __syncthreads();
        const int32_t i = top - (int32_t)tid;
        if (i >= range_start) {
            const int32_t gid = p.flatten_ids[i];
            s_gid[tid] = gid;
            s_rec[tid] = p.recs[gid];
        }
        __syncthreads();
        const int32_t cnt = min((int32_t)kRsThreads, top - range_start + 1);
        // a warp skips the slots behind its newest contributor (Bwd.cu:196-198)
        const int32_t wmax = __reduce_max_sync(0xffffffffu, bin_final);
        for (int32_t t = max(0, top - wmax); t < cnt; ++t) {
            const int32_t idx = top - t;
            const RsRec &g = s_rec[t];
            float v[16];
#pragma unroll
            for (int s = 0; s < 16; ++s) v[s] = 0.f;
            bool any = false;
            if (inside && idx <= bin_final) {
                const RsPair e = rs_pair(g, ray);
                if (e.ok) {
                    any = true;
                    const float ra = 1.0f / (1.0f - e.alpha);
                    T *= ra;
                    const float fac = e.alpha * T;
                    v[kR_C] = fac * vr; v[kR_C + 1] = fac * vg; v[kR_C + 2] = fac * vb;
                    float v_alpha = (g.r * T - buf_r * ra) * vr + (g.g * T - buf_g * ra) * vg + (g.b * T - buf_b * ra) * vb;
                    v_alpha += T_final * ra * va;
                    v_alpha -= T_final * ra * bgdot;
                    if (g.opac * e.vis <= kMaxAlpha) {
                        const float v_vis = g.opac * v_alpha;
                        const float s2 = -e.vis * v_vis; // dL/dk = 2 (-1/2 vis v_vis) k
                        const float vk[3] = {s2 * e.k[0], s2 * e.k[1], s2 * e.k[2]};
                        // k = grdn x gro:  dL/dgrdn = gro x vk,  dL/dgro = vk x grdn
                        const float vn[3] = {e.gro[1] * vk[2] - e.gro[2] * vk[1], e.gro[2] * vk[0] - e.gro[0] * vk[2],
                                             e.gro[0] * vk[1] - e.gro[1] * vk[0]};
                        const float vo[3] = {vk[1] * e.grdn[2] - vk[2] * e.grdn[1], vk[2] * e.grdn[0] - vk[0] * e.grdn[2],
                                             vk[0] * e.grdn[1] - vk[1] * e.grdn[0]};
                        // normalisation backward (Utils.cuh:186-194)
                        const float l2 = e.grd[0] * e.grd[0] + e.grd[1] * e.grd[1] + e.grd[2] * e.grd[2];
                        float vd[3] = {vn[0], vn[1], vn[2]};
                        if (l2 > 0.f) {
                            const float il = rsqrtf(l2), il3 = il * il * il;
                            const float dd = vn[0] * e.grd[0] + vn[1] * e.grd[1] + vn[2] * e.grd[2];
                            vd[0] = vn[0] * il - e.grd[0] * il3 * dd; vd[1] = vn[1] * il - e.grd[1] * il3 * dd;
                            vd[2] = vn[2] * il - e.grd[2] * il3 * dd;
                        }
                        // gro = M (o - mu), grd = M d:  dL/dM = vd (x) d + vo (x) (o - mu),  dL/dmu = -M^T vo
                        const float ex = ray.ox - g.mu[0], ey = ray.oy - g.mu[1], ez = ray.oz - g.mu[2];
#pragma unroll
                        for (int r = 0; r < 3; ++r) {
                            v[kR_M + r * 3] = vd[r] * ray.dx + vo[r] * ex;
                            v[kR_M + r * 3 + 1] = vd[r] * ray.dy + vo[r] * ey;
                            v[kR_M + r * 3 + 2] = vd[r] * ray.dz + vo[r] * ez;
                        }
#pragma unroll
                        for (int c = 0; c < 3; ++c)
                            v[kR_MU + c] = -(g.M[c] * vo[0] + g.M[3 + c] * vo[1] + g.M[6 + c] * vo[2]);
                        v[kR_OP] = e.vis * v_alpha;
                    }
                    buf_r += g.r * fac; buf_g += g.g * fac; buf_b += g.b * fac;
                }
            }
            if (!__any_sync(0xffffffffu, any)) continue;
            // butterfly16_preswapped wants lanes 16..31 to hold slot (i ^ 8) in register i
            float R[16];
#pragma unroll
            for (int s = 0; s < 16; ++s) R[s] = hi16 ? v[s ^ 8] : v[s];
            butterfly16_preswapped(R);
            if ((tid & 1) == 0) red_add_f32(moments + (size_t)s_gid[t] * kMomFloats + ((tid & 31) >> 1), R[0]);
        }
    }
}

__global__ void __launch_bounds__(kRsThreads) rs_finalize_kernel(uint32_t N, const float *__restrict__ quats,
                                                                 const float *__restrict__ scales,
                                                                 const float *__restrict__ moments,
                                                                 float *__restrict__ v_means, float *__restrict__ v_quats,
                                                                 float *__restrict__ v_scales, float *__restrict__ v_colors,
                                                                 float *__restrict__ v_opacities) {
    const uint32_t g = blockIdx.x * kRsThreads + threadIdx.x;
    if (g >= N) return;
    float m[16];
    const float4 *m4 = reinterpret_cast<const float4 *>(moments + (size_t)g * kMomFloats);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float4 v = m4[i];
        m[i * 4] = v.x; m[i * 4 + 1] = v.y; m[i * 4 + 2] = v.z; m[i * 4 + 3] = v.w;
    }
    float inv_norm, qn[4];
    const M3<float> R = rotmat_normalized(quats + (size_t)g * 4, inv_norm, qn);
    // M_ij = R_ji / s_i:  dL/dR_ji = dL/dM_ij / s_i,  dL/ds_i = -(1 / s_i^2) sum_j R_ji dL/dM_ij   (Utils.cuh:128-158)
    float vR[3][3], vs[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const float is = 1.0f / scales[(size_t)g * 3 + i];
        float acc = 0.f;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            vR[j][i] = m[kR_M + i * 3 + j] * is;
            acc += R.m[j][i] * m[kR_M + i * 3 + j];
        }
        vs[i] = -is * is * acc;
    }
    float vq[4];
    quat_vjp_from_rotmat_grad(qn, inv_norm, vR, vq);
    v_means[(size_t)g * 3] = m[kR_MU]; v_means[(size_t)g * 3 + 1] = m[kR_MU + 1]; v_means[(size_t)g * 3 + 2] = m[kR_MU + 2];
    v_scales[(size_t)g * 3] = vs[0]; v_scales[(size_t)g * 3 + 1] = vs[1]; v_scales[(size_t)g * 3 + 2] = vs[2];
    reinterpret_cast<float4 *>(v_quats)[g] = make_float4(vq[0], vq[1], vq[2], vq[3]);
    v_colors[(size_t)g * 3] = m[kR_C]; v_colors[(size_t)g * 3 + 1] = m[kR_C + 1]; v_colors[(size_t)g * 3 + 2] = m[kR_C + 2];
    v_opacities[g] = m[kR_OP];
}

static void fill_rs(RsParams &p, uint32_t N, uint64_t n_isects, uint32_t W, uint32_t H, const GsbCamera *cam,
                    const RsRec *recs, const float *backgrounds, const uint8_t *masks, const int32_t *tile_offsets,
                    const int32_t *flatten_ids) {
    p.N = N; p.n_isects = (uint32_t)n_isects; p.W = W; p.H = H;
    p.tile_w = (W + 15) / 16; p.tile_h = (H + 15) / 16;
    p.recs = recs; p.backgrounds = backgrounds; p.masks = masks; p.tile_offsets = tile_offsets; p.flatten_ids = flatten_ids;
    p.viewmat0 = cam->viewmats0; p.viewmat1 = cam->viewmats1 ? cam->viewmats1 : cam->viewmats0; p.Ks = cam->Ks;
    p.camera_model = cam->camera_model; p.shutter = cam->shutter_type;
    p.radial = cam->radial_coeffs; p.tangential = cam->tangential_coeffs; p.thin_prism = cam->thin_prism_coeffs;
    p.n_radial = cam->radial_count; p.n_tangential = cam->tangential_count; p.n_thin_prism = cam->thin_prism_count;
}

// Entry points used by gsb_raster_fwd / gsb_raster_bwd when the camera has a rolling shutter (same workspace layout:
// [records][moments]).
int raster_rs_fwd(uint32_t N, uint64_t n_isects, const float *means, const float *quats, const float *scales,
                  const float *colors, const float *opacities, const float *backgrounds, const uint8_t *masks, uint32_t W,
                  uint32_t H, const GsbCamera *cam, const int32_t *tile_offsets, const int32_t *flatten_ids, float *renders,
                  float *alphas, int32_t *last_ids, void *workspace, cudaStream_t s) {
    RsRec *recs = reinterpret_cast<RsRec *>(workspace);
    if (N > 0 && n_isects > 0) {
        ProfScope ps("raster_prep", s);
        rs_prep_kernel<<<(N + kRsThreads - 1) / kRsThreads, kRsThreads, 0, s>>>(N, means, quats, scales, colors, opacities,
                                                                                recs, nullptr);
        GSB_LAUNCH_CHECK();
    }
    RsParams p;
    fill_rs(p, N, n_isects, W, H, cam, recs, backgrounds, masks, tile_offsets, flatten_ids);
    {
        ProfScope ps("raster_fwd", s);
        rs_fwd_kernel<<<p.tile_w * p.tile_h, kRsThreads, 0, s>>>(p, renders, alphas, last_ids);
    }
    GSB_LAUNCH_CHECK();
    return GSB_OK;
}

int raster_rs_bwd(uint32_t N, uint64_t n_isects, const float *means, const float *quats, const float *scales,
                  const float *colors, const float *opacities, const float *backgrounds, const uint8_t *masks, uint32_t W,
                  uint32_t H, const GsbCamera *cam, const int32_t *tile_offsets, const int32_t *flatten_ids,
                  const float *render_alphas, const int32_t *last_ids, const float *v_render_colors,
                  const float *v_render_alphas, float *v_means, float *v_quats, float *v_scales, float *v_colors,
                  float *v_opacities, void *workspace, size_t rec_bytes, cudaStream_t s) {
    RsRec *recs = reinterpret_cast<RsRec *>(workspace);
    float *moments = reinterpret_cast<float *>(reinterpret_cast<char *>(workspace) + rec_bytes);
    {
        ProfScope ps("raster_prep", s);
        rs_prep_kernel<<<(N + kRsThreads - 1) / kRsThreads, kRsThreads, 0, s>>>(N, means, quats, scales, colors, opacities,
                                                                                recs, moments);
    }
    GSB_LAUNCH_CHECK();
    RsParams p;
    fill_rs(p, N, n_isects, W, H, cam, recs, backgrounds, masks, tile_offsets, flatten_ids);
    {
        ProfScope ps("raster_bwd", s);
        rs_bwd_kernel<<<p.tile_w * p.tile_h, kRsThreads, 0, s>>>(p, render_alphas, last_ids, v_render_colors, v_render_alphas,
                                                                 moments);
    }
    GSB_LAUNCH_CHECK();
    {
        ProfScope ps("raster_finalize", s);
        rs_finalize_kernel<<<(N + kRsThreads - 1) / kRsThreads, kRsThreads, 0, s>>>(N, quats, scales, moments, v_means, v_quats,
                                                                                    v_scales, v_colors, v_opacities);
    }
    GSB_LAUNCH_CHECK();
    return GSB_OK;
}

} // namespace gsb
