/*
 * gsb200.h -- C ABI of the B200-native (sm_100a) 3DGUT rasterizer hot path.
 *
 * This is the drop-in boundary underneath the reference's C++ operator API
 * (namespace gsplat, /root/reference/gsplat/Ops.h:12-165).  Every entry point
 * below states which reference operator it implements; the thin libtorch shim
 * in gaussian-splatting-cuda_b200/shim/Ops.cpp exports the identical gsplat::
 * signatures on top of it, so the reference's src/training and src/rendering
 * link against it unchanged.
 *
 * Conventions
 *  - plain C: raw DEVICE pointers, sizes, POD structs, a cudaStream_t.  No torch
 *    types, no exceptions.  All tensors are contiguous row-major float32 unless
 *    stated; layouts are exactly the reference's (Ops.h comments).
 *  - nothing here allocates device memory: outputs and workspaces are provided
 *    by the caller (the shim allocates them through the torch caching allocator,
 *    like gsplat/Common.h:23-30 does for CUB).  Workspace sizes come from the
 *    *_workspace() queries; workspaces need 256-byte alignment.
 *  - every call is asynchronous on `stream` and re-entrant (no static mutable
 *    state) -- the reference calls these ops from the training thread and the
 *    viewer thread.
 *  - return value: GSB_OK (0), a negative GSB_E_* code, or a positive
 *    cudaError_t from a failed launch.  gsb_error_string() names them.
 *  - there is NO CPU fallback: on a machine without an sm_100 device every
 *    compute entry point returns a CUDA error.
 */
#ifndef GSB200_H_
#define GSB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef GSB_API
#define GSB_API __attribute__((visibility("default")))
#endif

/* Opaque to C callers that do not include cuda_runtime.h. */
typedef struct CUstream_st *gsb_stream_t;

enum {
    GSB_OK = 0,
    GSB_E_INVALID = -1,     /* bad argument (null pointer, zero size where not allowed) */
    GSB_E_UNSUPPORTED = -2, /* valid in the reference but not implemented here */
    GSB_E_WORKSPACE = -3    /* workspace too small / misaligned */
};

/* gsplat::CameraModelType, gsplat/Common.h:46-50 */
enum { GSB_CAMERA_PINHOLE = 0, GSB_CAMERA_ORTHO = 1, GSB_CAMERA_FISHEYE = 2 };

/* ShutterType, gsplat/Cameras.h:16-22 */
enum {
    GSB_SHUTTER_ROLLING_TOP_TO_BOTTOM = 0,
    GSB_SHUTTER_ROLLING_LEFT_TO_RIGHT = 1,
    GSB_SHUTTER_ROLLING_BOTTOM_TO_TOP = 2,
    GSB_SHUTTER_ROLLING_RIGHT_TO_LEFT = 3,
    GSB_SHUTTER_GLOBAL = 4
};

/* UnscentedTransformParameters, gsplat/Cameras.h:27-44 */
typedef struct GsbUTParams {
    float alpha;
    float beta;
    float kappa;
    float in_image_margin_factor;
    int32_t require_all_sigma_points_valid;
} GsbUTParams;

/* Camera block shared by projection and the from-world rasterizer.
 * viewmats1 / distortion pointers are nullable; *_count gives the number of
 * floats per camera behind each distortion pointer (the reference reads a fixed
 * 6/2/4 regardless of what the caller allocated; here the
 * count is explicit and missing coefficients are zero-filled). */
typedef struct GsbCamera {
    const float *viewmats0; /* [C,4,4] world->camera, row-major */
    const float *viewmats1; /* [C,4,4] end-of-frame pose for rolling shutter, or NULL */
    const float *Ks;        /* [C,3,3] */
    int32_t camera_model;   /* GSB_CAMERA_* */
    int32_t shutter_type;   /* GSB_SHUTTER_* */
    const float *radial_coeffs;
    int32_t radial_count;
    const float *tangential_coeffs;
    int32_t tangential_count;
    const float *thin_prism_coeffs;
    int32_t thin_prism_count;
    GsbUTParams ut;
} GsbCamera;

GSB_API const char *gsb_error_string(int code);
GSB_API int gsb_version(void);

/* Diagnostics (no effect on results).  gsb_launch_count: number of this library's own kernel
 * launches so far in the process (CUB / memset launches are not counted).  gsb_profile_enable(1)
 * makes the hot kernels record CUDA events on their launch stream; gsb_profile_read sums the
 * recorded durations of one kernel ("projection_ut", "sh_fwd", "sh_bwd", "isect_count",
 * "isect_emit", "isect_sort", "isect_offsets", "raster_prep", "raster_fwd", "raster_bwd",
 * "raster_finalize") and returns how many launches were found. */
GSB_API uint64_t gsb_launch_count(void);
GSB_API void gsb_profile_enable(int on);
GSB_API int gsb_profile_read(const char *kernel, double *total_ms);

/* ---- a1: gsplat::projection_ut_3dgs_fused (Ops.h:69-98, Projection.cpp:16-110,
 *      ProjectionUT3DGSFused.cu:17-203) ------------------------------------------
 * radii [C,N,2] int32 (0,0 = culled), means2d [C,N,2], depths [C,N], conics [C,N,3],
 * compensations [C,N] or NULL.  Culled rows of means2d/depths/conics are left
 * untouched, like the reference (Projection.cpp:70-73).  Inputs need only their natural 4-byte
 * alignment (e.g. a contiguous view means[1:]); 16-byte aligned base pointers get the TMA staging path. */
GSB_API int gsb_projection_ut(
    uint32_t C, uint32_t N,
    const float *means, const float *quats, const float *scales, const float *opacities /*nullable*/,
    const GsbCamera *cam, uint32_t image_width, uint32_t image_height,
    float eps2d, float near_plane, float far_plane, float radius_clip,
    int32_t *radii, float *means2d, float *depths, float *conics, float *compensations /*nullable*/,
    gsb_stream_t stream);

/* ---- a3/a4: gsplat::spherical_harmonics_fwd / _bwd (Ops.h:12-25,
 *      SphericalHarmonics.cpp:15-75, SphericalHarmonicsCUDA.cu:374-481) -----------
 * M = number of (camera, Gaussian) elements, K = coefficients per element.
 * masks: [M] bytes (bool) or NULL.  fwd leaves masked-out rows of `colors`
 * untouched.  bwd writes EVERY element of v_coeffs (zeros for masked rows and for
 * k >= (degree+1)^2) and of v_dirs -- the caller need not zero-fill them
 * (the reference memsets 192 B/Gaussian first, SphericalHarmonics.cpp:58). */
GSB_API int gsb_sh_fwd(uint32_t M, uint32_t K, uint32_t degrees_to_use, const float *dirs,
                       const float *coeffs, const uint8_t *masks, float *colors, gsb_stream_t stream);
GSB_API int gsb_sh_bwd(uint32_t M, uint32_t K, uint32_t degrees_to_use, const float *dirs,
                       const float *coeffs, const uint8_t *masks, const float *v_colors,
                       float *v_coeffs, float *v_dirs /*nullable*/, gsb_stream_t stream);

/* ---- (e) multi-GPU exchange step: SH backward of V views at once ------------------------------
 * The reference has no distributed mode; this entry point belongs to the view-sharded
 * data-parallel step of SURVEY.md 8(e).  Per view the SH gradient row is the outer product of the basis
 * at that view's direction and a 12-byte colour gradient, so ranks all-gather the colour gradients
 * (v_colors [V, M, 3], zero where a Gaussian was not blended) and each expands and sums all views:
 *   v_coeffs [M, K, 3]  = sum_v Y(normalize(means - campos[v])) (x) v_colors[v]      (fully written)
 *   v_means  [M, 3]    += sum_v d(colour_v)/d(mean)                                   (accumulated)
 * Equal to summing gsb_sh_bwd over the views with dirs = means - campos[v]. */
GSB_API int gsb_sh_bwd_views(uint32_t M, uint32_t K, uint32_t degrees_to_use, uint32_t V,
                             const float *means, const float *campos /* [V,3] */, const float *coeffs,
                             const float *v_colors /* [V,M,3] */, float *v_coeffs, float *v_means,
                             gsb_stream_t stream);
/* The same step with the all-gather INSIDE the kernel: view v's colour gradients [M,3] and camera position [3] are
 * read through per-view device pointers (host arrays of V <= 16 pointers), which may point into PEER memory --
 * NVLink-mapped buffers of the other ranks (CUDA IPC / symmetric memory).  The remote reads overlap the expansion
 * and the gathered [V,M,3] tensor is never materialised.  The caller orders the kernel after every rank's write of
 * its buffer (a device-side barrier) and must not overwrite its buffer before all ranks have finished reading. */
GSB_API int gsb_sh_bwd_views_peer(uint32_t M, uint32_t K, uint32_t degrees_to_use, uint32_t V, const float *means,
                                  const float *const *campos_views /* host array [V] of device ptrs to [3] */,
                                  const float *coeffs,
                                  const float *const *v_colors_views /* host array [V] of device ptrs to [M,3] */,
                                  float *v_coeffs, float *v_means, gsb_stream_t stream);

/* ---- a5: gsplat::intersect_tile (Ops.h:28-38, Intersect.cpp:15-122,
 *      IntersectTile.cu:24-114,290-328) ----------------------------------------------
 * Three device steps around the one host read-back the API forces (the op
 * returns exactly-sized tensors):
 *   gsb_isect_count      tiles_per_gauss [C*N] int32 and its inclusive int64 scan
 *                        cum_tiles [C*N]; n_isects = cum_tiles[C*N-1].
 *   gsb_isect_emit       unsorted isect_ids [I] int64 / flatten_ids [I] int32.
 *   gsb_isect_sort       stable LSD radix sort (own kernels) by the low 32+tile_bits+cam_bits
 *                        key bits; equal keys keep emission order, like the reference's CUB sort.
 * Non-packed ([C,N,...]) layout only; packed mode is rejected by the reference's
 * caller too (rasterizer.cpp:56). */
GSB_API size_t gsb_isect_count_workspace(uint64_t n_elements);
GSB_API int gsb_isect_count(uint32_t C, uint32_t N, const float *means2d, const int32_t *radii,
                            uint32_t tile_size, uint32_t tile_width, uint32_t tile_height,
                            int32_t *tiles_per_gauss, int64_t *cum_tiles,
                            void *workspace, size_t workspace_bytes, gsb_stream_t stream);
GSB_API int gsb_isect_emit(uint32_t C, uint32_t N, const float *means2d, const int32_t *radii,
                           const float *depths, const int64_t *cum_tiles,
                           uint32_t tile_size, uint32_t tile_width, uint32_t tile_height,
                           int64_t *isect_ids, int32_t *flatten_ids, gsb_stream_t stream);
GSB_API size_t gsb_isect_sort_workspace(uint64_t n_isects);
GSB_API int gsb_isect_sort(uint64_t n_isects, uint32_t C, uint32_t tile_width, uint32_t tile_height,
                           const int64_t *isect_ids_in, const int32_t *flatten_ids_in,
                           int64_t *isect_ids_out, int32_t *flatten_ids_out,
                           void *workspace, size_t workspace_bytes, gsb_stream_t stream);

/* Planned sorted path for sort == true: same outputs as gsb_isect_emit + gsb_isect_sort, bit for bit,
 * without sorting any intersection.  The Gaussians are ordered by depth once (N elements, own radix
 * sort), the runs in that order are cut into chunks, a per-(chunk, tile) histogram is scanned, every
 * intersection is written DIRECTLY into the final range of its (chunk, tile) group and the groups
 * (1-2 entries on average) are put into depth order in place (csrc/gsb_intersect.cu).  It is arranged around the one host read-back the
 * operator API forces: everything that does not need the host to know n_isects runs in the PLAN.
 *   gsb_isect_plan          tiles_per_gauss [C*N]; *n_isects_out (device OR pinned host int64,
 *                           written asynchronously on `stream`); tile_offsets_out [C*th*tw] int32 or
 *                           NULL -- the a6 result, a by-product of the plan's scan (with
 *                           tile_offsets_total != 0 one more entry, [C*th*tw] = n_isects, so a
 *                           consumer needs no host-side count); plan_workspace keeps the run table
 *                           and the per-chunk first slots for the emit.
 *   gsb_isect_emit_planned  flatten_ids and, unless NULL, isect_ids, sorted.  `n_isects` is the CAPACITY
 *                           of those arrays: the true count is read on the device, slots beyond the
 *                           capacity are dropped (a caller that allocates from a guess checks the
 *                           count afterwards) -- with the exact count it is the reference's behaviour.
 * Limits: C*N < 2^31, C*th*tw < 2^31, tile grid sides <= 65535, n_isects < 2^31. */
GSB_API size_t gsb_isect_plan_workspace(uint32_t C, uint32_t N, uint32_t tile_width, uint32_t tile_height);
GSB_API int gsb_isect_plan(uint32_t C, uint32_t N, const float *means2d, const int32_t *radii,
                           const float *depths, uint32_t tile_size, uint32_t tile_width,
                           uint32_t tile_height, int32_t *tiles_per_gauss, int64_t *n_isects_out,
                           int32_t *tile_offsets_out /*nullable*/, int tile_offsets_total,
                           void *plan_workspace, size_t plan_workspace_bytes, gsb_stream_t stream);
GSB_API int gsb_isect_emit_planned(uint32_t C, uint32_t N, const float *depths,
                                   uint32_t tile_width, uint32_t tile_height,
                                   uint64_t n_isects, const void *plan_workspace,
                                   size_t plan_workspace_bytes, int64_t *isect_ids /*nullable*/,
                                   int32_t *flatten_ids, gsb_stream_t stream);

/* ---- a6: gsplat::intersect_offset (Ops.h:39-43, IntersectTile.cu:206-288) --------
 * offsets [C,tile_height,tile_width] int32; all zero when n_isects == 0. */
GSB_API int gsb_isect_offsets(uint64_t n_isects, const int64_t *isect_ids_sorted, uint32_t C,
                              uint32_t tile_width, uint32_t tile_height, int32_t *offsets,
                              gsb_stream_t stream);

/* ---- a7: gsplat::rasterize_to_pixels_from_world_3dgs_fwd (Ops.h:100-129,
 *      Rasterization.cpp:20-132, RasterizeToPixelsFromWorld3DGSFwd.cu:20-279) ---------
 * colors [C,N,3] (3 channels only, as the reference asserts), opacities [C,N],
 * backgrounds [C,3] or NULL, masks [C,th,tw] bytes or NULL, tile_offsets [C,th,tw],
 * flatten_ids [I].  C must be 1 (the reference's kernels index means[] with the
 * flattened id, Fwd.cu:197-200).  Outputs: renders [C,H,W,3], alphas [C,H,W,1],
 * last_ids [C,H,W] int32. */
GSB_API size_t gsb_raster_fwd_workspace(uint32_t N);
GSB_API int gsb_raster_fwd(
    uint32_t C, uint32_t N, uint64_t n_isects,
    const float *means, const float *quats, const float *scales, const float *colors,
    const float *opacities, const float *backgrounds, const uint8_t *masks,
    uint32_t image_width, uint32_t image_height, uint32_t tile_size, const GsbCamera *cam,
    const int32_t *tile_offsets, const int32_t *flatten_ids,
    float *renders, float *alphas, int32_t *last_ids,
    void *workspace, size_t workspace_bytes, gsb_stream_t stream);

/* ---- a8: gsplat::rasterize_to_pixels_from_world_3dgs_bwd (Ops.h:131-165,
 *      Rasterization.cpp:134-261, RasterizeToPixelsFromWorld3DGSBwd.cu:17-373) --------
 * Writes every element of v_means [N,3], v_quats [N,4], v_scales [N,3],
 * v_colors [C,N,3], v_opacities [C,N] (zeros for untouched Gaussians): the caller
 * need not zero-fill (the reference does, Rasterization.cpp:190-194). */
GSB_API size_t gsb_raster_bwd_workspace(uint32_t N);
GSB_API int gsb_raster_bwd(
    uint32_t C, uint32_t N, uint64_t n_isects,
    const float *means, const float *quats, const float *scales, const float *colors,
    const float *opacities, const float *backgrounds, const uint8_t *masks,
    uint32_t image_width, uint32_t image_height, uint32_t tile_size, const GsbCamera *cam,
    const int32_t *tile_offsets, const int32_t *flatten_ids,
    const float *render_alphas, const int32_t *last_ids,
    const float *v_render_colors, const float *v_render_alphas,
    float *v_means, float *v_quats, float *v_scales, float *v_colors, float *v_opacities,
    void *workspace, size_t workspace_bytes, gsb_stream_t stream);

/* ---- the L3 glue folded into the path -- an EXTENDED operator set over the raw SplatData --------
 * The reference's caller activates the parameters and builds SH inputs with ~10 torch kernels around the
 * operators every step (src/core/splat_data.cpp:267-286: exp / sigmoid / normalize / cat(sh0, shN);
 * src/training/rasterization/rasterizer.cpp:250-266: inverse(viewmat), dirs, masks, clamp_min(SH + 0.5)) and
 * their autograd backward.  These entry points take the RAW tensors instead (fastgs already does,
 * fastgs/rasterization/include/rasterization_api.h:25-75) and keep the unfused operators above untouched:
 *
 *   gsb_fused_front   activations -> UT projection (bit-identical radii / means2d / depths to a1) -> SH colour of
 *                     the view direction (+0.5, clamp) -> blend record; zeroes the backward's moment rows.
 *   gsb_isect_plan / gsb_isect_emit_planned          (above; tile_offsets_total = 1, isect_ids = NULL)
 *   gsb_raster_fwd_recs / gsb_raster_bwd_recs        a7 / a8 on the records of gsb_fused_front
 *   gsb_fused_back    moments -> gradients of the RAW parameters: finalize chain rule, exp / sigmoid /
 *                     normalise backward, clamp mask, SH backward straight into the sh0 / shN layouts,
 *                     direction gradient added to the position gradient.  Every output element is written.
 *
 * No step of this sequence needs a host-side value: flatten_ids may be allocated from a capacity guess
 * (intersections beyond it are dropped; compare *n_isects_out with the capacity afterwards), so a whole
 * training step can be enqueued -- or captured in a CUDA graph -- without a synchronisation. */
typedef struct GsbSplatRaw {
    uint32_t N;
    uint32_t sh_coeffs;       /* K: coefficients per Gaussian, sh0 included (shN holds K - 1) */
    uint32_t sh_degree;       /* active degree, (degree + 1)^2 <= K */
    float scaling_modifier;
    const float *means;        /* [N,3] */
    const float *sh0;          /* [N,1,3] */
    const float *shN;          /* [N,K-1,3], may be NULL when K == 1 */
    const float *scaling_raw;  /* [N,3] log-scales */
    const float *rotation_raw; /* [N,4] (w,x,y,z), not normalised; 16-byte aligned */
    const float *opacity_raw;  /* [N] logits */
} GsbSplatRaw;

GSB_API size_t gsb_fused_workspace(uint32_t N); /* [records][moments]; == gsb_raster_bwd_workspace(N) */
GSB_API int gsb_fused_front(const GsbSplatRaw *splats, const GsbCamera *cam, uint32_t image_width,
                            uint32_t image_height, float eps2d, float near_plane, float far_plane,
                            float radius_clip, int zero_moments,
                            int32_t *radii /*[N,2]*/, float *means2d /*[N,2]*/, float *depths /*[N]*/,
                            void *workspace, size_t workspace_bytes, gsb_stream_t stream);
GSB_API int gsb_raster_fwd_recs(uint32_t N, uint32_t capacity, const void *workspace, size_t workspace_bytes,
                                const float *backgrounds, const uint8_t *masks, uint32_t image_width,
                                uint32_t image_height, const GsbCamera *cam,
                                const int32_t *tile_offsets /*[th*tw + 1]*/, const int32_t *flatten_ids,
                                float *renders, float *alphas, int32_t *last_ids, gsb_stream_t stream);
GSB_API int gsb_raster_bwd_recs(uint32_t N, uint32_t capacity, void *workspace, size_t workspace_bytes,
                                const float *backgrounds, const uint8_t *masks, uint32_t image_width,
                                uint32_t image_height, const GsbCamera *cam,
                                const int32_t *tile_offsets /*[th*tw + 1]*/, const int32_t *flatten_ids,
                                const float *render_alphas, const int32_t *last_ids,
                                const float *v_render_colors, const float *v_render_alphas, gsb_stream_t stream);
GSB_API int gsb_fused_back(const GsbSplatRaw *splats, const GsbCamera *cam, uint32_t image_width,
                           uint32_t image_height, const int32_t *radii, const void *workspace,
                           size_t workspace_bytes, float *v_means, float *v_sh0, float *v_shN /*NULL iff K == 1*/,
                           float *v_scaling, float *v_rotation /*16-byte aligned*/, float *v_opacity,
                           gsb_stream_t stream);

/* ---- photometric loss of the training step and its gradient, one kernel -------------------------
 * loss = (1 - lambda) * mean|clamp(render,0,1) - target| + lambda * (1 - mean SSIM_valid)
 * (src/training/trainer.cpp:103-126; fused_ssim(..., "valid"): 11x11 Gaussian window, zero padding, map cropped by 5
 * per side: src/training/kernels/ssim.cu:64-420, include/kernels/fused_ssim.cuh:27-117).
 * renders [H,W,3] is the blend's output (unclamped); target is [3,H,W] (the reference's layout) or [H,W,3].
 * target_chw is a bit set: GSB_LOSS_TARGET_CHW (1) = target is [3,H,W]; GSB_LOSS_RENDERS_CHW (2) = renders AND v_renders
 * are [3,H,W] planes (the fastgs path's image).
 * v_renders (laid out like renders) = grad_scale * dLoss/d(renders) including the clamp mask, or NULL to evaluate only.
 * loss_out: DEVICE float[3] = (loss, l1 mean, ssim mean).  workspace: gsb_ssim_l1_workspace() bytes, 256-aligned. */
enum { GSB_LOSS_TARGET_CHW = 1, GSB_LOSS_RENDERS_CHW = 2 };
GSB_API size_t gsb_ssim_l1_workspace(void);
GSB_API int gsb_ssim_l1(uint32_t image_width, uint32_t image_height, const float *renders, const float *target,
                        int target_chw, float lambda_dssim, float grad_scale, float *v_renders /*nullable*/,
                        float *loss_out, void *workspace, size_t workspace_bytes, gsb_stream_t stream);

/* ---- Adam over all parameter groups in one launch ---------------------------------------------------
 * The update of fastgs/optimizer/include/adam_kernels.cuh:13-36 (param -= lr * bc1_rcp * m / (sqrt(v) * bc2_sqrt_rcp
 * + eps)); the caller keeps the step counts and learning-rate schedule (src/training/optimizers/fused_adam.cpp:22-95,
 * strategies/strategy_utils.cpp:27-55).  Up to 8 groups; groups with n == 0 are skipped. */
typedef struct GsbAdamGroup {
    float *param;
    const float *grad;
    float *exp_avg;
    float *exp_avg_sq;
    uint64_t n; /* elements */
    float lr, beta1, beta2, eps;
    float bias_correction1_rcp;      /* 1 / (1 - beta1^t) */
    float bias_correction2_sqrt_rcp; /* 1 / sqrt(1 - beta2^t) */
} GsbAdamGroup;
GSB_API int gsb_adam_step(const GsbAdamGroup *groups, uint32_t n_groups, gsb_stream_t stream);
/* Same, with the step-dependent scalars read from DEVICE memory at run time: dynamic_scalars [n_groups][4] floats =
 * (lr, bias_correction1_rcp, bias_correction2_sqrt_rcp, enabled != 0), 16-byte aligned; the lr / bias fields of
 * `groups` are ignored.  Lets a whole training iteration be captured once in a CUDA graph and replayed. */
GSB_API int gsb_adam_step_dynamic(const GsbAdamGroup *groups, uint32_t n_groups, const float *dynamic_scalars,
                                  gsb_stream_t stream);

/* ---- the reference's default rasterizer ("fastgs", EWA splatting of 2-D conics) ----------------------
 * fast_gs::rasterization::forward / backward (fastgs/rasterization/src/forward.cu:15-199, src/backward.cu:14-116;
 * API fastgs/rasterization/include/rasterization_api.h:25-75 -> include/fastgs/rasterization_api.h + shim/FastGs.cpp).
 * RAW parameters in (log-scales, un-normalised (w,x,y,z) quaternions, logit opacities, sh0 [N,1,3], shN [N,rest,3]),
 * image [3,H,W] and alpha [1,H,W] out (no background: the caller composites it, fast_rasterizer.cpp:71).
 * The three opaque buffers play the role of the reference's per_primitive / per_tile / per_instance blobs: the caller
 * allocates them (sizes below; the instance buffer holds n_instances int32), keeps them from forward to backward and
 * never looks inside.  Forward is two calls because the instance count sizes the third buffer:
 *   gsb_fastgs_forward_plan   per-primitive set-up, depth order, tile histogram; *n_instances_out (DEVICE or PINNED HOST
 *                             int64, written asynchronously on `stream`) = number of (primitive, tile) instances
 *   gsb_fastgs_forward_blend  instances placed in (tile, depth) order, blended; `capacity` = entries of `instances`
 *                             (>= n_instances for an exact result; smaller capacities drop the farthest-sorted slots)
 *   gsb_fastgs_backward       blend gradient + the set-up's chain rule; writes EVERY element of the six gradients
 *                             (the reference zero-fills them first), grad_w2c [4,4] if non-NULL, and adds to
 *                             densification_info [2,N] if non-NULL (kernels_backward.cuh:252-255). */
typedef struct GsbFastgsView {
    const float *w2c;          /* [4,4] world->camera, row-major, DEVICE */
    const float *cam_position; /* [3] DEVICE */
    uint32_t width, height;
    float focal_x, focal_y, center_x, center_y, near_plane, far_plane;
    uint32_t active_sh_bases;     /* 1, 4, 9 or 16 */
    uint32_t total_bases_sh_rest; /* rows of shN per primitive */
} GsbFastgsView;
GSB_API size_t gsb_fastgs_primitive_bytes(uint32_t N, uint32_t width, uint32_t height);
GSB_API size_t gsb_fastgs_tile_bytes(uint32_t width, uint32_t height);
GSB_API int gsb_fastgs_forward_plan(uint32_t N, const float *means, const float *scales_raw,
                                    const float *rotations_raw /*16-byte aligned*/, const float *opacities_raw,
                                    const float *sh0, const float *shN, const GsbFastgsView *view, void *per_primitive,
                                    size_t per_primitive_bytes, void *per_tile, size_t per_tile_bytes,
                                    int64_t *n_instances_out, gsb_stream_t stream);
GSB_API int gsb_fastgs_forward_blend(uint32_t N, const GsbFastgsView *view, void *per_primitive,
                                     size_t per_primitive_bytes, void *per_tile, size_t per_tile_bytes, int32_t *instances,
                                     uint64_t capacity, float *image, float *alpha, gsb_stream_t stream);
GSB_API int gsb_fastgs_backward(uint32_t N, const float *means, const float *scales_raw, const float *rotations_raw,
                                const float *shN, const GsbFastgsView *view,
                                void *per_primitive, size_t per_primitive_bytes, const void *per_tile,
                                size_t per_tile_bytes, const int32_t *instances, uint64_t capacity, const float *alpha,
                                const float *grad_image, const float *grad_alpha, float *grad_means,
                                float *grad_scales_raw, float *grad_rotations_raw /*16-byte aligned*/,
                                float *grad_opacities_raw, float *grad_sh0, float *grad_shN, float *grad_w2c /*nullable*/,
                                float *densification_info /*nullable*/, gsb_stream_t stream);

/* ---- link-surface ops used by the densification strategies -------------------------
 * gsplat::quats_to_rotmats (Ops.h:46-48, QuatToRotmatCUDA.cu:14-39): [N,4] -> [N,3,3] */
GSB_API int gsb_quat_to_rotmat(uint32_t N, const float *quats, float *rotmats, gsb_stream_t stream);
/* gsplat::relocation (Ops.h:52-57, RelocationCUDA.cu:12-43) */
GSB_API int gsb_relocation(uint32_t N, const float *opacities, const float *scales, const int32_t *ratios,
                           const float *binoms, int32_t n_max, float *new_opacities, float *new_scales,
                           gsb_stream_t stream);
/* gsplat::add_noise (Ops.h:59-65, RelocationCUDA.cu:113-144): means updated in place */
GSB_API int gsb_add_noise(uint32_t N, const float *raw_opacities, const float *raw_scales,
                          const float *raw_quats, const float *noise, float *means, float current_lr,
                          gsb_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* GSB200_H_ */
