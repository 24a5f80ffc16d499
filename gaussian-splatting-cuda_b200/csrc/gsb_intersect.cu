// gsb_intersect.cu -- a5/a6: Gaussian -> tile expansion, (camera, tile, depth) sort, tile offsets.
//
// Implements gsplat::intersect_tile / intersect_offset (reference: gsplat/IntersectTile.cu:24-114
// count+emit, :206-252 offsets, :290-328 radix sort; gsplat/Intersect.cpp:15-137 host side).
// All results are integer functions of (means2d, radii, depths) and are bit-exact with the
// reference: same float tile-bbox arithmetic (division by the tile size, floor/ceil, CUDA's
// saturating float->uint32 conversion), same 64-bit key layout
//      cam_id << (32 + tile_bits) | tile_id << 32 | float_bits(depth)
// and the same final order: ascending (camera, tile, depth bits), ties in emission order
// (ascending flattened Gaussian index), which is what the reference's stable
// cub::DeviceRadixSort::SortPairs over 32+tile_bits+cam_bits key bits produces.
//
// How the sorted lists are built (gsb_isect_plan + gsb_isect_emit_planned): the reference radix-sorts all I
// intersections on 46 bits (six 8-bit passes over 12-byte pairs, ~150 B of HBM traffic per
// intersection).  Here the depth order is established ONCE per Gaussian instead of once per
// intersection:
//   1. stable radix sort of the C*N Gaussians by (camera, depth bits)      [N elements, not I]
//   2. scan of tiles_per_gauss in that order, emit the intersections in that order
//   3. stable radix sort of the I intersections on the (camera, tile) bits ONLY (13-14 bits:
//      two passes instead of six).
// A stable sort by tile of a sequence already ordered by (depth, index) is ordered by
// (tile, depth, index): the same permutation as the reference's single 46-bit sort.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <thrust/iterator/permutation_iterator.h>
#include <thrust/iterator/transform_iterator.h>

#include "gsb_common.cuh"

namespace gsb {

constexpr int kIsectThreads = 256;

struct TileBox {
    uint32_t x0, y0, x1, y1;
    bool active;
};

// IntersectTile.cu:54-76.  This file is compiled with -fmad=false and uses IEEE division so the
// tile bounds are a pure function of the inputs (the reference's fast-math `x / 16.f` is exact
// for power-of-two tile sizes, which is what its callers use: rasterizer.cpp:180).
__device__ __forceinline__ TileBox tile_box(const float *__restrict__ means2d, const int32_t *__restrict__ radii,
                                            size_t idx, uint32_t tile_size, uint32_t tile_width,
                                            uint32_t tile_height) {
    TileBox b;
    const int2 r = reinterpret_cast<const int2 *>(radii)[idx];
    const float radius_x = (float)r.x, radius_y = (float)r.y;
    b.active = !(radius_x <= 0 || radius_y <= 0);
    if (!b.active) { b.x0 = b.x1 = b.y0 = b.y1 = 0; return b; }
    const float2 m = reinterpret_cast<const float2 *>(means2d)[idx];
    const float ts = (float)tile_size;
    const float trx = __fdiv_rn(radius_x, ts), try_ = __fdiv_rn(radius_y, ts);
    const float tx = __fdiv_rn(m.x, ts), ty = __fdiv_rn(m.y, ts);
    // (uint32_t) of a negative float saturates to 0 on the GPU (cvt.rzi.u32.f32)
    b.x0 = min((uint32_t)floorf(tx - trx), tile_width);
    b.y0 = min((uint32_t)floorf(ty - try_), tile_height);
    b.x1 = min((uint32_t)ceilf(tx + trx), tile_width);
    b.y1 = min((uint32_t)ceilf(ty + try_), tile_height);
    return b;
}

__global__ void __launch_bounds__(kIsectThreads) isect_count_kernel(uint64_t n, const float *__restrict__ means2d,
                                                                     const int32_t *__restrict__ radii,
                                                                     uint32_t tile_size, uint32_t tile_width,
                                                                     uint32_t tile_height,
                                                                     int32_t *__restrict__ tiles_per_gauss) {
    const uint64_t idx = (uint64_t)blockIdx.x * kIsectThreads + threadIdx.x;
    if (idx >= n) return;
    const TileBox b = tile_box(means2d, radii, idx, tile_size, tile_width, tile_height);
    tiles_per_gauss[idx] = b.active ? (int32_t)((b.y1 - b.y0) * (b.x1 - b.x0)) : 0;
}

// Emits the intersections of Gaussian perm[i] (or i when perm == nullptr) at cum[i-1].
__global__ void __launch_bounds__(kIsectThreads) isect_emit_kernel(uint64_t n, uint32_t N,
                                                                    const float *__restrict__ means2d,
                                                                    const int32_t *__restrict__ radii,
                                                                    const float *__restrict__ depths,
                                                                    const uint32_t *__restrict__ perm,
                                                                    const int64_t *__restrict__ cum_tiles,
                                                                    uint32_t tile_size, uint32_t tile_width,
                                                                    uint32_t tile_height, uint32_t tile_n_bits,
                                                                    int64_t *__restrict__ isect_ids,
                                                                    int32_t *__restrict__ flatten_ids) {
    const uint64_t i = (uint64_t)blockIdx.x * kIsectThreads + threadIdx.x;
    if (i >= n) return;
    const uint64_t idx = perm ? (uint64_t)perm[i] : i;
    const TileBox b = tile_box(means2d, radii, idx, tile_size, tile_width, tile_height);
    if (!b.active) return;
    const int64_t cid = (int64_t)(idx / N);
    const int64_t cid_enc = cid << (32 + tile_n_bits);
    const int64_t depth_enc = (int64_t)__float_as_uint(depths[idx]); // zero-extended bit pattern (:98-99)
    int64_t cur = (i == 0) ? 0 : cum_tiles[i - 1];
    for (uint32_t y = b.y0; y < b.y1; ++y)
        for (uint32_t x = b.x0; x < b.x1; ++x) {
            const int64_t tile_id = (int64_t)y * tile_width + x;
            isect_ids[cur] = cid_enc | (tile_id << 32) | depth_enc;
            flatten_ids[cur] = (int32_t)idx;
            ++cur;
        }
}

// Load-balanced emit for the sorted path: one CTA per kEmitSpan consecutive OUTPUT slots instead of one
// thread per Gaussian, so a Gaussian covering 400 tiles costs 400 slots of work spread over 50 threads
// rather than one thread's 400-iteration loop, and the span leaves through coalesced warp stores.
// Run i (Gaussian perm[i]) owns the slots [cum[i-1], cum[i]); Gaussians without tiles are parked behind
// all others by their depth key, so every run that intersects [0, I) is non-empty.
constexpr int kEmitItems = 8;
constexpr int kEmitSpan = kIsectThreads * kEmitItems;

__global__ void __launch_bounds__(kIsectThreads) isect_emit_balanced_kernel(
    uint64_t n, uint32_t N, uint64_t n_isects, const float *__restrict__ means2d, const int32_t *__restrict__ radii,
    const float *__restrict__ depths, const uint32_t *__restrict__ perm, const int64_t *__restrict__ cum,
    uint32_t tile_size, uint32_t tile_width, uint32_t tile_height, uint32_t tile_n_bits,
    int64_t *__restrict__ isect_ids, int32_t *__restrict__ flatten_ids) {
    __shared__ int32_t s_end[kEmitSpan + 1]; // run ends relative to the CTA's first slot, clamped to the span
    __shared__ int64_t s_keys[kEmitSpan];
    __shared__ int32_t s_vals[kEmitSpan];
    __shared__ uint64_t s_first;             // index of the run that owns the CTA's first slot
    __shared__ int64_t s_first_start;        // and its first slot
    const int64_t o0 = (int64_t)blockIdx.x * kEmitSpan;
    const int32_t span = (int32_t)min((int64_t)kEmitSpan, (int64_t)n_isects - o0);
    const uint32_t tid = threadIdx.x;

    if (tid < 32) {
        // first i with cum[i] > o0: 32-ary search, one probe per lane per round
        uint64_t lo = 0, hi = n; // answer in [lo, hi), cum[n-1] = n_isects > o0
        while (hi - lo > 1) {
            const uint64_t len = hi - lo;
            const uint64_t step = (len + 32) / 33; // 32 interior probes split the range into <= 33 pieces
            const uint64_t p = lo + (uint64_t)(tid + 1) * step - 1; // probe positions lo+step-1, lo+2step-1, ...
            const bool le = (p < hi - 1) ? (cum[p] <= o0) : false;   // position hi-1 is known to be > o0
            const uint32_t m = __ballot_sync(0xffffffffu, le);
            const uint32_t k = __popc(m); // probes are monotone: the first k satisfy cum <= o0
            const uint64_t nlo = (k == 0) ? lo : lo + (uint64_t)k * step;
            const uint64_t nhi = (k == 32) ? hi : min(hi, lo + (uint64_t)(k + 1) * step);
            lo = nlo; hi = nhi;
        }
        if (tid == 0) {
            s_first = lo;
            s_first_start = (lo == 0) ? 0 : cum[lo - 1];
        }
    }
    __syncthreads();
    const uint64_t g_lo = s_first;
    for (int32_t r = (int32_t)tid; r <= kEmitSpan; r += kIsectThreads) {
        const uint64_t g = g_lo + (uint64_t)r;
        int64_t e = (g < n) ? cum[g] - o0 : (int64_t)kEmitSpan;
        s_end[r] = (int32_t)min(e, (int64_t)kEmitSpan);
    }
    __syncthreads();

    // Each thread walks its kEmitItems consecutive slots (one run lookup, then increments) into shared
    // memory; the CTA then writes the span out with fully coalesced 256-byte warp stores -- per-thread
    // 8-byte global stores would touch one 32-byte sector each and run at the L2's transaction rate.
    auto phys = [](int32_t e) { return (e & ~7) | ((e & 7) ^ ((e >> 3) & 7)); }; // bank swizzle
    int32_t o = (int32_t)tid * kEmitItems;
    if (o < span) {
        // run of this thread's first slot: first r with s_end[r] > o
        int32_t lo = 0, hi = kEmitSpan;
        while (lo < hi) {
            const int32_t mid = (lo + hi) >> 1;
            if (s_end[mid] > o) hi = mid; else lo = mid + 1;
        }
        int32_t r = lo;
        const int32_t o_stop = min(o + kEmitItems, span);
        while (o < o_stop) {
            const uint64_t idx = perm[g_lo + (uint64_t)r];
            const TileBox b = tile_box(means2d, radii, idx, tile_size, tile_width, tile_height);
            const uint32_t w = b.x1 - b.x0;
            const int64_t run_start = (r == 0) ? s_first_start - o0 : (int64_t)s_end[r - 1];
            const uint32_t j = (uint32_t)((int64_t)o - run_start);
            uint32_t y = b.y0 + j / w, x = b.x0 + j % w;
            const int64_t cid_enc = (int64_t)(idx / N) << (32 + tile_n_bits);
            const int64_t depth_enc = (int64_t)__float_as_uint(depths[idx]);
            const int32_t run_stop = min(s_end[r], o_stop);
            for (; o < run_stop; ++o) {
                const int64_t tile_id = (int64_t)y * tile_width + x;
                s_keys[phys(o)] = cid_enc | (tile_id << 32) | depth_enc;
                s_vals[phys(o)] = (int32_t)idx;
                if (++x == b.x1) { x = b.x0; ++y; }
            }
            ++r;
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kEmitItems; ++k) {
        const int32_t e = k * kIsectThreads + (int32_t)tid;
        if (e < span) {
            isect_ids[o0 + e] = s_keys[phys(e)];
            flatten_ids[o0 + e] = s_vals[phys(e)];
        }
    }
}

// Count and depth key in one pass over the Gaussians (the plan phase of the two-phase sorted path).
__global__ void __launch_bounds__(kIsectThreads) isect_plan_kernel(uint64_t n, uint32_t N,
                                                                    const float *__restrict__ means2d,
                                                                    const int32_t *__restrict__ radii,
                                                                    const float *__restrict__ depths,
                                                                    uint32_t tile_size, uint32_t tile_width,
                                                                    uint32_t tile_height,
                                                                    int32_t *__restrict__ tiles_per_gauss,
                                                                    uint64_t *__restrict__ keys64,
                                                                    uint32_t *__restrict__ keys32,
                                                                    uint32_t *__restrict__ vals) {
    const uint64_t idx = (uint64_t)blockIdx.x * kIsectThreads + threadIdx.x;
    if (idx >= n) return;
    const TileBox b = tile_box(means2d, radii, idx, tile_size, tile_width, tile_height);
    const int32_t cnt = b.active ? (int32_t)((b.y1 - b.y0) * (b.x1 - b.x0)) : 0;
    tiles_per_gauss[idx] = cnt;
    // Gaussians without tiles never emit: park them behind ALL others (camera field = C, one past the last)
    const uint32_t d = cnt > 0 ? __float_as_uint(depths[idx]) : 0xffffffffu;
    const uint64_t cam = cnt > 0 ? idx / N : n / N;
    if (keys64) keys64[idx] = (cam << 32) | d;
    else keys32[idx] = d;
    vals[idx] = (uint32_t)idx;
}

// IntersectTile.cu:206-252, restated as "first sorted position whose (cam, tile) >= id".
__global__ void __launch_bounds__(kIsectThreads) isect_offsets_kernel(uint64_t n_isects,
                                                                       const int64_t *__restrict__ isect_ids,
                                                                       uint32_t total_tiles, uint32_t n_tiles,
                                                                       uint32_t tile_n_bits,
                                                                       int32_t *__restrict__ offsets) {
    const uint64_t idx = (uint64_t)blockIdx.x * kIsectThreads + threadIdx.x;
    if (idx >= n_isects) return;
    const int64_t cur = isect_ids[idx] >> 32;
    const int64_t id_curr = (cur >> tile_n_bits) * n_tiles + (cur & ((1ll << tile_n_bits) - 1));
    if (idx == 0)
        for (int64_t i = 0; i < id_curr + 1 && i < (int64_t)total_tiles; ++i) offsets[i] = 0;
    if (idx == n_isects - 1)
        for (int64_t i = id_curr + 1; i < (int64_t)total_tiles; ++i) offsets[i] = (int32_t)n_isects;
    if (idx > 0) {
        const int64_t prev = isect_ids[idx - 1] >> 32;
        if (prev == cur) return;
        const int64_t id_prev = (prev >> tile_n_bits) * n_tiles + (prev & ((1ll << tile_n_bits) - 1));
        for (int64_t i = id_prev + 1; i < id_curr + 1 && i < (int64_t)total_tiles; ++i) offsets[i] = (int32_t)idx;
    }
}

struct CastI64 {
    __host__ __device__ __forceinline__ int64_t operator()(const int32_t &v) const { return (int64_t)v; }
};
using CountIter = thrust::transform_iterator<CastI64, const int32_t *, int64_t>;
using PermCountIter =
    thrust::transform_iterator<CastI64, thrust::permutation_iterator<const int32_t *, const uint32_t *>, int64_t>;

static inline size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

} // namespace gsb

extern "C" size_t gsb_isect_count_workspace(uint64_t n_elements) {
    size_t bytes = 0;
    gsb::CountIter it(nullptr, gsb::CastI64());
    cub::DeviceScan::InclusiveSum(nullptr, bytes, it, (int64_t *)nullptr, (int64_t)n_elements);
    return gsb::align256(bytes) + 256;
}

extern "C" int gsb_isect_count(uint32_t C, uint32_t N, const float *means2d, const int32_t *radii,
                               uint32_t tile_size, uint32_t tile_width, uint32_t tile_height,
                               int32_t *tiles_per_gauss, int64_t *cum_tiles, void *workspace,
                               size_t workspace_bytes, gsb_stream_t stream) {
    const uint64_t n = (uint64_t)C * N;
    if (n == 0) return GSB_OK;
    if (!means2d || !radii || !tiles_per_gauss || !cum_tiles || tile_size == 0) return GSB_E_INVALID;
    if (!workspace || workspace_bytes < gsb_isect_count_workspace(n)) return GSB_E_WORKSPACE;
    cudaStream_t s = gsb::as_stream(stream);
    const uint32_t grid = (uint32_t)((n + gsb::kIsectThreads - 1) / gsb::kIsectThreads);
    gsb::ProfScope ps("isect_count", s); // count kernel + scan
    gsb::isect_count_kernel<<<grid, gsb::kIsectThreads, 0, s>>>(n, means2d, radii, tile_size, tile_width,
                                                               tile_height, tiles_per_gauss);
    GSB_LAUNCH_CHECK();
    size_t bytes = workspace_bytes;
    gsb::CountIter it(tiles_per_gauss, gsb::CastI64());
    GSB_CUDA_TRY(cub::DeviceScan::InclusiveSum(workspace, bytes, it, cum_tiles, (int64_t)n, s));
    return GSB_OK;
}

extern "C" int gsb_isect_emit(uint32_t C, uint32_t N, const float *means2d, const int32_t *radii,
                              const float *depths, const int64_t *cum_tiles, uint32_t tile_size,
                              uint32_t tile_width, uint32_t tile_height, int64_t *isect_ids,
                              int32_t *flatten_ids, gsb_stream_t stream) {
    const uint64_t n = (uint64_t)C * N;
    if (n == 0) return GSB_OK;
    if (!means2d || !radii || !depths || !cum_tiles || !isect_ids || !flatten_ids) return GSB_E_INVALID;
    const uint32_t tile_n_bits = gsb::bit_width_u32(tile_width * tile_height);
    const uint32_t cam_n_bits = gsb::bit_width_u32(C);
    if (tile_n_bits + cam_n_bits > 32) return GSB_E_INVALID; // Intersect.cpp:50
    const uint32_t grid = (uint32_t)((n + gsb::kIsectThreads - 1) / gsb::kIsectThreads);
    gsb::ProfScope ps("isect_emit", gsb::as_stream(stream));
    gsb::isect_emit_kernel<<<grid, gsb::kIsectThreads, 0, gsb::as_stream(stream)>>>(
        n, N, means2d, radii, depths, nullptr, cum_tiles, tile_size, tile_width, tile_height, tile_n_bits, isect_ids,
        flatten_ids);
    GSB_LAUNCH_CHECK();
    return GSB_OK;
}

extern "C" size_t gsb_isect_sort_workspace(uint64_t n_isects) {
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const int64_t *)nullptr, (int64_t *)nullptr,
                                    (const int32_t *)nullptr, (int32_t *)nullptr, (int64_t)n_isects, 0, 64);
    return gsb::align256(bytes) + 256;
}

extern "C" int gsb_isect_sort(uint64_t n_isects, uint32_t C, uint32_t tile_width, uint32_t tile_height,
                              const int64_t *isect_ids_in, const int32_t *flatten_ids_in,
                              int64_t *isect_ids_out, int32_t *flatten_ids_out, void *workspace,
                              size_t workspace_bytes, gsb_stream_t stream) {
    if (n_isects == 0) return GSB_OK;
    if (!isect_ids_in || !flatten_ids_in || !isect_ids_out || !flatten_ids_out) return GSB_E_INVALID;
    if (!workspace || workspace_bytes < gsb_isect_sort_workspace(n_isects)) return GSB_E_WORKSPACE;
    const uint32_t tile_n_bits = gsb::bit_width_u32(tile_width * tile_height);
    const uint32_t cam_n_bits = gsb::bit_width_u32(C);
    size_t bytes = workspace_bytes;
    gsb::ProfScope ps("isect_sort", gsb::as_stream(stream));
    GSB_CUDA_TRY(cub::DeviceRadixSort::SortPairs(workspace, bytes, isect_ids_in, isect_ids_out, flatten_ids_in,
                                                 flatten_ids_out, (int64_t)n_isects, 0,
                                                 (int)(32 + tile_n_bits + cam_n_bits), gsb::as_stream(stream)));
    return GSB_OK;
}

namespace gsb {
// Workspace of the plan: [perm N u32][cum N i64] survive until the emit; the rest is scratch.
struct PlanWs {
    size_t perm, cum, keys_a, keys_b, vals_a, cub, cub_bytes, total;
};
static PlanWs plan_ws(uint64_t n, bool multi_cam) {
    PlanWs w;
    const size_t kb = multi_cam ? 8 : 4;
    size_t off = 0;
    w.perm = off; off += align256(n * 4);
    w.cum = off; off += align256(n * 8);
    w.keys_a = off; off += align256(n * kb);
    w.keys_b = off; off += align256(n * kb);
    w.vals_a = off; off += align256(n * 4);
    size_t b1 = 0, b2 = 0;
    if (multi_cam)
        cub::DeviceRadixSort::SortPairs(nullptr, b1, (const uint64_t *)nullptr, (uint64_t *)nullptr,
                                        (const uint32_t *)nullptr, (uint32_t *)nullptr, (int64_t)n, 0, 64);
    else
        cub::DeviceRadixSort::SortPairs(nullptr, b1, (const uint32_t *)nullptr, (uint32_t *)nullptr,
                                        (const uint32_t *)nullptr, (uint32_t *)nullptr, (int64_t)n, 0, 32);
    PermCountIter it(thrust::permutation_iterator<const int32_t *, const uint32_t *>(nullptr, nullptr), CastI64());
    cub::DeviceScan::InclusiveSum(nullptr, b2, it, (int64_t *)nullptr, (int64_t)n);
    w.cub_bytes = align256(b1 > b2 ? b1 : b2) + 256;
    w.cub = off; off += w.cub_bytes;
    w.total = off + 256;
    return w;
}
struct PlannedWs {
    size_t tmp_keys, tmp_vals, cub, cub_bytes, total;
};
static PlannedWs planned_ws(uint64_t n_isects) {
    PlannedWs w;
    size_t off = 0;
    w.tmp_keys = off; off += align256(n_isects * 8);
    w.tmp_vals = off; off += align256(n_isects * 4);
    size_t b = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, b, (const int64_t *)nullptr, (int64_t *)nullptr, (const int32_t *)nullptr,
                                    (int32_t *)nullptr, (int64_t)n_isects, 32, 64);
    w.cub_bytes = align256(b) + 256;
    w.cub = off; off += w.cub_bytes;
    w.total = off + 256;
    return w;
}
} // namespace gsb

extern "C" size_t gsb_isect_plan_workspace(uint32_t C, uint32_t N) {
    return gsb::plan_ws((uint64_t)C * N, C > 1).total;
}

extern "C" int gsb_isect_plan(uint32_t C, uint32_t N, const float *means2d, const int32_t *radii, const float *depths,
                              uint32_t tile_size, uint32_t tile_width, uint32_t tile_height,
                              int32_t *tiles_per_gauss, int64_t *n_isects_out, void *plan_workspace,
                              size_t plan_workspace_bytes, gsb_stream_t stream) {
    using namespace gsb;
    const uint64_t n = (uint64_t)C * N;
    if (!n_isects_out) return GSB_E_INVALID;
    cudaStream_t s = as_stream(stream);
    if (n == 0) {
        GSB_CUDA_TRY(cudaMemsetAsync(n_isects_out, 0, sizeof(int64_t), s));
        return GSB_OK;
    }
    if (!means2d || !radii || !depths || !tiles_per_gauss || tile_size == 0) return GSB_E_INVALID;
    const uint32_t tile_n_bits = bit_width_u32(tile_width * tile_height);
    const uint32_t cam_n_bits = bit_width_u32(C);
    if (tile_n_bits + cam_n_bits > 32) return GSB_E_INVALID; // Intersect.cpp:50
    const bool multi = C > 1;
    const PlanWs w = plan_ws(n, multi);
    if (!plan_workspace || (reinterpret_cast<uintptr_t>(plan_workspace) & 255) || plan_workspace_bytes < w.total)
        return GSB_E_WORKSPACE;
    char *base = reinterpret_cast<char *>(plan_workspace);
    uint32_t *perm = reinterpret_cast<uint32_t *>(base + w.perm);
    int64_t *cum = reinterpret_cast<int64_t *>(base + w.cum);
    uint32_t *vals_a = reinterpret_cast<uint32_t *>(base + w.vals_a);
    void *cub_tmp = base + w.cub;
    const uint32_t grid = (uint32_t)((n + kIsectThreads - 1) / kIsectThreads);
    {
        ProfScope ps("isect_count", s);
        uint64_t *k64 = multi ? reinterpret_cast<uint64_t *>(base + w.keys_a) : nullptr;
        uint32_t *k32 = multi ? nullptr : reinterpret_cast<uint32_t *>(base + w.keys_a);
        isect_plan_kernel<<<grid, kIsectThreads, 0, s>>>(n, N, means2d, radii, depths, tile_size, tile_width,
                                                        tile_height, tiles_per_gauss, k64, k32, vals_a);
        GSB_LAUNCH_CHECK();
    }
    {
        ProfScope ps("isect_depth_sort", s);
        size_t bytes = w.cub_bytes;
        if (multi) {
            uint64_t *ka = reinterpret_cast<uint64_t *>(base + w.keys_a), *kb = reinterpret_cast<uint64_t *>(base + w.keys_b);
            GSB_CUDA_TRY(cub::DeviceRadixSort::SortPairs(cub_tmp, bytes, ka, kb, vals_a, perm, (int64_t)n, 0,
                                                         (int)(32 + cam_n_bits), s));
        } else {
            uint32_t *ka = reinterpret_cast<uint32_t *>(base + w.keys_a), *kb = reinterpret_cast<uint32_t *>(base + w.keys_b);
            GSB_CUDA_TRY(cub::DeviceRadixSort::SortPairs(cub_tmp, bytes, ka, kb, vals_a, perm, (int64_t)n, 0, 32, s));
        }
        bytes = w.cub_bytes;
        PermCountIter it(thrust::permutation_iterator<const int32_t *, const uint32_t *>(tiles_per_gauss, perm), CastI64());
        GSB_CUDA_TRY(cub::DeviceScan::InclusiveSum(cub_tmp, bytes, it, cum, (int64_t)n, s));
    }
    // device or pinned-host destination alike
    GSB_CUDA_TRY(cudaMemcpyAsync(n_isects_out, cum + (n - 1), sizeof(int64_t), cudaMemcpyDefault, s));
    return GSB_OK;
}

extern "C" size_t gsb_isect_emit_planned_workspace(uint64_t n_isects) { return gsb::planned_ws(n_isects).total; }

extern "C" int gsb_isect_emit_planned(uint32_t C, uint32_t N, const float *means2d, const int32_t *radii,
                                      const float *depths, uint32_t tile_size, uint32_t tile_width,
                                      uint32_t tile_height, uint64_t n_isects, const void *plan_workspace,
                                      int64_t *isect_ids, int32_t *flatten_ids, void *workspace,
                                      size_t workspace_bytes, gsb_stream_t stream) {
    using namespace gsb;
    const uint64_t n = (uint64_t)C * N;
    if (n == 0 || n_isects == 0) return GSB_OK;
    if (!means2d || !radii || !depths || !plan_workspace || !isect_ids || !flatten_ids) return GSB_E_INVALID;
    const uint32_t tile_n_bits = bit_width_u32(tile_width * tile_height);
    const uint32_t cam_n_bits = bit_width_u32(C);
    if (tile_n_bits + cam_n_bits > 32) return GSB_E_INVALID;
    const PlanWs pw = plan_ws(n, C > 1);
    const PlannedWs w = planned_ws(n_isects);
    if (!workspace || (reinterpret_cast<uintptr_t>(workspace) & 255) || workspace_bytes < w.total)
        return GSB_E_WORKSPACE;
    cudaStream_t s = as_stream(stream);
    const char *pbase = reinterpret_cast<const char *>(plan_workspace);
    const uint32_t *perm = reinterpret_cast<const uint32_t *>(pbase + pw.perm);
    const int64_t *cum = reinterpret_cast<const int64_t *>(pbase + pw.cum);
    char *base = reinterpret_cast<char *>(workspace);
    int64_t *tmp_keys = reinterpret_cast<int64_t *>(base + w.tmp_keys);
    int32_t *tmp_vals = reinterpret_cast<int32_t *>(base + w.tmp_vals);
    {
        ProfScope ps("isect_emit", s);
        const uint32_t egrid = (uint32_t)((n_isects + kEmitSpan - 1) / kEmitSpan);
        isect_emit_balanced_kernel<<<egrid, kIsectThreads, 0, s>>>(n, N, n_isects, means2d, radii, depths, perm, cum,
                                                                  tile_size, tile_width, tile_height, tile_n_bits,
                                                                  tmp_keys, tmp_vals);
        GSB_LAUNCH_CHECK();
    }
    {
        ProfScope ps("isect_sort", s);
        size_t bytes = w.cub_bytes;
        GSB_CUDA_TRY(cub::DeviceRadixSort::SortPairs(base + w.cub, bytes, tmp_keys, isect_ids, tmp_vals, flatten_ids,
                                                     (int64_t)n_isects, 32, (int)(32 + tile_n_bits + cam_n_bits), s));
    }
    return GSB_OK;
}

extern "C" int gsb_isect_offsets(uint64_t n_isects, const int64_t *isect_ids_sorted, uint32_t C,
                                 uint32_t tile_width, uint32_t tile_height, int32_t *offsets,
                                 gsb_stream_t stream) {
    const uint32_t n_tiles = tile_width * tile_height;
    const uint64_t total = (uint64_t)C * n_tiles;
    if (total == 0) return GSB_OK;
    if (!offsets) return GSB_E_INVALID;
    cudaStream_t s = gsb::as_stream(stream);
    if (n_isects == 0) { // IntersectTile.cu:268-271
        GSB_CUDA_TRY(cudaMemsetAsync(offsets, 0, total * sizeof(int32_t), s));
        return GSB_OK;
    }
    if (!isect_ids_sorted) return GSB_E_INVALID;
    const uint32_t tile_n_bits = gsb::bit_width_u32(n_tiles);
    const uint32_t grid = (uint32_t)((n_isects + gsb::kIsectThreads - 1) / gsb::kIsectThreads);
    gsb::ProfScope ps("isect_offsets", s);
    gsb::isect_offsets_kernel<<<grid, gsb::kIsectThreads, 0, s>>>(n_isects, isect_ids_sorted, (uint32_t)total,
                                                                  n_tiles, tile_n_bits, offsets);
    GSB_LAUNCH_CHECK();
    return GSB_OK;
}
