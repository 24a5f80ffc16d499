// gsb_intersect.cu -- a5/a6: Gaussian -> tile expansion, (camera, tile, depth) sort, tile offsets.
//
// Implements gsplat::intersect_tile / intersect_offset (reference: gsplat/IntersectTile.cu:24-114
// count+emit, :206-252 offsets, :290-328 radix sort; gsplat/Intersect.cpp:15-137 host side).
// All results are integer functions of (means2d, radii, depths) and are bit-exact with the
// reference: same float tile-bbox arithmetic (division by the tile size, floor/ceil, CUDA's
// saturating float->uint32 conversion), same 64-bit key layout
//      cam_id << (32 + tile_bits) | tile_id << 32 | float_bits(depth)
// and a stable LSD radix sort over the same low 32+tile_bits+cam_bits bits, so ties keep the
// emission order exactly like cub::DeviceRadixSort does in the reference.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/iterator/transform_input_iterator.cuh>

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

__global__ void __launch_bounds__(kIsectThreads) isect_emit_kernel(uint64_t n, uint32_t N,
                                                                    const float *__restrict__ means2d,
                                                                    const int32_t *__restrict__ radii,
                                                                    const float *__restrict__ depths,
                                                                    const int64_t *__restrict__ cum_tiles,
                                                                    uint32_t tile_size, uint32_t tile_width,
                                                                    uint32_t tile_height, uint32_t tile_n_bits,
                                                                    int64_t *__restrict__ isect_ids,
                                                                    int32_t *__restrict__ flatten_ids) {
    const uint64_t idx = (uint64_t)blockIdx.x * kIsectThreads + threadIdx.x;
    if (idx >= n) return;
    const TileBox b = tile_box(means2d, radii, idx, tile_size, tile_width, tile_height);
    if (!b.active) return;
    const int64_t cid = (int64_t)(idx / N);
    const int64_t cid_enc = cid << (32 + tile_n_bits);
    const int64_t depth_enc = (int64_t)__float_as_uint(depths[idx]); // zero-extended bit pattern (:98-99)
    int64_t cur = (idx == 0) ? 0 : cum_tiles[idx - 1];
    for (uint32_t i = b.y0; i < b.y1; ++i)
        for (uint32_t j = b.x0; j < b.x1; ++j) {
            const int64_t tile_id = (int64_t)i * tile_width + j;
            isect_ids[cur] = cid_enc | (tile_id << 32) | depth_enc;
            flatten_ids[cur] = (int32_t)idx;
            ++cur;
        }
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
using CountIter = cub::TransformInputIterator<int64_t, CastI64, const int32_t *>;

static inline size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

// ------------------------------------------------------------------------------------------
// Binned path (sort == true): instead of a 46-bit global radix sort over 12-byte pairs
// (6 CUB passes, ~150 B of HBM traffic per intersection) the intersections are
//   1. counted per Gaussian AND per (camera, tile) bucket   (isect_count_hist_kernel)
//   2. bucket starts = exclusive scan of the per-tile counts (tile_scan_kernel, one CTA)
//   3. scattered into their tile's segment as (depth_bits << 32 | flat_index) keys
//   4. sorted inside each tile by that 64-bit key in shared memory (tile_sort_kernel), which
//      writes the final isect_ids / flatten_ids.
// Because the key orders by depth first and by the flattened Gaussian index second, the result is
// exactly the stable sort of the emission sequence by (camera, tile, depth) that the reference's
// cub::DeviceRadixSort produces -- including ties -- but with ~36 B of traffic per intersection
// and no dependence on the scatter order (atomics only hand out slots).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kIsectThreads) isect_count_hist_kernel(
    uint64_t n, uint32_t N, const float *__restrict__ means2d, const int32_t *__restrict__ radii, uint32_t tile_size,
    uint32_t tile_width, uint32_t tile_height, int32_t *__restrict__ tiles_per_gauss,
    uint32_t *__restrict__ tile_counts) {
    const uint64_t idx = (uint64_t)blockIdx.x * kIsectThreads + threadIdx.x;
    if (idx >= n) return;
    const TileBox b = tile_box(means2d, radii, idx, tile_size, tile_width, tile_height);
    tiles_per_gauss[idx] = b.active ? (int32_t)((b.y1 - b.y0) * (b.x1 - b.x0)) : 0;
    if (!b.active) return;
    uint32_t *tc = tile_counts + (size_t)(idx / N) * tile_width * tile_height;
    for (uint32_t i = b.y0; i < b.y1; ++i)
        for (uint32_t j = b.x0; j < b.x1; ++j) atomicAdd(tc + i * tile_width + j, 1u);
}

// One CTA: exclusive scan of the per-tile counts, total and maximum.  totals = {n_isects, max count}.
constexpr int kScanThreads = 1024;
__global__ void __launch_bounds__(kScanThreads) tile_scan_kernel(uint32_t total_tiles,
                                                                 const uint32_t *__restrict__ tile_counts,
                                                                 uint32_t *__restrict__ tile_starts,
                                                                 uint64_t *__restrict__ totals) {
    __shared__ uint32_t s_warp[kScanThreads / 32];
    __shared__ uint32_t s_carry, s_max;
    const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) { s_carry = 0; s_max = 0; }
    __syncthreads();
    uint32_t lmax = 0;
    for (uint32_t base = 0; base < total_tiles; base += kScanThreads) {
        const uint32_t i = base + tid;
        const uint32_t v = (i < total_tiles) ? tile_counts[i] : 0;
        lmax = max(lmax, v);
        uint32_t x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= (uint32_t)o) x += y;
        }
        if (lane == 31) s_warp[wid] = x;
        __syncthreads();
        if (wid == 0) {
            uint32_t w = s_warp[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= (uint32_t)o) w += y;
            }
            s_warp[lane] = w; // inclusive scan of warp totals
        }
        __syncthreads();
        const uint32_t carry = s_carry;
        const uint32_t excl = carry + (wid ? s_warp[wid - 1] : 0) + (x - v);
        if (i < total_tiles) tile_starts[i] = excl;
        __syncthreads();
        if (tid == kScanThreads - 1) s_carry = excl + v;
        __syncthreads();
    }
    lmax = __reduce_max_sync(0xffffffffu, lmax);
    if (lane == 0) atomicMax(&s_max, lmax);
    __syncthreads();
    if (tid == 0) { totals[0] = s_carry; totals[1] = s_max; }
}

__global__ void __launch_bounds__(kIsectThreads) isect_scatter_kernel(
    uint64_t n, uint32_t N, const float *__restrict__ means2d, const int32_t *__restrict__ radii,
    const float *__restrict__ depths, uint32_t tile_size, uint32_t tile_width, uint32_t tile_height,
    const uint32_t *__restrict__ tile_starts, uint32_t *__restrict__ tile_cursor, uint64_t *__restrict__ bucket) {
    const uint64_t idx = (uint64_t)blockIdx.x * kIsectThreads + threadIdx.x;
    if (idx >= n) return;
    const TileBox b = tile_box(means2d, radii, idx, tile_size, tile_width, tile_height);
    if (!b.active) return;
    const size_t cam_off = (size_t)(idx / N) * tile_width * tile_height;
    const uint64_t key = ((uint64_t)__float_as_uint(depths[idx]) << 32) | (uint64_t)(uint32_t)idx;
    for (uint32_t i = b.y0; i < b.y1; ++i)
        for (uint32_t j = b.x0; j < b.x1; ++j) {
            const size_t t = cam_off + (size_t)i * tile_width + j;
            const uint32_t pos = tile_starts[t] + atomicAdd(tile_cursor + t, 1u);
            bucket[pos] = key;
        }
}

// One CTA per (camera, tile): bitonic sort of the tile's 64-bit keys in shared memory.
constexpr int kTileSortThreads = 256;
__global__ void __launch_bounds__(kTileSortThreads) tile_sort_kernel(
    uint32_t n_tiles, uint32_t tile_n_bits, const uint32_t *__restrict__ tile_starts,
    const uint32_t *__restrict__ tile_counts, const uint64_t *__restrict__ bucket, int64_t *__restrict__ isect_ids,
    int32_t *__restrict__ flatten_ids) {
    extern __shared__ __align__(16) uint64_t s_keys[];
    const uint32_t t = blockIdx.x;
    const uint32_t n = tile_counts[t];
    if (n == 0) return;
    const uint32_t start = tile_starts[t];
    const uint32_t tid = threadIdx.x;
    uint32_t m = 1;
    while (m < n) m <<= 1;
    for (uint32_t i = tid; i < m; i += kTileSortThreads) s_keys[i] = (i < n) ? bucket[start + i] : ~0ull;
    __syncthreads();
    for (uint32_t k = 2; k <= m; k <<= 1) {
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            for (uint32_t p = tid; p < (m >> 1); p += kTileSortThreads) {
                // p-th compare-exchange of this stage: i has bit j clear
                const uint32_t i = ((p & ~(j - 1)) << 1) | (p & (j - 1));
                const uint32_t l = i | j;
                const uint64_t a = s_keys[i], b = s_keys[l];
                const bool up = (i & k) == 0;
                if ((a > b) == up) { s_keys[i] = b; s_keys[l] = a; }
            }
            __syncthreads();
        }
    }
    const int64_t hi = ((int64_t)(t / n_tiles) << (32 + tile_n_bits)) | ((int64_t)(t % n_tiles) << 32);
    for (uint32_t i = tid; i < n; i += kTileSortThreads) {
        const uint64_t key = s_keys[i];
        isect_ids[start + i] = hi | (int64_t)(key >> 32);
        flatten_ids[start + i] = (int32_t)(uint32_t)key;
    }
}

constexpr uint32_t kMaxTileSortElems = 16 * 1024; // power of two: 128 KB of shared memory

} // namespace gsb

// tile workspace: [tile_counts | tile_starts | tile_cursor], each total_tiles x uint32
extern "C" size_t gsb_isect_binned_tile_workspace(uint64_t total_tiles) {
    return 3 * gsb::align256(total_tiles * 4) + 256;
}
// bucket workspace: n_isects x uint64 keys
extern "C" size_t gsb_isect_binned_bucket_workspace(uint64_t n_isects) { return gsb::align256(n_isects * 8) + 256; }

// Step 1+2: tiles_per_gauss, per-tile counts and starts; totals_out (DEVICE, 2 x uint64) receives
// {n_isects, largest per-tile count}.  `workspace` must be the buffer later handed to
// gsb_isect_binned_sort (sized with n_isects = 0 for this call is fine: only the header is used).
extern "C" int gsb_isect_binned_count(uint32_t C, uint32_t N, const float *means2d, const int32_t *radii,
                                      uint32_t tile_size, uint32_t tile_width, uint32_t tile_height,
                                      int32_t *tiles_per_gauss, uint64_t *totals_out, void *tile_workspace,
                                      size_t tile_workspace_bytes, gsb_stream_t stream) {
    const uint64_t n = (uint64_t)C * N;
    const uint64_t total_tiles = (uint64_t)C * tile_width * tile_height;
    if (!totals_out || !tile_workspace) return GSB_E_INVALID;
    if (tile_workspace_bytes < 3 * gsb::align256(total_tiles * 4)) return GSB_E_WORKSPACE;
    cudaStream_t s = gsb::as_stream(stream);
    char *w = reinterpret_cast<char *>(tile_workspace);
    uint32_t *tile_counts = reinterpret_cast<uint32_t *>(w);
    uint32_t *tile_starts = reinterpret_cast<uint32_t *>(w + gsb::align256(total_tiles * 4));
    uint32_t *tile_cursor = reinterpret_cast<uint32_t *>(w + 2 * gsb::align256(total_tiles * 4));
    GSB_CUDA_TRY(cudaMemsetAsync(tile_counts, 0, total_tiles * 4, s));
    GSB_CUDA_TRY(cudaMemsetAsync(tile_cursor, 0, total_tiles * 4, s));
    if (n > 0) {
        if (!means2d || !radii || !tiles_per_gauss || tile_size == 0) return GSB_E_INVALID;
        gsb::ProfScope ps("isect_count", s);
        const uint32_t grid = (uint32_t)((n + gsb::kIsectThreads - 1) / gsb::kIsectThreads);
        gsb::isect_count_hist_kernel<<<grid, gsb::kIsectThreads, 0, s>>>(n, N, means2d, radii, tile_size, tile_width,
                                                                        tile_height, tiles_per_gauss, tile_counts);
        GSB_LAUNCH_CHECK();
    }
    {
        gsb::ProfScope ps("isect_scan", s);
        gsb::tile_scan_kernel<<<1, gsb::kScanThreads, 0, s>>>((uint32_t)total_tiles, tile_counts, tile_starts, totals_out);
    }
    GSB_LAUNCH_CHECK();
    return GSB_OK;
}

// Step 3+4.  Returns GSB_E_UNSUPPORTED when a tile holds more keys than fit in shared memory
// (the caller then falls back to gsb_isect_emit + gsb_isect_sort).
extern "C" int gsb_isect_binned_sort(uint32_t C, uint32_t N, const float *means2d, const int32_t *radii,
                                     const float *depths, uint32_t tile_size, uint32_t tile_width,
                                     uint32_t tile_height, uint64_t n_isects, uint64_t max_tile_count,
                                     void *tile_workspace, void *bucket_workspace, size_t bucket_workspace_bytes,
                                     int64_t *isect_ids, int32_t *flatten_ids, gsb_stream_t stream) {
    const uint64_t n = (uint64_t)C * N;
    if (n == 0 || n_isects == 0) return GSB_OK;
    if (max_tile_count > gsb::kMaxTileSortElems) return GSB_E_UNSUPPORTED;
    if (!means2d || !radii || !depths || !tile_workspace || !bucket_workspace || !isect_ids || !flatten_ids)
        return GSB_E_INVALID;
    if (bucket_workspace_bytes < n_isects * 8) return GSB_E_WORKSPACE;
    const uint32_t n_tiles = tile_width * tile_height;
    const uint64_t total_tiles = (uint64_t)C * n_tiles;
    const uint32_t tile_n_bits = gsb::bit_width_u32(n_tiles);
    if (tile_n_bits + gsb::bit_width_u32(C) > 32) return GSB_E_INVALID;
    cudaStream_t s = gsb::as_stream(stream);
    char *w = reinterpret_cast<char *>(tile_workspace);
    uint32_t *tile_counts = reinterpret_cast<uint32_t *>(w);
    uint32_t *tile_starts = reinterpret_cast<uint32_t *>(w + gsb::align256(total_tiles * 4));
    uint32_t *tile_cursor = reinterpret_cast<uint32_t *>(w + 2 * gsb::align256(total_tiles * 4));
    uint64_t *bucket = reinterpret_cast<uint64_t *>(bucket_workspace);
    {
        gsb::ProfScope ps("isect_emit", s);
        const uint32_t grid = (uint32_t)((n + gsb::kIsectThreads - 1) / gsb::kIsectThreads);
        gsb::isect_scatter_kernel<<<grid, gsb::kIsectThreads, 0, s>>>(n, N, means2d, radii, depths, tile_size, tile_width,
                                                                     tile_height, tile_starts, tile_cursor, bucket);
    }
    GSB_LAUNCH_CHECK();
    uint32_t m = 1;
    while (m < max_tile_count) m <<= 1;
    const size_t smem = (size_t)m * 8;
    // always the same value, so concurrent callers cannot lower each other's limit
    GSB_CUDA_TRY(cudaFuncSetAttribute(gsb::tile_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)(gsb::kMaxTileSortElems * 8)));
    {
        gsb::ProfScope ps("isect_sort", s);
        gsb::tile_sort_kernel<<<(uint32_t)total_tiles, gsb::kTileSortThreads, smem, s>>>(
            n_tiles, tile_n_bits, tile_starts, tile_counts, bucket, isect_ids, flatten_ids);
    }
    GSB_LAUNCH_CHECK();
    return GSB_OK;
}

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
        n, N, means2d, radii, depths, cum_tiles, tile_size, tile_width, tile_height, tile_n_bits, isect_ids,
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
