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
// How the sorted lists are built (gsb_isect_plan + gsb_isect_emit_planned).  The reference radix-sorts
// all I intersections on 46 bits (six 8-bit passes over 12-byte pairs, ~150 B of HBM traffic per
// intersection).  Here no intersection is ever sorted:
//   1. plan      one thread per Gaussian: tile box, tile count, depth key.
//   2. depth     stable LSD radix sort of the C*N Gaussians by depth bits (N elements, not I;
//                gsb_devsort.cuh), passes whose digit is constant are skipped.
//   3. runs      scan of the tile counts in depth order -> compact run table (Gaussians with >= 1 tile:
//                cumulative end slot, index, box, depth key); n_isects goes to the host here.
//   4. histogram the I intersection slots, in depth order, are cut into P equal chunks, one warp each.  A
//                warp walks its chunk 32 slots per step (slot -> run by a ballot over the run ends, run ->
//                tile by the box) and counts the tiles in shared-memory counters: matrix M[P][tiles].
//   5. column scan over (tile, chunk): M[p][t] becomes the first output slot of chunk p in tile t;
//                the tile offsets (a6) fall out of the same scan.
//   -- host reads n_isects, allocates the outputs --
//   6. scatter   the same walk again; slot ranks within a step come from match.any on the tile id, so
//                every (tile, depth, index) lands DIRECTLY in its final position: 4 (+8 for the key) bytes
//                written per intersection, nothing read back.
// A stable placement by tile of a sequence ordered by (depth, index) is ordered by (tile, depth, index):
// the same permutation as the reference's single 46-bit sort, ties included.
#include "gsb_devsort.cuh"

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

// Unsorted path (sort == false): Gaussian i writes its intersections at cum[i-1], row-major over its box.
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
    const int64_t cid_enc = (int64_t)(idx / N) << (32 + tile_n_bits);
    const int64_t depth_enc = (int64_t)__float_as_uint(depths[idx]); // zero-extended bit pattern (:98-99)
    int64_t cur = (idx == 0) ? 0 : cum_tiles[idx - 1];
    for (uint32_t y = b.y0; y < b.y1; ++y)
        for (uint32_t x = b.x0; x < b.x1; ++x) {
            const int64_t tile_id = (int64_t)y * tile_width + x;
            isect_ids[cur] = cid_enc | (tile_id << 32) | depth_enc;
            flatten_ids[cur] = (int32_t)idx;
            ++cur;
        }
}

// ---- planned path, step 1: tile box, count and depth key of every Gaussian ---------------------------------
// boxes[idx] = (x0 | y0 << 16, w | h << 16) in tiles.  Gaussians without tiles get the key 0xffffffff: they
// end up behind every Gaussian that has tiles and never enter the run table.
__global__ void __launch_bounds__(kIsectThreads) isect_plan_kernel(uint64_t n, const float *__restrict__ means2d,
                                                                    const int32_t *__restrict__ radii,
                                                                    const float *__restrict__ depths,
                                                                    uint32_t tile_size, uint32_t tile_width,
                                                                    uint32_t tile_height,
                                                                    int32_t *__restrict__ tiles_per_gauss,
                                                                    uint32_t *__restrict__ keys,
                                                                    uint2 *__restrict__ boxes,
                                                                    SortCtl *__restrict__ ctl) {
    __shared__ uint32_t s_or[kIsectThreads / 32], s_nor[kIsectThreads / 32];
    const uint64_t idx = (uint64_t)blockIdx.x * kIsectThreads + threadIdx.x;
    uint32_t k_or = 0u, k_nor = 0u;
    if (idx < n) {
        const TileBox b = tile_box(means2d, radii, idx, tile_size, tile_width, tile_height);
        const uint32_t w = b.x1 - b.x0, h = b.y1 - b.y0;
        const int32_t cnt = b.active ? (int32_t)(w * h) : 0;
        tiles_per_gauss[idx] = cnt;
        uint32_t key = 0xffffffffu;
        if (cnt > 0) {
            key = __float_as_uint(depths[idx]);
            k_or = key;
            k_nor = ~key;
        }
        keys[idx] = key;
        boxes[idx] = make_uint2(b.x0 | (b.y0 << 16), w | (h << 16));
    }
    k_or = __reduce_or_sync(0xffffffffu, k_or);
    k_nor = __reduce_or_sync(0xffffffffu, k_nor);
    if ((threadIdx.x & 31) == 0) { s_or[threadIdx.x >> 5] = k_or; s_nor[threadIdx.x >> 5] = k_nor; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < kIsectThreads / 32; ++w) { k_or |= s_or[w]; k_nor |= s_nor[w]; }
        if (k_or | k_nor) { // at least one Gaussian of this CTA touches a tile
            atomicOr(&ctl->key_or, k_or);
            atomicOr(&ctl->key_nor, k_nor);
        }
    }
}

// ---- step 3: run table ---------------------------------------------------------------------------------------
// One scan over the Gaussians in depth order of the pair (has tiles ? 1 : 0, tile count).
struct RunAcc {
    unsigned long long cnt;
    uint32_t runs;
    __device__ __forceinline__ RunAcc &operator+=(const RunAcc &o) { cnt += o.cnt; runs += o.runs; return *this; }
};
__device__ __forceinline__ RunAcc operator+(RunAcc a, const RunAcc &b) { a += b; return a; }

__device__ __forceinline__ RunAcc run_shfl_up(RunAcc v, int o) {
    RunAcc r;
    r.cnt = __shfl_up_sync(0xffffffffu, v.cnt, o);
    r.runs = __shfl_up_sync(0xffffffffu, v.runs, o);
    return r;
}
__device__ __forceinline__ RunAcc run_block_scan(RunAcc v, RunAcc *s_warp, RunAcc &total) {
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const RunAcc t = run_shfl_up(v, o);
        if (lane >= (uint32_t)o) v += t;
    }
    __syncthreads();
    if (lane == 31) s_warp[warp] = v;
    __syncthreads();
    RunAcc add = {0ull, 0u}, tot = {0ull, 0u};
    for (int w = 0; w < kSortWarps; ++w) {
        const RunAcc x = s_warp[w];
        if ((uint32_t)w < warp) add += x;
        tot += x;
    }
    total = tot;
    return v + add;
}

struct RunTable {
    uint32_t *end;  // cumulative number of intersections up to and including this run
    uint32_t *idx;  // flattened Gaussian index (camera * N + gaussian)
    uint2 *box;
    uint32_t *key;  // depth bits
};

__device__ __forceinline__ const uint32_t *sorted_vals(const SortCtl *ctl, const uint32_t *v0, const uint32_t *v1) {
    return (radix_passes_done(ctl, 4) & 1) ? v1 : v0;
}

__global__ void __launch_bounds__(kSortThreads) runs_blocksum_kernel(const SortCtl *__restrict__ ctl,
                                                                     const uint32_t *__restrict__ v0,
                                                                     const uint32_t *__restrict__ v1,
                                                                     const int32_t *__restrict__ counts, uint64_t n,
                                                                     uint32_t seg, RunAcc *__restrict__ bsum) {
    __shared__ RunAcc s_warp[kSortWarps];
    const uint32_t *perm = sorted_vals(ctl, v0, v1);
    const SegRange r = cta_segment(n, seg);
    RunAcc acc = {0ull, 0u};
    for (uint64_t i = r.lo + threadIdx.x; i < r.hi; i += kSortThreads) {
        const int32_t c = counts[perm[i]];
        acc.cnt += (unsigned long long)c;
        acc.runs += c > 0 ? 1u : 0u;
    }
    RunAcc tot;
    run_block_scan(acc, s_warp, tot);
    if (threadIdx.x == 0) bsum[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(kSortThreads) runs_build_kernel(SortCtl *__restrict__ ctl,
                                                                  const uint32_t *__restrict__ v0,
                                                                  const uint32_t *__restrict__ v1,
                                                                  const uint32_t *__restrict__ k0,
                                                                  const uint32_t *__restrict__ k1,
                                                                  const int32_t *__restrict__ counts,
                                                                  const uint2 *__restrict__ boxes, uint64_t n,
                                                                  uint32_t seg, uint32_t nblocks,
                                                                  const RunAcc *__restrict__ bsum, RunTable rt) {
    __shared__ RunAcc s_warp[kSortWarps];
    const bool odd = (radix_passes_done(ctl, 4) & 1) != 0;
    const uint32_t *perm = odd ? v1 : v0;
    const uint32_t *keys = odd ? k1 : k0;
    RunAcc carry;
    {
        RunAcc part = {0ull, 0u};
        for (uint32_t bb = threadIdx.x; bb < blockIdx.x; bb += kSortThreads) part += bsum[bb];
        run_block_scan(part, s_warp, carry);
    }
    const SegRange r = cta_segment(n, seg);
    for (uint64_t i0 = r.lo; i0 < r.hi; i0 += kSortThreads) {
        const uint64_t i = i0 + threadIdx.x;
        RunAcc v = {0ull, 0u};
        uint32_t idx = 0;
        if (i < r.hi) {
            idx = perm[i];
            const int32_t c = counts[idx];
            v.cnt = (unsigned long long)c;
            v.runs = c > 0 ? 1u : 0u;
        }
        RunAcc tot;
        const RunAcc inc = run_block_scan(v, s_warp, tot);
        if (v.runs) {
            const uint32_t k = carry.runs + inc.runs - 1u;
            rt.end[k] = (uint32_t)(carry.cnt + inc.cnt);
            rt.idx[k] = idx;
            rt.box[k] = boxes[idx];
            rt.key[k] = keys[i];
        }
        carry += tot;
    }
    if (blockIdx.x == nblocks - 1 && threadIdx.x == 0) {
        ctl->n_runs = carry.runs;
        ctl->n_isects = carry.cnt;
    }
}

// ---- steps 4 and 6: the chunk walk -----------------------------------------------------------------------------
struct BinArgs {
    RunTable rt;
    const SortCtl *ctl;
    uint32_t N;          // Gaussians per camera
    uint32_t n_tiles;    // tiles per camera
    uint32_t tile_width;
    uint32_t tile_n_bits;
    uint32_t multi_cam;
    uint32_t t_lo, t_cnt; // window of global tile ids (camera * n_tiles + tile) counted by this launch
    uint32_t T_total, P;
    uint32_t *M;          // [P][T_total]
    int32_t *flatten_ids;
    int64_t *isect_ids;   // nullable
    uint32_t cap;         // capacity of flatten_ids / isect_ids (>= n_isects unless the caller under-allocated)
};

// first r in [0, n_runs) with end[r] > s (exists: s < end[n_runs - 1]); one probe per lane and round
__device__ __forceinline__ uint32_t first_run_after(const uint32_t *__restrict__ end, uint32_t n_runs, uint32_t s) {
    const uint32_t lane = threadIdx.x & 31;
    uint32_t lo = 0, hi = n_runs; // answer in [lo, hi)
    while (hi - lo > 1) {
        const uint32_t len = hi - lo;
        const uint32_t step = (len + 32) / 33;
        const uint32_t p = lo + (lane + 1) * step - 1;
        const bool le = (p < hi - 1) ? (end[p] <= s) : false; // position hi-1 is known to be > s
        const uint32_t k = __popc(__ballot_sync(0xffffffffu, le));
        const uint32_t nlo = (k == 0) ? lo : lo + k * step;
        const uint32_t nhi = (k == 32) ? hi : min(hi, lo + (k + 1) * step);
        lo = nlo; hi = nhi;
    }
    return lo;
}

// The walk of one chunk.  Per step the warp handles 32 consecutive slots: slot -> run -> tile (the MAP half) and
// then rank / count / store (the COMMIT half).  The run table entries the next steps need sit in a per-warp ring in
// shared memory, refilled 32 runs at a time from loads issued a refill earlier, so no global-memory latency is on the
// serial chain (run pointer -> next step); the MAP of step k+1 is issued before the COMMIT of step k so their
// shared-memory / shuffle latencies overlap.
constexpr uint32_t kRing = 128; // run-table entries per warp in shared memory (>= 96 live at any time)

struct RunEnt {
    uint32_t end, idx, bx, by, key;
};
template <bool kScatter>
__device__ __forceinline__ RunEnt load_run(const RunTable &rt, uint32_t r, uint32_t n_runs) {
    RunEnt e;
    e.end = 0xffffffffu; e.idx = 0; e.bx = 0; e.by = 0; e.key = 0;
    if (r < n_runs) {
        e.end = rt.end[r];
        e.idx = rt.idx[r];
        const uint2 b = rt.box[r];
        e.bx = b.x; e.by = b.y;
        if (kScatter) e.key = rt.key[r];
    }
    return e;
}

struct StepMap {
    uint32_t tt;   // tile id relative to the window, or a per-lane dummy when the lane has nothing to place
    uint32_t idx, key, tile, cam;
    bool in;
};

template <bool kScatter>
__global__ void __launch_bounds__(32) tile_bin_kernel(const BinArgs a) {
    extern __shared__ uint32_t s_mem[]; // [t_cnt counters][5 x kRing run entries]
    uint32_t *s_cnt = s_mem;
    uint32_t *s_end = s_mem + ((a.t_cnt + 31u) & ~31u);
    uint32_t *s_idx = s_end + kRing, *s_bx = s_idx + kRing, *s_by = s_bx + kRing, *s_key = s_by + kRing;
    const uint32_t lane = threadIdx.x;
    const uint32_t p = blockIdx.x;
    uint32_t *row = a.M + (size_t)p * a.T_total + a.t_lo;
    const unsigned long long I = a.ctl->n_isects;
    const uint32_t n_runs = a.ctl->n_runs;
    const uint32_t s_begin = (uint32_t)(((unsigned long long)p * I) / a.P);
    const uint32_t s_stop = (uint32_t)(((unsigned long long)(p + 1) * I) / a.P);
    if (kScatter) {
        if (s_begin >= s_stop) return;
        for (uint32_t t = lane; t < a.t_cnt; t += 32) s_cnt[t] = row[t];
    } else {
        for (uint32_t t = lane; t < a.t_cnt; t += 32) s_cnt[t] = 0;
    }
    if (s_begin < s_stop) {
        uint32_t r0 = first_run_after(a.rt.end, n_runs, s_begin);
        uint32_t start0 = r0 ? a.rt.end[r0 - 1] : 0u; // first slot of run r0
        // ring = runs [r0, r0 + 64); `pend` = the 32 runs after those, in flight
        {
            const RunEnt e0 = load_run<kScatter>(a.rt, r0 + lane, n_runs), e1 = load_run<kScatter>(a.rt, r0 + 32 + lane, n_runs);
            const uint32_t w0 = (r0 + lane) & (kRing - 1), w1 = (r0 + 32 + lane) & (kRing - 1);
            s_end[w0] = e0.end; s_idx[w0] = e0.idx; s_bx[w0] = e0.bx; s_by[w0] = e0.by; s_key[w0] = e0.key;
            s_end[w1] = e1.end; s_idx[w1] = e1.idx; s_bx[w1] = e1.bx; s_by[w1] = e1.by; s_key[w1] = e1.key;
        }
        uint32_t ring_end = r0 + 64;
        RunEnt pend = load_run<kScatter>(a.rt, ring_end + lane, n_runs);
        __syncwarp();

        // MAP: slots [s0, s0 + 32) -> (run, tile); advances (r0, start0) to the run that holds slot s0 + 32
        auto map_step = [&](uint32_t s0) {
            if (ring_end < r0 + 64) { // keep 64 runs ahead: a step consumes at most 32
                const uint32_t w = (ring_end + lane) & (kRing - 1);
                s_end[w] = pend.end; s_idx[w] = pend.idx; s_bx[w] = pend.bx; s_by[w] = pend.by; s_key[w] = pend.key;
                ring_end += 32;
                pend = load_run<kScatter>(a.rt, ring_end + lane, n_runs);
                __syncwarp();
            }
            // the run ends inside the step are strictly increasing (no empty runs): bit e of endmask <=> a run ends
            // after slot s0 + e - 1
            const uint32_t rel = s_end[(r0 + lane) & (kRing - 1)] - s0; // > 0 for lane 0
            const uint32_t endmask = __reduce_or_sync(0xffffffffu, rel <= 31u ? (1u << rel) : 0u);
            const uint32_t c = __popc(endmask & ((2u << lane) - 1u)); // runs finished at or before this lane's slot
            const uint32_t w = (r0 + c) & (kRing - 1);
            const uint32_t prev_end = s_end[(r0 + c - 1u) & (kRing - 1)];
            StepMap m;
            m.idx = s_idx[w];
            m.key = kScatter ? s_key[w] : 0u;
            const uint32_t bx = s_bx[w], by = s_by[w];
            const uint32_t slot = s0 + lane;
            const uint32_t run_start = c ? prev_end : start0;
            const uint32_t j = slot - run_start; // position inside the run's box, row-major
            const uint32_t bw = by & 0xffffu;
            m.tt = 0x80000000u | lane;
            m.tile = 0; m.cam = 0; m.in = false;
            if (slot < s_stop) {
                const uint32_t dy = j / bw, dx = j - dy * bw;
                m.tile = ((bx >> 16) + dy) * a.tile_width + (bx & 0xffffu) + dx;
                m.cam = a.multi_cam ? m.idx / a.N : 0u;
                const uint32_t g = m.cam * a.n_tiles + m.tile - a.t_lo;
                m.in = g < a.t_cnt;
                if (m.in) m.tt = g;
            }
            const uint32_t adv = __popc(__ballot_sync(0xffffffffu, rel <= 32u));
            if (adv) {
                start0 = s_end[(r0 + adv - 1u) & (kRing - 1)];
                r0 += adv;
            }
            return m;
        };
        // COMMIT: stable rank of equal tiles inside the step (lane order = depth order), counters, stores
        auto commit_step = [&](const StepMap &m) {
            const uint32_t peers = __match_any_sync(0xffffffffu, m.tt);
            uint32_t base = 0;
            if (m.in) base = s_cnt[m.tt];
            __syncwarp();
            if (m.in) {
                if (kScatter) {
                    const uint32_t pos = base + __popc(peers & ((1u << lane) - 1u));
                    if (pos < a.cap) a.flatten_ids[pos] = (int32_t)m.idx;
                }
                if ((peers >> lane) == 1u) s_cnt[m.tt] = base + __popc(peers);
            }
            __syncwarp();
        };

        StepMap cur = map_step(s_begin);
        for (uint32_t s0 = s_begin; s0 < s_stop; s0 += 32) {
            StepMap nxt = cur;
            if (s0 + 32 < s_stop) nxt = map_step(s0 + 32);
            commit_step(cur);
            cur = nxt;
        }
    }
    if (!kScatter) {
        __syncwarp();
        for (uint32_t t = lane; t < a.t_cnt; t += 32) row[t] = s_cnt[t];
    }
}

// isect_ids of the sorted list, written in order (coalesced) instead of scattered with the values:
//   key[pos] = cam << (32 + tile_bits) | tile << 32 | depth bits of flatten_ids[pos]
// the tile of a position is found in the closed offsets table tile_off[0 .. T] (tile_off[T] = n_isects).
__global__ void __launch_bounds__(kIsectThreads) isect_keys_kernel(uint32_t n, const uint32_t *__restrict__ tile_off,
                                                                   uint32_t T, uint32_t n_tiles, uint32_t tile_n_bits,
                                                                   const int32_t *__restrict__ flatten_ids,
                                                                   const float *__restrict__ depths,
                                                                   int64_t *__restrict__ isect_ids) {
    const uint32_t pos = blockIdx.x * kIsectThreads + threadIdx.x;
    if (pos >= n) return;
    // last t in [0, T) with tile_off[t] <= pos  (offsets are non-decreasing; empty tiles repeat a value)
    uint32_t lo = 0, hi = T;
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (__ldg(tile_off + mid) <= pos) lo = mid; else hi = mid;
    }
    const uint32_t cam = lo / n_tiles, tile = lo - cam * n_tiles;
    const uint32_t idx = (uint32_t)flatten_ids[pos];
    isect_ids[pos] = ((int64_t)cam << (32 + tile_n_bits)) | ((int64_t)tile << 32) | (int64_t)__float_as_uint(depths[idx]);
}

// ---- step 5: exclusive scan of M over (tile, chunk) ----------------------------------------------------------
// col_segsum: seg[s][t] = sum of M[p][t] over the chunks p of segment s.
__global__ void __launch_bounds__(kIsectThreads) col_segsum_kernel(const uint32_t *__restrict__ M, uint32_t T,
                                                                   uint32_t P, uint32_t seg_len,
                                                                   uint32_t *__restrict__ seg) {
    const uint32_t t = blockIdx.x * kIsectThreads + threadIdx.x;
    if (t >= T) return;
    const uint32_t s = blockIdx.y;
    const uint32_t p0 = s * seg_len, p1 = min(P, p0 + seg_len);
    uint32_t acc = 0;
    for (uint32_t p = p0; p < p1; ++p) acc += M[(size_t)p * T + t];
    seg[(size_t)s * T + t] = acc;
}
// col_tilescan (one CTA): seg[s][t] -> exclusive prefix over s plus the tile's first slot; tile offsets out.
__global__ void __launch_bounds__(kSortThreads) col_tilescan_kernel(uint32_t *__restrict__ seg, uint32_t T, uint32_t S,
                                                                    uint32_t *__restrict__ toff /* [T + 1] */,
                                                                    int32_t *__restrict__ tile_offsets /*nullable*/,
                                                                    int write_total) {
    __shared__ uint32_t s_warp[kSortWarps];
    uint32_t carry = 0;
    for (uint32_t t0 = 0; t0 < T; t0 += kSortThreads) {
        const uint32_t t = t0 + threadIdx.x;
        uint32_t tot = 0;
        if (t < T)
            for (uint32_t s = 0; s < S; ++s) tot += seg[(size_t)s * T + t];
        uint32_t all;
        const uint32_t inc = block_scan_inclusive<uint32_t>(tot, s_warp, all);
        if (t < T) {
            uint32_t run = carry + inc - tot; // first slot of tile t
            toff[t] = run;
            if (tile_offsets) tile_offsets[t] = (int32_t)run;
            for (uint32_t s = 0; s < S; ++s) {
                const uint32_t v = seg[(size_t)s * T + t];
                seg[(size_t)s * T + t] = run;
                run += v;
            }
        }
        carry += all;
    }
    if (threadIdx.x == 0) {
        toff[T] = carry; // = n_isects
        if (tile_offsets && write_total) tile_offsets[T] = (int32_t)carry;
    }
}
// col_apply: M[p][t] = first output slot of chunk p in tile t.
__global__ void __launch_bounds__(kIsectThreads) col_apply_kernel(uint32_t *__restrict__ M, uint32_t T, uint32_t P,
                                                                  uint32_t seg_len, const uint32_t *__restrict__ seg) {
    const uint32_t t = blockIdx.x * kIsectThreads + threadIdx.x;
    if (t >= T) return;
    const uint32_t s = blockIdx.y;
    const uint32_t p0 = s * seg_len, p1 = min(P, p0 + seg_len);
    uint32_t run = seg[(size_t)s * T + t];
    for (uint32_t p = p0; p < p1; ++p) {
        const uint32_t v = M[(size_t)p * T + t];
        M[(size_t)p * T + t] = run;
        run += v;
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

// ---- host-side geometry ------------------------------------------------------------------------------------------
static inline size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

struct DeviceShape {
    int sms;
    int smem_optin; // max dynamic shared memory per block
    int smem_sm;    // shared memory per SM
};
static const DeviceShape &device_shape() {
    // immutable after the first call (C++11 thread-safe initialisation); sm_100a: 148 SMs, 227 KB / 228 KB
    static const DeviceShape s = [] {
        DeviceShape d{148, 227 * 1024, 228 * 1024};
        int dev = 0;
        if (cudaGetDevice(&dev) == cudaSuccess) {
            int v = 0;
            if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0) d.sms = v;
            if (cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) == cudaSuccess && v > 0)
                d.smem_optin = v;
            if (cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev) == cudaSuccess && v > 0)
                d.smem_sm = v;
        }
        return d;
    }();
    return s;
}

// segment length / CTA count of the 1024-thread sort and scan kernels
struct SegPlan {
    uint32_t nblocks, seg;
};
static SegPlan seg_plan(uint64_t n) {
    const DeviceShape &d = device_shape();
    uint64_t nb = (n + 4095) / 4096;
    if (nb < 1) nb = 1;
    if (nb > (uint64_t)d.sms) nb = (uint64_t)d.sms;
    SegPlan s;
    s.nblocks = (uint32_t)nb;
    s.seg = (uint32_t)((n + nb - 1) / nb);
    return s;
}

constexpr uint32_t kMaxWindowTiles = 48 * 1024; // 192 KB of shared-memory counters per warp at most

struct BinPlan {
    uint32_t T_total, t_win, n_win, P, S, seg_len;
    size_t smem;
};
static BinPlan bin_plan(uint32_t C, uint32_t N, uint32_t tile_width, uint32_t tile_height) {
    const DeviceShape &d = device_shape();
    BinPlan b;
    b.T_total = C * tile_width * tile_height;
    b.t_win = b.T_total < kMaxWindowTiles ? b.T_total : kMaxWindowTiles;
    if (b.t_win == 0) b.t_win = 1;
    b.n_win = (b.T_total + b.t_win - 1) / b.t_win;
    if (b.n_win == 0) b.n_win = 1;
    b.smem = (size_t)((b.t_win + 31u) & ~31u) * 4 + 5 * kRing * 4;
    uint32_t per_sm = (uint32_t)((size_t)d.smem_sm / (b.smem + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 16) per_sm = 16;
    uint64_t P = (uint64_t)d.sms * per_sm;
    const uint64_t n = (uint64_t)C * N;
    const uint64_t cap = (n + 255) / 256; // no point in chunks of a handful of Gaussians
    if (P > cap) P = cap;
    if (P < 1) P = 1;
    b.P = (uint32_t)P;
    uint32_t S = b.T_total ? (65536u + b.T_total - 1) / b.T_total : 1u;
    if (S > 64) S = 64;
    if (S > b.P) S = b.P;
    if (S < 1) S = 1;
    b.seg_len = (b.P + S - 1) / S;
    b.S = (b.P + b.seg_len - 1) / b.seg_len;
    return b;
}

// Workspace of the plan; everything the emit needs afterwards lives here too.
struct PlanWs {
    size_t ctl, keys0, keys1, vals0, vals1, boxes, hist, bsum, rt_end, rt_idx, rt_box, rt_key, M, seg, toff, total;
};
static PlanWs plan_ws(uint64_t n, const BinPlan &b, const SegPlan &sp) {
    PlanWs w;
    size_t off = 0;
    auto take = [&](size_t bytes) { const size_t o = off; off += align256(bytes); return o; };
    w.ctl = take(sizeof(SortCtl));
    w.keys0 = take(n * 4); w.keys1 = take(n * 4);
    w.vals0 = take(n * 4); w.vals1 = take(n * 4);
    w.boxes = take(n * 8);
    w.hist = take((size_t)sp.nblocks * kRadixBins * 4);
    w.bsum = take((size_t)sp.nblocks * sizeof(RunAcc));
    w.rt_end = take(n * 4); w.rt_idx = take(n * 4); w.rt_box = take(n * 8); w.rt_key = take(n * 4);
    w.M = take((size_t)b.P * b.T_total * 4);
    w.seg = take((size_t)b.S * b.T_total * 4);
    w.toff = take(((size_t)b.T_total + 1) * 4);
    w.total = off + 256;
    return w;
}

template <typename KeyT>
static int radix_sort_launch(KeyT *k0, KeyT *k1, uint32_t *v0, uint32_t *v1, uint64_t n, const SegPlan &sp, int passes,
                             uint32_t begin_bit, uint32_t end_bit, const SortCtl *ctl, uint32_t *H, bool iota_first,
                             cudaStream_t s) {
    for (int pass = 0; pass < passes; ++pass) {
        const uint32_t shift = begin_bit + 8u * (uint32_t)pass;
        const uint32_t bits = end_bit - shift < 8u ? end_bit - shift : 8u;
        const uint32_t mask = (1u << bits) - 1u;
        // without a control block the host alternates the buffers itself
        KeyT *ks = (!ctl && (pass & 1)) ? k1 : k0, *kd = (!ctl && (pass & 1)) ? k0 : k1;
        uint32_t *vs = (!ctl && (pass & 1)) ? v1 : v0, *vd = (!ctl && (pass & 1)) ? v0 : v1;
        radix_count_kernel<KeyT><<<sp.nblocks, kSortThreads, 0, s>>>(ks, kd, n, sp.seg, pass, shift, mask, ctl, H);
        GSB_LAUNCH_CHECK();
        radix_scatter_kernel<KeyT><<<sp.nblocks, kSortThreads, 0, s>>>(ks, kd, vs, vd, n, sp.seg, sp.nblocks, pass, shift,
                                                                     mask, ctl, H, iota_first && pass == 0 ? 1 : 0);
        GSB_LAUNCH_CHECK();
    }
    return GSB_OK;
}

} // namespace gsb

extern "C" size_t gsb_isect_count_workspace(uint64_t n_elements) {
    const gsb::SegPlan sp = gsb::seg_plan(n_elements);
    return gsb::align256((size_t)sp.nblocks * sizeof(long long)) + 256;
}

extern "C" int gsb_isect_count(uint32_t C, uint32_t N, const float *means2d, const int32_t *radii,
                               uint32_t tile_size, uint32_t tile_width, uint32_t tile_height,
                               int32_t *tiles_per_gauss, int64_t *cum_tiles, void *workspace,
                               size_t workspace_bytes, gsb_stream_t stream) {
    using namespace gsb;
    const uint64_t n = (uint64_t)C * N;
    if (n == 0) return GSB_OK;
    if (!means2d || !radii || !tiles_per_gauss || !cum_tiles || tile_size == 0) return GSB_E_INVALID;
    if (!workspace || workspace_bytes < gsb_isect_count_workspace(n)) return GSB_E_WORKSPACE;
    cudaStream_t s = as_stream(stream);
    const uint32_t grid = (uint32_t)((n + kIsectThreads - 1) / kIsectThreads);
    ProfScope ps("isect_count", s); // count kernel + scan
    isect_count_kernel<<<grid, kIsectThreads, 0, s>>>(n, means2d, radii, tile_size, tile_width, tile_height,
                                                      tiles_per_gauss);
    GSB_LAUNCH_CHECK();
    const SegPlan sp = seg_plan(n);
    long long *bsum = reinterpret_cast<long long *>(workspace);
    scan_blocksum_kernel<<<sp.nblocks, kSortThreads, 0, s>>>(tiles_per_gauss, n, sp.seg, bsum);
    GSB_LAUNCH_CHECK();
    scan_apply_kernel<<<sp.nblocks, kSortThreads, 0, s>>>(tiles_per_gauss, n, sp.seg, bsum, cum_tiles);
    GSB_LAUNCH_CHECK();
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

// Generic stable sort of unsorted (isect_id, flatten_id) pairs on the low 32+tile_bits+cam_bits key bits
// (IntersectTile.cu:290-328): ceil(bits / 8) passes of the radix kernels of gsb_devsort.cuh.
extern "C" size_t gsb_isect_sort_workspace(uint64_t n_isects) {
    const gsb::SegPlan sp = gsb::seg_plan(n_isects);
    return gsb::align256(n_isects * 8) + gsb::align256(n_isects * 4) +
           gsb::align256((size_t)sp.nblocks * gsb::kRadixBins * 4) + 256;
}

extern "C" int gsb_isect_sort(uint64_t n_isects, uint32_t C, uint32_t tile_width, uint32_t tile_height,
                              const int64_t *isect_ids_in, const int32_t *flatten_ids_in,
                              int64_t *isect_ids_out, int32_t *flatten_ids_out, void *workspace,
                              size_t workspace_bytes, gsb_stream_t stream) {
    using namespace gsb;
    if (n_isects == 0) return GSB_OK;
    if (!isect_ids_in || !flatten_ids_in || !isect_ids_out || !flatten_ids_out) return GSB_E_INVALID;
    if (!workspace || (reinterpret_cast<uintptr_t>(workspace) & 255) || workspace_bytes < gsb_isect_sort_workspace(n_isects))
        return GSB_E_WORKSPACE;
    const uint32_t tile_n_bits = bit_width_u32(tile_width * tile_height);
    const uint32_t cam_n_bits = bit_width_u32(C);
    const uint32_t end_bit = 32 + tile_n_bits + cam_n_bits;
    if (end_bit > 64) return GSB_E_INVALID;
    const int passes = (int)((end_bit + 7) / 8);
    cudaStream_t s = as_stream(stream);
    ProfScope ps("isect_sort", s);
    const SegPlan sp = seg_plan(n_isects);
    char *base = reinterpret_cast<char *>(workspace);
    unsigned long long *tmp_k = reinterpret_cast<unsigned long long *>(base);
    uint32_t *tmp_v = reinterpret_cast<uint32_t *>(base + align256(n_isects * 8));
    uint32_t *H = reinterpret_cast<uint32_t *>(base + align256(n_isects * 8) + align256(n_isects * 4));
    unsigned long long *out_k = reinterpret_cast<unsigned long long *>(isect_ids_out);
    uint32_t *out_v = reinterpret_cast<uint32_t *>(flatten_ids_out);
    // The input is read-only: copy it into the buffer from which `passes` alternations end in the outputs.
    unsigned long long *k0 = (passes & 1) ? tmp_k : out_k, *k1 = (passes & 1) ? out_k : tmp_k;
    uint32_t *v0 = (passes & 1) ? tmp_v : out_v, *v1 = (passes & 1) ? out_v : tmp_v;
    GSB_CUDA_TRY(cudaMemcpyAsync(k0, isect_ids_in, n_isects * 8, cudaMemcpyDeviceToDevice, s));
    GSB_CUDA_TRY(cudaMemcpyAsync(v0, flatten_ids_in, n_isects * 4, cudaMemcpyDeviceToDevice, s));
    return radix_sort_launch<unsigned long long>(k0, k1, v0, v1, n_isects, sp, passes, 0, end_bit, nullptr, H, false, s);
}

extern "C" size_t gsb_isect_plan_workspace(uint32_t C, uint32_t N, uint32_t tile_width, uint32_t tile_height) {
    const uint64_t n = (uint64_t)C * N;
    return gsb::plan_ws(n, gsb::bin_plan(C, N, tile_width, tile_height), gsb::seg_plan(n)).total;
}

extern "C" int gsb_isect_plan(uint32_t C, uint32_t N, const float *means2d, const int32_t *radii, const float *depths,
                              uint32_t tile_size, uint32_t tile_width, uint32_t tile_height,
                              int32_t *tiles_per_gauss, int64_t *n_isects_out, int32_t *tile_offsets_out,
                              int tile_offsets_total, void *plan_workspace, size_t plan_workspace_bytes,
                              gsb_stream_t stream) {
    using namespace gsb;
    const uint64_t n = (uint64_t)C * N;
    if (!n_isects_out) return GSB_E_INVALID;
    cudaStream_t s = as_stream(stream);
    const uint64_t T64 = (uint64_t)C * tile_width * tile_height;
    if (n == 0 || T64 == 0) {
        GSB_CUDA_TRY(cudaMemsetAsync(n_isects_out, 0, sizeof(int64_t), s));
        if (tile_offsets_out && T64)
            GSB_CUDA_TRY(cudaMemsetAsync(tile_offsets_out, 0, (T64 + (tile_offsets_total ? 1 : 0)) * 4, s));
        if (n && tiles_per_gauss) GSB_CUDA_TRY(cudaMemsetAsync(tiles_per_gauss, 0, n * 4, s));
        return GSB_OK;
    }
    if (!means2d || !radii || !depths || !tiles_per_gauss || tile_size == 0) return GSB_E_INVALID;
    const uint32_t tile_n_bits = bit_width_u32(tile_width * tile_height);
    const uint32_t cam_n_bits = bit_width_u32(C);
    if (tile_n_bits + cam_n_bits > 32) return GSB_E_INVALID; // Intersect.cpp:50
    if (tile_width > 0xffffu || tile_height > 0xffffu || n > 0x7fffffffull || T64 > 0x7fffffffull) return GSB_E_INVALID;
    const BinPlan bp = bin_plan(C, N, tile_width, tile_height);
    const SegPlan sp = seg_plan(n);
    const PlanWs w = plan_ws(n, bp, sp);
    if (!plan_workspace || (reinterpret_cast<uintptr_t>(plan_workspace) & 255) || plan_workspace_bytes < w.total)
        return GSB_E_WORKSPACE;
    char *base = reinterpret_cast<char *>(plan_workspace);
    SortCtl *ctl = reinterpret_cast<SortCtl *>(base + w.ctl);
    uint32_t *k0 = reinterpret_cast<uint32_t *>(base + w.keys0), *k1 = reinterpret_cast<uint32_t *>(base + w.keys1);
    uint32_t *v0 = reinterpret_cast<uint32_t *>(base + w.vals0), *v1 = reinterpret_cast<uint32_t *>(base + w.vals1);
    uint2 *boxes = reinterpret_cast<uint2 *>(base + w.boxes);
    uint32_t *H = reinterpret_cast<uint32_t *>(base + w.hist);
    RunAcc *bsum = reinterpret_cast<RunAcc *>(base + w.bsum);
    RunTable rt{reinterpret_cast<uint32_t *>(base + w.rt_end), reinterpret_cast<uint32_t *>(base + w.rt_idx),
                reinterpret_cast<uint2 *>(base + w.rt_box), reinterpret_cast<uint32_t *>(base + w.rt_key)};
    uint32_t *M = reinterpret_cast<uint32_t *>(base + w.M);
    uint32_t *seg = reinterpret_cast<uint32_t *>(base + w.seg);
    {
        ProfScope ps("isect_count", s);
        GSB_CUDA_TRY(cudaMemsetAsync(ctl, 0, sizeof(SortCtl), s));
        const uint32_t grid = (uint32_t)((n + kIsectThreads - 1) / kIsectThreads);
        isect_plan_kernel<<<grid, kIsectThreads, 0, s>>>(n, means2d, radii, depths, tile_size, tile_width, tile_height,
                                                        tiles_per_gauss, k0, boxes, ctl);
        GSB_LAUNCH_CHECK();
    }
    {
        ProfScope ps("isect_depth_sort", s);
        if (int rc = radix_sort_launch<uint32_t>(k0, k1, v0, v1, n, sp, 4, 0, 32, ctl, H, true, s)) return rc;
        runs_blocksum_kernel<<<sp.nblocks, kSortThreads, 0, s>>>(ctl, v0, v1, tiles_per_gauss, n, sp.seg, bsum);
        GSB_LAUNCH_CHECK();
        runs_build_kernel<<<sp.nblocks, kSortThreads, 0, s>>>(ctl, v0, v1, k0, k1, tiles_per_gauss, boxes, n, sp.seg,
                                                             sp.nblocks, bsum, rt);
        GSB_LAUNCH_CHECK();
    }
    // device or pinned-host destination alike
    GSB_CUDA_TRY(cudaMemcpyAsync(n_isects_out, &ctl->n_isects, sizeof(int64_t), cudaMemcpyDefault, s));
    {
        ProfScope ps("isect_tile_hist", s);
        BinArgs a;
        a.rt = rt; a.ctl = ctl; a.N = N; a.n_tiles = tile_width * tile_height; a.tile_width = tile_width;
        a.tile_n_bits = tile_n_bits; a.multi_cam = C > 1 ? 1u : 0u;
        a.T_total = bp.T_total; a.P = bp.P; a.M = M; a.flatten_ids = nullptr; a.isect_ids = nullptr; a.cap = 0;
        if (bp.smem > 48 * 1024)
            GSB_CUDA_TRY(cudaFuncSetAttribute(tile_bin_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              (int)bp.smem));
        for (uint32_t wnd = 0; wnd < bp.n_win; ++wnd) {
            a.t_lo = wnd * bp.t_win;
            a.t_cnt = min(bp.t_win, bp.T_total - a.t_lo);
            tile_bin_kernel<false><<<bp.P, 32, bp.smem, s>>>(a);
            GSB_LAUNCH_CHECK();
        }
        const dim3 cgrid((bp.T_total + kIsectThreads - 1) / kIsectThreads, bp.S);
        col_segsum_kernel<<<cgrid, kIsectThreads, 0, s>>>(M, bp.T_total, bp.P, bp.seg_len, seg);
        GSB_LAUNCH_CHECK();
        col_tilescan_kernel<<<1, kSortThreads, 0, s>>>(seg, bp.T_total, bp.S, reinterpret_cast<uint32_t *>(base + w.toff),
                                                    tile_offsets_out, tile_offsets_total);
        GSB_LAUNCH_CHECK();
        col_apply_kernel<<<cgrid, kIsectThreads, 0, s>>>(M, bp.T_total, bp.P, bp.seg_len, seg);
        GSB_LAUNCH_CHECK();
    }
    return GSB_OK;
}

extern "C" int gsb_isect_emit_planned(uint32_t C, uint32_t N, const float *depths, uint32_t tile_width,
                                      uint32_t tile_height, uint64_t n_isects, const void *plan_workspace,
                                      size_t plan_workspace_bytes, int64_t *isect_ids /*nullable*/, int32_t *flatten_ids,
                                      gsb_stream_t stream) {
    using namespace gsb;
    const uint64_t n = (uint64_t)C * N;
    if (n == 0 || n_isects == 0) return GSB_OK;
    if (!plan_workspace || !flatten_ids || (isect_ids && !depths)) return GSB_E_INVALID;
    if (n_isects > 0x7fffffffull) return GSB_E_INVALID;
    const BinPlan bp = bin_plan(C, N, tile_width, tile_height);
    const SegPlan sp = seg_plan(n);
    const PlanWs w = plan_ws(n, bp, sp);
    if ((reinterpret_cast<uintptr_t>(plan_workspace) & 255) || plan_workspace_bytes < w.total) return GSB_E_WORKSPACE;
    cudaStream_t s = as_stream(stream);
    char *base = const_cast<char *>(reinterpret_cast<const char *>(plan_workspace));
    BinArgs a;
    a.rt = RunTable{reinterpret_cast<uint32_t *>(base + w.rt_end), reinterpret_cast<uint32_t *>(base + w.rt_idx),
                    reinterpret_cast<uint2 *>(base + w.rt_box), reinterpret_cast<uint32_t *>(base + w.rt_key)};
    a.ctl = reinterpret_cast<const SortCtl *>(base + w.ctl);
    a.N = N; a.n_tiles = tile_width * tile_height; a.tile_width = tile_width;
    a.tile_n_bits = bit_width_u32(tile_width * tile_height); a.multi_cam = C > 1 ? 1u : 0u;
    a.T_total = bp.T_total; a.P = bp.P; a.M = reinterpret_cast<uint32_t *>(base + w.M);
    a.flatten_ids = flatten_ids; a.isect_ids = nullptr; a.cap = (uint32_t)n_isects;
    ProfScope ps("isect_emit", s);
    if (bp.smem > 48 * 1024)
        GSB_CUDA_TRY(cudaFuncSetAttribute(tile_bin_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bp.smem));
    for (uint32_t wnd = 0; wnd < bp.n_win; ++wnd) {
        a.t_lo = wnd * bp.t_win;
        a.t_cnt = min(bp.t_win, bp.T_total - a.t_lo);
        tile_bin_kernel<true><<<bp.P, 32, bp.smem, s>>>(a);
        GSB_LAUNCH_CHECK();
    }
    if (isect_ids) { // only the operator API wants the 64-bit keys back (intersect_offset consumes them)
        const uint32_t grid = (uint32_t)((n_isects + kIsectThreads - 1) / kIsectThreads);
        isect_keys_kernel<<<grid, kIsectThreads, 0, s>>>((uint32_t)n_isects, reinterpret_cast<const uint32_t *>(base + w.toff),
                                                        bp.T_total, a.n_tiles, a.tile_n_bits, flatten_ids, depths, isect_ids);
        GSB_LAUNCH_CHECK();
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
