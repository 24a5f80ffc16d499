// gsb_ewa.cuh -- the exact tile test of the EWA ("fastgs") rasterizer, SURVEY.md 8 f4, shared by the per-primitive
// front kernel (which counts a primitive's tiles, gsb_fastgs.cu) and the intersect stage (which walks them twice more,
// gsb_intersect.cu).  All three evaluations must agree bit for bit on every (primitive, tile) pair -- the slot counts are
// exact -- so every floating-point operation below is an explicitly rounded intrinsic: the result does not depend on the
// translation unit's contraction / fast-math flags.
//
// The test (reference: fastgs/rasterization/include/kernel_utils.cuh:105-139, after StopThePop): the primitive
// contributes to a tile iff the smallest value of sigma/2 = 1/2 (ca dx^2 + cc dy^2) + cb dx dy it finds on the tile's
// rectangle of pixel centres is at most ln(255 opacity).  The candidate point is the rectangle corner nearest to the
// centre, moved along the two edges that leave that corner by the clamped 1-D minimisers of the quadratic.
#pragma once

#include "gsb_common.cuh"

namespace gsb {

// centre shifted by -0.5 px (pixel centres become integers), conic, power threshold ln(255 opacity)
struct EwaFilter {
    float mx, my, ca, cb, cc, thr;
};

__device__ __forceinline__ bool ewa_tile_contributes(const EwaFilter &f, uint32_t tx, uint32_t ty) {
    const float lo_x = (float)(tx * 16u), lo_y = (float)(ty * 16u);
    const float hi_x = lo_x + 15.0f, hi_y = lo_y + 15.0f; // exact: small integers
    const bool left = lo_x > f.mx, above = lo_y > f.my;
    const bool out_x = left || f.mx > hi_x, out_y = above || f.my > hi_y;
    if (!out_x && !out_y) return true;
    const float cx = left ? lo_x : hi_x, cy = above ? lo_y : hi_y;
    const float dfx = __fsub_rn(f.mx, cx), dfy = __fsub_rn(f.my, cy);
    const float ex = (lo_x >= f.mx) ? 15.0f : -15.0f, ey = (lo_y >= f.my) ? 15.0f : -15.0f;
    // 1-D minimisers along the two edges through the corner, clamped to the edge; an edge is only walked when the
    // centre lies outside the rectangle's range in the OTHER coordinate
    const float sx = __fdiv_rn(__fmaf_rn(f.ca, dfx, __fmul_rn(f.cb, dfy)), __fmul_rn(f.ca, ex));
    const float sy = __fdiv_rn(__fmaf_rn(f.cb, dfx, __fmul_rn(f.cc, dfy)), __fmul_rn(f.cc, ey));
    const float tx_ = out_y ? __saturatef(sx) : 0.0f, ty_ = out_x ? __saturatef(sy) : 0.0f;
    const float dx = __fsub_rn(f.mx, __fmaf_rn(tx_, ex, cx)), dy = __fsub_rn(f.my, __fmaf_rn(ty_, ey, cy));
    const float q = __fmaf_rn(f.ca, __fmul_rn(dx, dx), __fmul_rn(f.cc, __fmul_rn(dy, dy)));
    const float power = __fmaf_rn(0.5f, q, __fmul_rn(f.cb, __fmul_rn(dx, dy)));
    return power <= f.thr;
}

} // namespace gsb
