// Microbenchmark: issue rate of FFMA vs FFMA2 (fma.rn.f32x2), alone and interleaved with SHFL / SEL.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d;
}
template <int MODE> __global__ void k(float *out, int iters, float s) {
    float a[8]; uint64_t p[8];
    for (int i = 0; i < 8; ++i) { a[i] = threadIdx.x * 0.001f + i; float2 t = make_float2(a[i], a[i] + 1); p[i] = *reinterpret_cast<uint64_t *>(&t); }
    float2 s2 = make_float2(s, s); uint64_t sp = *reinterpret_cast<uint64_t *>(&s2);
    float sh = threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) a[i] = __fmaf_rn(a[i], s, s);
            if (MODE == 1) p[i] = ffma2(p[i], sp, sp);
            if (MODE == 2) { a[i] = __fmaf_rn(a[i], s, s); sh += __shfl_xor_sync(0xffffffffu, sh, 1 + (i & 3)); }
            if (MODE == 3) { p[i] = ffma2(p[i], sp, sp); sh += __shfl_xor_sync(0xffffffffu, sh, 1 + (i & 3)); }
            if (MODE == 4) { a[i] = __fmaf_rn(a[i], s, s); asm volatile("" ::: "memory"); sh = (__float_as_int(sh) & (1 << i)) ? sh + 1.f : sh * 0.5f; }
            if (MODE == 5) { p[i] = ffma2(p[i], sp, sp); sh = (__float_as_int(sh) & (1 << i)) ? sh + 1.f : sh * 0.5f; }
        }
    }
    float r = sh;
    for (int i = 0; i < 8; ++i) { float2 t = *reinterpret_cast<float2 *>(&p[i]); r += a[i] + t.x + t.y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int MODE> void run(const char *name, float *out) {
    const int iters = 4096, blocks = 148 * 4, threads = 256;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, threads>>>(out, 16, 0.999f);
    cudaEventRecord(e0); k<MODE><<<blocks, threads>>>(out, iters, 0.999f); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double winst = (double)blocks * threads / 32 * iters * 8;
    printf("%-24s %.3f ms  %.1f G loop-bodies/s (warp-level)\n", name, ms, winst / ms * 1e-6);
}
int main() {
    float *out; cudaMalloc(&out, 148 * 4 * 256 * 4);
    run<0>("FFMA", out); run<1>("FFMA2", out); run<2>("FFMA+SHFL+FADD", out); run<3>("FFMA2+SHFL+FADD", out);
    run<4>("FFMA+alu", out); run<5>("FFMA2+alu", out);
    return 0;
}
