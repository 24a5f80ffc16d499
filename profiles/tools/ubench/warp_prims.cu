// Latency (cycles per dependent operation, one warp) of the warp-level primitives the intersect walk is built from.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a warp_prims.cu -o warp_prims && ./warp_prims
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(unsigned *out, long long *cyc, unsigned seed) {
    const unsigned lane = threadIdx.x;
    unsigned v = seed + lane * 2654435761u;
    __shared__ unsigned s[1024];
    for (int i = lane; i < 1024; i += 32) s[i] = (i * 7 + 3) & 1023;
    __syncwarp();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < 2048; ++it) {
        if (MODE == 0) v = __match_any_sync(0xffffffffu, v ^ lane) + lane;            // 32 distinct values
        if (MODE == 1) v = __match_any_sync(0xffffffffu, (v & 3u)) + it;               // 4 groups
        if (MODE == 2) v = __match_any_sync(0xffffffffu, 7u) ^ v;                      // one group
        if (MODE == 3) v = __reduce_or_sync(0xffffffffu, v) + lane;
        if (MODE == 4) v = __ballot_sync(0xffffffffu, v & 1u) + lane;
        if (MODE == 5) v = __shfl_sync(0xffffffffu, v, (lane + 1) & 31) + 1;
        if (MODE == 6) v = s[v & 1023];                                                // dependent LDS
        if (MODE == 7) v = (v | 1u) / ((lane & 7u) + 3u) + it;                         // u32 division by a variable
        if (MODE == 8) v = __popc(v) + v;
        if (MODE == 9) { unsigned b = s[v & 1023]; __syncwarp(); if (lane == (v & 31)) s[(v + lane) & 1023] = b + 1; __syncwarp(); v += b; }
    }
    const long long t1 = clock64();
    out[lane] = v;
    if (lane == 0) cyc[0] = t1 - t0;
}

int main() {
    unsigned *out; long long *cyc, h;
    cudaMalloc(&out, 128); cudaMalloc(&cyc, 8);
    const char *names[] = {"match_any 32 distinct", "match_any 4 groups", "match_any 1 group", "redux.or", "ballot",
                           "shfl", "dependent LDS", "u32 div by variable", "popc", "LDS+syncwarp+STS+syncwarp"};
#define RUN(M) k<M><<<1, 32>>>(out, cyc, 12345u); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); \
    printf("%-28s %7.1f cycles/iter\n", names[M], h / 2048.0);
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8) RUN(9)
    return 0;
}
