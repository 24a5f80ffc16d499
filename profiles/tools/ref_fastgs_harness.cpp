// Debug harness: calls the reference's fast_gs::rasterization::forward directly (oracle/_ref/libfastgs_ref.so) with
// cudaMalloc-backed allocators that print what is requested.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <tuple>
#include <vector>
#include <cmath>
namespace fast_gs::rasterization {
std::tuple<int, int, int, int, int> forward(std::function<char*(size_t)>, std::function<char*(size_t)>, std::function<char*(size_t)>,
    std::function<char*(size_t)>, const float3*, const float3*, const float4*, const float*, const float3*, const float3*,
    const float4*, const float3*, float*, float*, const int, const int, const int, const int, const int, const float, const float,
    const float, const float, const float, const float);
}
static std::function<char*(size_t)> mk(const char* name) {
    return [name](size_t n) { char* p = nullptr; cudaError_t e = cudaMalloc(&p, n ? n : 1); printf("  alloc %-14s %zu bytes -> %p (%s)\n", name, n, (void*)p, cudaGetErrorString(e)); fflush(stdout); return p; };
}
int main(int argc, char** argv) {
    const int W = argc > 1 ? atoi(argv[1]) : 1280, H = argc > 2 ? atoi(argv[2]) : 720, N = argc > 3 ? atoi(argv[3]) : 100000;
    std::vector<float> means(3 * N), sc(3 * N), rot(4 * N), op(N), sh0(3 * N), shN(45 * (size_t)N, 0.01f);
    srand(1);
    auto u = []() { return rand() / (float)RAND_MAX; };
    for (int i = 0; i < N; ++i) {
        means[3*i] = u()*3-1.5f; means[3*i+1] = u()*1.7f-0.85f; means[3*i+2] = u()*2+2;
        for (int k = 0; k < 3; ++k) sc[3*i+k] = logf(0.003f) + u()*(logf(0.02f)-logf(0.003f));
        rot[4*i] = 1; rot[4*i+1] = u()-0.5f; rot[4*i+2] = u()-0.5f; rot[4*i+3] = u()-0.5f;
        op[i] = u()*2-0.5f; sh0[3*i] = sh0[3*i+1] = sh0[3*i+2] = 0.3f;
    }
    float w2c[16] = {1,0,0,0, 0,1,0,0, 0,0,1,0, 0,0,0,1}, cam[3] = {0,0,0};
    auto up = [](const void* h, size_t b) { void* d; cudaMalloc(&d, b); cudaMemcpy(d, h, b, cudaMemcpyHostToDevice); return d; };
    float *image, *alpha; cudaMalloc(&image, 3ull*W*H*4); cudaMalloc(&alpha, 1ull*W*H*4);
    auto r = fast_gs::rasterization::forward(mk("per_primitive"), mk("per_tile"), mk("per_instance"), mk("per_bucket"),
        (const float3*)up(means.data(), means.size()*4), (const float3*)up(sc.data(), sc.size()*4), (const float4*)up(rot.data(), rot.size()*4),
        (const float*)up(op.data(), op.size()*4), (const float3*)up(sh0.data(), sh0.size()*4), (const float3*)up(shN.data(), shN.size()*4),
        (const float4*)up(w2c, 64), (const float3*)up(cam, 12), image, alpha, N, 16, 15, W, H, 1600.f*W/1920, 1600.f*W/1920, W/2.f, H/2.f, 0.01f, 1e10f);
    cudaError_t e = cudaDeviceSynchronize();
    printf("W %d H %d N %d -> n_visible %d n_instances %d n_buckets %d sel %d %d; sync: %s; last: %s\n", W, H, N, std::get<0>(r), std::get<1>(r),
           std::get<2>(r), std::get<3>(r), std::get<4>(r), cudaGetErrorString(e), cudaGetErrorString(cudaGetLastError()));
    return 0;
}
