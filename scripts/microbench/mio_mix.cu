// Microbenchmark: how MUFU.EX2, LDS.128 (broadcast) and SHFL share the MIO path on sm_100a.
// Each "position" = 4 EX2 + kLds LDS.128 + kShfl SHFL + 8 FFMA, 4 independent chains per thread.
#include <cstdio>
#include <cuda_runtime.h>

template <int kEx2, int kLds, int kShfl>
__global__ void k(float* out, int iters, float c) {
    __shared__ float4 s[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) s[i] = make_float4(c, 0.5f * c, 0.25f * c, 0.125f * c);
    __syncthreads();
    float h[4] = {0.f, 0.f, 0.f, 0.f}, acc = 0.f, x = -0.001f * threadIdx.x;
    const int q = threadIdx.x & 3;
    for (int it = 0; it < iters; ++it) {
        float4 b = make_float4(1.f, 1.f, 1.f, 1.f), cc = b;
        if (kLds >= 1) b = s[((it * 4) & 1020) + q];
        if (kLds >= 2) cc = s[((it * 4 + 512) & 1020) + q];
        if (kLds >= 3) { float4 d = s[((it * 4 + 256) & 1020) + (q ^ 1)]; x += d.x * 1e-9f; }
        const float bm[4] = {b.x, b.y, b.z, b.w}, cm[4] = {cc.x, cc.y, cc.z, cc.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float e = (i < kEx2) ? exp2f(x * bm[i]) : x * bm[i];
            h[i] = fmaf(e, h[i], bm[i]);
            acc = fmaf(cm[i], h[i], acc);
        }
#pragma unroll
        for (int j = 0; j < kShfl; ++j) acc += __shfl_xor_sync(0xffffffffu, acc, 1 + j);
        x -= 1e-7f;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc + h[0] + h[1] + h[2] + h[3];
}

template <int kEx2, int kLds, int kShfl>
void run(int blocks_per_sm, int threads) {
    float* out;
    int grid = 148 * blocks_per_sm, iters = 4000;
    cudaMalloc(&out, sizeof(float) * grid * threads);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<kEx2, kLds, kShfl><<<grid, threads>>>(out, 10, 0.5f);
    cudaEventRecord(e0);
    k<kEx2, kLds, kShfl><<<grid, threads>>>(out, iters, 0.5f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double warp_pos = (double)grid * threads / 32 * iters;
    printf("ex2=%d lds128=%d shfl=%d warps/SM=%2d : %.1f cycles per warp-position per SM (@1.965GHz)\n", kEx2, kLds, kShfl,
           blocks_per_sm * threads / 32, ms * 1e-3 * 1.965e9 / (warp_pos / 148));
    cudaFree(out);
}

int main() {
    run<4, 0, 0>(8, 128); run<4, 1, 0>(8, 128); run<4, 2, 0>(8, 128); run<4, 3, 0>(8, 128);
    run<0, 1, 0>(8, 128); run<0, 2, 0>(8, 128); run<0, 3, 0>(8, 128);
    run<4, 2, 1>(8, 128); run<4, 2, 2>(8, 128); run<0, 0, 2>(8, 128);
    run<2, 2, 0>(8, 128); run<4, 2, 0>(4, 128); run<4, 2, 0>(16, 128);
    return 0;
}
