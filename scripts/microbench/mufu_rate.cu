// Microbenchmark: sustained MUFU.EX2 rate on sm_100a, alone and mixed with FFMA / LDS.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 --use_fast_math -o mufu_rate mufu_rate.cu && ./mufu_rate
#include <cstdio>
#include <cuda_runtime.h>

template <int kFma, bool kLds>
__global__ void k(float* out, int iters, float c) {
    __shared__ float4 s[256];
    s[threadIdx.x & 255] = make_float4(c, c, c, c);
    __syncthreads();
    float x[8], a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { x[i] = -0.001f * (threadIdx.x + i); a[i] = 0.f; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float e = exp2f(x[i] * c);
            float m = 1.0f;
            if (kLds) m = s[(it + i) & 255].x;
#pragma unroll
            for (int f = 0; f < kFma; ++f) a[i] = fmaf(e, a[i], m + f);
            x[i] = kFma ? x[i] - 1e-6f : e;
        }
    }
    float r = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) r += x[i] + a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int kFma, bool kLds>
void run(const char* name, int blocks_per_sm, int threads) {
    float* out;
    int grid = 148 * blocks_per_sm, iters = 2000;
    cudaMalloc(&out, sizeof(float) * grid * threads);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<kFma, kLds><<<grid, threads>>>(out, 10, 0.5f);
    cudaEventRecord(e0);
    k<kFma, kLds><<<grid, threads>>>(out, iters, 0.5f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double ex2 = (double)grid * threads * iters * 8;
    printf("%-28s warps/SM=%2d  %.3f ms  EX2 lanes/clk/SM @1.965GHz = %.2f\n", name, blocks_per_sm * threads / 32, ms,
           ex2 / (ms * 1e-3) / 148 / 1.965e9);
    cudaFree(out);
}

int main() {
    run<0, false>("ex2 only", 4, 256);
    run<0, false>("ex2 only", 1, 256);
    run<0, false>("ex2 only", 1, 128);
    run<1, false>("ex2 + 1 ffma", 4, 256);
    run<4, false>("ex2 + 4 ffma", 4, 256);
    run<4, false>("ex2 + 4 ffma", 1, 256);
    run<4, true>("ex2 + 4 ffma + lds", 4, 256);
    run<4, true>("ex2 + 4 ffma + lds", 2, 128);
    run<6, true>("ex2 + 6 ffma + lds", 4, 256);
    return 0;
}
