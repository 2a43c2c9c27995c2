// Microbenchmark: packed fp32 FMA (fma.rn.f32x2 -> FFMA2) issue rate on sm_100a, alone and next to
// scalar FFMA and MUFU.EX2.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_rate ffma2_rate.cu && ./ffma2_rate
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b),
                       rc = *reinterpret_cast<unsigned long long*>(&c), rd;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    return *reinterpret_cast<float2*>(&rd);
}

// kMode 0: 8 independent scalar FFMA chains x kPer;  1: 8 independent FFMA2 chains x kPer;
// kEx2: additionally one MUFU.EX2 per chain per iteration
template <int kMode, int kPer, int kEx2>
__global__ void k(float* out, int iters, float c) {
    float2 x[8], m = make_float2(c, c * 0.5f), b = make_float2(1e-3f, 2e-3f);
    float e[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { x[i] = make_float2(0.001f * (threadIdx.x + i), 0.5f); e[i] = -0.01f * i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (kEx2) {
                asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(e[i]));
                e[i] -= 1.0f;   // 1 FADD keeps the argument bounded
            }
#pragma unroll
            for (int f = 0; f < kPer; ++f) {
                if (kMode == 0) {
                    asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[i].x) : "f"(m.x), "f"(b.x));
                } else {
                    x[i] = ffma2(x[i], m, b);
                }
            }
        }
    }
    float r = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) r += x[i].x + x[i].y + e[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int kMode, int kPer, int kEx2>
void run(const char* name, int blocks_per_sm, int threads) {
    float* out;
    int grid = 148 * blocks_per_sm, iters = 2000;
    cudaMalloc(&out, sizeof(float) * grid * threads);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<kMode, kPer, kEx2><<<grid, threads>>>(out, 10, 0.5f);
    cudaEventRecord(e0);
    k<kMode, kPer, kEx2><<<grid, threads>>>(out, iters, 0.5f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double clk = ms * 1e-3 * 1.965e9;
    const double winst = (double)grid * threads / 32 * iters * 8 * kPer / 148;   // FMA-type warp instructions per SM
    printf("%-34s warps/SM=%2d  %.3f ms  FMA-inst/clk/SM=%.2f  fp32 FMA lanes/clk/SM=%.1f  cycles/iter/warp-slot=%.2f\n", name,
           blocks_per_sm * threads / 32, ms, winst / clk, winst / clk * 32 * (kMode ? 2 : 1),
           clk / iters / 8);
    cudaFree(out);
}

int main() {
    run<0, 4, 0>("FFMA x4", 4, 256);
    run<1, 4, 0>("FFMA2 x4", 4, 256);
    run<1, 4, 0>("FFMA2 x4", 2, 128);
    run<0, 4, 1>("EX2 + FFMA x4", 4, 256);
    run<1, 4, 1>("EX2 + FFMA2 x4", 4, 256);
    run<0, 8, 1>("EX2 + FFMA x8", 4, 256);
    run<1, 8, 1>("EX2 + FFMA2 x8", 4, 256);
    run<0, 14, 1>("EX2 + FFMA x14", 4, 256);
    run<1, 7, 1>("EX2 + FFMA2 x7 (=14 fp32 FMA)", 4, 256);
    run<1, 14, 1>("EX2 + FFMA2 x14", 4, 256);
    return 0;
}
