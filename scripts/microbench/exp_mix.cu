// Microbenchmark: moving a fraction of the scan's decays exp2(dt * A) from MUFU.EX2 (16 lanes/clk/SM, on the MIO path
// it shares with LDS / SHFL) to a degree-5 polynomial on the FMA pipe (packed fp32x2), next to the shared-memory
// traffic of the forward walk.  Two positions per iteration = 4 state pairs; kPolyPairs of them use the polynomial.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 --use_fast_math -o exp_mix exp_mix.cu && ./exp_mix
#include <cmath>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float2 exp2_poly2(float2 x) {
    x.x = fmaxf(x.x, -126.f);
    x.y = fmaxf(x.y, -126.f);
    const float2 t = __fadd2_rn(x, make_float2(12582912.f, 12582912.f));          // 1.5 * 2^23: round(x) in the low mantissa bits
    const float2 r = __fadd2_rn(t, make_float2(-12582912.f, -12582912.f));
    const float2 f = __ffma2_rn(r, make_float2(-1.f, -1.f), x);                   // x - round(x), exact, in [-0.5, 0.5]
    float2 p = __ffma2_rn(f, make_float2(0.0013276472f, 0.0013276472f), make_float2(0.009675541f, 0.009675541f));
    p = __ffma2_rn(p, f, make_float2(0.05550713f, 0.05550713f));
    p = __ffma2_rn(p, f, make_float2(0.2402212f, 0.2402212f));
    p = __ffma2_rn(p, f, make_float2(0.69314694f, 0.69314694f));
    p = __ffma2_rn(p, f, make_float2(1.0000001f, 1.0000001f));
    return make_float2(__int_as_float(__float_as_int(p.x) + (__float_as_int(t.x) << 23)),
                       __int_as_float(__float_as_int(p.y) + (__float_as_int(t.y) << 23)));
}

__global__ void accuracy(const float* x, float* mufu, float* poly, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        mufu[i] = exp2f(x[i]);
        poly[i] = exp2_poly2(make_float2(x[i], x[i])).x;
    }
}

template <int kPolyPairs, int kLds>
__global__ void k(float* out, int iters, float c) {
    __shared__ float4 s[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) s[i] = make_float4(c, 0.5f * c, 0.25f * c, 0.125f * c);
    __syncthreads();
    float2 h[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
    float acc = 0.f, x = -0.001f * threadIdx.x;
    const int q = threadIdx.x & 3;
    const float2 A2[2] = {make_float2(-1.f, -2.f), make_float2(-3.f, -4.f)};
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int pos = 0; pos < 2; ++pos) {
            float4 b = make_float4(1.f, 1.f, 1.f, 1.f), cc = b;
            if (kLds >= 1) b = s[((it * 8 + pos * 4) & 1020) + q];
            if (kLds >= 2) cc = s[((it * 8 + pos * 4 + 512) & 1020) + q];
            const float2 bm[2] = {make_float2(b.x, b.y), make_float2(b.z, b.w)};
            const float2 cm[2] = {make_float2(cc.x, cc.y), make_float2(cc.z, cc.w)};
            const float2 dt2 = make_float2(x, x);
            float2 a2 = make_float2(0.f, 0.f);
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const float2 arg = __fmul2_rn(dt2, A2[i]);
                const float2 e = (pos * 2 + i < kPolyPairs) ? exp2_poly2(arg) : make_float2(exp2f(arg.x), exp2f(arg.y));
                h[i] = __ffma2_rn(e, h[i], __fmul2_rn(dt2, bm[i]));
                a2 = __ffma2_rn(cm[i], h[i], a2);
            }
            acc += a2.x + a2.y;
            x += 1e-7f;
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc + h[0].x + h[0].y + h[1].x + h[1].y;
}

template <int kPolyPairs, int kLds>
void run(int blocks_per_sm, int threads) {
    float* out;
    int grid = 148 * blocks_per_sm, iters = 2000;
    cudaMalloc(&out, sizeof(float) * grid * threads);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<kPolyPairs, kLds><<<grid, threads>>>(out, 10, 0.5f);
    cudaEventRecord(e0);
    k<kPolyPairs, kLds><<<grid, threads>>>(out, iters, 0.5f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double warp_pos = (double)grid * threads / 32 * iters * 2;
    printf("poly %d of 8 exps, lds128 per position=%d, warps/SM=%2d : %.2f cycles per warp-position per SM (@1.965GHz)\n",
           kPolyPairs * 2, kLds, blocks_per_sm * threads / 32, ms * 1e-3 * 1.965e9 / (warp_pos / 148));
    cudaFree(out);
}

int main() {
    // accuracy of both evaluations against exp2 in double, x in [-130, 2]
    const int n = 1 << 22;
    float *x, *m, *p;
    cudaMallocManaged(&x, n * 4); cudaMallocManaged(&m, n * 4); cudaMallocManaged(&p, n * 4);
    for (int i = 0; i < n; ++i) x[i] = -130.f + 132.f * (float)i / (float)n + 1e-4f * (float)(i % 97);
    accuracy<<<n / 256, 256>>>(x, m, p, n);
    cudaDeviceSynchronize();
    double em = 0, ep = 0, ep_small = 0;
    for (int i = 0; i < n; ++i) {
        const double ref = exp2((double)x[i]);
        if (x[i] > -125.9f) {
            em = fmax(em, fabs(m[i] - ref) / ref);
            ep = fmax(ep, fabs(p[i] - ref) / ref);
        } else {
            ep_small = fmax(ep_small, fabs((double)p[i]));
        }
    }
    printf("max relative error on [-125.9, 2]: MUFU.EX2 %.3e   polynomial %.3e   (largest polynomial value below -125.9: %.3e)\n", em, ep, ep_small);
    run<0, 2>(8, 128); run<1, 2>(8, 128); run<2, 2>(8, 128); run<3, 2>(8, 128); run<4, 2>(8, 128);
    run<0, 1>(8, 128); run<2, 1>(8, 128); run<3, 1>(8, 128); run<4, 1>(8, 128);
    run<0, 2>(6, 128); run<2, 2>(6, 128); run<3, 2>(6, 128);
    return 0;
}
