// dwconv3d.cuh -- depthwise 3x3x3 convolution over (frame, y, x), forward / backward, for sm_100a.
//
// The Mlp of every Temporal Mamba block of the reference runs nn.Conv3d(C, C, 3, 1, 1, groups=C) on its
// tokens (modeling/vivim.py:57-68, 99-106): (B, N, C) -> transpose -> (B, C, nf, H, W) -> cuDNN -> flatten ->
// transpose.  cuDNN serves the depthwise 3-D case with per-group convolveNd engines: in a Vivim training step
// (batch 3, 256x256, clip 5) its dgrad + wgrad kernels are 10752 launches and 90 % of the GPU time
// (profiles/r01_vivim_step.md).  The op is a 27-tap stencil with no reuse across channels, i.e. bandwidth
// bound: 2 tensors in/out forward, 3 + a (C, 27) reduction backward.
//
// Design (B200-first): work directly on the TOKEN layout (B, nf, H, W, C), channels innermost, so neither
// transpose exists.  One thread owns 8 channels (one 128-bit access) of kDwX consecutive x positions of one
// (b, frame, y) row and slides a 3-wide window along x: 9 rows x (kDwX + 2) vector loads for kDwX outputs; a
// warp covers 256 contiguous channels (512 B per request).  Neighbouring rows / frames are re-read through
// L1 / L2 (the whole activation is < 32 MB, L2 is 126 MB).  The weights of the thread's 8 channels are read
// through the read-only path (27 x 32 B per thread, L1 resident).
//   forward   out = bias + sum_tap w[c, tap] x[p + off(tap)]
//   backward  dx  = sum_tap w[c, tap] dout[p - off(tap)]          (same kernel, mirrored taps, no bias)
//             dw[c, tap] = sum_p dout[p] x[p + off(tap)],  db[c] = sum_p dout[p]
//   the weight gradient keeps 27 x 2 fp32 accumulators per thread (2 channels, 32-bit accesses, a warp = 64
//   contiguous channels), strides over the positions, reduces the 8 position slots of a CTA in shared memory
//   and issues one fp32 atomicAdd per (CTA, channel, tap).
#pragma once

#include "../../include/vivim_b200.h"
#include "common.cuh"

namespace vv {

constexpr int kDwX = 4;            // x positions per thread (sliding window)
constexpr int kDwThreads = 256;

struct DwGeom {
    int B, T, H, W, C;
};

// 8 channels of one position (zeros outside the volume)
template <typename T, bool kVec>
__device__ __forceinline__ void dw_load(const T* __restrict__ base, const DwGeom& g, int b, int t, int y, int x, int c0,
                                        float (&v)[8]) {
    if ((unsigned)t < (unsigned)g.T && (unsigned)y < (unsigned)g.H && (unsigned)x < (unsigned)g.W) {
        const T* p = base + ((((int64_t)b * g.T + t) * g.H + y) * g.W + x) * g.C + c0;
        if (kVec) {
            load8_vec<T>(p, v);
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = (c0 + i < g.C) ? to_f32<T>(p[i]) : 0.f;
        }
    } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = 0.f;
    }
}

// kMirror = false: forward (taps as stored, + bias).  kMirror = true: input gradient (taps mirrored, no bias).
// weight: fp32 (C, 27), tap = (dt * 3 + dy) * 3 + dx.
template <typename T, bool kVec, bool kMirror>
__global__ void __launch_bounds__(kDwThreads) dwconv3d_kernel(const T* __restrict__ in, const float* __restrict__ weight,
                                                              const float* __restrict__ bias, T* __restrict__ out,
                                                              const DwGeom g) {
    const int cvecs = (g.C + 7) / 8;
    const int xt = (g.W + kDwX - 1) / kDwX;
    const int64_t total = (int64_t)g.B * g.T * g.H * xt * cvecs;
    const int64_t idx = (int64_t)blockIdx.x * kDwThreads + threadIdx.x;
    if (idx >= total) return;
    const int cv = (int)(idx % cvecs);
    int64_t r = idx / cvecs;
    const int xb = (int)(r % xt); r /= xt;
    const int y = (int)(r % g.H); r /= g.H;
    const int t = (int)(r % g.T);
    const int b = (int)(r / g.T);
    const int c0 = cv * 8, x0 = xb * kDwX;

    float acc[kDwX][8];
#pragma unroll
    for (int j = 0; j < kDwX; ++j)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[j][i] = (!kMirror && bias && c0 + i < g.C) ? __ldg(bias + c0 + i) : 0.f;

#pragma unroll 1
    for (int row = 0; row < 9; ++row) {
        const int dt = row / 3, dy = row - dt * 3;
        float win[kDwX + 2][8];
#pragma unroll
        for (int j = 0; j < kDwX + 2; ++j) dw_load<T, kVec>(in, g, b, t + dt - 1, y + dy - 1, x0 + j - 1, c0, win[j]);
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
            const int tap = kMirror ? 26 - (row * 3 + dx) : row * 3 + dx;
            float w[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) w[i] = (c0 + i < g.C) ? __ldg(weight + (int64_t)(c0 + i) * 27 + tap) : 0.f;
#pragma unroll
            for (int j = 0; j < kDwX; ++j)
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[j][i] = fmaf(w[i], win[j + dx][i], acc[j][i]);
        }
    }
#pragma unroll
    for (int j = 0; j < kDwX; ++j) {
        const int x = x0 + j;
        if (x < g.W) {
            T* p = out + ((((int64_t)b * g.T + t) * g.H + y) * g.W + x) * g.C + c0;
            if (kVec) {
                store8_vec<T>(p, acc[j]);
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    if (c0 + i < g.C) p[i] = from_f32<T>(acc[j][i]);
            }
        }
    }
}

// weight / bias gradient.  grid (ceil(C / 64), position blocks); block (32 lanes = 64 channels, 8 position slots).
constexpr int kDwSlots = 8;

template <typename T>
__device__ __forceinline__ float2 dw_load2(const T* __restrict__ p, int c, int C) {
    if (sizeof(T) == 4 || c + 1 >= C)   // fp32: the pair need not be 8-byte aligned when C is odd
        return make_float2(c < C ? to_f32<T>(p[0]) : 0.f, c + 1 < C ? to_f32<T>(p[1]) : 0.f);
    return load_pair<T>(p);
}

template <typename T>
__global__ void __launch_bounds__(32 * kDwSlots) dwconv3d_wgrad_kernel(const T* __restrict__ x, const T* __restrict__ dout,
                                                                       float* __restrict__ dweight, float* __restrict__ dbias,
                                                                       const DwGeom g) {
    __shared__ float2 red[kDwSlots][32];
    const int lane = threadIdx.x, slot = threadIdx.y;
    const int c = (blockIdx.x * 32 + lane) * 2;
    const bool live = c < g.C;
    const int64_t npos = (int64_t)g.B * g.T * g.H * g.W;
    float2 acc[27], accb = make_float2(0.f, 0.f);
#pragma unroll
    for (int k = 0; k < 27; ++k) acc[k] = make_float2(0.f, 0.f);
    if (live) {
        for (int64_t p = (int64_t)blockIdx.y * kDwSlots + slot; p < npos; p += (int64_t)gridDim.y * kDwSlots) {
            int64_t r = p;
            const int xx = (int)(r % g.W); r /= g.W;
            const int y = (int)(r % g.H); r /= g.H;
            const int t = (int)(r % g.T);
            const float2 go = dw_load2<T>(dout + p * g.C + c, c, g.C);
            accb.x += go.x;
            accb.y += go.y;
#pragma unroll
            for (int dt = 0; dt < 3; ++dt)
#pragma unroll
                for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                    for (int dx = 0; dx < 3; ++dx) {
                        const int tt = t + dt - 1, yy = y + dy - 1, xq = xx + dx - 1;
                        if ((unsigned)tt < (unsigned)g.T && (unsigned)yy < (unsigned)g.H && (unsigned)xq < (unsigned)g.W) {
                            const int64_t q = p + ((int64_t)(dt - 1) * g.H + (dy - 1)) * g.W + (dx - 1);
                            const float2 xv = dw_load2<T>(x + q * g.C + c, c, g.C);
                            const int k = (dt * 3 + dy) * 3 + dx;
                            acc[k].x = fmaf(go.x, xv.x, acc[k].x);
                            acc[k].y = fmaf(go.y, xv.y, acc[k].y);
                        }
                    }
        }
    }
    // reduce the position slots of the CTA, one tap at a time (fully unrolled: acc[] stays in registers)
#pragma unroll
    for (int k = 0; k < 28; ++k) {
        red[slot][lane] = k < 27 ? acc[k < 27 ? k : 0] : accb;
        __syncthreads();
        if (slot == 0 && live) {
            float2 s = red[0][lane];
#pragma unroll
            for (int j = 1; j < kDwSlots; ++j) {
                s.x += red[j][lane].x;
                s.y += red[j][lane].y;
            }
            if (k < 27) {
                atomicAdd(dweight + (int64_t)c * 27 + k, s.x);
                if (c + 1 < g.C) atomicAdd(dweight + (int64_t)(c + 1) * 27 + k, s.y);
            } else if (dbias) {
                atomicAdd(dbias + c, s.x);
                if (c + 1 < g.C) atomicAdd(dbias + c + 1, s.y);
            }
        }
        __syncthreads();
    }
}

}  // namespace vv
