// conv1d.cuh -- causal depthwise conv1d forward / backward for sm_100a.
//
// Replaces the channel-first kernels of the reference (causal-conv1d/csrc/causal_conv1d_fwd.cu:39-130,
// causal_conv1d_bwd.cu:46-240).  Design (B200-first, not a port):
//   * pure streaming op, HBM-bound: 2T bytes fwd, 3T bytes bwd (T = B*D*L*sizeof(io)).
//   * one thread owns 2 chunks of 8 consecutive positions (1024 positions apart) = two 128-bit accesses per
//     tensor, all issued before the first use (memory-level parallelism is what bounds a streaming kernel
//     this short); a warp touches 512 contiguous bytes per request.  No shared-memory staging on the data path:
//     the (K-1)-element halo travels lane-to-lane by warp shuffle, only lane 0 / lane 31 of a
//     warp fetch their halo from global memory (L1/L2 hits: the neighbouring warp streams it).
//   * the grid is (rows = B*D, tiles of 2048 positions): every CTA is independent, there is no
//     serial chunk loop, so even the B=1 stage-1 shape (128 rows x 20480) gives 1280 CTAs.
//   * K in {2,3,4} is served by one 4-tap code path (leading taps zero-padded).
//   * dweight/dbias: per-thread partial sums -> warp shuffle reduce -> 4-warp shared-memory
//     reduce -> one fp32 atomicAdd per (CTA, tap); accumulators are zeroed by the caller exactly
//     like the reference's torch::zeros (causal_conv1d.cpp:247-249).
#pragma once

#include "../../include/vivim_b200.h"
#include "common.cuh"

namespace vv {

constexpr int kConvThreads = 128;
constexpr int kConvChunks = 2;                        // 8-position chunks per thread: 2 x 16 B of every tensor in flight
constexpr int kConvSpan = kConvThreads * kVecElems;   // 1024 positions: one chunk of every thread of the CTA
constexpr int kConvTile = kConvSpan * kConvChunks;    // positions per CTA
constexpr int kTaps = 4;

__device__ __forceinline__ float load_weight(const void* p, int dtype, int idx) {
    if (dtype == VV_F32) return reinterpret_cast<const float*>(p)[idx];
    if (dtype == VV_F16) return __half2float(reinterpret_cast<const __half*>(p)[idx]);
    return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[idx]);
}

// taps[j] multiplies x[t - 3 + j]; bias in *bias_out
__device__ __forceinline__ void load_taps(const vv_conv1d_args& a, int d, float (&taps)[kTaps], float& bias) {
    const int K = a.width;
#pragma unroll
    for (int j = 0; j < kTaps; ++j) {
        const int k = j - (kTaps - K);
        taps[j] = k >= 0 ? load_weight(a.weight, a.w_dtype, d * K + k) : 0.f;
    }
    bias = a.bias ? load_weight(a.bias, a.w_dtype, d) : 0.f;
}

// Halo of the 3 positions before t0 (chronological order), taken from the lane below; lane 0
// reads global memory.  `v` are this thread's own 8 positions.
template <typename T>
__device__ __forceinline__ void halo_before(const T* __restrict__ row, int t0, int L, const float (&v)[8],
                                            float (&h)[kTaps - 1]) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int j = 0; j < kTaps - 1; ++j) h[j] = __shfl_up_sync(0xffffffffu, v[8 - (kTaps - 1) + j], 1);
    if (lane == 0) {
#pragma unroll
        for (int j = 0; j < kTaps - 1; ++j) {
            const int t = t0 - (kTaps - 1) + j;
            h[j] = (t >= 0 && t < L) ? to_f32<T>(row[t]) : 0.f;
        }
    }
}

template <typename T, bool kSilu, bool kVec>
__global__ void __launch_bounds__(kConvThreads) conv1d_fwd_kernel(const vv_conv1d_args a) {
    const int row = blockIdx.x;
    const int b = row / a.dim, d = row - b * a.dim;
    const int L = a.seqlen;
    const int tb = blockIdx.y * kConvTile + threadIdx.x * kVecElems;
    const T* __restrict__ x = reinterpret_cast<const T*>(a.x) + b * a.x_bs + d * a.x_ds;
    T* __restrict__ out = reinterpret_cast<T*>(a.out) + b * a.out_bs + d * a.out_ds;

    // all loads of the thread first (kConvChunks x 16 B in flight), then the arithmetic
    Raw8<T, kVec> rx[kConvChunks];
#pragma unroll
    for (int c = 0; c < kConvChunks; ++c) rx[c].load(x, tb + c * kConvSpan, L);
    float taps[kTaps], bias;
    load_taps(a, d, taps, bias);
#pragma unroll
    for (int c = 0; c < kConvChunks; ++c) {
        const int t0 = tb + c * kConvSpan;
        float xx[kTaps - 1 + 8];
        {
            float v[8], h[kTaps - 1];
            rx[c].unpack(v);
            halo_before<T>(x, t0, L, v, h);
#pragma unroll
            for (int j = 0; j < kTaps - 1; ++j) xx[j] = h[j];
#pragma unroll
            for (int i = 0; i < 8; ++i) xx[kTaps - 1 + i] = v[i];
        }
        float o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float acc = bias;
#pragma unroll
            for (int j = 0; j < kTaps; ++j) acc = fmaf(taps[j], xx[i + j], acc);
            o[i] = kSilu ? acc * sigmoid_f(acc) : acc;
        }
        store8<T, kVec>(out, t0, L, o);
    }
}

// d(pre-activation) at position t, recomputed from global memory (used for the 3-position halo to
// the right of a warp, where no lane holds it).
template <typename T, bool kSilu>
__device__ __forceinline__ float dpre_at(const T* __restrict__ x, const T* __restrict__ dout, int t, int L,
                                         const float (&taps)[kTaps], float bias) {
    if (t >= L) return 0.f;
    float g = to_f32<T>(dout[t]);
    if (kSilu) {
        float pre = bias;
#pragma unroll
        for (int j = 0; j < kTaps; ++j) {
            const int s = t - (kTaps - 1) + j;
            if (s >= 0) pre = fmaf(taps[j], to_f32<T>(x[s]), pre);
        }
        const float sg = sigmoid_f(pre);
        g *= sg * (1.f + pre * (1.f - sg));
    }
    return g;
}

template <typename T, bool kSilu, bool kVec>
__global__ void __launch_bounds__(kConvThreads) conv1d_bwd_kernel(const vv_conv1d_args a) {
    __shared__ float s_red[kConvThreads / 32][kTaps + 1];
    __shared__ float s_halo[kConvChunks][kConvThreads / 32][kTaps - 1];
    const int row = blockIdx.x;
    const int b = row / a.dim, d = row - b * a.dim;
    const int L = a.seqlen;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tb = blockIdx.y * kConvTile + threadIdx.x * kVecElems;
    const T* __restrict__ x = reinterpret_cast<const T*>(a.x) + b * a.x_bs + d * a.x_ds;
    const T* __restrict__ dout = reinterpret_cast<const T*>(a.dout) + b * a.dout_bs + d * a.dout_ds;
    T* __restrict__ dx = reinterpret_cast<T*>(a.dx) + b * a.dx_bs + d * a.dx_ds;

    // all loads of the thread first (2 tensors x kConvChunks x 16 B in flight), then the arithmetic
    Raw8<T, kVec> rx[kConvChunks], rg[kConvChunks];
#pragma unroll
    for (int c = 0; c < kConvChunks; ++c) {
        rx[c].load(x, tb + c * kConvSpan, L);
        rg[c].load(dout, tb + c * kConvSpan, L);
    }
    float taps[kTaps], bias;
    load_taps(a, d, taps, bias);

    // pass 1: own positions -- x window and d(pre-activation) of every chunk
    float xx[kConvChunks][kTaps - 1 + 8], dd[kConvChunks][8 + kTaps - 1];
#pragma unroll
    for (int c = 0; c < kConvChunks; ++c) {
        const int t0 = tb + c * kConvSpan;
        {
            float v[8], h[kTaps - 1];
            rx[c].unpack(v);
            halo_before<T>(x, t0, L, v, h);
#pragma unroll
            for (int j = 0; j < kTaps - 1; ++j) xx[c][j] = h[j];
#pragma unroll
            for (int i = 0; i < 8; ++i) xx[c][kTaps - 1 + i] = v[i];
        }
        float g[8];
        rg[c].unpack(g);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (kSilu) {
                float pre = bias;
#pragma unroll
                for (int j = 0; j < kTaps; ++j) pre = fmaf(taps[j], xx[c][i + j], pre);
                const float sg = sigmoid_f(pre);
                dd[c][i] = g[i] * sg * (1.f + pre * (1.f - sg));
            } else {
                dd[c][i] = g[i];
            }
        }
        // the first 3 values of every warp are the right halo of the warp before it: through shared memory
        if (lane == 0) {
#pragma unroll
            for (int j = 0; j < kTaps - 1; ++j) s_halo[c][warp][j] = dd[c][j];
        }
    }
    __syncthreads();
    // pass 2: dd[8..10] = d(pre-activation) of the 3 positions to the right (next lane / next warp / next chunk;
    // only the last warp of the last chunk recomputes them from global memory), dx, parameter-gradient partials
    // dtaps[j] = sum_t x[t - 3 + j] * dpre[t];  dbias = sum_t dpre[t]
    float part[kTaps + 1];
#pragma unroll
    for (int j = 0; j <= kTaps; ++j) part[j] = 0.f;
#pragma unroll
    for (int c = 0; c < kConvChunks; ++c) {
        const int t0 = tb + c * kConvSpan;
#pragma unroll
        for (int j = 0; j < kTaps - 1; ++j) dd[c][8 + j] = __shfl_down_sync(0xffffffffu, dd[c][j], 1);
        if (lane == 31) {
            constexpr int kWarps = kConvThreads / 32;
            if (warp + 1 < kWarps || c + 1 < kConvChunks) {
                const int cn = warp + 1 < kWarps ? c : c + 1, wn = warp + 1 < kWarps ? warp + 1 : 0;
#pragma unroll
                for (int j = 0; j < kTaps - 1; ++j) dd[c][8 + j] = s_halo[cn < kConvChunks ? cn : 0][wn][j];
            } else {
#pragma unroll
                for (int j = 0; j < kTaps - 1; ++j) dd[c][8 + j] = dpre_at<T, kSilu>(x, dout, t0 + 8 + j, L, taps, bias);
            }
        }
        // dx[s] = sum_j taps[j] * dpre[s + 3 - j]
        float o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float acc = 0.f;
#pragma unroll
            for (int j = 0; j < kTaps; ++j) acc = fmaf(taps[j], dd[c][i + (kTaps - 1) - j], acc);
            o[i] = acc;
        }
        store8<T, kVec>(dx, t0, L, o);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
#pragma unroll
            for (int j = 0; j < kTaps; ++j) part[j] = fmaf(xx[c][i + j], dd[c][i], part[j]);
            part[kTaps] += dd[c][i];
        }
    }
#pragma unroll
    for (int j = 0; j <= kTaps; ++j) part[j] = warp_sum(part[j]);
    if (lane == 0) {
#pragma unroll
        for (int j = 0; j <= kTaps; ++j) s_red[warp][j] = part[j];
    }
    __syncthreads();
    if (threadIdx.x <= kTaps) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < kConvThreads / 32; ++w) s += s_red[w][threadIdx.x];
        const int j = threadIdx.x;
        if (j == kTaps) {
            if (a.dbias) atomicAdd(a.dbias + d, s);
        } else {
            const int k = j - (kTaps - a.width);
            if (k >= 0) atomicAdd(a.dweight + d * a.width + k, s);
        }
    }
}

}  // namespace vv
