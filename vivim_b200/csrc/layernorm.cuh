// layernorm.cuh -- LayerNorm over the channel axis of token tensors (rows x channels), forward and backward, sm_100a.
//
// The glue around the Mamba path inside a Temporal Mamba block (reference: modeling/vivim.py:153-157 --
// `x_mamba = x_flat + drop_path(mamba(norm1(x_flat)))`, `x_mamba + drop_path(mlp(norm2(x_mamba), ...))`).  The reference runs
// nn.LayerNorm; at Vivim's shapes (61 440 tokens x 64 channels at stage 1, batch 3) torch's backward spends 70 us per call
// in its gamma / beta reduction alone -- more than the selective scan of the same layer.  Here:
//   * one warp per row, the row lives in registers (channels <= 1024), two-pass mean / variance in fp32, 128-bit accesses;
//   * the output can be written directly in the dtype the consumer GEMM wants (bf16 under autocast): the separate
//     fp32 -> bf16 cast pass of the activations disappears;
//   * backward: dx and the gamma / beta partial sums in ONE pass over x and dout -- every warp keeps its channels' partial
//     sums in registers across the rows it walks, the CTA folds its warps through shared memory and issues one fp32
//     atomicAdd per channel (HBM-bound: 2 reads + 1 write of the tensor).
#pragma once

#include "../../include/vivim_b200.h"
#include "common.cuh"

namespace vv {

constexpr int kLnThreads = 256;                 // 8 warps = 8 rows in flight per CTA
constexpr int kLnWarps = kLnThreads / 32;

// 4 consecutive elements (V = 4) or one (V = 1) of a row, as fp32
template <typename T, int V> struct LnVec;
template <typename T> struct LnVec<T, 1> {
    static __device__ __forceinline__ void load(const T* p, float (&v)[1]) { v[0] = to_f32<T>(*p); }
    static __device__ __forceinline__ void store(T* p, const float (&v)[1]) { *p = from_f32<T>(v[0]); }
};
template <> struct LnVec<float, 4> {
    static __device__ __forceinline__ void load(const float* p, float (&v)[4]) {
        const float4 x = *reinterpret_cast<const float4*>(p);
        v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w;
    }
    static __device__ __forceinline__ void store(float* p, const float (&v)[4]) {
        *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    }
};
template <> struct LnVec<__nv_bfloat16, 4> {
    static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[4]) {
        const uint2 r = *reinterpret_cast<const uint2*>(p);
        const float2 a = unpack_pair<__nv_bfloat16>(r.x), b = unpack_pair<__nv_bfloat16>(r.y);
        v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
    }
    static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&v)[4]) {
        const __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
        uint2 r;
        r.x = *reinterpret_cast<const uint32_t*>(&a);
        r.y = *reinterpret_cast<const uint32_t*>(&b);
        *reinterpret_cast<uint2*>(p) = r;
    }
};
template <> struct LnVec<__half, 4> {
    static __device__ __forceinline__ void load(const __half* p, float (&v)[4]) {
        const uint2 r = *reinterpret_cast<const uint2*>(p);
        const float2 a = unpack_pair<__half>(r.x), b = unpack_pair<__half>(r.y);
        v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
    }
    static __device__ __forceinline__ void store(__half* p, const float (&v)[4]) {
        const __half2 a = __floats2half2_rn(v[0], v[1]), b = __floats2half2_rn(v[2], v[3]);
        uint2 r;
        r.x = *reinterpret_cast<const uint32_t*>(&a);
        r.y = *reinterpret_cast<const uint32_t*>(&b);
        *reinterpret_cast<uint2*>(p) = r;
    }
};

// sum over the LPR (8, 16 or 32) adjacent lanes that share a row
template <int LPR>
__device__ __forceinline__ float row_sum(float v) {
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// LPR lanes share a row (LPR < 32 only with K = 1: a 64-channel row is 16 groups of 4, so a warp serves two rows at a
// time instead of idling half of its lanes).  Lane `lane` of the row owns the V-element groups g = lane + 32 k (k < K) of a row, i.e. channels [g V, g V + V); groups beyond
// the row (g V >= C) are padding.  Vivim: C = 64 / 128 / 320 / 512 -> (V, K) = (4, 1) / (4, 1) / (4, 4) / (4, 4).
template <typename TI, typename TO, int V, int K, int LPR>
__global__ void __launch_bounds__(kLnThreads) layernorm_fwd_kernel(const vv_layernorm_args a) {
    constexpr int RPW = 32 / LPR;                 // rows per warp
    const int lane = (threadIdx.x & 31) % LPR, warp = (threadIdx.x >> 5) * RPW + (threadIdx.x & 31) / LPR;
    const int C = a.channels;
    const float inv_c = 1.f / (float)C;
    float w[K][V], bsh[K][V];
#pragma unroll
    for (int k = 0; k < K; ++k)
#pragma unroll
        for (int i = 0; i < V; ++i) {
            const int c = (lane + LPR * k) * V + i;
            w[k][i] = (a.weight && c < C) ? a.weight[c] : 1.f;
            bsh[k][i] = (a.bias && c < C) ? a.bias[c] : 0.f;
        }
    // every lane of a warp makes the same number of trips (the row sums are full-mask shuffles): the rows of a trip that
    // do not exist are computed on zeros and not stored
    for (int64_t row0 = (int64_t)blockIdx.x * kLnWarps * RPW; row0 < a.rows; row0 += (int64_t)gridDim.x * kLnWarps * RPW) {
        const int64_t row = row0 + warp;
        const bool valid = row < a.rows;
        const TI* x = reinterpret_cast<const TI*>(a.x) + (valid ? row : 0) * a.x_rs;
        float v[K][V];
        float sum = 0.f;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int c0 = (lane + LPR * k) * V;
            if (c0 < C && valid) LnVec<TI, V>::load(x + c0, v[k]);
            else {
#pragma unroll
                for (int i = 0; i < V; ++i) v[k][i] = 0.f;
            }
#pragma unroll
            for (int i = 0; i < V; ++i) sum += v[k][i];
        }
        const float mean = row_sum<LPR>(sum) * inv_c;
        float sq = 0.f;
#pragma unroll
        for (int k = 0; k < K; ++k)
#pragma unroll
            for (int i = 0; i < V; ++i) {
                const float dlt = ((lane + LPR * k) * V + i < C) ? v[k][i] - mean : 0.f;
                sq = fmaf(dlt, dlt, sq);
            }
        const float rstd = rsqrtf(row_sum<LPR>(sq) * inv_c + a.eps);
        TO* out = reinterpret_cast<TO*>(a.out) + (valid ? row : 0) * a.out_rs;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int c0 = (lane + LPR * k) * V;
            if (c0 < C && valid) {
                float o[V];
#pragma unroll
                for (int i = 0; i < V; ++i) o[i] = fmaf((v[k][i] - mean) * rstd, w[k][i], bsh[k][i]);
                LnVec<TO, V>::store(out + c0, o);
            }
        }
        if (lane == 0 && valid) {
            a.mean[row] = mean;
            a.rstd[row] = rstd;
        }
    }
}

// dx = rstd (g w - mean(g w) - xhat mean(g w xhat));  dweight += sum_rows g xhat;  dbias += sum_rows g
template <typename TI, typename TO, int V, int K, int LPR>
__global__ void __launch_bounds__(kLnThreads) layernorm_bwd_kernel(const vv_layernorm_args a) {
    constexpr int RPW = 32 / LPR;
    __shared__ float red[2][kLnWarps * RPW][LPR * K * V + 1];
    const int lane = (threadIdx.x & 31) % LPR, warp = (threadIdx.x >> 5) * RPW + (threadIdx.x & 31) / LPR;
    const int C = a.channels;
    const float inv_c = 1.f / (float)C;
    float w[K][V], pw[K][V], pb[K][V];
#pragma unroll
    for (int k = 0; k < K; ++k)
#pragma unroll
        for (int i = 0; i < V; ++i) {
            const int c = (lane + LPR * k) * V + i;
            w[k][i] = (a.weight && c < C) ? a.weight[c] : 1.f;
            pw[k][i] = pb[k][i] = 0.f;
        }
    for (int64_t row0 = (int64_t)blockIdx.x * kLnWarps * RPW; row0 < a.rows; row0 += (int64_t)gridDim.x * kLnWarps * RPW) {
        const int64_t row = row0 + warp;
        const bool valid = row < a.rows;
        const TI* x = reinterpret_cast<const TI*>(a.x) + (valid ? row : 0) * a.x_rs;
        const TO* g = reinterpret_cast<const TO*>(a.dout) + (valid ? row : 0) * a.dout_rs;
        const float mean = valid ? a.mean[row] : 0.f, rstd = valid ? a.rstd[row] : 0.f;
        float xh[K][V], gw[K][V];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int c0 = (lane + LPR * k) * V;
            float gv[V];
            if (c0 < C && valid) {
                LnVec<TI, V>::load(x + c0, xh[k]);
                LnVec<TO, V>::load(g + c0, gv);
            } else {
#pragma unroll
                for (int i = 0; i < V; ++i) { xh[k][i] = mean; gv[i] = 0.f; }
            }
#pragma unroll
            for (int i = 0; i < V; ++i) {
                xh[k][i] = (xh[k][i] - mean) * rstd;
                gw[k][i] = gv[i] * w[k][i];
                s1 = fmaf(gw[k][i], xh[k][i], s1);
                s2 += gw[k][i];
                pw[k][i] = fmaf(gv[i], xh[k][i], pw[k][i]);
                pb[k][i] += gv[i];
            }
        }
        s1 = row_sum<LPR>(s1) * inv_c;
        s2 = row_sum<LPR>(s2) * inv_c;
        if (a.dx && valid) {
            TI* dx = reinterpret_cast<TI*>(a.dx) + row * a.dx_rs;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const int c0 = (lane + LPR * k) * V;
                if (c0 < C) {
                    float o[V];
#pragma unroll
                    for (int i = 0; i < V; ++i) o[i] = rstd * (gw[k][i] - s2 - xh[k][i] * s1);
                    LnVec<TI, V>::store(dx + c0, o);
                }
            }
        }
    }
    if (a.dweight == nullptr && a.dbias == nullptr) return;
    // ---- fold the partial sums of the CTA's (sub-)warps, one atomic per channel and CTA
#pragma unroll
    for (int k = 0; k < K; ++k)
#pragma unroll
        for (int i = 0; i < V; ++i) {
            red[0][warp][(lane + LPR * k) * V + i] = pw[k][i];
            red[1][warp][(lane + LPR * k) * V + i] = pb[k][i];
        }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += kLnThreads) {
        float sw = 0.f, sb = 0.f;
#pragma unroll
        for (int wv = 0; wv < kLnWarps * RPW; ++wv) {
            sw += red[0][wv][c];
            sb += red[1][wv][c];
        }
        if (a.dweight) atomicAdd(a.dweight + c, sw);
        if (a.dbias) atomicAdd(a.dbias + c, sb);
    }
}

}  // namespace vv
