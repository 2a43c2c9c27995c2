// conv1d_dirs.cuh -- causal depthwise conv1d of all scan directions of a Temporal Mamba block in one launch (sm_100a).
//
// Reference: Mamba.forward v3 (mamba/mamba_ssm/modules/mamba_simple.py:217-260) runs causal_conv1d_fwd three times, on
// xz, on xz.flip([-1]) and on the frame-interleaved copy of xz, each with its own weights
// (selective_scan_interface.py:177), and the backward runs the conv forward again plus causal_conv1d_bwd three times
// (:239, :281-283) and sums the three dx through the flip / permute backward copies.
//
// Here a direction is an addressing mode (Trav, common.cuh): a CTA stages ONE fp32 tile of x -- all frames x a run of
// pixels, with a 3-position halo on both sides -- in shared memory and produces from it the conv output of every
// direction, written in MEMORY order:
//     forward   reads T bytes, writes ndirs * T          (the reference: ndirs * 2T plus 4T of flip / interleave copies)
//     backward  reads (1 + ndirs) * T, writes T          (the reference: ndirs * 3T plus the copies and the dx sums)
// Layout of the work: x is (B, D, L) with L = nf * HW tokens (frame, pixel); a CTA owns (row, pixels [p0, p0 + pt)) for
// all nf frames (pt = 512, or 1024 without a frame-interleaved direction).  128-bit route (HW % 8 == 0): a thread owns 4
// consecutive pixels of a frame -- 128-bit shared loads at 16-byte stride (conflict-free), 64-bit global stores; the
// element-wise route (ragged HW) owns single pixels, consecutive threads on consecutive pixels.
//   FWD     out[m] = bias + sum_i w[i] x[m - 3 + i]                      taps to the left in memory
//   REV     out[m] = bias + sum_i w[i] x[m + 3 - i]                      taps to the right
//   FRAMES  token (t, p) is position j = p nf + t of the traversal; tap j - k lives at frame t - k of the same pixel,
//           or, wrapped, at frame t - k + s nf of pixel p - s
// The backward stages the tile of x and one tile of d(pre-activation) per direction, then every thread gathers the input
// gradient of its pixels from the three tiles; parameter gradients: per-thread partial sums -> warp shuffle -> shared
// memory -> one fp32 atomicAdd per (CTA, direction, tap).
#pragma once

#include "../../include/vivim_b200.h"
#include "common.cuh"

namespace vv {

constexpr int kDirsThreads = 128;
constexpr int kDirsHalo = 4;          // 3 halo positions + 1 pad word: the tile body starts 16-byte aligned
constexpr int kDirsMaxFrames = 16;

struct DirsGeom {
    int L, nf, hw, pt, pitch;   // pitch = pt + 2 * kDirsHalo floats per frame row of a tile
    int p0;                     // first pixel of the CTA
};

// x (or any tile) at frame t, local pixel q in [-kDirsHalo, pt + kDirsHalo)
__device__ __forceinline__ float tile_at(const float* __restrict__ tile, const DirsGeom& g, int t, int q) {
    return tile[t * g.pitch + kDirsHalo + q];
}

// Stage rows [t*hw + p0 - 3, t*hw + p0 + pt + 3) of `row` (memory order, zero outside [0, L)) for every frame t.
template <typename T, bool kVec>
__device__ __forceinline__ void stage_tile(float* __restrict__ tile, const T* __restrict__ row, const DirsGeom& g) {
    if (kVec) {
        const int chunks = g.pt / 8;
        for (int idx = threadIdx.x; idx < g.nf * chunks; idx += kDirsThreads) {
            const int t = idx / chunks, c = idx - t * chunks;
            const int m = t * g.hw + g.p0 + c * 8;
            float v[8];
            load8<T, true>(row, m, g.L, v);
            float* dst = tile + t * g.pitch + kDirsHalo + c * 8;
            reinterpret_cast<float4*>(dst)[0] = make_float4(v[0], v[1], v[2], v[3]);
            reinterpret_cast<float4*>(dst)[1] = make_float4(v[4], v[5], v[6], v[7]);
        }
        for (int idx = threadIdx.x; idx < g.nf * 2 * (kDirsHalo - 1); idx += kDirsThreads) {
            const int t = idx / (2 * (kDirsHalo - 1)), h = idx - t * 2 * (kDirsHalo - 1);
            const int q = h < kDirsHalo - 1 ? h - (kDirsHalo - 1) : g.pt + h - (kDirsHalo - 1);
            const int m = t * g.hw + g.p0 + q;
            tile[t * g.pitch + kDirsHalo + q] = (m >= 0 && m < g.L) ? to_f32<T>(row[m]) : 0.f;
        }
    } else {
        const int span = g.pt + 2 * (kDirsHalo - 1);
        for (int idx = threadIdx.x; idx < g.nf * span; idx += kDirsThreads) {
            const int t = idx / span, q = idx - t * span - (kDirsHalo - 1);
            const int m = t * g.hw + g.p0 + q;
            tile[t * g.pitch + kDirsHalo + q] = (m >= 0 && m < g.L) ? to_f32<T>(row[m]) : 0.f;
        }
    }
}

// Source of tap k (k positions back in traversal order) of token (t, q): frame ts, local pixel qs; false = before the
// start of the sequence (reads as zero).  FWD / REV stay in the frame row (rows are contiguous in memory).
__device__ __forceinline__ bool tap_source(int mode, const DirsGeom& g, int t, int q, int k, int& ts, int& qs) {
    if (mode == VV_DIR_FWD) { ts = t; qs = q - k; return t * g.hw + g.p0 + qs >= 0; }
    if (mode == VV_DIR_REV) { ts = t; qs = q + k; return t * g.hw + g.p0 + qs < g.L; }
    const int s = k > t ? (k - t + g.nf - 1) / g.nf : 0;     // pixels to step back
    ts = t - k + s * g.nf;
    qs = q - s;
    return g.p0 + qs >= 0;
}

// Token that reads token (t, q) through tap k (the inverse of tap_source): false = beyond the end of the sequence.
__device__ __forceinline__ bool tap_reader(int mode, const DirsGeom& g, int t, int q, int k, int& tr, int& qr) {
    if (mode == VV_DIR_FWD) { tr = t; qr = q + k; return t * g.hw + g.p0 + qr < g.L; }
    if (mode == VV_DIR_REV) { tr = t; qr = q - k; return t * g.hw + g.p0 + qr >= 0; }
    const int s = (t + k) / g.nf;
    tr = t + k - s * g.nf;
    qr = q + s;
    return g.p0 + qr < g.hw;
}

// pre-activation of direction `mode` at token (t, q); taps[i] multiplies the token kTaps-1-i positions back
__device__ __forceinline__ float pre_at(const float* __restrict__ xs, const DirsGeom& g, int mode, int t, int q,
                                        const float (&taps)[4], float bias) {
    float acc = bias;
#pragma unroll
    for (int k = 3; k >= 0; --k) {   // oldest tap first: the summation order of conv1d_fwd_kernel (bit-identical results)
        int ts, qs;
        if (tap_source(mode, g, t, q, k, ts, qs)) acc = fmaf(taps[3 - k], tile_at(xs, g, ts, qs), acc);
    }
    return acc;
}

__device__ __forceinline__ void load_dir_taps(const vv_conv1d_dirs_args& a, int k, int d, float (&taps)[4], float& bias) {
    const int K = a.width;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int i = j - (4 - K);
        taps[j] = i >= 0 ? a.weight[((int64_t)k * a.dim + d) * K + i] : 0.f;
    }
    bias = a.bias ? a.bias[(int64_t)k * a.dim + d] : 0.f;
}

__device__ __forceinline__ DirsGeom dirs_geom(const vv_conv1d_dirs_args& a, int pt) {
    DirsGeom g;
    g.L = a.seqlen;
    g.nf = a.nframes > 0 ? a.nframes : 1;
    g.hw = a.seqlen / g.nf;
    g.pt = pt;
    g.pitch = pt + 2 * kDirsHalo;
    g.p0 = blockIdx.y * pt;
    return g;
}

// ---------------------------------------------------------------- 128-bit path: 4 consecutive pixels per thread
// (hw % 8 == 0, pt % 8 == 0: a quad never straddles a frame or the end of the sequence).  W holds the own row
// x[q0-4 .. q0+7]; the taps of FWD / REV are compile-time offsets into it, a FRAMES tap is a window V = row[q0-4 .. q0+3] of
// another frame shifted by s pixels (warp-uniform), resolved by a 4-way switch so that register indices stay static.
__device__ __forceinline__ void load_f4(const float* __restrict__ p, float* v) {
    const float4 x = *reinterpret_cast<const float4*>(p);
    v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w;
}

template <int S>
__device__ __forceinline__ void add_shifted(float (&acc)[4], float w, const float (&V)[8]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i] = fmaf(w, V[4 + i - S], acc[i]);
}
__device__ __forceinline__ void add_shifted_by(int s, float (&acc)[4], float w, const float (&V)[8]) {
    switch (s) {
        case 0: add_shifted<0>(acc, w, V); break;
        case 1: add_shifted<1>(acc, w, V); break;
        case 2: add_shifted<2>(acc, w, V); break;
        default: add_shifted<3>(acc, w, V); break;
    }
}
// the same towards the readers (pixels to the right): V = row[q0 .. q0+7], value at index i + S
template <int S>
__device__ __forceinline__ void add_ahead(float (&acc)[4], float w, const float (&V)[8]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i] = fmaf(w, V[i + S], acc[i]);
}
__device__ __forceinline__ void add_ahead_by(int s, float (&acc)[4], float w, const float (&V)[8]) {
    switch (s) {
        case 0: add_ahead<0>(acc, w, V); break;
        case 1: add_ahead<1>(acc, w, V); break;
        case 2: add_ahead<2>(acc, w, V); break;
        default: add_ahead<3>(acc, w, V); break;
    }
}

// frame / pixel shift of tap kk (kk tokens back) of frame t in FRAMES order
__device__ __forceinline__ void frames_source(const DirsGeom& g, int t, int kk, int& ts, int& s) {
    s = kk > t ? (kk - t + g.nf - 1) / g.nf : 0;
    ts = t - kk + s * g.nf;
}
__device__ __forceinline__ void frames_reader(const DirsGeom& g, int t, int kk, int& tr, int& s) {
    s = (t + kk) / g.nf;
    tr = t + kk - s * g.nf;
}

// pre-activations of 4 consecutive pixels q0..q0+3 of frame t, direction `mode`, summed oldest tap first like
// conv1d_fwd_kernel (bit-identical results).  W = own row [q0-4, q0+8).
__device__ __forceinline__ void pre_quad(const float* __restrict__ xs, const DirsGeom& g, int mode, int t, int q0,
                                         const float (&W)[12], const float (&taps)[4], float bias, float (&acc)[4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i] = bias;
    if (mode == VV_DIR_FWD) {
#pragma unroll
        for (int kk = 3; kk >= 0; --kk)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[i] = fmaf(taps[3 - kk], W[4 + i - kk], acc[i]);
    } else if (mode == VV_DIR_REV) {
#pragma unroll
        for (int kk = 3; kk >= 0; --kk)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[i] = fmaf(taps[3 - kk], W[4 + i + kk], acc[i]);
    } else {
#pragma unroll
        for (int kk = 3; kk >= 1; --kk) {
            int ts, sft;
            frames_source(g, t, kk, ts, sft);
            float V[8];
            const float* r2 = xs + ts * g.pitch + kDirsHalo + q0;
            load_f4(r2 - 4, V);
            load_f4(r2, V + 4);
            if (g.p0 + q0 == 0) V[0] = V[1] = V[2] = V[3] = 0.f;   // pixels before the first one: start of the sequence
            add_shifted_by(sft, acc, taps[3 - kk], V);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[i] = fmaf(taps[3], W[4 + i], acc[i]);
    }
}

// src[kk][i] = the token kk positions back of pixel q0 + i (frame t) in direction `mode`
__device__ __forceinline__ void quad_sources(const float* __restrict__ xs, const DirsGeom& g, int mode, int t, int q0,
                                             const float (&W)[12], float (&src)[4][4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) src[0][i] = W[4 + i];
    if (mode == VV_DIR_FWD) {
#pragma unroll
        for (int kk = 1; kk < 4; ++kk)
#pragma unroll
            for (int i = 0; i < 4; ++i) src[kk][i] = W[4 + i - kk];
    } else if (mode == VV_DIR_REV) {
#pragma unroll
        for (int kk = 1; kk < 4; ++kk)
#pragma unroll
            for (int i = 0; i < 4; ++i) src[kk][i] = W[4 + i + kk];
    } else {
#pragma unroll
        for (int kk = 1; kk < 4; ++kk) {
            int ts, sft;
            frames_source(g, t, kk, ts, sft);
            float V[8];
            const float* r2 = xs + ts * g.pitch + kDirsHalo + q0;
            load_f4(r2 - 4, V);
            load_f4(r2, V + 4);
            if (g.p0 + q0 == 0) V[0] = V[1] = V[2] = V[3] = 0.f;
            float tmp[4] = {0.f, 0.f, 0.f, 0.f};
            add_shifted_by(sft, tmp, 1.f, V);          // tmp[i] = V[4 + i - sft]
#pragma unroll
            for (int i = 0; i < 4; ++i) src[kk][i] = tmp[i];
        }
    }
}

template <typename T>
__device__ __forceinline__ void store4(T* __restrict__ p, const float (&v)[4]);
template <> __device__ __forceinline__ void store4<float>(float* __restrict__ p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
template <> __device__ __forceinline__ void store4<__nv_bfloat16>(__nv_bfloat16* __restrict__ p, const float (&v)[4]) {
    const __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
    uint2 r;
    r.x = *reinterpret_cast<const uint32_t*>(&a);
    r.y = *reinterpret_cast<const uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p) = r;
}
template <> __device__ __forceinline__ void store4<__half>(__half* __restrict__ p, const float (&v)[4]) {
    const __half2 a = __floats2half2_rn(v[0], v[1]), b = __floats2half2_rn(v[2], v[3]);
    uint2 r;
    r.x = *reinterpret_cast<const uint32_t*>(&a);
    r.y = *reinterpret_cast<const uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p) = r;
}

// smem: [x tile: nf x pitch floats]
template <typename T, bool kSilu, bool kVec>
__global__ void __launch_bounds__(kDirsThreads) conv1d_dirs_fwd_kernel(const vv_conv1d_dirs_args a, const int pt) {
    extern __shared__ __align__(16) float dirs_smem[];
    float* xs = dirs_smem;
    const DirsGeom g = dirs_geom(a, pt);
    const int row = blockIdx.x;
    const int b = row / a.dim, d = row - b * a.dim;
    stage_tile<T, kVec>(xs, reinterpret_cast<const T*>(a.x) + b * a.x_bs + d * a.x_ds, g);
    __syncthreads();
    if (kVec) {
        float wd[VV_MAX_DIRS][4], bd[VV_MAX_DIRS];
#pragma unroll
        for (int k = 0; k < VV_MAX_DIRS; ++k)
            if (k < a.ndirs) load_dir_taps(a, k, d, wd[k], bd[k]);
        const int quads = g.pt / 4;
        for (int t = 0; t < g.nf; ++t) {
            for (int qi = threadIdx.x; qi < quads; qi += kDirsThreads) {
                const int q0 = qi * 4;
                if (g.p0 + q0 >= g.hw) break;
                const float* row = xs + t * g.pitch + kDirsHalo + q0;
                float W[12];
                load_f4(row - 4, W);
                load_f4(row, W + 4);
                load_f4(row + 4, W + 8);
#pragma unroll
                for (int k = 0; k < VV_MAX_DIRS; ++k) {
                    if (k < a.ndirs) {
                        float acc[4];
                        pre_quad(xs, g, a.dir_mode[k], t, q0, W, wd[k], bd[k], acc);
                        if (kSilu) {
#pragma unroll
                            for (int i = 0; i < 4; ++i) acc[i] *= sigmoid_f(acc[i]);
                        }
                        store4<T>(reinterpret_cast<T*>(a.out) + b * a.out_bs + ((int64_t)k * a.dim + d) * a.out_ds
                                      + t * g.hw + g.p0 + q0, acc);
                    }
                }
            }
        }
        return;
    }
    for (int k = 0; k < a.ndirs; ++k) {
        const int mode = a.dir_mode[k];
        float taps[4], bias;
        load_dir_taps(a, k, d, taps, bias);
        T* __restrict__ out = reinterpret_cast<T*>(a.out) + b * a.out_bs + ((int64_t)k * a.dim + d) * a.out_ds;
        for (int t = 0; t < g.nf; ++t) {
            for (int q = threadIdx.x; q < g.pt; q += kDirsThreads) {
                if (g.p0 + q >= g.hw) break;
                float v = pre_at(xs, g, mode, t, q, taps, bias);
                if (kSilu) v *= sigmoid_f(v);
                out[t * g.hw + g.p0 + q] = from_f32<T>(v);
            }
        }
    }
}

// smem: [x tile][d(pre) tile of direction 0][1][...]  (ndirs + 1 tiles of nf x pitch floats), [16 x 5 reduction words]
template <typename T, bool kSilu, bool kVec>
__global__ void __launch_bounds__(kDirsThreads) conv1d_dirs_bwd_kernel(const vv_conv1d_dirs_args a, const int pt) {
    extern __shared__ __align__(16) float dirs_smem[];
    const DirsGeom g = dirs_geom(a, pt);
    const int tile_words = g.nf * g.pitch;
    float* xs = dirs_smem;
    float* red = dirs_smem + (a.ndirs + 1) * tile_words;       // [warp][dir][5]
    const int row = blockIdx.x;
    const int b = row / a.dim, d = row - b * a.dim;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    stage_tile<T, kVec>(xs, reinterpret_cast<const T*>(a.x) + b * a.x_bs + d * a.x_ds, g);
    for (int k = 0; k < a.ndirs; ++k)
        stage_tile<T, kVec>(xs + (k + 1) * tile_words,
                            reinterpret_cast<const T*>(a.dout) + b * a.dout_bs + ((int64_t)k * a.dim + d) * a.dout_ds, g);
    __syncthreads();
    // ---- dout -> d(pre-activation), in place, body and halo; parameter-gradient partial sums over the body
    const int span = g.pt + 2 * (kDirsHalo - 1);
    const int qv = min(g.pt, g.hw - g.p0);       // pixels of the tile that exist
    for (int k = 0; k < a.ndirs; ++k) {
        const int mode = a.dir_mode[k];
        float taps[4], bias;
        load_dir_taps(a, k, d, taps, bias);
        float* dp = xs + (k + 1) * tile_words;
        float part[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
        if (kVec) {
            // body: 4 pixels per thread
            for (int t = 0; t < g.nf; ++t) {
                for (int qi = threadIdx.x; qi < qv / 4; qi += kDirsThreads) {
                    const int q0 = qi * 4;
                    const float* row = xs + t * g.pitch + kDirsHalo + q0;
                    float W[12], src[4][4], gr[4];
                    load_f4(row - 4, W);
                    load_f4(row, W + 4);
                    load_f4(row + 4, W + 8);
                    quad_sources(xs, g, mode, t, q0, W, src);
                    float* dq = dp + t * g.pitch + kDirsHalo + q0;
                    load_f4(dq, gr);
                    if (kSilu) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            float pre = bias;
#pragma unroll
                            for (int kk = 3; kk >= 0; --kk) pre = fmaf(taps[3 - kk], src[kk][i], pre);
                            const float sg = sigmoid_f(pre);
                            gr[i] *= sg * (1.f + pre * (1.f - sg));
                        }
                    }
                    *reinterpret_cast<float4*>(dq) = make_float4(gr[0], gr[1], gr[2], gr[3]);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        part[4] += gr[i];
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) part[3 - kk] = fmaf(gr[i], src[kk][i], part[3 - kk]);
                    }
                }
            }
        }
        // element-wise route: everything (kVec = false) or just the 3 positions on either side of the body.  The items
        // (frame, position) are dealt to threads 4 apart, so that the few halo items of the 128-bit route are one pass
        // for every warp instead of nf x 6 passes for warp 0 while the others wait at the barrier.
        const int nq = kVec ? 2 * (kDirsHalo - 1) : span;
        const int items = g.nf * nq;
        const int first = kVec ? ((threadIdx.x & 3) == 0 ? (int)(threadIdx.x >> 2) : items) : (int)threadIdx.x;
        for (int it = first; it < items; it += kVec ? kDirsThreads / 4 : kDirsThreads) {
            {
                const int t = it / nq, qq = it - t * nq;
                const int q = kVec ? (qq < kDirsHalo - 1 ? qq - (kDirsHalo - 1) : qv + qq - (kDirsHalo - 1))
                                   : qq - (kDirsHalo - 1);
                const int p = g.p0 + q, m = t * g.hw + p;
                // a token exists in this direction's sequence iff its memory index is in range (FWD / REV: rows are
                // contiguous, a halo position may belong to the neighbouring frame) / its pixel is (FRAMES)
                const bool exists = mode == VV_DIR_FRAMES ? (p >= 0 && p < g.hw) : (m >= 0 && m < g.L);
                // d(pre) is only consumed towards the readers of the body: to the right for FWD / FRAMES, to the left for REV
                const bool needed = mode == VV_DIR_REV ? q < g.pt : q >= 0;
                float gr = 0.f;
                if (exists && needed) {
                    gr = dp[t * g.pitch + kDirsHalo + q];
                    if (kSilu) {
                        const float pre = pre_at(xs, g, mode, t, q, taps, bias);
                        const float sg = sigmoid_f(pre);
                        gr *= sg * (1.f + pre * (1.f - sg));
                    }
                }
                dp[t * g.pitch + kDirsHalo + q] = gr;
                if (q >= 0 && q < g.pt && p < g.hw) {       // body: this CTA owns the token
                    part[4] += gr;
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        int ts, qs;
                        if (tap_source(mode, g, t, q, kk, ts, qs)) part[3 - kk] = fmaf(gr, tile_at(xs, g, ts, qs), part[3 - kk]);
                    }
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 5; ++j) part[j] = warp_sum(part[j]);
        if (lane == 0) {
#pragma unroll
            for (int j = 0; j < 5; ++j) red[(warp * VV_MAX_DIRS + k) * 5 + j] = part[j];
        }
    }
    __syncthreads();
    // ---- dx: every direction's readers of this token
    float wd[VV_MAX_DIRS][4];
#pragma unroll
    for (int k = 0; k < VV_MAX_DIRS; ++k) {
        float bias_unused;
        if (k < a.ndirs) load_dir_taps(a, k, d, wd[k], bias_unused);
    }
    T* __restrict__ dx = reinterpret_cast<T*>(a.dx) + b * a.dx_bs + d * a.dx_ds;
    if (kVec) {
        for (int t = 0; t < g.nf; ++t) {
            for (int qi = threadIdx.x; qi < qv / 4; qi += kDirsThreads) {
                const int q0 = qi * 4;
                float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int k = 0; k < VV_MAX_DIRS; ++k) {
                    if (k < a.ndirs) {
                        const int mode = a.dir_mode[k];
                        const float* dq = xs + (k + 1) * tile_words + t * g.pitch + kDirsHalo + q0;
                        float Dw[8];
                        if (mode == VV_DIR_FWD) {              // readers q + kk
                            load_f4(dq, Dw);
                            load_f4(dq + 4, Dw + 4);
#pragma unroll
                            for (int kk = 0; kk < 4; ++kk)
#pragma unroll
                                for (int i = 0; i < 4; ++i) acc[i] = fmaf(wd[k][3 - kk], Dw[i + kk], acc[i]);
                        } else if (mode == VV_DIR_REV) {       // readers q - kk
                            load_f4(dq - 4, Dw);
                            load_f4(dq, Dw + 4);
#pragma unroll
                            for (int kk = 0; kk < 4; ++kk)
#pragma unroll
                                for (int i = 0; i < 4; ++i) acc[i] = fmaf(wd[k][3 - kk], Dw[4 + i - kk], acc[i]);
                        } else {                               // readers in frame tr, s pixels to the right
#pragma unroll
                            for (int kk = 0; kk < 4; ++kk) {
                                int tr, sft;
                                frames_reader(g, t, kk, tr, sft);
                                const float* dr = xs + (k + 1) * tile_words + tr * g.pitch + kDirsHalo + q0;
                                load_f4(dr, Dw);
                                load_f4(dr + 4, Dw + 4);
                                add_ahead_by(sft, acc, wd[k][3 - kk], Dw);
                            }
                        }
                    }
                }
                store4<T>(dx + t * g.hw + g.p0 + q0, acc);
            }
        }
    }
    for (int t = 0; t < (kVec ? 0 : g.nf); ++t) {
        for (int q = threadIdx.x; q < g.pt; q += kDirsThreads) {
            if (g.p0 + q >= g.hw) break;
            float acc = 0.f;
#pragma unroll
            for (int k = 0; k < VV_MAX_DIRS; ++k) {
                if (k < a.ndirs) {
                    const int mode = a.dir_mode[k];
                    const float* dp = xs + (k + 1) * tile_words;
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        int tr, qr;
                        if (tap_reader(mode, g, t, q, kk, tr, qr)) acc = fmaf(wd[k][3 - kk], tile_at(dp, g, tr, qr), acc);
                    }
                }
            }
            dx[t * g.hw + g.p0 + q] = from_f32<T>(acc);
        }
    }
    // ---- parameter gradients of the CTA
    if (threadIdx.x < a.ndirs * 5) {
        const int k = threadIdx.x / 5, j = threadIdx.x - k * 5;
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < kDirsThreads / 32; ++w) s += red[(w * VV_MAX_DIRS + k) * 5 + j];
        if (j == 4) {
            if (a.dbias) atomicAdd(a.dbias + (int64_t)k * a.dim + d, s);
        } else {
            const int i = j - (4 - a.width);
            if (i >= 0) atomicAdd(a.dweight + ((int64_t)k * a.dim + d) * a.width + i, s);
        }
    }
}

}  // namespace vv
