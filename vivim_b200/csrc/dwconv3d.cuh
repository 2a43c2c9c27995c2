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
// weight: fp32 (27, C) -- tap-major, so that the 8 channels of a thread are 32 contiguous bytes and a warp reads
// 1 KB contiguous per tap; tap = (dt * 3 + dy) * 3 + dx.
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
        const int ts = t + dt - 1, ys = y + dy - 1;
        const bool row_ok = (unsigned)ts < (unsigned)g.T && (unsigned)ys < (unsigned)g.H;
        // the window stays packed in the I/O dtype (4 registers per 8 bf16 channels) and is widened at each use:
        // 80 registers, 3 CTAs per SM -- the kernel is latency bound, more warps in flight matter more than ALU ops
        Raw8<T, kVec> win[kDwX + 2];
        const T* rowp = in + (((int64_t)b * g.T + ts) * g.H + ys) * g.W * g.C + c0;
#pragma unroll
        for (int j = 0; j < kDwX + 2; ++j) {
            const int xs = x0 + j - 1;
            if (kVec) win[j].load(rowp + (int64_t)xs * g.C, (row_ok && (unsigned)xs < (unsigned)g.W) ? 0 : 1, 1);
            else win[j].load(rowp + (int64_t)xs * g.C, (row_ok && (unsigned)xs < (unsigned)g.W) ? 0 : 8, min(8, g.C - c0));
        }
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
            const int tap = kMirror ? 26 - (row * 3 + dx) : row * 3 + dx;
            float w[8];
            if (kVec) {
                load8_vec<float>(weight + (int64_t)tap * g.C + c0, w);
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) w[i] = (c0 + i < g.C) ? __ldg(weight + (int64_t)tap * g.C + c0 + i) : 0.f;
            }
#pragma unroll
            for (int j = 0; j < kDwX; ++j) {
                float v[8];
                win[j + dx].unpack(v);
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[j][i] = fmaf(w[i], v[i], acc[j][i]);
            }
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

// weight / bias gradient.  A thread owns 4 channels (64-bit accesses; a warp = 128 contiguous channels) and the 9
// taps of ONE frame offset dt; it walks rows of kDwX outputs with the same sliding window as the forward kernel
// (3 rows x (kDwX + 2) loads of x + kDwX loads of dout for 36 x kDwX FMAs), striding over the (b, frame, y, x-block)
// rows.  block (32 lanes, 3 frame offsets, kDwSlots row slots); grid (ceil(C / 128), row blocks).  The row slots of
// a CTA are reduced in shared memory, then one fp32 atomicAdd per (CTA, channel, tap).
constexpr int kDwSlots = 4;

template <typename T>
__device__ __forceinline__ void dw_load4(const T* __restrict__ p, int c, int C, bool vec, float (&v)[4]) {
    if (vec) {
        if (sizeof(T) == 4) {
            const float4 x = __ldg(reinterpret_cast<const float4*>(p));
            v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w;
        } else {
            const uint2 x = __ldg(reinterpret_cast<const uint2*>(p));
            const float2 lo = unpack_pair<T>(x.x), hi = unpack_pair<T>(x.y);
            v[0] = lo.x; v[1] = lo.y; v[2] = hi.x; v[3] = hi.y;
        }
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = (c + i < C) ? to_f32<T>(p[i]) : 0.f;
    }
}

template <typename T>
__global__ void __launch_bounds__(32 * 3 * kDwSlots) dwconv3d_wgrad_kernel(const T* __restrict__ x, const T* __restrict__ dout,
                                                                           float* __restrict__ dweight, float* __restrict__ dbias,
                                                                           const DwGeom g, const int vec) {
    __shared__ float4 red[kDwSlots][3][32];
    const int lane = threadIdx.x, dt = threadIdx.y, slot = threadIdx.z;
    const int c = (blockIdx.x * 32 + lane) * 4;
    const bool live = c < g.C;
    const int xt = (g.W + kDwX - 1) / kDwX;
    const int64_t nrows = (int64_t)g.B * g.T * g.H * xt;
    float acc[9][4], accb[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < 9; ++k)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[k][i] = 0.f;
    if (live) {
        for (int64_t rix = (int64_t)blockIdx.y * kDwSlots + slot; rix < nrows; rix += (int64_t)gridDim.y * kDwSlots) {
            int64_t r = rix;
            const int xb = (int)(r % xt); r /= xt;
            const int y = (int)(r % g.H); r /= g.H;
            const int t = (int)(r % g.T);
            const int b = (int)(r / g.T);
            const int x0 = xb * kDwX, ts = t + dt - 1;
            if ((unsigned)ts >= (unsigned)g.T) continue;       // this frame offset falls outside the clip: zero padding
            float go[kDwX][4];
#pragma unroll
            for (int j = 0; j < kDwX; ++j) {
                if (x0 + j < g.W) {
                    dw_load4<T>(dout + ((((int64_t)b * g.T + t) * g.H + y) * g.W + x0 + j) * g.C + c, c, g.C, vec, go[j]);
                } else {
#pragma unroll
                    for (int i = 0; i < 4; ++i) go[j][i] = 0.f;
                }
                if (dt == 1) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) accb[i] += go[j][i];
                }
            }
#pragma unroll
            for (int dy = 0; dy < 3; ++dy) {
                const int ys = y + dy - 1;
                if ((unsigned)ys >= (unsigned)g.H) continue;
                float win[kDwX + 2][4];
#pragma unroll
                for (int j = 0; j < kDwX + 2; ++j) {
                    const int xs = x0 + j - 1;
                    if ((unsigned)xs < (unsigned)g.W) {
                        dw_load4<T>(x + ((((int64_t)b * g.T + ts) * g.H + ys) * g.W + xs) * g.C + c, c, g.C, vec, win[j]);
                    } else {
#pragma unroll
                        for (int i = 0; i < 4; ++i) win[j][i] = 0.f;
                    }
                }
#pragma unroll
                for (int dx = 0; dx < 3; ++dx)
#pragma unroll
                    for (int j = 0; j < kDwX; ++j)
#pragma unroll
                        for (int i = 0; i < 4; ++i) acc[dy * 3 + dx][i] = fmaf(go[j][i], win[j + dx][i], acc[dy * 3 + dx][i]);
            }
        }
    }
    // reduce the row slots of the CTA, one tap at a time (fully unrolled: acc[] stays in registers)
#pragma unroll
    for (int k = 0; k < 10; ++k) {
        const float* src = k < 9 ? acc[k < 9 ? k : 0] : accb;
        red[slot][dt][lane] = make_float4(src[0], src[1], src[2], src[3]);
        __syncthreads();
        if (slot == 0 && live && (k < 9 || (dt == 1 && dbias))) {
            float4 s = red[0][dt][lane];
#pragma unroll
            for (int j = 1; j < kDwSlots; ++j) {
                const float4 o = red[j][dt][lane];
                s.x += o.x; s.y += o.y; s.z += o.z; s.w += o.w;
            }
            const float sv[4] = {s.x, s.y, s.z, s.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (c + i < g.C) {
                    if (k < 9) atomicAdd(dweight + (int64_t)(dt * 9 + k) * g.C + c + i, sv[i]);
                    else atomicAdd(dbias + c + i, sv[i]);
                }
            }
        }
        __syncthreads();
    }
}

}  // namespace vv
