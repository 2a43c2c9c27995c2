// dwconv3d.cuh -- depthwise 3x3x3 convolution over (frame, y, x), forward / backward, for sm_100a.
//
// The Mlp of every Temporal Mamba block of the reference runs nn.Conv3d(C, C, 3, 1, 1, groups=C) on its
// tokens (modeling/vivim.py:57-68, 99-106): (B, N, C) -> transpose -> (B, C, nf, H, W) -> cuDNN -> flatten ->
// transpose.  cuDNN serves the depthwise 3-D case with per-group convolveNd engines: in a Vivim training step
// (batch 3, 256x256, clip 5) its dgrad + wgrad kernels are 10752 launches and 90 % of the GPU time
// (profiles/r01_vivim_step.md).  The op is a 27-tap stencil with no reuse across channels, i.e. bandwidth
// bound: 2 tensors in/out forward, 3 + a (27, C) reduction backward.
//
// Design (B200-first): work directly on the TOKEN layout (B, nf, H, W, C), channels innermost, so neither
// transpose exists, and keep everything that is reused in REGISTERS -- shared memory and L1 hits go through the
// same 128 B/clk data pipe, which is what bounded the first version (27 window + 27 weight reads per output
// through L1: 69 % data-pipe utilisation at 13 % of the HBM roofline).  One thread owns 2 channels (a warp = 64
// contiguous channels = one 128-byte line per position) of a column of kDwX x-positions of one (b, y) row and
// walks the frames:
//   * its 27 x 2 weights live in registers for the whole walk;
//   * of every input frame it loads 3 rows x (kDwX + 2) positions ONCE and scatters them into the rolling
//     accumulators of the three output frames they touch (f-1, f, f+1); a finished frame is stored and its
//     accumulators recycled.  That is (kDwX + 2) * 3 / kDwX = 4.5 reads per output instead of 27, no weight traffic,
//     and the multiply-adds are packed fp32 pairs (FFMA2).
//   forward   out = bias + sum_tap w[tap, c] x[p + off(tap)]
//   backward  dx  = sum_tap w[tap, c] dout[p - off(tap)]          (same kernel, taps mirrored when loaded, no bias)
//             dw[tap, c] = sum_p dout[p] x[p + off(tap)],  db[c] = sum_p dout[p]
//   the weight gradient is the same walk with the roles swapped: 27 x 2 accumulators in registers, dout of frames
//   f-1, f, f+1 rolling in registers; the columns a CTA covers are reduced in shared memory, then one fp32
//   atomicAdd per (CTA, channel, tap).
#pragma once

#include "../../include/vivim_b200.h"
#include "common.cuh"

namespace vv {

constexpr int kDwX = 4;            // x positions per thread
constexpr int kDwThreads = 256;
constexpr int kDwCols = 8;         // columns per CTA of the weight-gradient kernel (block = 32 lanes x kDwCols)

struct DwGeom {
    int B, T, H, W, C;
};

// 2 adjacent channels of one position.  kPair: C is even and the tensor is aligned for one 4-byte (16-bit T) /
// 8-byte (fp32) access; otherwise element-wise with a channel bound.
template <typename T, bool kPair>
__device__ __forceinline__ float2 dw_ld(const T* __restrict__ p, int c, int C) {
    if (kPair) {
        if (sizeof(T) == 4) {
            const float2 v = __ldg(reinterpret_cast<const float2*>(p));
            return v;
        }
        return unpack_pair<T>(__ldg(reinterpret_cast<const uint32_t*>(p)));
    }
    return make_float2(to_f32<T>(p[0]), c + 1 < C ? to_f32<T>(p[1]) : 0.f);
}
template <typename T, bool kPair>
__device__ __forceinline__ void dw_st(T* __restrict__ p, int c, int C, float2 v) {
    if (kPair) {
        if (sizeof(T) == 4) {
            *reinterpret_cast<float2*>(p) = v;
        } else if (sizeof(T) == 2) {
            T tmp[2] = {from_f32<T>(v.x), from_f32<T>(v.y)};
            *reinterpret_cast<uint32_t*>(p) = *reinterpret_cast<const uint32_t*>(tmp);
        }
    } else {
        p[0] = from_f32<T>(v.x);
        if (c + 1 < C) p[1] = from_f32<T>(v.y);
    }
}
__device__ __forceinline__ float2 dw_ldw(const float* __restrict__ p, int c, int C, bool pair) {
    if (pair) return __ldg(reinterpret_cast<const float2*>(p));
    return make_float2(__ldg(p), c + 1 < C ? __ldg(p + 1) : 0.f);
}

struct DwCoord {
    int b, y, x0, c;
    bool live;
};
// thread index -> (channel pair fastest, x block, y, batch)
__device__ __forceinline__ DwCoord dw_coord(const DwGeom& g, int64_t idx) {
    const int cp = (g.C + 1) / 2, xt = (g.W + kDwX - 1) / kDwX;
    DwCoord k;
    k.live = idx < (int64_t)g.B * g.H * xt * cp;
    k.c = (int)(idx % cp) * 2;
    int64_t r = idx / cp;
    k.x0 = (int)(r % xt) * kDwX; r /= xt;
    k.y = (int)(r % g.H);
    k.b = (int)(r / g.H);
    return k;
}

// A pair of channels as loaded (one 32-bit register for 16-bit T): all 3 x (kDwX + 2) reads of a frame are issued
// before the first use, so that a thread has 72 bytes in flight -- the kernel is bound by memory-level parallelism.
template <typename T, bool kPair> struct DwRaw {
    float2 v;
    __device__ __forceinline__ void load(const T* __restrict__ p, int c, int C, bool ok) {
        v = ok ? dw_ld<T, kPair>(p, c, C) : make_float2(0.f, 0.f);
    }
    __device__ __forceinline__ float2 get() const { return v; }
};
template <> struct DwRaw<__nv_bfloat16, true> {
    uint32_t r;
    __device__ __forceinline__ void load(const __nv_bfloat16* __restrict__ p, int, int, bool ok) {
        r = ok ? __ldg(reinterpret_cast<const uint32_t*>(p)) : 0u;
    }
    __device__ __forceinline__ float2 get() const { return unpack_pair<__nv_bfloat16>(r); }
};
template <> struct DwRaw<__half, true> {
    uint32_t r;
    __device__ __forceinline__ void load(const __half* __restrict__ p, int, int, bool ok) {
        r = ok ? __ldg(reinterpret_cast<const uint32_t*>(p)) : 0u;
    }
    __device__ __forceinline__ float2 get() const { return unpack_pair<__half>(r); }
};

// Addressing of a thread's column, hoisted out of the frame walk (the first version recomputed a 64-bit address and
// the bounds of every read: 4600 instructions per warp, 42 % of them on the ALU pipe): the element offset of
// (b, frame 0, y, x = 0, c), the frame stride, and the frame-independent offsets / validity of the 3 rows and
// kDwX + 2 positions of the window.  A read is then one 32-bit add, one widening multiply-add and the load.
template <typename T, bool kPair>
struct DwWindow {
    int64_t col, fstride;
    int yoff[3], xoff[kDwX + 2];
    bool yok[3], xok[kDwX + 2];
    int c, C;
    __device__ __forceinline__ DwWindow(const DwGeom& g, const DwCoord& k) : c(k.c), C(g.C) {
        col = (((int64_t)k.b * g.T * g.H + k.y) * g.W) * g.C + k.c;
        fstride = (int64_t)g.H * g.W * g.C;
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
            yoff[dy] = (dy - 1) * g.W * g.C;
            yok[dy] = (unsigned)(k.y + dy - 1) < (unsigned)g.H;
        }
#pragma unroll
        for (int j = 0; j < kDwX + 2; ++j) {
            xoff[j] = (k.x0 + j - 1) * g.C;
            xok[j] = (unsigned)(k.x0 + j - 1) < (unsigned)g.W;
        }
    }
    // the 3 rows (y-1, y, y+1) of frame f: positions x0-1 .. x0+kDwX, zeros outside the volume
    __device__ __forceinline__ void load(const T* __restrict__ in, int f, DwRaw<T, kPair> (&win)[3][kDwX + 2]) const {
        const T* fb = in + col + f * fstride;
#pragma unroll
        for (int dy = 0; dy < 3; ++dy)
#pragma unroll
            for (int j = 0; j < kDwX + 2; ++j) win[dy][j].load(fb + (yoff[dy] + xoff[j]), c, C, yok[dy] && xok[j]);
    }
    // positions x0 .. x0+kDwX-1 of row y of frame f
    __device__ __forceinline__ void store(T* __restrict__ out, int f, const float2 (&v)[kDwX]) const {
        T* fb = out + col + f * fstride;
#pragma unroll
        for (int j = 0; j < kDwX; ++j)
            if (xok[j + 1]) dw_st<T, kPair>(fb + xoff[j + 1], c, C, v[j]);
    }
    __device__ __forceinline__ void load_row(const T* __restrict__ in, int f, bool fok, float2 (&v)[kDwX]) const {
        const T* fb = in + col + (fok ? f : 0) * fstride;
#pragma unroll
        for (int j = 0; j < kDwX; ++j) v[j] = (fok && xok[j + 1]) ? dw_ld<T, kPair>(fb + xoff[j + 1], c, C) : make_float2(0.f, 0.f);
    }
};

// kMirror = false: forward (+ bias).  kMirror = true: input gradient (weights mirrored, no bias).
// weight: fp32 (27, C), tap-major; tap = (dt * 3 + dy) * 3 + dx.
template <typename T, bool kPair, bool kMirror>
__global__ void __launch_bounds__(kDwThreads, 2) dwconv3d_kernel(const T* __restrict__ in, const float* __restrict__ weight,
                                                              const float* __restrict__ bias, T* __restrict__ out,
                                                              const DwGeom g) {
    const DwCoord k = dw_coord(g, (int64_t)blockIdx.x * kDwThreads + threadIdx.x);
    if (!k.live) return;
    const bool wpair = kPair;   // the host-side decision covers C even AND 8-byte aligned weight / bias pointers
    float2 w[27];
#pragma unroll
    for (int t = 0; t < 27; ++t) w[t] = dw_ldw(weight + (int64_t)(kMirror ? 26 - t : t) * g.C + k.c, k.c, g.C, wpair);
    const float2 b2 = (!kMirror && bias) ? dw_ldw(bias + k.c, k.c, g.C, wpair) : make_float2(0.f, 0.f);

    const DwWindow<T, kPair> win_at(g, k);
    // a0: output frame f-1, a1: frame f, a2: frame f+1 while input frame f is being scattered
    float2 a0[kDwX], a1[kDwX], a2[kDwX];
#pragma unroll
    for (int j = 0; j < kDwX; ++j) a0[j] = a1[j] = a2[j] = b2;

#pragma unroll 1
    for (int f = 0; f < g.T; ++f) {
        DwRaw<T, kPair> raw[3][kDwX + 2];
        win_at.load(in, f, raw);
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
            float2 win[kDwX + 2];
#pragma unroll
            for (int j = 0; j < kDwX + 2; ++j) win[j] = raw[dy][j].get();
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
                const float2 w_prev = w[(2 * 3 + dy) * 3 + dx], w_cur = w[(1 * 3 + dy) * 3 + dx], w_next = w[(0 * 3 + dy) * 3 + dx];
#pragma unroll
                for (int j = 0; j < kDwX; ++j) {
                    a0[j] = fma2(w_prev, win[j + dx], a0[j]);   // input frame f is frame t+1 of output t = f-1: dt = 2
                    a1[j] = fma2(w_cur, win[j + dx], a1[j]);
                    a2[j] = fma2(w_next, win[j + dx], a2[j]);
                }
            }
        }
        if (f > 0) win_at.store(out, f - 1, a0);
#pragma unroll
        for (int j = 0; j < kDwX; ++j) {
            a0[j] = a1[j];
            a1[j] = a2[j];
            a2[j] = b2;
        }
    }
    win_at.store(out, g.T - 1, a0);
}

// weight / bias gradient: block (32 lanes = 64 channels, kDwCols columns), grid (ceil(C / 64), column blocks)
template <typename T, bool kPair>
__global__ void __launch_bounds__(32 * kDwCols, 2) dwconv3d_wgrad_kernel(const T* __restrict__ x, const T* __restrict__ dout,
                                                                      float* __restrict__ dweight, float* __restrict__ dbias,
                                                                      const DwGeom g) {
    __shared__ float2 red[kDwCols][32];
    const int lane = threadIdx.x, col = threadIdx.y;
    const int xt = (g.W + kDwX - 1) / kDwX;
    const int ncols = g.B * g.H * xt;
    DwCoord k;
    k.c = (blockIdx.x * 32 + lane) * 2;
    k.live = k.c < g.C;
    float2 acc[27], accb = make_float2(0.f, 0.f);
#pragma unroll
    for (int t = 0; t < 27; ++t) acc[t] = make_float2(0.f, 0.f);
    if (k.live) {
        for (int ci = blockIdx.y * kDwCols + col; ci < ncols; ci += gridDim.y * kDwCols) {   // 32-bit: no division routines
            const int r1 = ci / xt, r2 = r1 / g.H;
            k.x0 = (ci - r1 * xt) * kDwX;
            k.y = r1 - r2 * g.H;
            k.b = r2;
            // dout of the column for frames f-1, f, f+1 (zeros outside the clip)
            const DwWindow<T, kPair> win_at(g, k);
            float2 g0[kDwX], g1[kDwX], g2[kDwX];
            auto load_g = [&](int f, float2 (&dst)[kDwX]) { win_at.load_row(dout, f, f < g.T, dst); };
#pragma unroll
            for (int j = 0; j < kDwX; ++j) g0[j] = make_float2(0.f, 0.f);
            load_g(0, g1);
#pragma unroll 1
            for (int f = 0; f < g.T; ++f) {
                DwRaw<T, kPair> raw[3][kDwX + 2];
                win_at.load(x, f, raw);
                load_g(f + 1, g2);
#pragma unroll
                for (int j = 0; j < kDwX; ++j) accb = add2(accb, g1[j]);
#pragma unroll
                for (int dy = 0; dy < 3; ++dy) {
                    float2 win[kDwX + 2];
#pragma unroll
                    for (int j = 0; j < kDwX + 2; ++j) win[j] = raw[dy][j].get();
#pragma unroll
                    for (int dx = 0; dx < 3; ++dx)
#pragma unroll
                        for (int j = 0; j < kDwX; ++j) {
                            // x frame f is frame t + dt - 1 of output t: t = f+1 (dt 0), f (dt 1), f-1 (dt 2)
                            acc[(0 * 3 + dy) * 3 + dx] = fma2(g2[j], win[j + dx], acc[(0 * 3 + dy) * 3 + dx]);
                            acc[(1 * 3 + dy) * 3 + dx] = fma2(g1[j], win[j + dx], acc[(1 * 3 + dy) * 3 + dx]);
                            acc[(2 * 3 + dy) * 3 + dx] = fma2(g0[j], win[j + dx], acc[(2 * 3 + dy) * 3 + dx]);
                        }
                }
#pragma unroll
                for (int j = 0; j < kDwX; ++j) {
                    g0[j] = g1[j];
                    g1[j] = g2[j];
                }
            }
        }
    }
    // reduce the columns of the CTA, one tap at a time (fully unrolled: acc[] stays in registers)
#pragma unroll
    for (int t = 0; t < 28; ++t) {
        red[col][lane] = t < 27 ? acc[t < 27 ? t : 0] : accb;
        __syncthreads();
        if (col == 0 && k.live && (t < 27 || dbias)) {
            float2 s = red[0][lane];
#pragma unroll
            for (int j = 1; j < kDwCols; ++j) s = add2(s, red[j][lane]);
            float* dst = t < 27 ? dweight + (int64_t)t * g.C + k.c : dbias + k.c;
            atomicAdd(dst, s.x);
            if (k.c + 1 < g.C) atomicAdd(dst + 1, s.y);
        }
        __syncthreads();
    }
}

}  // namespace vv
