// common.cuh -- element-type helpers shared by the conv1d and selective-scan kernels (sm_100a).
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/vivim_b200.h"

namespace vv {

constexpr int kWarp = 32;
constexpr int kVecElems = 8;          // sequence positions one thread owns per unit
constexpr float kLog2e = 1.4426950408889634f;

// ---------------------------------------------------------------- scalar conversions
template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// ---------------------------------------------------------------- 8-element row segments
// Vector path: `p` is 16-byte aligned and all 8 elements are in range (128-bit coalesced accesses).
template <typename T> __device__ __forceinline__ void load8_vec(const T* __restrict__ p, float (&v)[8]);

template <> __device__ __forceinline__ void load8_vec<float>(const float* __restrict__ p, float (&v)[8]) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p));
    const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
    v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <> __device__ __forceinline__ void load8_vec<__nv_bfloat16>(const __nv_bfloat16* __restrict__ p, float (&v)[8]) {
    const uint4 r = __ldg(reinterpret_cast<const uint4*>(p));
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {          // bf16 -> f32 is a 16-bit shift
        v[2 * i] = __uint_as_float(w[i] << 16);
        v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
}
template <> __device__ __forceinline__ void load8_vec<__half>(const __half* __restrict__ p, float (&v)[8]) {
    const uint4 r = __ldg(reinterpret_cast<const uint4*>(p));
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
        v[2 * i] = f.x;
        v[2 * i + 1] = f.y;
    }
}

// same unpacking from a plain (shared-memory or generic) 16-byte aligned address
template <typename T> __device__ __forceinline__ void load8_plain(const T* p, float (&v)[8]);
template <> __device__ __forceinline__ void load8_plain<float>(const float* p, float (&v)[8]) {
    const float4 a = reinterpret_cast<const float4*>(p)[0];
    const float4 b = reinterpret_cast<const float4*>(p)[1];
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
    v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <> __device__ __forceinline__ void load8_plain<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[8]) {
    const uint4 r = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        v[2 * i] = __uint_as_float(w[i] << 16);
        v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
}
template <> __device__ __forceinline__ void load8_plain<__half>(const __half* p, float (&v)[8]) {
    const uint4 r = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
        v[2 * i] = f.x;
        v[2 * i + 1] = f.y;
    }
}

template <typename T> __device__ __forceinline__ void store8_vec(T* __restrict__ p, const float (&v)[8]);

template <> __device__ __forceinline__ void store8_vec<float>(float* __restrict__ p, const float (&v)[8]) {
    reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
    reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
}
template <> __device__ __forceinline__ void store8_vec<__nv_bfloat16>(__nv_bfloat16* __restrict__ p, const float (&v)[8]) {
    uint4 r;
    uint32_t* w = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
        w[i] = *reinterpret_cast<const uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(p) = r;
}
template <> __device__ __forceinline__ void store8_vec<__half>(__half* __restrict__ p, const float (&v)[8]) {
    uint4 r;
    uint32_t* w = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const __half2 h = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
        w[i] = *reinterpret_cast<const uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(p) = r;
}

// Row segment [t0, t0+8) of a row of length L starting at `row`.  kVec: the row base is 16-byte
// aligned and L % 8 == 0, so a segment is either fully inside or fully outside.  Positions outside
// [0, L) read as `fill` and are not written.
template <typename T, bool kVec>
__device__ __forceinline__ void load8(const T* __restrict__ row, int t0, int L, float (&v)[8], float fill = 0.f) {
    if (kVec) {
        if (t0 >= 0 && t0 < L) {
            load8_vec<T>(row + t0, v);
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = fill;
        }
    } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int t = t0 + i;
            v[i] = (t >= 0 && t < L) ? to_f32<T>(row[t]) : fill;
        }
    }
}

template <typename T, bool kVec>
__device__ __forceinline__ void store8(T* __restrict__ row, int t0, int L, const float (&v)[8]) {
    if (kVec) {
        if (t0 < L) store8_vec<T>(row + t0, v);
    } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int t = t0 + i;
            if (t < L) row[t] = from_f32<T>(v[i]);
        }
    }
}

// ---------------------------------------------------------------- traversal order
// A scan / conv direction visits the L tokens of a row in TRAVERSAL order j = 0..L-1; every (.., L) tensor stays in
// MEMORY order.  Trav maps j to the memory index (reference: the three directions of Mamba.forward v3,
// mamba/mamba_ssm/modules/mamba_simple.py:217-264, which materialise xz.flip(-1) and the (t, hw) -> (hw, t) copy):
//   VV_DIR_FWD     m = j
//   VV_DIR_REV     m = L - 1 - j                                   (xz.flip([-1]))
//   VV_DIR_FRAMES  m = (j % nf) * (L / nf) + j / nf                (tokens (frame, pixel) visited pixel-major)
struct Trav {
    int mode, L, nf, hw;
    __device__ __forceinline__ int mem(int j) const {
        if (mode == VV_DIR_FWD) return j;
        if (mode == VV_DIR_REV) return L - 1 - j;
        const int p = j / nf;
        return (j - p * nf) * hw + p;
    }
};

__device__ __forceinline__ uint32_t swap_halves(uint32_t w) { return __byte_perm(w, 0, 0x1032); }

// positions [j0, j0+8) of traversal order -> memory (kVec: L % 8 == 0, j0 % 8 == 0, row 16-byte aligned)
template <typename T, bool kVec>
__device__ __forceinline__ void store8_trav(T* __restrict__ row, int j0, const Trav& tr, const float (&v)[8]) {
    if (tr.mode == VV_DIR_FWD) { store8<T, kVec>(row, j0, tr.L, v); return; }
    if (kVec && tr.mode == VV_DIR_REV) {
        if (j0 < tr.L) {
            const float r[8] = {v[7], v[6], v[5], v[4], v[3], v[2], v[1], v[0]};
            store8_vec<T>(row + (tr.L - 8 - j0), r);
        }
        return;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int j = j0 + i;
        if (j < tr.L) row[tr.mem(j)] = from_f32<T>(v[i]);
    }
}

// ---------------------------------------------------------------- FRAMES order, coalesced
// The 64 traversal positions [j0, j0+64) of a FRAMES-ordered row are, in memory, nf short runs of consecutive pixels
// (one per frame, ~64/nf pixels each).  Reading them position by position costs one 2-byte access per element and lane,
// every warp-wide request touching up to 32 sectors; here the lanes that share a row walk the runs instead (consecutive
// lanes = consecutive pixels: one or two sectors per run) and drop the elements into a shared-memory row in TRAVERSAL
// order, from which every lane takes its 8 positions with one 128-bit read.  FramesSpan holds the run boundaries of one
// segment (two integer divisions per CTA instead of two per element).
struct FramesSpan {
    int p_first, t_first;   // pixel / frame of position j0
    int p_last, t_last;     // pixel / frame of the last position of the span that exists (< L)
    int j0;
};
__device__ __forceinline__ FramesSpan frames_span(const Trav& tr, int j0, int len) {
    FramesSpan s;
    s.j0 = j0;
    s.p_first = j0 / tr.nf;
    s.t_first = j0 - s.p_first * tr.nf;
    const int je = min(j0 + len, tr.L) - 1;
    s.p_last = je / tr.nf;
    s.t_last = je - s.p_last * tr.nf;
    return s;
}
// `sub` of NL lanes: global row (memory order) -> shared row dst[0..len) (traversal order).  All loads are issued before
// the first store (a short-lived CTA must not serialise its global latencies): kFramesMaxNf frames x KP pixels per lane,
// statically unrolled, so the route serves nframes <= kFramesMaxNf with at most NL * KP pixels of a frame per segment
// (frames_staging_ok); anything else takes the element-wise route.
constexpr int kFramesMaxNf = 8;
template <int NL, int KP>
__host__ __device__ constexpr bool frames_fits(int nf, int len) { return nf >= 2 && nf <= kFramesMaxNf && (len / nf + 2) <= NL * KP; }

// Per-frame bookkeeping of one lane, shared by every tensor it gathers / scatters: first pixel, number of further pixels
// (cnt < 0: none), element offset in the global row and in the traversal-ordered shared row.  Per element that leaves one
// compare and constant offsets from two per-frame base pointers.
template <int NL, int NF>
struct FramesLanePlan {
    int cnt[NF], goff[NF], soff[NF];
    __device__ __forceinline__ FramesLanePlan(const FramesSpan& s, const Trav& tr, int sub) {
#pragma unroll
        for (int t = 0; t < NF; ++t) {
            const int plo = s.p_first + (t < s.t_first ? 1 : 0), phi = s.p_last - (t > s.t_last ? 1 : 0);
            const int p = plo + sub;
            cnt[t] = (t < tr.nf && s.j0 < tr.L) ? phi - p : -1;
            goff[t] = t * tr.hw + p;
            soff[t] = p * tr.nf + t - s.j0;
        }
    }
};

template <typename T, int NL, int KP, int NF>
struct FramesGather {
    T v[NF][KP];
    __device__ __forceinline__ void load(const T* __restrict__ row, const FramesLanePlan<NL, NF>& pl) {
#pragma unroll
        for (int t = 0; t < NF; ++t) {
            const T* src = row + pl.goff[t];
#pragma unroll
            for (int k = 0; k < KP; ++k)
                if (k * NL <= pl.cnt[t]) v[t][k] = src[k * NL];
        }
    }
    __device__ __forceinline__ void store(T* __restrict__ dst, const FramesLanePlan<NL, NF>& pl, int nf) const {
#pragma unroll
        for (int t = 0; t < NF; ++t) {
            T* d = dst + pl.soff[t];
#pragma unroll
            for (int k = 0; k < KP; ++k)
                if (k * NL <= pl.cnt[t]) d[k * NL * nf] = v[t][k];
        }
    }
};

// NF: static bound on nframes (5 covers Vivim's clips with 3/8 fewer registers in flight than kFramesMaxNf)
template <typename T, int NL, int KP, int NF>
__device__ __forceinline__ void frames_gather_n(T* __restrict__ dst, const T* __restrict__ row, const FramesSpan& s,
                                                const Trav& tr, int sub) {
    const FramesLanePlan<NL, NF> pl(s, tr, sub);
    FramesGather<T, NL, KP, NF> g;
    g.load(row, pl);
    g.store(dst, pl, tr.nf);
}
template <typename T, int NL, int KP>
__device__ __forceinline__ void frames_gather(T* __restrict__ dst, const T* __restrict__ row, const FramesSpan& s,
                                              const Trav& tr, int sub) {
    if (tr.nf <= 5) frames_gather_n<T, NL, KP, 5>(dst, row, s, tr, sub);
    else frames_gather_n<T, NL, KP, kFramesMaxNf>(dst, row, s, tr, sub);
}

// shared row src[0..len) (traversal order) -> global row (memory order)
template <typename T, int NL, int KP, int NF>
__device__ __forceinline__ void frames_scatter_n(T* __restrict__ row, const T* __restrict__ src, const FramesLanePlan<NL, NF>& pl,
                                                 int nf) {
#pragma unroll
    for (int t = 0; t < NF; ++t) {
        T* d = row + pl.goff[t];
        const T* sp = src + pl.soff[t];
#pragma unroll
        for (int k = 0; k < KP; ++k)
            if (k * NL <= pl.cnt[t]) d[k * NL] = sp[k * NL * nf];
    }
}
template <typename T, int NL, int KP>
__device__ __forceinline__ void frames_scatter(T* __restrict__ row, const T* __restrict__ src, const FramesSpan& s,
                                               const Trav& tr, int sub) {
    if (tr.nf <= 5) {
        const FramesLanePlan<NL, 5> pl(s, tr, sub);
        frames_scatter_n<T, NL, KP, 5>(row, src, pl, tr.nf);
    } else {
        const FramesLanePlan<NL, kFramesMaxNf> pl(s, tr, sub);
        frames_scatter_n<T, NL, KP, kFramesMaxNf>(row, src, pl, tr.nf);
    }
}

// ---------------------------------------------------------------- deferred unpacking
// Raw8 holds the 8 positions of a row segment as loaded (16 or 32 bytes of registers), so that a
// kernel can issue ALL of its global loads first and convert later: the DRAM latency of a
// short-lived CTA is then exposed once, not once per tensor.
template <typename T, bool kVec> struct Raw8;
template <typename T> struct Raw8<T, true> {
    uint4 lo, hi;   // hi only for 4-byte T
    __device__ __forceinline__ void load(const T* __restrict__ row, int t0, int L) {
        lo = hi = make_uint4(0, 0, 0, 0);
        if (t0 >= 0 && t0 < L) {
            lo = __ldg(reinterpret_cast<const uint4*>(row + t0));
            if (sizeof(T) == 4) hi = __ldg(reinterpret_cast<const uint4*>(row + t0) + 1);
        }
    }
    // 8 positions of a staged (traversal-ordered, 16-byte aligned) shared-memory row; positions >= L read as 0
    __device__ __forceinline__ void load_staged(const T* srow, int j, int j0, int L) {
        lo = hi = make_uint4(0, 0, 0, 0);
        if (j0 >= 0 && j0 < L) {
            lo = *reinterpret_cast<const uint4*>(srow + j);
            if (sizeof(T) == 4) hi = *(reinterpret_cast<const uint4*>(srow + j) + 1);
        }
    }
    // positions [j0, j0+8) in traversal order (all inside or all outside [0, L): L % 8 == 0, j0 % 8 == 0)
    __device__ __forceinline__ void load_trav(const T* __restrict__ row, int j0, const Trav& tr) {
        if (tr.mode == VV_DIR_FWD) { load(row, j0, tr.L); return; }
        lo = hi = make_uint4(0, 0, 0, 0);
        if (j0 < 0 || j0 >= tr.L) return;
        if (tr.mode == VV_DIR_REV) {
            const T* p = row + (tr.L - 8 - j0);          // the same 8 tokens, ascending memory order
            const uint4 a = __ldg(reinterpret_cast<const uint4*>(p));
            if (sizeof(T) == 4) {
                const uint4 b = __ldg(reinterpret_cast<const uint4*>(p) + 1);
                lo = make_uint4(b.w, b.z, b.y, b.x);
                hi = make_uint4(a.w, a.z, a.y, a.x);
            } else {
                lo = make_uint4(swap_halves(a.w), swap_halves(a.z), swap_halves(a.y), swap_halves(a.x));
            }
            return;
        }
        int p = j0 / tr.nf, t = j0 - p * tr.nf;
        uint32_t w[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const T* q = row + (int64_t)t * tr.hw + p;
            if (sizeof(T) == 4) w[i] = __ldg(reinterpret_cast<const uint32_t*>(q));
            else w[i] = __ldg(reinterpret_cast<const unsigned short*>(q));
            if (++t == tr.nf) { t = 0; ++p; }
        }
        if (sizeof(T) == 4) {
            lo = make_uint4(w[0], w[1], w[2], w[3]);
            hi = make_uint4(w[4], w[5], w[6], w[7]);
        } else {
            lo = make_uint4(w[0] | (w[1] << 16), w[2] | (w[3] << 16), w[4] | (w[5] << 16), w[6] | (w[7] << 16));
        }
    }
    __device__ __forceinline__ void unpack(float (&v)[8]) const {
        if (sizeof(T) == 4) {
            v[0] = __uint_as_float(lo.x); v[1] = __uint_as_float(lo.y); v[2] = __uint_as_float(lo.z); v[3] = __uint_as_float(lo.w);
            v[4] = __uint_as_float(hi.x); v[5] = __uint_as_float(hi.y); v[6] = __uint_as_float(hi.z); v[7] = __uint_as_float(hi.w);
        } else {
            load8_plain<T>(reinterpret_cast<const T*>(&lo), v);
        }
    }
    // verbatim 8 x T to a 16-byte aligned shared address
    __device__ __forceinline__ void store_raw(T* dst) const {
        reinterpret_cast<uint4*>(dst)[0] = lo;
        if (sizeof(T) == 4) reinterpret_cast<uint4*>(dst)[1] = hi;
    }
    // Hide the register contents from the optimiser: a later unpack() is recomputed from the packed
    // registers instead of keeping the 8 unpacked values alive in between.
    __device__ __forceinline__ void keep_packed() {
        asm volatile("" : "+r"(lo.x), "+r"(lo.y), "+r"(lo.z), "+r"(lo.w));
        if (sizeof(T) == 4) asm volatile("" : "+r"(hi.x), "+r"(hi.y), "+r"(hi.z), "+r"(hi.w));
    }
};
template <typename T> struct Raw8<T, false> {
    float f[8];
    __device__ __forceinline__ void load(const T* __restrict__ row, int t0, int L) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int t = t0 + i;
            f[i] = (t >= 0 && t < L) ? to_f32<T>(row[t]) : 0.f;
        }
    }
    __device__ __forceinline__ void load_trav(const T* __restrict__ row, int j0, const Trav& tr) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int j = j0 + i;
            f[i] = (j >= 0 && j < tr.L) ? to_f32<T>(row[tr.mem(j)]) : 0.f;
        }
    }
    __device__ __forceinline__ void load_staged(const T* srow, int j, int j0, int L) {   // element-wise twin (unused route)
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] = (j0 >= 0 && j0 + i < L) ? to_f32<T>(srow[j + i]) : 0.f;
    }
    __device__ __forceinline__ void unpack(float (&v)[8]) const {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = f[i];
    }
    __device__ __forceinline__ void store_raw(T* dst) const { store8_vec<T>(dst, f); }
    __device__ __forceinline__ void keep_packed() {}
};

// ---------------------------------------------------------------- math (fast-math forms, like the
// reference build's --use_fast_math: mamba/setup.py:145, causal-conv1d/setup.py:143)
__device__ __forceinline__ float sigmoid_f(float v) { return __fdividef(1.f, 1.f + __expf(-v)); }

// F.softplus with threshold 20 (selective_scan_fwd_kernel.cuh:153-156: log1pf(expf(v)) for v <= 20).
// log1p(e) is taken by series for small e and by lg2 otherwise: 1 MUFU.EX2 + 1 MUFU.LG2 + ~6 FMA,
// relative error < 3e-6 everywhere.  Every kernel uses this one form so that the forward pass and
// the backward recomputation see bit-identical dt.
__device__ __forceinline__ float softplus_f(float v) {
    const float e = __expf(v);
    const float small = e * fmaf(e, fmaf(e, 0.33333333f, -0.5f), 1.f);
    const float big = __logf(1.f + e);
    const float r = e < 0.02f ? small : big;
    return v <= 20.f ? r : v;
}

// Programmatic dependent launch (griddepcontrol): a kernel launched with the PDL attribute may begin
// before its stream predecessor has finished.  pdl_trigger() lets the successor's CTAs be
// scheduled as soon as every CTA of this grid has been scheduled and called it; pdl_wait() blocks
// until the predecessor grid has completed and its memory is visible.  Both are no-ops for a
// kernel launched without the attribute / without a PDL successor.
// L2 prefetch of the cache line holding `p`: a hint without architectural effect, so it may run ahead of pdl_wait() on
// tensors the predecessor kernel may still be writing (the later real load reads whatever L2 holds then).
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ---------------------------------------------------------------- TMA (cp.async.bulk.tensor) + mbarrier
// One elected thread arms an mbarrier with the byte count of the tile and issues the bulk tensor copy; the copy engine
// writes the box into shared memory and completes the barrier; every thread then waits on its phase.
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");   // make the init visible to the async proxy
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, unsigned parity) {
    unsigned ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// bounded wait: a copy that never lands must trap, not hang the GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    for (int it = 0; it < (1 << 24); ++it)
        if (mbar_try_wait(bar, parity)) return;
    __trap();
}
// 2-D box at (x = position, y = row) of the tensor described by `tmap` -> shared `dst` (128-byte aligned)
__device__ __forceinline__ void tma_load_2d(void* dst, const void* tmap, int x, int y, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(tmap), "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(x), "r"(y)
                 : "memory");
}

// packed fp32 pairs (FFMA2 / FMUL2 / FADD2 of sm_100): one issue slot for two lanes of fp32 math
__device__ __forceinline__ float2 mul2(float2 x, float2 y) { return __fmul2_rn(x, y); }
__device__ __forceinline__ float2 add2(float2 x, float2 y) { return __fadd2_rn(x, y); }
__device__ __forceinline__ float2 fma2(float2 x, float2 y, float2 z) { return __ffma2_rn(x, y, z); }

// two packed 16-bit elements -> fp32 pair
template <typename T> __device__ __forceinline__ float2 unpack_pair(uint32_t w);
template <> __device__ __forceinline__ float2 unpack_pair<__nv_bfloat16>(uint32_t w) {
    return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
}
template <> __device__ __forceinline__ float2 unpack_pair<float>(uint32_t w) { return make_float2(__uint_as_float(w), 0.f); }   // never used
template <> __device__ __forceinline__ float2 unpack_pair<__half>(uint32_t w) {
    return __half22float2(*reinterpret_cast<const __half2*>(&w));
}
// fp32 pair -> two packed 16-bit elements, and the packed add (HADD2 / HADD2.BF16: one issue slot for two sums)
template <typename T> __device__ __forceinline__ uint32_t pack_pair(float2 v);
template <> __device__ __forceinline__ uint32_t pack_pair<__nv_bfloat16>(float2 v) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(v.x, v.y);
    return *reinterpret_cast<const uint32_t*>(&h);
}
template <> __device__ __forceinline__ uint32_t pack_pair<__half>(float2 v) {
    const __half2 h = __floats2half2_rn(v.x, v.y);
    return *reinterpret_cast<const uint32_t*>(&h);
}
template <> __device__ __forceinline__ uint32_t pack_pair<float>(float2 v) { return __float_as_uint(v.x); }   // never used
template <typename T> __device__ __forceinline__ uint32_t add_pairs(uint32_t a, uint32_t b);
template <> __device__ __forceinline__ uint32_t add_pairs<__nv_bfloat16>(uint32_t a, uint32_t b) {
    const __nv_bfloat162 r = __hadd2(*reinterpret_cast<const __nv_bfloat162*>(&a), *reinterpret_cast<const __nv_bfloat162*>(&b));
    return *reinterpret_cast<const uint32_t*>(&r);
}
template <> __device__ __forceinline__ uint32_t add_pairs<__half>(uint32_t a, uint32_t b) {
    const __half2 r = __hadd2(*reinterpret_cast<const __half2*>(&a), *reinterpret_cast<const __half2*>(&b));
    return *reinterpret_cast<const uint32_t*>(&r);
}
template <> __device__ __forceinline__ uint32_t add_pairs<float>(uint32_t a, uint32_t) { return a; }   // never used
// a pair of adjacent elements (4-byte aligned for 16-bit T, 8-byte for fp32)
template <typename T> __device__ __forceinline__ float2 load_pair(const T* p);
template <> __device__ __forceinline__ float2 load_pair<float>(const float* p) { return *reinterpret_cast<const float2*>(p); }
template <> __device__ __forceinline__ float2 load_pair<__nv_bfloat16>(const __nv_bfloat16* p) {
    return unpack_pair<__nv_bfloat16>(*reinterpret_cast<const uint32_t*>(p));
}
template <> __device__ __forceinline__ float2 load_pair<__half>(const __half* p) {
    return unpack_pair<__half>(*reinterpret_cast<const uint32_t*>(p));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace vv
