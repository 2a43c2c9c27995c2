// scan_seq.cuh -- time-sequential selective-scan kernels for sm_100a.
//
// Used for everything that only needs the recurrence in ONE direction: the segment aggregates
// (forward and reverse) and the whole forward pass.  Layout:
//
//   * the sequence is cut into SEGMENTS of 64 positions; a CTA (4 warps) owns 32 channels x one
//     segment.  Four adjacent lanes share a channel and split its N states (N/4 each, in registers);
//     every lane walks the 64 positions in order, so a state recurrence is one FMA chain and there
//     is no cross-lane scan and no block barrier on the data path:
//         per (channel, position, state): 1 FMUL (dt*A), 1 MUFU.EX2, 1 FMUL (drive*B), 1 FFMA (h),
//         1 FFMA (y += C h)  +  1/4 LDS.128 each for B and C,
//     plus a 2-step shuffle reduce-scatter of the y partial sums per 8 positions.  That is ~half the
//     instructions of a warp-scan formulation.  Splitting the states over lanes (instead of one
//     lane per channel) quadruples the warps in flight -- 5120 at the B=1 stage-1 shape -- which is
//     what keeps the MUFU pipe fed; these kernels are bound by the MUFU.EX2 rate (16 lanes/clk/SM)
//     and by instruction issue, not by HBM: see DESIGN.md section 5.
//   * pre-pass: each lane reads two 128-bit chunks of its channel's rows straight from global memory
//     (the four lanes of a channel cover 64 contiguous bytes, so every 32-byte sector is fully
//     used), evaluates softplus(delta + bias) and the drive dt*u (or the gated upstream gradient)
//     ONCE per (channel, position), and parks them in fp32 shared tiles [channel][64] whose rows are
//     padded by 16 bytes, so the 128-bit reads of the walk are bank-conflict free.
//   * B / C of the segment are converted once per CTA to fp32 [position][state] shared tiles: every
//     lane of a quad-column reads the same address (broadcast), and the 16-bit -> fp32 conversion
//     is paid once per CTA instead of once per channel.
//   * the forward kernel writes its outputs into a shared tile in the I/O dtype and flushes it with
//     128-bit stores (a row segment of 128 bytes per 8 lanes).
//   * the aggregate kernels need 20 KB and 56 registers, so 9 CTAs (36 warps) are resident per SM and the 1280-CTA
//     grid of the B=1 stage-1 shape is a single wave; measured against the MUFU.EX2 rate they run at 77-89 % of
//     peak (bench.py "mufu_roofline").  The forward kernel (35 KB: B and C tiles plus the output staging) is bound
//     by shared-memory wavefronts instead, 11.75 per warp and position against 6 in the aggregate kernels.
//   * launched as programmatic dependents: a kernel's loads, pre-pass and tile fills overlap the tail of its
//     predecessor; its first access to what the predecessor wrote (or may still read) comes after the wait.  The FIRST
//     kernel of a call (seg_agg_kernel) is the exception: its predecessor is the caller's, so it waits before it reads
//     anything and only prefetches to L2 ahead of the wait.
//   * direction blocks (kGen = true): every load / store goes through a traversal order (Trav); B / C may be rows of
//     x_proj's output; z and the upstream gradient may be shared by the blocks.  kGen = false folds all of it away.
//
//   pass 1  seg_agg_kernel    (P, X) of every (row, segment, state), forward or reverse
//   pass 2  seg_carry_kernel  per row: fold the segment aggregates -> state entering each segment
//                             (forward: chk, saved for the backward; reverse: radj)
//   pass 3  seg_fwd_kernel    forward outputs, seeded by chk
#pragma once

#include "../../include/vivim_b200.h"
#include "common.cuh"

namespace vv {

constexpr int kSeg = 64;                      // positions per segment
constexpr int kSegRows = 32;                  // channels per CTA
constexpr int kSegThreads = 4 * kSegRows;     // 4 lanes per channel -> 2 warps
constexpr int kF32Pitch = kSeg * 4 + 16;      // bytes per row of an fp32 [channel][64] tile (17 x 16)
constexpr int kSegKP = 4;                     // FRAMES staging: pixels of a frame per lane (4 lanes per channel: 16 pixels, nframes >= 5)

template <typename T> struct SegTile {
    static constexpr int kPitch = kSeg * (int)sizeof(T) + 16;   // bytes per row of an I/O-dtype tile
};

// 8 consecutive positions of a staged I/O-dtype row -> fp32
template <typename T>
__device__ __forceinline__ void tile_read8(const unsigned char* __restrict__ row, int c, float (&v)[8]) {
    load8_plain<T>(reinterpret_cast<const T*>(row) + c * 8, v);
}

// NQ consecutive fp32 states of one position (16-byte aligned when NQ % 4 == 0, 8-byte when NQ == 2)
template <int NQ>
__device__ __forceinline__ void load_states(const float* __restrict__ p, float (&v)[NQ]) {
    if (NQ % 4 == 0) {
#pragma unroll
        for (int k = 0; k < NQ / 4; ++k) {
            const float4 x = reinterpret_cast<const float4*>(p)[k];
            v[4 * k] = x.x; v[4 * k + 1] = x.y; v[4 * k + 2] = x.z; v[4 * k + 3] = x.w;
        }
    } else {
        const float2 x = *reinterpret_cast<const float2*>(p);
        v[0] = x.x; v[1] = x.y;
    }
}

// the same NQ states as NQ / 2 register pairs, the operands of the packed fp32x2 instructions (FMUL2 / FFMA2)
template <int NQ>
__device__ __forceinline__ void load_state_pairs(const float* __restrict__ p, float2 (&v)[NQ / 2]) {
    static_assert(NQ % 2 == 0, "states per lane come in pairs");
    if (NQ % 4 == 0) {
#pragma unroll
        for (int k = 0; k < NQ / 4; ++k) {
            const float4 x = reinterpret_cast<const float4*>(p)[k];
            v[2 * k] = make_float2(x.x, x.y);
            v[2 * k + 1] = make_float2(x.z, x.w);
        }
    } else {
        v[0] = *reinterpret_cast<const float2*>(p);
    }
}

// decays of a state pair: exp2(dt * A * log2 e)
__device__ __forceinline__ float2 decay2(float2 dt2, float2 A2) {
    const float2 x = mul2(dt2, A2);
    return make_float2(exp2f(x.x), exp2f(x.y));
}

// Four consecutive states of one token, from any (state stride, sequence stride) layout: the (B,G,N,L) tensors of the
// reference op (ns = L, ls = 1) or rows of x_proj's GEMM output x_dbl (ns = 1, ls = R+2N; reference:
// selective_scan_interface.py:187-207 transposes those into (B,1,N,L) first).  Kept packed until store time.
template <typename T>
struct StateQuad {                    // 16-bit element types
    uint32_t w[2];
    __device__ __forceinline__ void load(const T* __restrict__ base, int64_t ns, int64_t ls, int N, int n0, int pos,
                                         bool valid, const Trav& tr) {
        w[0] = w[1] = 0u;
        if (!valid || pos >= tr.L || n0 >= N) return;
        const T* p = base + (int64_t)tr.mem(pos) * ls + (int64_t)n0 * ns;
        if (ns == 1 && n0 + 4 <= N && reinterpret_cast<uintptr_t>(p) % 8 == 0) {
            const uint2 v = __ldg(reinterpret_cast<const uint2*>(p));
            w[0] = v.x; w[1] = v.y;
            return;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (n0 + k < N)
                w[k >> 1] |= (uint32_t)__ldg(reinterpret_cast<const unsigned short*>(p + k * ns)) << (16 * (k & 1));
    }
    __device__ __forceinline__ float4 unpack() const {
        const float2 a = unpack_pair<T>(w[0]), b = unpack_pair<T>(w[1]);
        return make_float4(a.x, a.y, b.x, b.y);
    }
};
template <>
struct StateQuad<float> {
    float4 v;
    __device__ __forceinline__ void load(const float* __restrict__ base, int64_t ns, int64_t ls, int N, int n0, int pos,
                                         bool valid, const Trav& tr) {
        v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (!valid || pos >= tr.L || n0 >= N) return;
        const float* p = base + (int64_t)tr.mem(pos) * ls + (int64_t)n0 * ns;
        if (ns == 1 && n0 + 4 <= N && reinterpret_cast<uintptr_t>(p) % 16 == 0) {
            v = __ldg(reinterpret_cast<const float4*>(p));
            return;
        }
        if (n0 + 0 < N) v.x = __ldg(p);
        if (n0 + 1 < N) v.y = __ldg(p + ns);
        if (n0 + 2 < N) v.z = __ldg(p + 2 * ns);
        if (n0 + 3 < N) v.w = __ldg(p + 3 * ns);
    }
    __device__ __forceinline__ float4 unpack() const { return v; }
};

// B or C of the segment -> fp32 tile[position][NB] (states >= N are 0).  Split in two so that the loads can be issued
// early and consumed late.  Two routes, chosen per CTA: `rows` = the (.., N, L) layout walked left to right (one 128-bit
// load covers 8 positions of a state row), `quads` = anything else (position-major rows, reversed or frame-interleaved
// traversal): (position, 4 states) items, state index fastest so that a row of x_dbl is read by adjacent lanes.
template <typename T, bool kVec, int NB>
struct StateTileLoader {
    static constexpr int kChunks = kSeg / 8;
    static constexpr int kPer = (NB * kChunks + kSegThreads - 1) / kSegThreads;
    static constexpr int kQuads = kSeg * (NB / 4);
    static constexpr int kPerQ = (kQuads + kSegThreads - 1) / kSegThreads;
    Raw8<T, kVec> raw[kPer];
    StateQuad<T> quad[kPerQ];
    bool rows;
    __device__ __forceinline__ void load(const T* __restrict__ base, int64_t ns, int64_t ls, int N, int t0, const Trav& tr,
                                         bool gen = true) {
        rows = !gen || (ls <= 1 && tr.mode == VV_DIR_FWD);
        if (rows) {
#pragma unroll
            for (int j = 0; j < kPer; ++j) {
                const int idx = threadIdx.x + j * kSegThreads;
                const int c = idx / NB, n = idx - c * NB;   // n fastest: the transposed shared stores spread over banks
                raw[j].load(base + (n < N ? n : 0) * ns, (n < N && c < kChunks) ? t0 + c * 8 : tr.L, tr.L);
            }
        } else {
#pragma unroll
            for (int j = 0; j < kPerQ; ++j) {
                const int idx = threadIdx.x + j * kSegThreads;
                const int pos = idx / (NB / 4), n4 = idx - pos * (NB / 4);
                quad[j].load(base, ns, ls < 1 ? 1 : ls, N, n4 * 4, t0 + pos, idx < kQuads, tr);
            }
        }
    }
    __device__ __forceinline__ void store(float* __restrict__ tile) const {
        if (rows) {
#pragma unroll
            for (int j = 0; j < kPer; ++j) {
                const int idx = threadIdx.x + j * kSegThreads;
                const int c = idx / NB, n = idx - c * NB;
                if (c < kChunks) {
                    float v[8];
                    raw[j].unpack(v);
#pragma unroll
                    for (int i = 0; i < 8; ++i) tile[(c * 8 + i) * NB + n] = v[i];
                }
            }
        } else {
#pragma unroll
            for (int j = 0; j < kPerQ; ++j) {
                const int idx = threadIdx.x + j * kSegThreads;
                if (idx < kQuads) reinterpret_cast<float4*>(tile)[idx] = quad[j].unpack();   // tile[pos * NB + n4 * 4]
            }
        }
    }
};

struct SegCoord {
    int b, g, d0, nrows, seg, t0;   // d0: first channel of the CTA, nrows: valid channels in the CTA
    Trav tr;                        // traversal order of the CTA's direction block
};

// traversal order of group g (groups are split evenly over the direction blocks).  kGen = false: the launch is the
// plain reference op (one left-to-right direction, (.., N, L) B / C, no shared gate rows) -- everything below folds to
// constants and the kernels compile to the lean single-direction code.
template <bool kGen>
__device__ __forceinline__ Trav group_trav(const vv_scan_args& a, int g) {
    Trav tr;
    if (!kGen) {
        tr.mode = VV_DIR_FWD;
        tr.L = a.seqlen;
        tr.nf = 1;
        tr.hw = a.seqlen;
        return tr;
    }
    const int ndirs = a.ndirs > 1 ? a.ndirs : 1;
    tr.mode = a.dir_mode[g / (a.ngroups / ndirs)];
    tr.L = a.seqlen;
    tr.nf = a.nframes > 0 ? a.nframes : 1;
    tr.hw = a.seqlen / tr.nf;
    return tr;
}

// channel row of the tensors shared by the direction blocks (z, dout)
template <bool kGen>
__device__ __forceinline__ int gate_row(const vv_scan_args& a, int d) { return (kGen && a.gate_rows > 0) ? d % a.gate_rows : d; }

template <bool kGen>
__device__ __forceinline__ SegCoord seg_coord(const vv_scan_args& a) {
    SegCoord c;
    const int dpg = a.dim / a.ngroups;
    const int blocks_per_group = (dpg + kSegRows - 1) / kSegRows;
    c.seg = blockIdx.x;
    c.t0 = c.seg * kSeg;
    c.b = blockIdx.z;
    c.g = blockIdx.y / blocks_per_group;
    const int off = (blockIdx.y - c.g * blocks_per_group) * kSegRows;
    c.d0 = c.g * dpg + off;
    c.nrows = min(kSegRows, dpg - off);
    c.tr = group_trav<kGen>(a, c.g);
    return c;
}

// Pre-pass of lane (channel row, quad q): chunks {q, q + 4} of the row.  load() issues the global
// loads, finish() converts and writes the shared tiles:
//   f_dt   <- softplus?(delta + bias), exactly 0 outside [0, L) (padding = scan identity)
//   f_cf   <- kKind 0: dt * u (forward drive)          kKind 1: dout * silu(z) (gated upstream grad)
//   raw_a / raw_b (optional, I/O dtype tiles) <- verbatim copies of the `cf` row and of the z row
template <typename T, bool kVec, int kKind>
struct SegPrepass {
    Raw8<T, kVec> r_dt[2], r_cf[2], r_z[2];
    int t[2];
    // s_dt / s_cf / s_z: 16-byte aligned shared rows of kSeg elements of T (or NULL) that the FRAMES order may use to
    // stage this channel's rows (coalesced gather by the 4 lanes of the channel, see frames_gather); they may alias the
    // tiles that finish() writes later.  Must be called by whole warps.
    __device__ __forceinline__ void load(const T* __restrict__ g_dt, const T* __restrict__ g_cf, const T* __restrict__ g_z,
                                         bool live, int q, int t0, const Trav& tr, T* s_dt = nullptr, T* s_cf = nullptr,
                                         T* s_z = nullptr) {
        if (kVec && tr.mode == VV_DIR_FRAMES && s_dt != nullptr && s_cf != nullptr && frames_fits<4, kSegKP>(tr.nf, kSeg)) {
            const FramesSpan sp = frames_span(tr, t0, kSeg);
            if (live) {
                frames_gather<T, 4, kSegKP>(s_dt, g_dt, sp, tr, q);        // one row at a time: 56-register kernels
                frames_gather<T, 4, kSegKP>(s_cf, g_cf, sp, tr, q);
                if (g_z && s_z) frames_gather<T, 4, kSegKP>(s_z, g_z, sp, tr, q);
            }
            __syncwarp();
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                t[k] = live ? t0 + (q + 4 * k) * 8 : tr.L;
                r_dt[k].load_staged(s_dt, (q + 4 * k) * 8, t[k], tr.L);
                r_cf[k].load_staged(s_cf, (q + 4 * k) * 8, t[k], tr.L);
                if (g_z) {
                    if (s_z) r_z[k].load_staged(s_z, (q + 4 * k) * 8, t[k], tr.L);
                    else r_z[k].load_trav(g_z, t[k], tr);
                }
            }
            __syncwarp();   // every lane has its rows in registers before the staging space is reused
            return;
        }
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            t[k] = live ? t0 + (q + 4 * k) * 8 : tr.L;   // dead rows read as padding
            r_dt[k].load_trav(g_dt, t[k], tr);
            r_cf[k].load_trav(g_cf, t[k], tr);
            if (g_z) r_z[k].load_trav(g_z, t[k], tr);
        }
    }
    __device__ __forceinline__ void finish(bool has_z, unsigned char* __restrict__ f_dt, unsigned char* __restrict__ f_cf,
                                           unsigned char* __restrict__ raw_a, unsigned char* __restrict__ raw_b,
                                           int q, int L, float bias, bool sp) const {
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int ch = q + 4 * k;
            float dt[8], cf[8], zv[8];
            r_dt[k].unpack(dt);
            r_cf[k].unpack(cf);
            if (has_z) r_z[k].unpack(zv);
            if (raw_a) r_cf[k].store_raw(reinterpret_cast<T*>(raw_a) + ch * 8);
            if (raw_b && has_z) r_z[k].store_raw(reinterpret_cast<T*>(raw_b) + ch * 8);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float v = dt[i] + bias;
                if (sp) v = softplus_f(v);
                dt[i] = (t[k] + i < L) ? v : 0.f;
            }
            if (kKind == 0) {
#pragma unroll
                for (int i = 0; i < 8; ++i) cf[i] *= dt[i];
            } else if (has_z) {
#pragma unroll
                for (int i = 0; i < 8; ++i) cf[i] *= zv[i] * sigmoid_f(zv[i]);
            }
            store8_vec<float>(reinterpret_cast<float*>(f_dt) + ch * 8, dt);
            store8_vec<float>(reinterpret_cast<float*>(f_cf) + ch * 8, cf);
        }
    }
};

// L2 hints for the B or C values of one segment (memory positions [m0, m0 + kSeg)): state-major rows -- thread `id` < N
// takes state row `id` -- or, position-major, the contiguous span of the segment's x_dbl rows, 128 bytes per thread.
template <typename T>
__device__ __forceinline__ void prefetch_state_tile(const T* __restrict__ base, int64_t ns, int64_t ls, int N, int m0, int L,
                                                    int id, int nthreads) {
    if (ls <= 1) {
        if (id < N) {
#pragma unroll
            for (int h = 0; h < (int)sizeof(T) / 2; ++h)
                if (m0 + h * (kSeg / 2) < L) prefetch_l2(base + id * ns + m0 + h * (kSeg / 2));
        }
    } else {
        const char* p0 = reinterpret_cast<const char*>(base + (int64_t)m0 * ls);
        const int64_t bytes = ((int64_t)(min(kSeg, L - m0) - 1) * ls + N) * (int64_t)sizeof(T);
        for (int64_t off = (int64_t)id * 128; off < bytes; off += 128 * nthreads) prefetch_l2(p0 + off);
    }
}

// ================================================================ pass 1: segment aggregates
// Lane (channel r, quad q) owns states [q*NQ, (q+1)*NQ) of channel r, NQ = NB/4.
// smem: [f32 dt][f32 coef][state tile fp32]
// kRev = false: (P, X) of h_t = a_t h_{t-1} + dt_t B_t u_t over the segment      (uses u, B)
// kRev = true : (P, X) of e_t = a_t (e_{t+1} + g_t C_t), g = dout*silu(z)        (uses dout, z, C)
//               e_t = a_t r_t is the adjoint of h_t pushed through its own decay: its segment
//               aggregate has the SAME decay product as the forward one and needs nothing from
//               the neighbouring segment.
// Register budget: the 128-bit variants with <= 16 states fit 56 registers (9 CTAs / SM: one wave at the B = 1 stage-1
// shape); the element-wise, fp32 and 32-state variants would spill there and get 96.
template <typename T, bool kVec, int NB, bool kRev, bool kGen>
__global__ void __launch_bounds__(kSegThreads, (kVec && NB <= 16 && sizeof(T) == 2) ? 9 : 5) seg_agg_kernel(const vv_scan_args a) {
    constexpr int NQ = NB / 4;
    extern __shared__ __align__(16) unsigned char smem[];
    const int L = a.seqlen, N = a.dstate;
    const SegCoord c = seg_coord<kGen>(a);
    unsigned char* f_dt = smem;
    unsigned char* f_cf = f_dt + kSegRows * kF32Pitch;
    float* t_m = reinterpret_cast<float*>(f_cf + kSegRows * kF32Pitch);
    // FIRST kernel of vv_scan_fwd / vv_scan_bwd: its stream predecessor is whatever produced the caller's tensors
    // (a GEMM, a conv, another scan) and may itself have released its dependents early, so nothing the caller provides
    // -- inputs AND parameters -- is read before the predecessor has completed.  Only index arithmetic runs ahead.  The
    // dependents (carry, main) are released after that point, so their own pre-wait loads are safe as well.
    const int r = threadIdx.x >> 2, q = threadIdx.x & 3;
    const bool live = r < c.nrows;
    const int d = c.d0 + (live ? r : 0);
    const int dg = gate_row<kGen>(a, d);
    const T* g_dt = reinterpret_cast<const T*>(a.delta) + c.b * a.delta_bs + d * a.delta_ds;
    // While the predecessor drains: pull this CTA's rows towards L2 (hints only, see prefetch_l2).  One lane per channel
    // covers the 64-position row segment (one or two 128-byte lines); FRAMES rows are not contiguous and are skipped.
    if (live && q == 0 && c.tr.mode != VV_DIR_FRAMES && c.t0 < L) {
        const int m0 = c.tr.mode == VV_DIR_REV ? max(L - kSeg - c.t0, 0) : c.t0;
        const T* g_cf0 = kRev ? reinterpret_cast<const T*>(a.dout) + c.b * a.dout_bs + dg * a.dout_ds
                              : reinterpret_cast<const T*>(a.u) + c.b * a.u_bs + d * a.u_ds;
#pragma unroll
        for (int h = 0; h < (int)sizeof(T) / 2; ++h) {
            if (m0 + h * (kSeg / 2) >= L) break;   // never form an address past the row
            prefetch_l2(g_dt + m0 + h * (kSeg / 2));
            prefetch_l2(g_cf0 + m0 + h * (kSeg / 2));
            if (kRev && a.z) prefetch_l2(reinterpret_cast<const T*>(a.z) + c.b * a.z_bs + dg * a.z_ds + m0 + h * (kSeg / 2));
        }
    }
    const bool first_block = c.d0 == c.g * (a.dim / a.ngroups);   // one CTA per (batch, group, segment) hints B / C
    if (first_block && threadIdx.x >= kSegThreads / 2 && c.tr.mode != VV_DIR_FRAMES && c.t0 < L) {
        // the state tile this kernel itself reads (B forward, C reverse), by the second half of the CTA
        const int m0 = c.tr.mode == VV_DIR_REV ? max(L - kSeg - c.t0, 0) : c.t0;
        const T* base = kRev ? reinterpret_cast<const T*>(a.Cm) + c.b * a.C_bs + c.g * a.C_gs
                             : reinterpret_cast<const T*>(a.Bm) + c.b * a.B_bs + c.g * a.B_gs;
        prefetch_state_tile<T>(base, kRev ? a.C_ns : a.B_ns, kGen ? (kRev ? a.C_ls : a.B_ls) : 1, N, m0, L,
                               threadIdx.x - kSegThreads / 2, kSegThreads / 2);
    }
    pdl_wait();
    pdl_trigger();
    // Rows that only the MAIN kernel of this pass reads -- z and C in the forward, u and B in the backward -- start their
    // way to L2 now (hints): that kernel's CTAs all start at once and would otherwise wait for HBM together.
    if (c.tr.mode != VV_DIR_FRAMES && c.t0 < L) {
        const int m0 = c.tr.mode == VV_DIR_REV ? max(L - kSeg - c.t0, 0) : c.t0;
        if (live && q == 1) {
            const T* row = kRev ? reinterpret_cast<const T*>(a.u) + c.b * a.u_bs + d * a.u_ds
                                : (a.z ? reinterpret_cast<const T*>(a.z) + c.b * a.z_bs + dg * a.z_ds : nullptr);
            if (row != nullptr) {
#pragma unroll
                for (int h = 0; h < (int)sizeof(T) / 2; ++h)
                    if (m0 + h * (kSeg / 2) < L) prefetch_l2(row + m0 + h * (kSeg / 2));
            }
        }
        if (first_block) {   // B / C of the other kind
            const T* base = kRev ? reinterpret_cast<const T*>(a.Bm) + c.b * a.B_bs + c.g * a.B_gs
                                 : reinterpret_cast<const T*>(a.Cm) + c.b * a.C_bs + c.g * a.C_gs;
            prefetch_state_tile<T>(base, kRev ? a.B_ns : a.C_ns, kGen ? (kRev ? a.B_ls : a.C_ls) : 1, N, m0, L, threadIdx.x,
                                   kSegThreads);
        }
    }

    const float bias = a.delta_bias ? a.delta_bias[d] : 0.f;
    const bool sp = a.delta_softplus != 0;
    // every global load of the CTA is issued before the first use, so the latencies overlap
    SegPrepass<T, kVec, kRev ? 1 : 0> pre;
    StateTileLoader<T, kVec, NB> st;
    // FRAMES staging rows: the head of this channel's fp32 tile rows (a 16-bit row needs 128 of the 256 bytes, so z
    // shares the dt row); fp32 I/O stages dt and the coefficient only
    T* s_dt = reinterpret_cast<T*>(f_dt + r * kF32Pitch);
    T* s_cf = reinterpret_cast<T*>(f_cf + r * kF32Pitch);
    T* s_z = sizeof(T) == 2 ? s_dt + kSeg : nullptr;
    if (!kRev) {
        pre.load(g_dt, reinterpret_cast<const T*>(a.u) + c.b * a.u_bs + d * a.u_ds, nullptr, live, q, c.t0, c.tr, s_dt, s_cf, s_z);
        st.load(reinterpret_cast<const T*>(a.Bm) + c.b * a.B_bs + c.g * a.B_gs, a.B_ns, a.B_ls, N, c.t0, c.tr, kGen);
    } else {
        pre.load(g_dt, reinterpret_cast<const T*>(a.dout) + c.b * a.dout_bs + dg * a.dout_ds,
                 a.z ? reinterpret_cast<const T*>(a.z) + c.b * a.z_bs + dg * a.z_ds : nullptr, live, q, c.t0, c.tr, s_dt, s_cf, s_z);
        st.load(reinterpret_cast<const T*>(a.Cm) + c.b * a.C_bs + c.g * a.C_gs, a.C_ns, a.C_ls, N, c.t0, c.tr, kGen);
    }
    // two states per register pair: the recurrences run on the packed fp32x2 pipe (same roundings as the scalar form)
    constexpr int NP = NQ / 2;
    float2 A2[NP], h[NP];
#pragma unroll
    for (int k = 0; k < NP; ++k) {
        const int n = q * NQ + 2 * k;
        A2[k].x = n < N ? a.A[d * a.A_ds + n * a.A_ns] * kLog2e : 0.f;
        A2[k].y = n + 1 < N ? a.A[d * a.A_ds + (n + 1) * a.A_ns] * kLog2e : 0.f;
        h[k] = make_float2(0.f, 0.f);
    }
    float sum_dt = 0.f;
    st.store(t_m);
    pre.finish(kRev && a.z != nullptr, f_dt + r * kF32Pitch, f_cf + r * kF32Pitch, nullptr, nullptr, q, L, bias, sp);
    __syncthreads();
    if (!live) return;
    const float4* my_dt = reinterpret_cast<const float4*>(f_dt + r * kF32Pitch);
    const float4* my_cf = reinterpret_cast<const float4*>(f_cf + r * kF32Pitch);
#pragma unroll 2
    for (int jb = 0; jb < kSeg / 4; ++jb) {
        const int j = kRev ? kSeg / 4 - 1 - jb : jb;
        const float4 d4 = my_dt[j], c4 = my_cf[j];
        const float dts[4] = {d4.x, d4.y, d4.z, d4.w};
        const float cfs[4] = {c4.x, c4.y, c4.z, c4.w};
#pragma unroll
        for (int ii = 0; ii < 4; ++ii) {
            const int i = kRev ? 3 - ii : ii;
            const float dti = dts[i];
            sum_dt += dti;
            const float2 dt2 = make_float2(dti, dti), cf2 = make_float2(cfs[i], cfs[i]);
            float2 m[NP];
            load_state_pairs<NQ>(t_m + (j * 4 + i) * NB + q * NQ, m);
            if (!kRev) {
#pragma unroll
                for (int k = 0; k < NP; ++k) h[k] = fma2(decay2(dt2, A2[k]), h[k], mul2(cf2, m[k]));
            } else {
                // pushed adjoint e_t = a_t r_t:  e_t = a_t (e_{t+1} + g_t C_t)
#pragma unroll
                for (int k = 0; k < NP; ++k) h[k] = mul2(decay2(dt2, A2[k]), fma2(cf2, m[k], h[k]));
            }
        }
    }
    const int S = gridDim.x;
    float2* out = reinterpret_cast<float2*>(a.agg) + (((int64_t)c.b * a.dim + d) * S + c.seg) * N;
    if (kRev && a.zero_accumulators) {
        // The backward's fp32 accumulators are zero-filled here, by the first kernel of vv_scan_bwd and after the
        // dependency wait (whatever used the buffers before is complete), instead of by a memset launch of the caller:
        // the live threads of the first channel block of a group clear the segment's dB / dC (N rows x 64 positions),
        // lane (r, q) of the first segment of the first batch entry its channel's parameter gradients.
        if (c.d0 == c.g * (a.dim / a.ngroups)) {
            const int64_t base = ((int64_t)c.b * a.ngroups + c.g) * N;
            for (int idx = threadIdx.x; idx < N * kSeg; idx += 4 * c.nrows) {
                const int n = idx / kSeg, t = c.t0 + (idx - n * kSeg);
                if (t < L) {
                    a.dB[(base + n) * L + t] = 0.f;
                    a.dC[(base + n) * L + t] = 0.f;
                }
            }
        }
        if (c.seg == 0 && c.b == 0) {
#pragma unroll
            for (int k = 0; k < NQ; ++k)
                if (q * NQ + k < N) a.dA[(int64_t)d * N + q * NQ + k] = 0.f;
            if (q == 0) {
                if (a.dD) a.dD[d] = 0.f;
                if (a.ddelta_bias) a.ddelta_bias[d] = 0.f;
            }
        }
    }
#pragma unroll
    for (int k = 0; k < NP; ++k) {
        const int n = q * NQ + 2 * k;
        if (n < N) out[n] = make_float2(exp2f(A2[k].x * sum_dt), h[k].x);
        if (n + 1 < N) out[n + 1] = make_float2(exp2f(A2[k].y * sum_dt), h[k].y);
    }
}

// ================================================================ pass 2: fold segment aggregates
// A CTA of 512 threads serves `rows_per_cta` rows with `chunks` chunks each (chunks * N * rows_per_cta <= 512:
// one row with 32 chunks for the long stage-1 sequences, many rows per CTA for the short late-stage ones).
// Thread (row, k, n): chunk k of the row's segments (in scan order), state n.  Each thread folds its few segments, the chunk aggregates are combined through shared
// memory, and a second sweep writes the state entering every segment.  Forward: carry[s] = state
// entering segment s (chk), last_state = state after the last one; reverse: carry[s] = adjoint
// entering segment s from the right (radj).  All loads of a thread are independent of its FMA chain.
constexpr int kCarryThreads = 512;
constexpr int kCarryMaxPer = 16;   // segments a thread keeps in registers between the two sweeps

template <bool kRev>
__global__ void __launch_bounds__(kCarryThreads) seg_carry_kernel(const float2* __restrict__ agg, float* __restrict__ carry,
                                                                  float* __restrict__ last_state, const int S, const int N,
                                                                  const int chunks, const int rows_per_cta, const int64_t rows) {
    __shared__ float2 s_chunk[kCarryThreads];
    pdl_trigger();
    pdl_wait();   // segment aggregates of the preceding kernel
    const int rloc = threadIdx.x / (chunks * N);
    const int64_t row = (int64_t)blockIdx.x * rows_per_cta + rloc;
    const int k = (threadIdx.x - rloc * chunks * N) / N, n = threadIdx.x % N;
    const int per = (S + chunks - 1) / chunks;
    const int lo = min(k * per, S), hi = min(lo + per, S);   // this thread's segments, in scan order
    const float2* __restrict__ ag = agg + row * S * N + n;
    float* __restrict__ cr = carry + row * S * N + n;
    const bool active = rloc < rows_per_cta && row < rows;
    const bool cached = per <= kCarryMaxPer;
    float2 v[kCarryMaxPer];
    float P = 1.f, X = 0.f;
    if (active) {
        if (cached) {
#pragma unroll
            for (int j = 0; j < kCarryMaxPer; ++j) {
                const int q = lo + j;
                v[j] = q < hi ? __ldg(ag + (int64_t)(kRev ? S - 1 - q : q) * N) : make_float2(1.f, 0.f);
            }
#pragma unroll
            for (int j = 0; j < kCarryMaxPer; ++j) {
                X = fmaf(v[j].x, X, v[j].y);
                P *= v[j].x;
            }
        } else {
#pragma unroll 4
            for (int q = lo; q < hi; ++q) {
                const float2 w = __ldg(ag + (int64_t)(kRev ? S - 1 - q : q) * N);
                X = fmaf(w.x, X, w.y);
                P *= w.x;
            }
        }
        s_chunk[threadIdx.x] = make_float2(P, X);
    }
    __syncthreads();
    if (!active) return;
    float E = 0.f;
    for (int kk = 0; kk < k; ++kk) {
        const float2 w = s_chunk[(rloc * chunks + kk) * N + n];
        E = fmaf(w.x, E, w.y);
    }
    if (cached) {
#pragma unroll
        for (int j = 0; j < kCarryMaxPer; ++j) {
            const int q = lo + j;
            if (q < hi) cr[(int64_t)(kRev ? S - 1 - q : q) * N] = E;
            E = fmaf(v[j].x, E, v[j].y);
        }
    } else {
#pragma unroll 4
        for (int q = lo; q < hi; ++q) {
            const int s = kRev ? S - 1 - q : q;
            const float2 w = __ldg(ag + (int64_t)s * N);
            cr[(int64_t)s * N] = E;
            E = fmaf(w.x, E, w.y);
        }
    }
    if (!kRev && last_state && hi == S && lo < S) last_state[row * N + n] = E;
}

// ================================================================ pass 3: forward outputs
// smem: [f32 dt][f32 drive][B tile][C tile][raw u -> gated y][raw z -> pre-gate y]
template <typename T, bool kVec, int NB, bool kGen>
__global__ void __launch_bounds__(kSegThreads, (kVec && NB <= 16) ? 7 : 4) seg_fwd_kernel(const vv_scan_args a) {
    constexpr int NQ = NB / 4;
    extern __shared__ __align__(16) unsigned char smem[];
    const int L = a.seqlen, N = a.dstate;
    const SegCoord c = seg_coord<kGen>(a);
    unsigned char* f_dt = smem;
    unsigned char* f_dr = f_dt + kSegRows * kF32Pitch;
    float* t_B = reinterpret_cast<float*>(f_dr + kSegRows * kF32Pitch);
    float* t_C = t_B + kSeg * NB;
    unsigned char* t_u = reinterpret_cast<unsigned char*>(t_C + kSeg * NB);
    unsigned char* t_z = t_u + kSegRows * SegTile<T>::kPitch;

    const int r = threadIdx.x >> 2, q = threadIdx.x & 3;
    const bool live = r < c.nrows;
    const int d = c.d0 + (live ? r : 0);
    const float bias = a.delta_bias ? a.delta_bias[d] : 0.f;
    const float Dv = a.D ? a.D[d] : 0.f;
    const bool sp = a.delta_softplus != 0;
    const int S = gridDim.x;
    // every global load of the CTA is issued before the first use, so the latencies overlap
    SegPrepass<T, kVec, 0> pre;
    StateTileLoader<T, kVec, NB> stB, stC;
    pre.load(reinterpret_cast<const T*>(a.delta) + c.b * a.delta_bs + d * a.delta_ds,
             reinterpret_cast<const T*>(a.u) + c.b * a.u_bs + d * a.u_ds,
             a.z ? reinterpret_cast<const T*>(a.z) + c.b * a.z_bs + gate_row<kGen>(a, d) * a.z_ds : nullptr, live, q, c.t0, c.tr,
             reinterpret_cast<T*>(f_dt + r * kF32Pitch), reinterpret_cast<T*>(t_u + r * SegTile<T>::kPitch),
             reinterpret_cast<T*>(t_z + r * SegTile<T>::kPitch));
    stB.load(reinterpret_cast<const T*>(a.Bm) + c.b * a.B_bs + c.g * a.B_gs, a.B_ns, a.B_ls, N, c.t0, c.tr, kGen);
    stC.load(reinterpret_cast<const T*>(a.Cm) + c.b * a.C_bs + c.g * a.C_gs, a.C_ns, a.C_ls, N, c.t0, c.tr, kGen);
    constexpr int NP = NQ / 2;   // state pairs: packed fp32x2 recurrences, as in seg_agg_kernel
    float2 A2[NP], h[NP];
#pragma unroll
    for (int k = 0; k < NP; ++k) {
        const int n = q * NQ + 2 * k;
        A2[k].x = n < N ? a.A[d * a.A_ds + n * a.A_ns] * kLog2e : 0.f;
        A2[k].y = n + 1 < N ? a.A[d * a.A_ds + (n + 1) * a.A_ns] * kLog2e : 0.f;
    }
    pdl_trigger();   // a dependent launch (the reverse aggregate of a backward that follows directly) may start its loads
    stB.store(t_B);
    stC.store(t_C);
    pre.finish(a.z != nullptr, f_dt + r * kF32Pitch, f_dr + r * kF32Pitch, t_u + r * SegTile<T>::kPitch,
               t_z + r * SegTile<T>::kPitch, q, L, bias, sp);
    pdl_wait();   // chk comes from the carry kernel; the loads, the softplus pre-pass and the tile fills above overlap it
    {
        const float* E = a.chk + (((int64_t)c.b * a.dim + d) * S + c.seg) * N;
#pragma unroll
        for (int k = 0; k < NP; ++k) {
            const int n = q * NQ + 2 * k;
            h[k].x = n < N ? E[n] : 0.f;
            h[k].y = n + 1 < N ? E[n + 1] : 0.f;
        }
    }
    __syncthreads();
    {
        // the raw u / z rows receive the gated / pre-gate output in place: positions {2q, 2q+1} of a
        // chunk are read and later overwritten by the same lane only
        T* my_u = reinterpret_cast<T*>(t_u + r * SegTile<T>::kPitch);
        T* my_z = reinterpret_cast<T*>(t_z + r * SegTile<T>::kPitch);
        const float4* my_dt = reinterpret_cast<const float4*>(f_dt + r * kF32Pitch);
        const float4* my_dr = reinterpret_cast<const float4*>(f_dr + r * kF32Pitch);
        const bool b1 = (q & 2) != 0, b0 = (q & 1) != 0;
#pragma unroll 1
        for (int ch = 0; ch < kSeg / 8; ++ch) {
            // this lane finalises positions {2q, 2q+1} of the chunk: D*u skip and gate, independent of the recurrences
            const int p0 = ch * 8 + 2 * q;
            const float u0 = to_f32<T>(my_u[p0]), u1 = to_f32<T>(my_u[p0 + 1]);
            float g0 = 1.f, g1 = 1.f;
            if (a.z) {
                g0 = to_f32<T>(my_z[p0]);
                g1 = to_f32<T>(my_z[p0 + 1]);
                g0 *= sigmoid_f(g0);
                g1 *= sigmoid_f(g1);
            }
            float y[8];
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const float4 d4 = my_dt[ch * 2 + half], r4 = my_dr[ch * 2 + half];
                const float dts[4] = {d4.x, d4.y, d4.z, d4.w};
                const float drs[4] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int t = ch * 8 + half * 4 + i;
                    const float2 dt2 = make_float2(dts[i], dts[i]), dr2 = make_float2(drs[i], drs[i]);
                    float2 bm[NP], cm[NP];
                    load_state_pairs<NQ>(t_B + t * NB + q * NQ, bm);
                    load_state_pairs<NQ>(t_C + t * NB + q * NQ, cm);
                    float2 acc = make_float2(0.f, 0.f);   // even / odd states; the 4 lanes of the channel are summed below
#pragma unroll
                    for (int k = 0; k < NP; ++k) {
                        h[k] = fma2(decay2(dt2, A2[k]), h[k], mul2(dr2, bm[k]));
                        acc = fma2(cm[k], h[k], acc);
                    }
                    y[half * 4 + i] = acc.x + acc.y;
                }
            }
            // reduce-scatter of the 8 partial sums over the 4 lanes of the channel: lane q ends with {2q, 2q+1}
            float keep[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float send = b1 ? y[j] : y[4 + j];
                const float recv = __shfl_xor_sync(0xffffffffu, send, 2);
                keep[j] = (b1 ? y[4 + j] : y[j]) + recv;
            }
            float fin[2];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const float send = b0 ? keep[j] : keep[2 + j];
                const float recv = __shfl_xor_sync(0xffffffffu, send, 1);
                fin[j] = (b0 ? keep[2 + j] : keep[j]) + recv;
            }
            fin[0] = fmaf(Dv, u0, fin[0]);
            fin[1] = fmaf(Dv, u1, fin[1]);
            if (a.out) { my_z[p0] = from_f32<T>(fin[0]); my_z[p0 + 1] = from_f32<T>(fin[1]); }
            if (a.z) { my_u[p0] = from_f32<T>(fin[0] * g0); my_u[p0 + 1] = from_f32<T>(fin[1] * g1); }
        }
    }
    __syncwarp();   // the four lanes of a channel (same warp) wrote its row; each lane now flushes two chunks of it
    if (live && kVec && c.tr.mode == VV_DIR_FRAMES && frames_fits<4, kSegKP>(c.tr.nf, kSeg)) {
        // the rows hold the outputs in traversal order: the 4 lanes of the channel write them out run by run
        const FramesSpan sp = frames_span(c.tr, c.t0, kSeg);
        if (a.out) frames_scatter<T, 4, kSegKP>(reinterpret_cast<T*>(a.out) + c.b * a.out_bs + d * a.out_ds,
                                                reinterpret_cast<const T*>(t_z + r * SegTile<T>::kPitch), sp, c.tr, q);
        if (a.z) frames_scatter<T, 4, kSegKP>(reinterpret_cast<T*>(a.out_z) + c.b * a.outz_bs + d * a.outz_ds,
                                              reinterpret_cast<const T*>(t_u + r * SegTile<T>::kPitch), sp, c.tr, q);
    } else if (live) {
        const unsigned char* my_u = t_u + r * SegTile<T>::kPitch;
        const unsigned char* my_z = t_z + r * SegTile<T>::kPitch;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int ch = q + 4 * k;
            float v[8];
            if (a.out) {
                tile_read8<T>(my_z, ch, v);
                store8_trav<T, kVec>(reinterpret_cast<T*>(a.out) + c.b * a.out_bs + d * a.out_ds, c.t0 + ch * 8, c.tr, v);
            }
            if (a.z) {
                tile_read8<T>(my_u, ch, v);
                store8_trav<T, kVec>(reinterpret_cast<T*>(a.out_z) + c.b * a.outz_bs + d * a.outz_ds, c.t0 + ch * 8, c.tr, v);
            }
        }
    }
}

}  // namespace vv
