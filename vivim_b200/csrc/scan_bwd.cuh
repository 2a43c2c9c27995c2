// scan_bwd.cuh -- backward main kernel of the fused selective scan for sm_100a.
//
// Replaces selective_scan_bwd_kernel of the reference
// (mamba/csrc/selective_scan/selective_scan_bwd_kernel.cuh:75-489).  The algebra is the reference's:
// the associative operator (a0,b0)o(a1,b1) = (a1 a0, a1 b0 + b1) (selective_scan_common.h:110-115),
// softplus with threshold 20, exp2f(dt * A * log2e), gradient formulas of :279-295 and :439-453.
//
// Design (B200-first, not a port).  The reference walks a whole (batch, channel) row inside one CTA,
// chunk after chunk, last to first, and reduces dB/dC with D-way contended scalar atomics.  Here every
// (batch, 16 channels, 64-position segment) is an independent CTA: the forward state entering the
// segment comes from the checkpoint tensor `chk` written by the forward pass (checkpointed chunk
// states, recomputed in the backward -- a saved `out` is never read), the adjoint entering it from the
// right comes from `radj`, produced by the reverse segment-aggregate + carry passes of scan_seq.cuh.
//
// Inside a CTA (4 warps): a warp owns 4 channels; 8 adjacent lanes share a channel and lane `tb` of
// the group owns positions [8 tb, 8 tb + 8) of the segment in registers (one 128-bit access per
// streamed tensor, 128 contiguous bytes per channel).  For one state n a lane
//   1. evaluates the 8 decays (one MUFU.EX2 each, kept in registers) and the lane-local aggregates
//      of both recurrences (h left->right; the pushed adjoint e_t = a_t r_t right->left, which needs
//      no decay from the neighbouring segment and shares the lane's decay product with h),
//   2. combines the 8 lane aggregates of its channel by two interleaved 3-step shuffle scans,
//   3. sweeps its 8 positions once more producing states, adjoints and every gradient term.
// B and C of the segment are staged once per CTA as fp32 (a 128-bit shared load costs 4 wavefronts per warp
// whether or not its four quarter-warps read the same bytes, so the four channel groups gain nothing from
// reading the same row -- measured).  dB/dC are reduced over the 4 channels of a
// warp by a transposing shuffle reduce-scatter (12 shuffles for 16 values), parked with ONE 128-bit
// store per state in a warp-private tile -- no atomics, no block barrier inside the state loop --
// and leave the CTA summed over its 16 channels as 128-bit red.global.add: D/16-way contention in
// vector units instead of the reference's D-way scalar atomics (selective_scan_bwd_kernel.cuh:298-316).
#pragma once

#include "../../include/vivim_b200.h"
#include "common.cuh"
#include "scan_seq.cuh"

namespace vv {

constexpr int kBwdWarps = 4;
constexpr int kBwdGroups = 4;                          // channels per warp (8 lanes each)
constexpr int kBwdRows = kBwdWarps * kBwdGroups;       // channels per CTA
constexpr int kBwdThreads = kBwdWarps * 32;
constexpr int kMaxState = 32;
constexpr int kBwdSlots = kSeg / 4;                    // float4 slots per state row of a tile
#ifndef VV_BWD_TRANSPOSE
#define VV_BWD_TRANSPOSE 1
#endif
#ifndef VV_BWD_PACK_BC
#define VV_BWD_PACK_BC 1
#endif
constexpr bool kBwdTranspose = VV_BWD_TRANSPOSE != 0;  // dB/dC over a warp's channels: shuffle reduce-scatter (1) or staggered smem RMW (0)

static_assert(kSeg == 64, "a channel group is 8 lanes x 8 positions");

// float4 units of shared memory per warp / per CTA for a state block of NB rows
__host__ __device__ constexpr int bwd_warp_f4(int NB) { return NB * (2 * kBwdSlots + kBwdGroups + 8); }
__host__ __device__ constexpr size_t bwd_smem_bytes(int NB) {
    return 16 * (size_t)(2 * NB * kBwdSlots + kBwdWarps * bwd_warp_f4(NB));
}

__device__ __forceinline__ void acc4(float4& v, float2 lo, float2 hi) {
    asm("add.rn.ftz.f32x2 %0, %0, %1;" : "+l"(*reinterpret_cast<unsigned long long*>(&v.x)) : "l"(*reinterpret_cast<const unsigned long long*>(&lo)));
    asm("add.rn.ftz.f32x2 %0, %0, %1;" : "+l"(*reinterpret_cast<unsigned long long*>(&v.z)) : "l"(*reinterpret_cast<const unsigned long long*>(&hi)));
}
__device__ __forceinline__ float4 add4(float4 v, float2 lo, float2 hi) {
    const float2 p = add2(make_float2(v.x, v.y), lo), q = add2(make_float2(v.z, v.w), hi);
    return make_float4(p.x, p.y, q.x, q.y);
}

// tile[n * 16 + tb]     = positions 8 tb .. 8 tb + 3 of state row n
// tile[n * 16 + 8 + tb] = positions 8 tb + 4 .. 8 tb + 7
// so the 8 lanes of a channel group (a quarter warp) read 128 contiguous bytes per access.
template <typename T, bool kVec, int NB>
struct BwdTileLoader {
    static constexpr int kPer = (NB * 8 + kBwdThreads - 1) / kBwdThreads;
    static constexpr int kQuads = kSeg * (NB / 4);
    static constexpr int kPerQ = (kQuads + kBwdThreads - 1) / kBwdThreads;
    Raw8<T, kVec> raw[kPer];
    StateQuad<T> quad[kPerQ];
    bool rows;   // see StateTileLoader (scan_seq.cuh): `rows` = (.., N, L) layout walked left to right, else (position, 4 states) items
    __device__ __forceinline__ void load(const T* __restrict__ base, int64_t ns, int64_t ls, int N, int t0, const Trav& tr,
                                         bool gen = true) {
        rows = !gen || (ls <= 1 && tr.mode == VV_DIR_FWD);
        if (rows) {
#pragma unroll
            for (int j = 0; j < kPer; ++j) {
                const int idx = threadIdx.x + j * kBwdThreads;
                const int n = idx >> 3, tb = idx & 7;
                raw[j].load(base + (n < N ? n : 0) * ns, (n < N && n < NB) ? t0 + tb * 8 : tr.L, tr.L);   // rows >= N read as 0
            }
        } else {
#pragma unroll
            for (int j = 0; j < kPerQ; ++j) {
                const int idx = threadIdx.x + j * kBwdThreads;
                const int n4 = idx / kSeg, pos = idx - n4 * kSeg;   // position fastest: 2-way bank conflicts on the transposed stores
                quad[j].load(base, ns, ls < 1 ? 1 : ls, N, n4 * 4, t0 + pos, idx < kQuads, tr);
            }
        }
    }
    __device__ __forceinline__ void store(float4* __restrict__ tile) const {
        if (rows) {
#pragma unroll
            for (int j = 0; j < kPer; ++j) {
                const int idx = threadIdx.x + j * kBwdThreads;
                const int n = idx >> 3, tb = idx & 7;
                if (n < NB) {
                    float v[8];
                    raw[j].unpack(v);
                    tile[n * kBwdSlots + tb] = make_float4(v[0], v[1], v[2], v[3]);
                    tile[n * kBwdSlots + 8 + tb] = make_float4(v[4], v[5], v[6], v[7]);
                }
            }
        } else {
            float* tf = reinterpret_cast<float*>(tile);
#pragma unroll
            for (int j = 0; j < kPerQ; ++j) {
                const int idx = threadIdx.x + j * kBwdThreads;
                if (idx < kQuads) {
                    const int n4 = idx / kSeg, pos = idx - n4 * kSeg;
                    const int slot = (pos >> 3) + ((pos & 4) ? 8 : 0);       // float4 slot of the position inside a state row
                    const float4 v = quad[j].unpack();
                    float* dst = tf + (n4 * 4 * kBwdSlots + slot) * 4 + (pos & 3);
                    dst[0] = v.x;
                    dst[4 * kBwdSlots] = v.y;
                    dst[8 * kBwdSlots] = v.z;
                    dst[12 * kBwdSlots] = v.w;
                }
            }
        }
    }
};

// Gradient formulas (real A, variable B and C): selective_scan_bwd_kernel.cuh:279-295, 439-453.
// With r_t the adjoint of h_t and e_t = a_t r_t:
//   r_t = e_{t+1} + g_t C_t            g = dout * silu(z) (or dout)
//   w_t = e_t h_{t-1}                  (= r_t a_t h_{t-1})
//   du = D g + dt sum_n r B            ddt = u sum_n r B + sum_n A_n w
//   dA_n = sum_t dt w                  dB_n = sum_d r dt u          dC_n = sum_d g h
// NB: compile-time state block (8, 16, 32) >= N; rows N..NB-1 are zero padding.
// Tensor maps of B and C for the TMA variant of the tile fill (kTma): rank 2, dims (L, B*G*N), box (64, N).
struct BwdTmaMaps {
    alignas(64) unsigned char B[128];
    alignas(64) unsigned char C[128];
};

// kTma (measurement variant, VV_TMA_BC=1; plain launches, 16-bit I/O): the B / C tiles of the segment arrive by
// cp.async.bulk.tensor into a raw 16-bit staging area (aliasing warp 0's dB / dC tile) instead of through registers, and
// are then expanded into the fp32 [state][position] tiles.  Measured against the register route in profiles/r02_tma.md.
template <typename T, bool kVec, int NB, bool kGen, bool kTma = false>
__global__ void __launch_bounds__(kBwdThreads, 4) seg_bwd_kernel(const vv_scan_args a, const __grid_constant__ BwdTmaMaps maps) {
    extern __shared__ __align__(128) float4 smem4[];
    __shared__ __align__(8) uint64_t tma_bar;
    constexpr bool kPackBC = sizeof(T) == 2 && VV_BWD_PACK_BC != 0;
    const int L = a.seqlen, N = a.dstate;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int cg = lane >> 3, tb = lane & 7;
    float4* tB = smem4;
    float4* tC = tB + NB * kBwdSlots;
    float4* tdB = tC + NB * kBwdSlots + warp * bwd_warp_f4(NB);
    float4* tdC = tdB + NB * kBwdSlots;
    float2* tabER = reinterpret_cast<float2*>(tdC + NB * kBwdSlots);   // [cg][n] = (state entering, pushed adjoint entering)
    float* tabA = reinterpret_cast<float*>(tabER + kBwdGroups * NB);   // [cg][n] = A * log2(e)
    float* dAs = reinterpret_cast<float*>(tdC + NB * kBwdSlots + kBwdGroups * NB);   // [n][lane rotated by 8 n]

    // ---- coordinates
    const int dpg = a.dim / a.ngroups;
    const int blocks_per_group = (dpg + kBwdRows - 1) / kBwdRows;
    const int seg = blockIdx.x, b = blockIdx.z;
    const int grp = blockIdx.y / blocks_per_group;
    const int off = (blockIdx.y - grp * blocks_per_group) * kBwdRows;
    const int nrows = min(kBwdRows, dpg - off);
    const int r = warp * kBwdGroups + cg;
    const bool live = r < nrows;
    const int d = grp * dpg + off + (live ? r : 0);
    const int S = gridDim.x;
    const int t0s = seg * kSeg;
    const int t0 = live ? t0s + tb * 8 : L;   // dead channels read as padding
    const Trav tr = group_trav<kGen>(a, grp);   // t0 / t0s are TRAVERSAL positions of the direction block
    const int dg = gate_row<kGen>(a, d);

    // ---- every global load that does not depend on the preceding kernels, issued up front
    Raw8<T, kVec> r_dt, r_u, r_g, r_z;
    constexpr int kKP = 3, kNF = 5;                           // FRAMES staging: 8 lanes x 3 pixels of a frame, nframes <= 5
    // FRAMES order: rows gathered run by run through shared memory (all loads in flight before the first store)
    const bool staged = kGen && kVec && tr.mode == VV_DIR_FRAMES && tr.nf <= kNF && frames_fits<8, kKP>(tr.nf, kSeg);
    const FramesSpan span = frames_span(tr, t0s, kSeg);
    FramesGather<T, 8, kKP, kNF> fg_dt, fg_u, fg_g, fg_z;
    if (staged) {
        if (live) {
            const FramesLanePlan<8, kNF> pl(span, tr, tb);
            fg_dt.load(reinterpret_cast<const T*>(a.delta) + b * a.delta_bs + d * a.delta_ds, pl);
            fg_u.load(reinterpret_cast<const T*>(a.u) + b * a.u_bs + d * a.u_ds, pl);
            fg_g.load(reinterpret_cast<const T*>(a.dout) + b * a.dout_bs + dg * a.dout_ds, pl);
            if (a.z) fg_z.load(reinterpret_cast<const T*>(a.z) + b * a.z_bs + dg * a.z_ds, pl);
        }
    } else {
        r_dt.load_trav(reinterpret_cast<const T*>(a.delta) + b * a.delta_bs + d * a.delta_ds, t0, tr);
        r_u.load_trav(reinterpret_cast<const T*>(a.u) + b * a.u_bs + d * a.u_ds, t0, tr);
        r_g.load_trav(reinterpret_cast<const T*>(a.dout) + b * a.dout_bs + dg * a.dout_ds, t0, tr);
        if (a.z) r_z.load_trav(reinterpret_cast<const T*>(a.z) + b * a.z_bs + dg * a.z_ds, t0, tr);
    }
    BwdTileLoader<T, kVec, NB> lB, lC;
    T* tma_stage = reinterpret_cast<T*>(tC + NB * kBwdSlots);      // warp 0's dB / dC tile: free until the state loop
    if (kTma) {
        if (threadIdx.x == 0) mbar_init(&tma_bar, 1);
        __syncthreads();
        if (threadIdx.x == 0) {
            const int row0 = (b * a.ngroups + grp) * N;             // first state row of this (batch, group) in (B*G*N, L)
            mbar_expect_tx(&tma_bar, 2u * (unsigned)(N * kSeg * sizeof(T)));
            tma_load_2d(tma_stage, maps.B, t0s, row0, &tma_bar);
            tma_load_2d(tma_stage + N * kSeg, maps.C, t0s, row0, &tma_bar);
        }
    } else {
        lB.load(reinterpret_cast<const T*>(a.Bm) + b * a.B_bs + grp * a.B_gs, a.B_ns, a.B_ls, N, t0s, tr, kGen);
        lC.load(reinterpret_cast<const T*>(a.Cm) + b * a.C_bs + grp * a.C_gs, a.C_ns, a.C_ls, N, t0s, tr, kGen);
    }
    const float bias = a.delta_bias ? a.delta_bias[d] : 0.f;
    const float Dv = a.D ? a.D[d] : 0.f;
    const bool sp = a.delta_softplus != 0;
    // table entries this lane fills: (channel group, state) pairs o = lane, lane + 32, ...
    constexpr int kTab = kBwdGroups * NB / 32;
    float tA[kTab], tE[kTab];
    int64_t tck[kTab];
#pragma unroll
    for (int j = 0; j < kTab; ++j) {
        const int o = lane + 32 * j;
        const int c2 = o / NB, n2 = o % NB;
        const int r2 = warp * kBwdGroups + c2;
        const int d2 = grp * dpg + off + (r2 < nrows ? r2 : 0);
        tA[j] = n2 < N ? a.A[d2 * a.A_ds + n2 * a.A_ns] : 0.f;
        tck[j] = (((int64_t)b * a.dim + d2) * S + seg) * N + (n2 < N ? n2 : 0);
        tE[j] = n2 < N ? a.chk[tck[j]] : 0.f;   // written by the forward pass: no dependency on the predecessor kernel
    }
    pdl_trigger();
    if (staged) {
        // staging rows of the warp's 4 channels x 4 tensors: the warp's own dB / dC tile, not used before the state loop
        T* stg = reinterpret_cast<T*>(tdB) + cg * kSeg;
        if (live) {
            const FramesLanePlan<8, kNF> pl(span, tr, tb);
            fg_dt.store(stg, pl, tr.nf);
            fg_u.store(stg + 4 * kSeg, pl, tr.nf);
            fg_g.store(stg + 8 * kSeg, pl, tr.nf);
            if (a.z) fg_z.store(stg + 12 * kSeg, pl, tr.nf);
        }
        __syncwarp();
        r_dt.load_staged(stg, tb * 8, t0, L);
        r_u.load_staged(stg + 4 * kSeg, tb * 8, t0, L);
        r_g.load_staged(stg + 8 * kSeg, tb * 8, t0, L);
        if (a.z) r_z.load_staged(stg + 12 * kSeg, tb * 8, t0, L);
        __syncwarp();
    }

    // ---- per-position quantities of this lane's channel (fp32 pairs, registers)
    float2 dt2[4], drive2[4], g2[4];
    float dzf[8];
    float sum_dt = 0.f;
    {
        float dt[8], u[8], g[8];
        r_dt.unpack(dt);
        r_u.unpack(u);
        r_g.unpack(g);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float v = dt[i] + bias;
            if (sp) v = softplus_f(v);
            dt[i] = (t0 + i < L) ? v : 0.f;   // padding = scan identity (decay 1, drive 0)
            sum_dt += dt[i];
        }
        if (a.z) {
            float zv[8];
            r_z.unpack(zv);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float sg = sigmoid_f(zv[i]);
                dzf[i] = g[i] * sg * (1.f + zv[i] * (1.f - sg));   // dz = dzf * y
                g[i] *= zv[i] * sg;                                  // grad w.r.t. pre-gate y
            }
        }
#pragma unroll
        for (int jp = 0; jp < 4; ++jp) {
            dt2[jp] = make_float2(dt[2 * jp], dt[2 * jp + 1]);
            drive2[jp] = make_float2(dt[2 * jp] * u[2 * jp], dt[2 * jp + 1] * u[2 * jp + 1]);
            g2[jp] = make_float2(g[2 * jp], g[2 * jp + 1]);
        }
    }
    if (kTma) {
        // the boxes have landed as [state][64 positions] in the I/O dtype (positions >= L zero-filled by the copy engine):
        // expand them into the fp32 tiles, thread = (state, 8 positions)
        mbar_wait(&tma_bar, 0);
#pragma unroll
        for (int j = 0; j < (NB * 8 + kBwdThreads - 1) / kBwdThreads; ++j) {
            const int idx = threadIdx.x + j * kBwdThreads;
            const int n = idx >> 3, c8 = idx & 7;
            if (n < NB) {
                float vb[8], vc[8];
                if (n < N) {
                    load8_plain<T>(tma_stage + n * kSeg + c8 * 8, vb);
                    load8_plain<T>(tma_stage + (N + n) * kSeg + c8 * 8, vc);
                } else {
#pragma unroll
                    for (int i = 0; i < 8; ++i) vb[i] = vc[i] = 0.f;
                }
                tB[n * kBwdSlots + c8] = make_float4(vb[0], vb[1], vb[2], vb[3]);
                tB[n * kBwdSlots + 8 + c8] = make_float4(vb[4], vb[5], vb[6], vb[7]);
                tC[n * kBwdSlots + c8] = make_float4(vc[0], vc[1], vc[2], vc[3]);
                tC[n * kBwdSlots + 8 + c8] = make_float4(vc[4], vc[5], vc[6], vc[7]);
            }
        }
    } else {
        lB.store(tB);
        lC.store(tC);
    }
    pdl_wait();   // chk is from the forward pass, radj from the reverse carry kernel just before us
#pragma unroll
    for (int j = 0; j < kTab; ++j) {
        const int o = lane + 32 * j;
        tabA[o] = tA[j] * kLog2e;
        tabER[o] = make_float2(tE[j], (o % NB) < N ? a.radj[tck[j]] : 0.f);
    }
    __syncthreads();

    float2 y2[4], s12[4], ddt2[4];
#pragma unroll
    for (int jp = 0; jp < 4; ++jp) y2[jp] = s12[jp] = ddt2[jp] = make_float2(0.f, 0.f);
    if (!kBwdTranspose) {
#pragma unroll
        for (int k = 0; k < 2 * NB * kBwdSlots / 32; ++k) tdB[lane + 32 * k] = make_float4(0.f, 0.f, 0.f, 0.f);
        __syncwarp();
    }
    const bool hi16 = (lane & 16) != 0, mid8 = (lane & 8) != 0;
    float4* my_d = (hi16 ? tdC : tdB) + (mid8 ? 8 : 0) + tb;

    // The four channel groups of a warp work on DIFFERENT states at any time (group cg is NB/4 states
    // ahead of group cg-1), so their read-modify-writes of the warp's dB / dC tile never touch the
    // same row: no atomics, no shuffles, no barrier.
    float A2_next = tabA[cg * NB + (kBwdTranspose ? 0 : (cg * (NB / 4)) & (NB - 1))];
#pragma unroll 1
    for (int j = 0; j < NB; ++j) {
        const int n = kBwdTranspose ? j : (j + cg * (NB / 4)) & (NB - 1);
        const float A2 = A2_next;
        A2_next = tabA[cg * NB + ((n + 1) & (NB - 1))];   // one iteration ahead: the load latency is off the critical path
        const float2 er = tabER[cg * NB + n];
        const float An = A2 * (1.f / kLog2e);
        const float2 A22 = make_float2(A2, A2), An2 = make_float2(An, An);
        float2 bm2[4], cm2[4], hs2[4];
        float dec[8];
        {
            const float4 lo = tB[n * kBwdSlots + tb], hi = tB[n * kBwdSlots + 8 + tb];
            bm2[0] = make_float2(lo.x, lo.y); bm2[1] = make_float2(lo.z, lo.w);
            bm2[2] = make_float2(hi.x, hi.y); bm2[3] = make_float2(hi.z, hi.w);
        }
        {
            const float4 lo = tC[n * kBwdSlots + tb], hi = tC[n * kBwdSlots + 8 + tb];
            cm2[0] = make_float2(lo.x, lo.y); cm2[1] = make_float2(lo.z, lo.w);
            cm2[2] = make_float2(hi.x, hi.y); cm2[3] = make_float2(hi.z, hi.w);
        }
        // ---- lane-local aggregates: h left->right, pushed adjoint e right->left
        float X = 0.f;
#pragma unroll
        for (int jp = 0; jp < 4; ++jp) {
            const float2 x2 = mul2(dt2[jp], A22);
            dec[2 * jp] = exp2f(x2.x);
            dec[2 * jp + 1] = exp2f(x2.y);
            hs2[jp] = mul2(drive2[jp], bm2[jp]);        // drive, replaced by the state below
            X = fmaf(dec[2 * jp], X, hs2[jp].x);
            X = fmaf(dec[2 * jp + 1], X, hs2[jp].y);
        }
        const float P = exp2f(A2 * sum_dt);            // decay product of the lane's 8 positions
        float XE = 0.f;
#pragma unroll
        for (int jp = 3; jp >= 0; --jp) {
            XE = dec[2 * jp + 1] * fmaf(g2[jp].y, cm2[jp].y, XE);
            XE = dec[2 * jp] * fmaf(g2[jp].x, cm2[jp].x, XE);
        }
        // ---- two independent 3-step scans over the 8 lanes of the channel, shuffle chains interleaved
        float Pf = P, Pr = P;
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
            const float Pu = __shfl_up_sync(0xffffffffu, Pf, o, 8);
            const float Xu = __shfl_up_sync(0xffffffffu, X, o, 8);
            const float Pd = __shfl_down_sync(0xffffffffu, Pr, o, 8);
            const float Xd = __shfl_down_sync(0xffffffffu, XE, o, 8);
            if (tb >= o) {
                X = fmaf(Pf, Xu, X);
                Pf *= Pu;
            }
            if (tb + o < 8) {
                XE = fmaf(Pr, Xd, XE);
                Pr *= Pd;
            }
        }
        float Pfx = __shfl_up_sync(0xffffffffu, Pf, 1, 8);
        float Xx = __shfl_up_sync(0xffffffffu, X, 1, 8);
        float Prx = __shfl_down_sync(0xffffffffu, Pr, 1, 8);
        float XEx = __shfl_down_sync(0xffffffffu, XE, 1, 8);
        if (tb == 0) { Pfx = 1.f; Xx = 0.f; }
        if (tb == 7) { Prx = 1.f; XEx = 0.f; }
        const float h_in = fmaf(Pfx, er.x, Xx);    // state entering this lane's first position
        float e = fmaf(Prx, er.y, XEx);            // pushed adjoint entering from the right
        {
            float h = h_in;
#pragma unroll
            for (int jp = 0; jp < 4; ++jp) {
                h = fmaf(dec[2 * jp], h, hs2[jp].x);
                hs2[jp].x = h;
                h = fmaf(dec[2 * jp + 1], h, hs2[jp].y);
                hs2[jp].y = h;
            }
        }
        float2 dA2 = make_float2(0.f, 0.f);
        float2 dB2[4], dC2[4];
#pragma unroll
        for (int jp = 3; jp >= 0; --jp) {
            float2 rr, w;
            rr.y = fmaf(g2[jp].y, cm2[jp].y, e);        // adjoint of h at position 2 jp + 1
            e = dec[2 * jp + 1] * rr.y;
            w.y = e * hs2[jp].x;
            rr.x = fmaf(g2[jp].x, cm2[jp].x, e);
            e = dec[2 * jp] * rr.x;
            w.x = e * (jp > 0 ? hs2[jp > 0 ? jp - 1 : 0].y : h_in);
            s12[jp] = fma2(rr, bm2[jp], s12[jp]);
            ddt2[jp] = fma2(An2, w, ddt2[jp]);
            dA2 = fma2(dt2[jp], w, dA2);
            dB2[jp] = mul2(rr, drive2[jp]);
            dC2[jp] = mul2(g2[jp], hs2[jp]);
            y2[jp] = fma2(cm2[jp], hs2[jp], y2[jp]);
        }
        dAs[n * 32 + ((lane + 8 * n) & 31)] = dA2.x + dA2.y;
        if (kBwdTranspose && kPackBC) {
            // 16-bit I/O: dB / dC leave the kernel in that dtype anyway, so the 4-channel partial sums of the warp travel
            // as packed pairs -- half the shuffles, selects and adds of the fp32 route below, and a 64-bit tile store.
            uint32_t pB[4], pC[4], k2[4], k1[2];
#pragma unroll
            for (int jp = 0; jp < 4; ++jp) {
                pB[jp] = pack_pair<T>(dB2[jp]);
                pC[jp] = pack_pair<T>(dC2[jp]);
            }
#pragma unroll
            for (int jp = 0; jp < 4; ++jp) {
                const uint32_t send = hi16 ? pB[jp] : pC[jp], keep = hi16 ? pC[jp] : pB[jp];
                k2[jp] = add_pairs<T>(keep, __shfl_xor_sync(0xffffffffu, send, 16));
            }
#pragma unroll
            for (int jp = 0; jp < 2; ++jp) {
                const uint32_t send = mid8 ? k2[jp] : k2[2 + jp], keep = mid8 ? k2[2 + jp] : k2[jp];
                k1[jp] = add_pairs<T>(keep, __shfl_xor_sync(0xffffffffu, send, 8));
            }
            *reinterpret_cast<uint2*>(my_d + n * kBwdSlots) = make_uint2(k1[0], k1[1]);   // head of the fp32 route's slot
        } else if (kBwdTranspose) {
            // reduce dB / dC over the 4 channels of the warp: transposing reduce-scatter.  Lanes 0-15 end
            // with dB, lanes 16-31 with dC; (lane & 8) selects the half of the lane's 8 positions.
            float2 k2[4];
#pragma unroll
            for (int jp = 0; jp < 4; ++jp) {
                const float2 send = hi16 ? dB2[jp] : dC2[jp];
                const float2 keep = hi16 ? dC2[jp] : dB2[jp];
                k2[jp].x = keep.x + __shfl_xor_sync(0xffffffffu, send.x, 16);
                k2[jp].y = keep.y + __shfl_xor_sync(0xffffffffu, send.y, 16);
            }
            float2 k1[2];
#pragma unroll
            for (int jp = 0; jp < 2; ++jp) {
                const float2 send = mid8 ? k2[jp] : k2[2 + jp];
                const float2 keep = mid8 ? k2[2 + jp] : k2[jp];
                k1[jp].x = keep.x + __shfl_xor_sync(0xffffffffu, send.x, 8);
                k1[jp].y = keep.y + __shfl_xor_sync(0xffffffffu, send.y, 8);
            }
            my_d[n * kBwdSlots] = make_float4(k1[0].x, k1[0].y, k1[1].x, k1[1].y);
        } else {
            float4* pB = tdB + n * kBwdSlots + tb;
            float4* pC = tdC + n * kBwdSlots + tb;
            float4 b0 = pB[0], b1 = pB[8], c0 = pC[0], c1 = pC[8];
            acc4(b0, dB2[0], dB2[1]);
            acc4(b1, dB2[2], dB2[3]);
            acc4(c0, dC2[0], dC2[1]);
            acc4(c1, dC2[2], dC2[3]);
            pB[0] = b0;
            pB[8] = b1;
            pC[0] = c0;
            pC[8] = c1;
        }
    }

    // ---- per-position outputs of this lane's channel
    {
        float u[8], du_o[8], ddt_o[8];
        r_u.keep_packed();
        r_u.unpack(u);
        float dD_loc = 0.f, dbias_loc = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float dti = (i & 1) ? dt2[i >> 1].y : dt2[i >> 1].x;
            const float gi = (i & 1) ? g2[i >> 1].y : g2[i >> 1].x;
            const float s1i = (i & 1) ? s12[i >> 1].y : s12[i >> 1].x;
            const float ddi = (i & 1) ? ddt2[i >> 1].y : ddt2[i >> 1].x;
            du_o[i] = fmaf(Dv, gi, dti * s1i);
            float dd = fmaf(u[i], s1i, ddi);
            // d softplus(v)/dv = sigmoid(v) = 1 - exp(-softplus(v)); dt == 0 marks padding
            if (sp) dd *= (1.f - __expf(-dti));
            ddt_o[i] = dd;
            dbias_loc += (t0 + i < L) ? dd : 0.f;
            dD_loc = fmaf(gi, u[i], dD_loc);
        }
        if (a.z) {
#pragma unroll
            for (int i = 0; i < 8; ++i) dzf[i] *= fmaf(Dv, u[i], (i & 1) ? y2[i >> 1].y : y2[i >> 1].x);
        }
        if (staged && 4 * (int)sizeof(T) <= NB) {
            // FRAMES order: one tensor at a time through a 4-row staging area (the warp's state tables, done with)
            T* ostg = reinterpret_cast<T*>(tabER) + cg * kSeg;
            const FramesLanePlan<8, kNF> pl(span, tr, tb);
#pragma unroll
            for (int which = 0; which < 3; ++which) {
                if (which == 2 && !a.z) break;
                __syncwarp();
                if (kVec) store8_vec<T>(ostg + tb * 8, which == 0 ? du_o : (which == 1 ? ddt_o : dzf));
                __syncwarp();
                T* grow = which == 0 ? reinterpret_cast<T*>(a.du) + b * a.du_bs + d * a.du_ds
                        : which == 1 ? reinterpret_cast<T*>(a.ddelta) + b * a.ddelta_bs + d * a.ddelta_ds
                                     : reinterpret_cast<T*>(a.dz) + b * a.dz_bs + d * a.dz_ds;
                if (live) frames_scatter_n<T, 8, kKP, kNF>(grow, ostg, pl, tr.nf);
            }
        } else if (live) {
            store8_trav<T, kVec>(reinterpret_cast<T*>(a.du) + b * a.du_bs + d * a.du_ds, t0, tr, du_o);
            store8_trav<T, kVec>(reinterpret_cast<T*>(a.ddelta) + b * a.ddelta_bs + d * a.ddelta_ds, t0, tr, ddt_o);
            if (a.z) store8_trav<T, kVec>(reinterpret_cast<T*>(a.dz) + b * a.dz_bs + d * a.dz_ds, t0, tr, dzf);
        }
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
            dD_loc += __shfl_xor_sync(0xffffffffu, dD_loc, o);
            dbias_loc += __shfl_xor_sync(0xffffffffu, dbias_loc, o);
        }
        if (live && tb == 0) {
            if (a.dD) atomicAdd(a.dD + d, dD_loc);
            if (a.ddelta_bias) atomicAdd(a.ddelta_bias + d, dbias_loc);
        }
    }
    // ---- dA: sum the 8 lanes of each (channel, state) of this warp
    __syncwarp();
#pragma unroll
    for (int j = 0; j < kTab; ++j) {
        const int o = lane + 32 * j;
        const int c2 = o / NB, n2 = o % NB;
        const int r2 = warp * kBwdGroups + c2;
        if (r2 < nrows && n2 < N) {
            const float* src = dAs + n2 * 32 + ((c2 * 8 + 8 * n2) & 31);
            const float4 p0 = *reinterpret_cast<const float4*>(src);
            const float4 p1 = *reinterpret_cast<const float4*>(src + 4);
            const float sum = ((p0.x + p0.y) + (p0.z + p0.w)) + ((p1.x + p1.y) + (p1.z + p1.w));
            atomicAdd(a.dA + (int64_t)(grp * dpg + off + r2) * N + n2, sum);
        }
    }
    // ---- the CTA's dB / dC: sum the four warp tiles -> global (fp32, 128-bit reductions)
    __syncthreads();
    {
        const float4* w0 = tC + NB * kBwdSlots;
        const int64_t bc_base = ((int64_t)b * a.ngroups + grp) * N;
#pragma unroll
        for (int it = 0; it < 2 * NB * kBwdSlots / kBwdThreads; ++it) {
            const int idx = threadIdx.x + it * kBwdThreads;
            const int tensor = idx / (NB * kBwdSlots);
            const int n = (idx / kBwdSlots) % NB, slot = idx % kBwdSlots;   // consecutive threads, consecutive slots
            const int pc = (slot & 7) * 2 + (slot >> 3);                      // the 4-position chunk of the segment it holds
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int w = 0; w < kBwdWarps; ++w) {
                const float4 x = w0[w * bwd_warp_f4(NB) + tensor * (NB * kBwdSlots) + n * kBwdSlots + slot];
                if (kPackBC) v = add4(v, unpack_pair<T>(__float_as_uint(x.x)), unpack_pair<T>(__float_as_uint(x.y)));
                else v = add4(v, make_float2(x.x, x.y), make_float2(x.z, x.w));
            }
            const int t = t0s + pc * 4;
            if (n < N) {
                float* p = (tensor ? a.dC : a.dB) + (bc_base + n) * L + t;
                if (kVec) {
                    if (t < L) atomicAdd(reinterpret_cast<float4*>(p), v);
                } else {
                    const float ev[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (t + k < L) atomicAdd(p + k, ev[k]);
                }
            }
        }
    }
}

// fp32 dB / dC -> I/O dtype, 8 elements per thread; chained to seg_bwd_kernel as a programmatic dependent
template <typename T>
__global__ void __launch_bounds__(256) cast_bc_kernel(const float* __restrict__ dB, const float* __restrict__ dC,
                                                      T* __restrict__ oB, T* __restrict__ oC, const int64_t n) {
    pdl_wait();
    const int64_t i = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 8;
    const float* src = blockIdx.y ? dC : dB;
    T* dst = blockIdx.y ? oC : oB;
    if (i + 8 <= n && (reinterpret_cast<uintptr_t>(src + i) % 16 == 0) && (reinterpret_cast<uintptr_t>(dst + i) % 16 == 0)) {
        float v[8];
        load8_vec<float>(src + i, v);
        store8_vec<T>(dst + i, v);
    } else {
        for (int64_t k = i; k < n && k < i + 8; ++k) dst[k] = from_f32<T>(src[k]);
    }
}

// The same cast for any output layout and traversal order: the fp32 accumulators are (B,G,N,L) in TRAVERSAL order, the
// destination is indexed by MEMORY position with arbitrary strides -- e.g. the B / C column blocks of dx_dbl
// (B*L, R+2N), which the reference fills through `rearrange(dB, "b 1 dstate l -> (b l) dstate")` and a slice copy
// (selective_scan_interface.py:255-271).  One CTA per (64-position segment, group, batch): the accumulator rows are read
// along L (coalesced), transposed through shared memory and written with the state index fastest.
constexpr int kCastThreads = 256;
template <typename T>
__global__ void __launch_bounds__(kCastThreads) cast_bc_strided_kernel(const vv_scan_args a) {
    __shared__ float tile[2][kMaxState][kSeg + 1];
    __shared__ int s_mem[kSeg];                       // memory position of each traversal position of the segment
    pdl_wait();
    const int N = a.dstate, L = a.seqlen;
    const int seg = blockIdx.x, g = blockIdx.y, b = blockIdx.z;
    const int t0 = seg * kSeg;
    const int tid = threadIdx.x;
    const Trav tr = group_trav<true>(a, g);
    const int64_t acc_base = (((int64_t)b * a.ngroups + g) * N) * L;
    if (tid < kSeg) s_mem[tid] = t0 + tid < L ? tr.mem(t0 + tid) : -1;
    {   // rows of the accumulators: thread = (row n0 + 4 i, position j), 64 consecutive floats per row
        const int j = tid & (kSeg - 1), n0 = tid >> 6;
        const bool ok = t0 + j < L;
        for (int n = n0; n < N; n += kCastThreads / kSeg) {
            const int64_t o = acc_base + (int64_t)n * L + t0 + j;
            tile[0][n][j] = ok ? a.dB[o] : 0.f;
            tile[1][n][j] = ok ? a.dC[o] : 0.f;
        }
    }
    __syncthreads();
    {   // a warp writes one position at a time: lane = (tensor, state), the state index fastest
        const int lane = tid & 31, w = tid >> 5;
        for (int e = lane; e < 2 * N; e += 32) {
            const int which = e >= N ? 1 : 0, n = e - which * N;
            T* dst = which ? reinterpret_cast<T*>(a.dC_io) + b * a.dCio_bs + g * a.dCio_gs + n * a.dCio_ns
                           : reinterpret_cast<T*>(a.dB_io) + b * a.dBio_bs + g * a.dBio_gs + n * a.dBio_ns;
            const int64_t ls = which ? a.dCio_ls : a.dBio_ls;
#pragma unroll 4
            for (int j = w; j < kSeg; j += kCastThreads / 32) {
                const int m = s_mem[j];
                if (m >= 0) dst[m * ls] = from_f32<T>(tile[which][n][j]);
            }
        }
    }
}

}  // namespace vv
