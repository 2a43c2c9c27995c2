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
// B and C of the segment are staged once per CTA as fp32; the four channel groups of a warp read the
// same 128 bytes (broadcast, one wavefront per access).  dB/dC are reduced over the 4 channels of a
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

static_assert(kSeg == 64, "a channel group is 8 lanes x 8 positions");

// bytes of dynamic shared memory for N states
__host__ __device__ constexpr size_t bwd_smem_bytes(int N) {
    return (size_t)N * (2 * kBwdSlots * 16                    // B, C tiles (CTA)
                        + kBwdWarps * (2 * kBwdSlots * 16     // dB, dC tiles (per warp)
                                       + kBwdGroups * 16      // (A2, A, E, R) table (per warp)
                                       + 32 * 4));            // dA scratch (per warp)
}

// tile[n * 16 + tb]     = positions 8 tb .. 8 tb + 3 of state row n
// tile[n * 16 + 8 + tb] = positions 8 tb + 4 .. 8 tb + 7
// so the 8 lanes of a channel group read 128 contiguous bytes per access.
template <typename T, bool kVec>
struct BwdTileLoader {
    static constexpr int kPer = (kMaxState * 8 + kBwdThreads - 1) / kBwdThreads;
    Raw8<T, kVec> raw[kPer];
    __device__ __forceinline__ void load(const T* __restrict__ base, int64_t ns, int N, int t0, int L) {
#pragma unroll
        for (int j = 0; j < kPer; ++j) {
            const int idx = threadIdx.x + j * kBwdThreads;
            const int n = idx >> 3, tb = idx & 7;
            raw[j].load(base + (n < N ? n : 0) * ns, n < N ? t0 + tb * 8 : L, L);
        }
    }
    __device__ __forceinline__ void store(float4* __restrict__ tile, int N) const {
#pragma unroll
        for (int j = 0; j < kPer; ++j) {
            const int idx = threadIdx.x + j * kBwdThreads;
            const int n = idx >> 3, tb = idx & 7;
            if (n < N) {
                float v[8];
                raw[j].unpack(v);
                tile[n * kBwdSlots + tb] = make_float4(v[0], v[1], v[2], v[3]);
                tile[n * kBwdSlots + 8 + tb] = make_float4(v[4], v[5], v[6], v[7]);
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
template <typename T, bool kVec>
__global__ void __launch_bounds__(kBwdThreads, 4) seg_bwd_kernel(const vv_scan_args a) {
    extern __shared__ float4 smem4[];
    const int L = a.seqlen, N = a.dstate;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int cg = lane >> 3, tb = lane & 7;
    float4* tB = smem4;
    float4* tC = tB + N * kBwdSlots;
    float4* wbase = tC + N * kBwdSlots + warp * (N * (2 * kBwdSlots + kBwdGroups + 8));
    float4* tdB = wbase;
    float4* tdC = tdB + N * kBwdSlots;
    float4* tab = tdC + N * kBwdSlots;                           // [cg][n] = (A2, A, E, R)
    float* dAs = reinterpret_cast<float*>(tab + kBwdGroups * N);  // [n][lane]

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
    const int64_t row = (int64_t)b * a.dim + d;

    // ---- every global load that does not depend on the preceding kernels, issued up front
    Raw8<T, kVec> r_dt, r_u, r_g, r_z;
    r_dt.load(reinterpret_cast<const T*>(a.delta) + b * a.delta_bs + d * a.delta_ds, t0, L);
    r_u.load(reinterpret_cast<const T*>(a.u) + b * a.u_bs + d * a.u_ds, t0, L);
    r_g.load(reinterpret_cast<const T*>(a.dout) + b * a.dout_bs + d * a.dout_ds, t0, L);
    if (a.z) r_z.load(reinterpret_cast<const T*>(a.z) + b * a.z_bs + d * a.z_ds, t0, L);
    BwdTileLoader<T, kVec> lB, lC;
    lB.load(reinterpret_cast<const T*>(a.Bm) + b * a.B_bs + grp * a.B_gs, a.B_ns, N, t0s, L);
    lC.load(reinterpret_cast<const T*>(a.Cm) + b * a.C_bs + grp * a.C_gs, a.C_ns, N, t0s, L);
    const float bias = a.delta_bias ? a.delta_bias[d] : 0.f;
    const float Dv = a.D ? a.D[d] : 0.f;
    const bool sp = a.delta_softplus != 0;
    // table entries this lane fills: (channel group, state) pairs o = lane, lane + 32, ...
    float tA[(kBwdGroups * kMaxState + 31) / 32];
#pragma unroll
    for (int j = 0; j < (kBwdGroups * kMaxState + 31) / 32; ++j) {
        const int o = lane + 32 * j;
        const int c2 = o / N, n2 = o - c2 * N;
        const int r2 = warp * kBwdGroups + c2;
        const int d2 = grp * dpg + off + (r2 < nrows ? r2 : 0);
        tA[j] = (c2 < kBwdGroups) ? a.A[d2 * a.A_ds + n2 * a.A_ns] : 0.f;
    }
    pdl_trigger();

    // ---- per-position quantities of this lane's channel (fp32, registers)
    float dt[8], drive[8], g[8], dzf[8];
    {
        float u[8];
        r_dt.unpack(dt);
        r_u.unpack(u);
        r_g.unpack(g);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float v = dt[i] + bias;
            if (sp) v = softplus_f(v);
            dt[i] = (t0 + i < L) ? v : 0.f;   // padding = scan identity (decay 1, drive 0)
            drive[i] = dt[i] * u[i];
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
    }
    lB.store(tB, N);
    lC.store(tC, N);
    pdl_wait();   // chk is from the forward pass, radj from the reverse carry kernel just before us
#pragma unroll
    for (int j = 0; j < (kBwdGroups * kMaxState + 31) / 32; ++j) {
        const int o = lane + 32 * j;
        const int c2 = o / N, n2 = o - c2 * N;
        if (c2 < kBwdGroups) {
            const int r2 = warp * kBwdGroups + c2;
            const int d2 = grp * dpg + off + (r2 < nrows ? r2 : 0);
            const int64_t ck = (((int64_t)b * a.dim + d2) * S + seg) * N + n2;
            tab[c2 * N + n2] = make_float4(tA[j] * kLog2e, tA[j], a.chk[ck], a.radj[ck]);
        }
    }
    __syncthreads();

    float y[8], s1[8], ddt[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        y[i] = 0.f;
        s1[i] = 0.f;
        ddt[i] = 0.f;
    }
    const bool hi = (lane & 16) != 0, mid = (lane & 8) != 0;
    float4* my_d = (hi ? tdC : tdB) + (mid ? 8 : 0) + tb;

#pragma unroll 1
    for (int n = 0; n < N; ++n) {
        const float4 q = tab[cg * N + n];
        const float A2 = q.x, An = q.y;
        float bm[8], cm[8], dec[8], hs[8];
        {
            const float4 lo = tB[n * kBwdSlots + tb], hi4 = tB[n * kBwdSlots + 8 + tb];
            bm[0] = lo.x; bm[1] = lo.y; bm[2] = lo.z; bm[3] = lo.w;
            bm[4] = hi4.x; bm[5] = hi4.y; bm[6] = hi4.z; bm[7] = hi4.w;
        }
        {
            const float4 lo = tC[n * kBwdSlots + tb], hi4 = tC[n * kBwdSlots + 8 + tb];
            cm[0] = lo.x; cm[1] = lo.y; cm[2] = lo.z; cm[3] = lo.w;
            cm[4] = hi4.x; cm[5] = hi4.y; cm[6] = hi4.z; cm[7] = hi4.w;
        }
        // ---- lane-local aggregates: h left->right, pushed adjoint e right->left
        float X = 0.f, P = 1.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            dec[i] = exp2f(dt[i] * A2);
            hs[i] = drive[i] * bm[i];        // drive, replaced by the state below
            X = fmaf(dec[i], X, hs[i]);
            P *= dec[i];
        }
        float XE = 0.f;
#pragma unroll
        for (int i = 7; i >= 0; --i) XE = dec[i] * fmaf(g[i], cm[i], XE);
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
        const float h_in = fmaf(Pfx, q.z, Xx);    // state entering this lane's first position
        float e = fmaf(Prx, q.w, XEx);            // pushed adjoint entering from the right
        {
            float h = h_in;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                h = fmaf(dec[i], h, hs[i]);
                hs[i] = h;
            }
        }
        float dA_loc = 0.f;
        float dBv[8], dCv[8];
#pragma unroll
        for (int i = 7; i >= 0; --i) {
            const float rr = fmaf(g[i], cm[i], e);     // adjoint of h_t
            e = dec[i] * rr;
            const float w = e * (i > 0 ? hs[i - 1] : h_in);
            s1[i] = fmaf(rr, bm[i], s1[i]);
            ddt[i] = fmaf(An, w, ddt[i]);
            dA_loc = fmaf(dt[i], w, dA_loc);
            dBv[i] = rr * drive[i];
            dCv[i] = g[i] * hs[i];
            y[i] = fmaf(cm[i], hs[i], y[i]);
        }
        dAs[n * 32 + lane] = dA_loc;
        // ---- reduce dB / dC over the 4 channels of the warp: transposing reduce-scatter.
        // lanes 0-15 end with dB, lanes 16-31 with dC; (lane & 8) selects the half of the 8 positions.
        float k8[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float send = hi ? dBv[i] : dCv[i];
            const float keep = hi ? dCv[i] : dBv[i];
            k8[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
        }
        float k4[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float send = mid ? k8[j] : k8[4 + j];
            const float keep = mid ? k8[4 + j] : k8[j];
            k4[j] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
        }
        my_d[n * kBwdSlots] = make_float4(k4[0], k4[1], k4[2], k4[3]);
    }

    // ---- per-position outputs of this lane's channel
    {
        float u[8], du_o[8], ddt_o[8];
        r_u.unpack(u);
        float dD_loc = 0.f, dbias_loc = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            du_o[i] = fmaf(Dv, g[i], dt[i] * s1[i]);
            float dd = fmaf(u[i], s1[i], ddt[i]);
            // d softplus(v)/dv = sigmoid(v) = 1 - exp(-softplus(v)); dt == 0 marks padding
            if (sp) dd *= (1.f - __expf(-dt[i]));
            ddt_o[i] = dd;
            dbias_loc += (t0 + i < L) ? dd : 0.f;
            dD_loc = fmaf(g[i], u[i], dD_loc);
        }
        if (live) {
            store8<T, kVec>(reinterpret_cast<T*>(a.du) + b * a.du_bs + d * a.du_ds, t0, L, du_o);
            store8<T, kVec>(reinterpret_cast<T*>(a.ddelta) + b * a.ddelta_bs + d * a.ddelta_ds, t0, L, ddt_o);
            if (a.z) {
#pragma unroll
                for (int i = 0; i < 8; ++i) dzf[i] *= fmaf(Dv, u[i], y[i]);
                store8<T, kVec>(reinterpret_cast<T*>(a.dz) + b * a.dz_bs + d * a.dz_ds, t0, L, dzf);
            }
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
    for (int o = lane; o < kBwdGroups * N; o += 32) {
        const int c2 = o / N, n2 = o - c2 * N;
        const int r2 = warp * kBwdGroups + c2;
        if (r2 < nrows) {
            const float4 p0 = *reinterpret_cast<const float4*>(dAs + n2 * 32 + c2 * 8);
            const float4 p1 = *reinterpret_cast<const float4*>(dAs + n2 * 32 + c2 * 8 + 4);
            const float sum = ((p0.x + p0.y) + (p0.z + p0.w)) + ((p1.x + p1.y) + (p1.z + p1.w));
            atomicAdd(a.dA + (int64_t)(grp * dpg + off + r2) * N + n2, sum);
        }
    }
    // ---- the CTA's dB / dC: sum the four warp tiles -> global (fp32, 128-bit reductions)
    __syncthreads();
    {
        const float4* w0 = tC + N * kBwdSlots;
        const int wstride = N * (2 * kBwdSlots + kBwdGroups + 8);
        const int64_t bc_base = ((int64_t)b * a.ngroups + grp) * N;
        for (int idx = threadIdx.x; idx < 2 * N * kBwdSlots; idx += kBwdThreads) {
            const int tensor = idx / (N * kBwdSlots);
            const int rem = idx - tensor * (N * kBwdSlots);
            const int n = rem / kBwdSlots, pc = rem - n * kBwdSlots;   // pc: 4-position chunk of the segment
            const int slot = (pc & 1) * 8 + (pc >> 1);
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int w = 0; w < kBwdWarps; ++w) {
                const float4 x = w0[w * wstride + tensor * (N * kBwdSlots) + n * kBwdSlots + slot];
                v.x += x.x; v.y += x.y; v.z += x.z; v.w += x.w;
            }
            const int t = t0s + pc * 4;
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

}  // namespace vv
