// scan.cuh -- backward main kernel of the fused selective scan for sm_100a.
//
// Replaces selective_scan_bwd_kernel of the reference
// (mamba/csrc/selective_scan/selective_scan_bwd_kernel.cuh:75-489).  The algebra is the reference's:
// the associative operator (a0,b0)o(a1,b1) = (a1 a0, a1 b0 + b1) (selective_scan_common.h:110-115),
// softplus with threshold 20, exp2f(dt * A * log2e), gradient formulas of :279-295 and :439-453.
//
// Design (B200-first, not a port).  The reference walks a whole (batch, channel) row inside one
// CTA, chunk after chunk, last to first.  Here every (batch, channel, UNIT of 256 positions) is an
// independent piece of work for one warp: the forward state entering the unit comes from the
// checkpoint tensor `chk` written by the forward pass (checkpointed chunk states, recomputed in
// the backward -- a saved `out` is never read), the adjoint entering it from the right comes from
// `radj`, produced by the reverse segment-aggregate + carry passes of scan_seq.cuh.  So
// B*D*ceil(L/256) warps are in flight (10240 at the B=1 stage-1 shape) and no CTA waits on another.
//
// Inside a warp: lane l owns positions [8l, 8l+8) of the unit (one 128-bit access per streamed
// tensor, 512 contiguous bytes per warp request).  For one state n the lane runs the recurrence over
// its 8 positions in registers (exp2 evaluated once per element and kept), the 32 lane aggregates
// are combined by a 5-step warp-shuffle scan (once left-to-right for h, once right-to-left for the
// adjoint), and a second in-register sweep produces states, adjoints and all gradient terms.
// B and C rows of the unit are staged once per CTA in shared memory as fp32 (the bf16->fp32
// conversion is paid once per CTA, not once per channel) in a slot order that makes each lane's two
// 128-bit reads bank-conflict free, and are reused by every channel the CTA walks.
// dB/dC are reduced over the channels of a CTA in shared memory (each warp owns a different state
// row between two barriers, so the read-modify-write needs no atomics) and leave the CTA as one
// 128-bit red.global.add per 4 positions -- D/(rows per CTA)-way contention instead of the
// reference's D-way scalar atomics (selective_scan_bwd_kernel.cuh:298-316).
#pragma once

#include "../../include/vivim_b200.h"
#include "common.cuh"
#include "scan_seq.cuh"

namespace vv {

constexpr int kUnit = VV_SCAN_UNIT;       // positions per unit = 32 lanes x 8
constexpr int kSlots = 64;                // float4 slots per state row of a staged tile
constexpr int kMaxState = 32;             // states are held one per lane
constexpr int kDaPitch = 33;              // padded pitch of the per-lane dA scratch

static_assert(kUnit == kWarp * kVecElems, "a unit is one warp x 8 positions");

// ---------------------------------------------------------------- warp scans of (P, X) pairs
// forward: lane l ends with the aggregate of lanes 0..l
__device__ __forceinline__ void warp_scan_fwd(float& P, float& X, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const float Pu = __shfl_up_sync(0xffffffffu, P, o);
        const float Xu = __shfl_up_sync(0xffffffffu, X, o);
        if (lane >= o) {
            X = fmaf(P, Xu, X);
            P *= Pu;
        }
    }
}
// reverse: lane l ends with the aggregate of lanes l..31 (recurrence runs right to left)
__device__ __forceinline__ void warp_scan_rev(float& P, float& X, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const float Pd = __shfl_down_sync(0xffffffffu, P, o);
        const float Xd = __shfl_down_sync(0xffffffffu, X, o);
        if (lane + o < 32) {
            X = fmaf(P, Xd, X);
            P *= Pd;
        }
    }
}

// ---------------------------------------------------------------- staging of B / C unit tiles
// tile[n * 64 + l]      = positions 8l .. 8l+3 of state row n
// tile[n * 64 + 32 + l] = positions 8l+4 .. 8l+7
template <typename T, bool kVec>
__device__ __forceinline__ void fill_tile(float4* __restrict__ tile, const T* __restrict__ base, int64_t row_stride,
                                          int N, int unit, int L) {
    for (int idx = threadIdx.x; idx < N * 32; idx += blockDim.x) {
        const int n = idx >> 5, l = idx & 31;
        float v[8];
        load8<T, kVec>(base + n * row_stride, unit * kUnit + l * kVecElems, L, v);
        tile[n * kSlots + l] = make_float4(v[0], v[1], v[2], v[3]);
        tile[n * kSlots + 32 + l] = make_float4(v[4], v[5], v[6], v[7]);
    }
}

__device__ __forceinline__ void read_tile(const float4* __restrict__ tile, int n, int lane, float (&v)[8]) {
    const float4 lo = tile[n * kSlots + lane];
    const float4 hi = tile[n * kSlots + 32 + lane];
    v[0] = lo.x; v[1] = lo.y; v[2] = lo.z; v[3] = lo.w;
    v[4] = hi.x; v[5] = hi.y; v[6] = hi.z; v[7] = hi.w;
}

// dt = softplus?(delta + bias) for this lane's 8 positions; exactly 0 outside [0, L) so that padded
// positions are the scan identity (decay 1, drive 0).
template <typename T, bool kVec>
__device__ __forceinline__ void load_dt(const T* __restrict__ row, int t0, int L, float bias, bool softplus,
                                        float (&dt)[8]) {
    load8<T, kVec>(row, t0, L, dt);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        float v = dt[i] + bias;
        if (softplus) v = softplus_f(v);
        dt[i] = (t0 + i < L) ? v : 0.f;
    }
}

struct RowCoord {
    int b, d, g, unit, lane, warp;
    int64_t row;  // b * dim + d
};

__device__ __forceinline__ RowCoord row_coord(const vv_scan_args& a, int rows_seq, int k) {
    RowCoord c;
    c.lane = threadIdx.x & 31;
    c.warp = threadIdx.x >> 5;
    const int W = blockDim.x >> 5;
    c.unit = blockIdx.x;
    c.b = blockIdx.z;
    c.d = (blockIdx.y * rows_seq + k) * W + c.warp;
    c.g = c.d / (a.dim / a.ngroups);
    c.row = (int64_t)c.b * a.dim + c.d;
    return c;
}

// ================================================================ pass 3 (backward)
// Gradient formulas (real A, variable B and C): selective_scan_bwd_kernel.cuh:279-295, 439-453.
//
// Per-row inputs are software-pipelined: while a warp works on row k, the four 128-bit chunks per
// lane of row k+1 (delta, u, dout, z) travel global->shared with cp.async into the lane's own slots,
// and its per-row scalars (A, checkpoint, reverse carry, bias, D) sit in registers, so no row starts
// with an exposed DRAM round trip.
template <typename T>
struct RowPrefetch {
    static constexpr int kVB = 8 * (int)sizeof(T);              // bytes per lane per tensor
    static constexpr int kWarpBytes = 4 * 32 * kVB + 32;        // + the chunk holding the next unit's first delta
};

template <typename T, bool kVec>
__global__ void __launch_bounds__(128) scan_bwd_main_kernel(const vv_scan_args a, const int rows_seq) {
    extern __shared__ float4 smem4[];
    const int L = a.seqlen, N = a.dstate;
    const int W = blockDim.x >> 5;
    float4* tB = smem4;
    float4* tC = tB + N * kSlots;
    float4* tdB = tC + N * kSlots;
    float4* tdC = tdB + N * kSlots;
    unsigned char* pf_base = reinterpret_cast<unsigned char*>(tdC + N * kSlots);
    const int S = (L + kSeg - 1) / kSeg;   // chk / radj are indexed by 64-position segments
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char* pf = pf_base + warp * RowPrefetch<T>::kWarpBytes;
    constexpr int kVB = RowPrefetch<T>::kVB;
    const int unit = blockIdx.x;
    const int t0 = unit * kUnit + lane * kVecElems;
    const int tn = unit * kUnit + kUnit;          // first position of the next unit
    const int seg_first = unit * kSegPerUnit;
    const int seg_last = min(seg_first + kSegPerUnit - 1, S - 1);

    // asynchronous copy of one row's chunks into this lane's slots
    auto prefetch_row = [&](const RowCoord& c) {
        if (!kVec) return;
        const T* rows[4] = {reinterpret_cast<const T*>(a.delta) + c.b * a.delta_bs + c.d * a.delta_ds,
                            reinterpret_cast<const T*>(a.u) + c.b * a.u_bs + c.d * a.u_ds,
                            reinterpret_cast<const T*>(a.dout) + c.b * a.dout_bs + c.d * a.dout_ds,
                            a.z ? reinterpret_cast<const T*>(a.z) + c.b * a.z_bs + c.d * a.z_ds : nullptr};
        const bool ok = t0 < L;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (rows[j] == nullptr) continue;
            const unsigned char* src = reinterpret_cast<const unsigned char*>(ok ? rows[j] + t0 : rows[j]);
            unsigned char* dst = pf + (j * 32 + lane) * kVB;
            cp_async16(dst, src, ok ? 16 : 0);
            if (kVB == 32) cp_async16(dst + 16, src + (ok ? 16 : 0), ok ? 16 : 0);
        }
        if (lane == 31) cp_async16(pf + 4 * 32 * kVB, tn < L ? rows[0] + tn : rows[0], tn < L ? 16 : 0);
        cp_async_commit();
    };
    struct RowScalars { float A_l, E_l, R_l, bias, Dv; };
    auto load_scalars = [&](const RowCoord& c) {
        RowScalars s;
        s.bias = a.delta_bias ? a.delta_bias[c.d] : 0.f;
        s.Dv = a.D ? a.D[c.d] : 0.f;
        s.A_l = lane < N ? a.A[c.d * a.A_ds + lane * a.A_ns] : 0.f;
        s.E_l = lane < N ? a.chk[(c.row * S + seg_first) * N + lane] : 0.f;
        s.R_l = lane < N ? a.radj[(c.row * S + seg_last) * N + lane] : 0.f;
        return s;
    };

    RowCoord c = row_coord(a, rows_seq, 0);
    prefetch_row(c);
    fill_tile<T, kVec>(tB, reinterpret_cast<const T*>(a.Bm) + c.b * a.B_bs + c.g * a.B_gs, a.B_ns, N, unit, L);
    fill_tile<T, kVec>(tC, reinterpret_cast<const T*>(a.Cm) + c.b * a.C_bs + c.g * a.C_gs, a.C_ns, N, unit, L);
    for (int idx = threadIdx.x; idx < 2 * N * kSlots; idx += blockDim.x) tdB[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
    pdl_wait();   // radj comes from the reverse carry kernel; the fills above overlap its tail
    RowScalars sc = load_scalars(c);
    __syncthreads();

    for (int k = 0; k < rows_seq; ++k) {
        const float bias = sc.bias, Dv = sc.Dv, A_l = sc.A_l, E_l = sc.E_l, R_l = sc.R_l;
        const bool sp = a.delta_softplus != 0;
        const T* delta_row = reinterpret_cast<const T*>(a.delta) + c.b * a.delta_bs + c.d * a.delta_ds;
        float dt[8], u[8], g[8], dzf[8];
        float dt_next = 0.f;
        if (kVec) {
            cp_async_wait_all();   // this lane's own slots only: no cross-lane hazard, no barrier
            load8_plain<T>(reinterpret_cast<const T*>(pf + (0 * 32 + lane) * kVB), dt);
            load8_plain<T>(reinterpret_cast<const T*>(pf + (1 * 32 + lane) * kVB), u);
            load8_plain<T>(reinterpret_cast<const T*>(pf + (2 * 32 + lane) * kVB), g);
            if (a.z) load8_plain<T>(reinterpret_cast<const T*>(pf + (3 * 32 + lane) * kVB), dzf);
            if (lane == 31 && tn < L) dt_next = to_f32<T>(*reinterpret_cast<const T*>(pf + 4 * 32 * kVB));
        } else {
            load8<T, kVec>(delta_row, t0, L, dt);
            load8<T, kVec>(reinterpret_cast<const T*>(a.u) + c.b * a.u_bs + c.d * a.u_ds, t0, L, u);
            load8<T, kVec>(reinterpret_cast<const T*>(a.dout) + c.b * a.dout_bs + c.d * a.dout_ds, t0, L, g);
            if (a.z) load8<T, kVec>(reinterpret_cast<const T*>(a.z) + c.b * a.z_bs + c.d * a.z_ds, t0, L, dzf);
            if (lane == 31 && tn < L) dt_next = to_f32<T>(delta_row[tn]);
        }
        // rows of this warp after the current one: start their copies / scalar loads now
        const RowCoord c_cur = c;
        RowScalars sc_next = sc;
        if (k + 1 < rows_seq) {
            c = row_coord(a, rows_seq, k + 1);
            prefetch_row(c);
            sc_next = load_scalars(c);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float v = dt[i] + bias;
            if (sp) v = softplus_f(v);
            dt[i] = (t0 + i < L) ? v : 0.f;
        }
        if (a.z) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float zv = dzf[i];
                const float sg = sigmoid_f(zv);
                dzf[i] = g[i] * sg * (1.f + zv * (1.f - sg));   // dz = dzf * y
                g[i] *= zv * sg;                                  // grad w.r.t. pre-gate y
            }
        }
        if (lane == 31) {
            if (tn < L) {
                const float v = dt_next + bias;
                dt_next = sp ? softplus_f(v) : v;
            } else {
                dt_next = 0.f;
            }
        }
        {
            const float from_above = __shfl_down_sync(0xffffffffu, dt[0], 1);
            if (lane != 31) dt_next = from_above;
        }
        float sum_dt = 0.f;
        float y[8], s1[8], ddt[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            sum_dt += dt[i];
            y[i] = Dv * u[i];
            s1[i] = 0.f;
            ddt[i] = 0.f;
        }
        const float sum_dt_rev = sum_dt - dt[0] + dt_next;
        float dA_acc = 0.f;   // lane n accumulates dA[d, n]

        for (int j = 0; j < N; ++j) {
            // each warp of the CTA works on a different state row between two barriers
            int n = j + warp;
            if (n >= N) n -= N;
            const float An = __shfl_sync(0xffffffffu, A_l, n);
            const float E = __shfl_sync(0xffffffffu, E_l, n);
            const float R = __shfl_sync(0xffffffffu, R_l, n);
            const float A2 = An * kLog2e;
            float bm[8], cm[8], dec[8], hs[8];
            read_tile(tB, n, lane, bm);
            read_tile(tC, n, lane, cm);
            // ---- lane-local aggregates of both recurrences (h left->right, adjoint right->left)
            float X = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                dec[i] = exp2f(dt[i] * A2);
                hs[i] = dt[i] * u[i] * bm[i];        // drive, replaced by the state below
                X = fmaf(dec[i], X, hs[i]);
            }
            const float dec_next = exp2f(dt_next * A2);
            float RX = g[7] * cm[7];
#pragma unroll
            for (int i = 6; i >= 0; --i) RX = fmaf(dec[i + 1], RX, g[i] * cm[i]);
            float P = exp2f(A2 * sum_dt);
            float RP = exp2f(A2 * sum_dt_rev);
            // ---- the two 5-step warp scans are independent: interleave their shuffle chains
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const float Pu = __shfl_up_sync(0xffffffffu, P, o);
                const float Xu = __shfl_up_sync(0xffffffffu, X, o);
                const float Pd = __shfl_down_sync(0xffffffffu, RP, o);
                const float Xd = __shfl_down_sync(0xffffffffu, RX, o);
                if (lane >= o) {
                    X = fmaf(P, Xu, X);
                    P *= Pu;
                }
                if (lane + o < 32) {
                    RX = fmaf(RP, Xd, RX);
                    RP *= Pd;
                }
            }
            float Pex = __shfl_up_sync(0xffffffffu, P, 1);
            float Xex = __shfl_up_sync(0xffffffffu, X, 1);
            float RPex = __shfl_down_sync(0xffffffffu, RP, 1);
            float RXex = __shfl_down_sync(0xffffffffu, RX, 1);
            if (lane == 0) { Pex = 1.f; Xex = 0.f; }
            if (lane == 31) { RPex = 1.f; RXex = 0.f; }
            float h = fmaf(Pex, E, Xex);    // state entering this lane's first position
            float r = fmaf(RPex, R, RXex);  // adjoint of the position right after this lane's last one
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                h = fmaf(dec[i], h, hs[i]);
                hs[i] = h;
            }
            float dA_loc = 0.f;
            float dBv[8], dCv[8];
#pragma unroll
            for (int i = 7; i >= 0; --i) {
                r = fmaf(i == 7 ? dec_next : dec[i + 1], r, g[i] * cm[i]);
                const float drive = dt[i] * u[i];
                const float ah = fmaf(-drive, bm[i], hs[i]);  // = a_t h_{t-1}
                const float w = r * ah;
                s1[i] = fmaf(r, bm[i], s1[i]);
                ddt[i] = fmaf(An, w, ddt[i]);
                dA_loc = fmaf(dt[i], w, dA_loc);
                dBv[i] = r * drive;
                dCv[i] = g[i] * hs[i];
                y[i] = fmaf(cm[i], hs[i], y[i]);
            }
            dA_loc = warp_sum(dA_loc);
            if (lane == n) dA_acc += dA_loc;
            if (W > 1) __syncthreads();
            {
                float4 v = tdB[n * kSlots + lane];
                v.x += dBv[0]; v.y += dBv[1]; v.z += dBv[2]; v.w += dBv[3];
                tdB[n * kSlots + lane] = v;
                v = tdB[n * kSlots + 32 + lane];
                v.x += dBv[4]; v.y += dBv[5]; v.z += dBv[6]; v.w += dBv[7];
                tdB[n * kSlots + 32 + lane] = v;
                v = tdC[n * kSlots + lane];
                v.x += dCv[0]; v.y += dCv[1]; v.z += dCv[2]; v.w += dCv[3];
                tdC[n * kSlots + lane] = v;
                v = tdC[n * kSlots + 32 + lane];
                v.x += dCv[4]; v.y += dCv[5]; v.z += dCv[6]; v.w += dCv[7];
                tdC[n * kSlots + 32 + lane] = v;
            }
        }

        // ---- per-position outputs of this row
        float du_o[8], ddt_o[8];
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
        store8<T, kVec>(reinterpret_cast<T*>(a.du) + c_cur.b * a.du_bs + c_cur.d * a.du_ds, t0, L, du_o);
        store8<T, kVec>(reinterpret_cast<T*>(a.ddelta) + c_cur.b * a.ddelta_bs + c_cur.d * a.ddelta_ds, t0, L, ddt_o);
        if (a.z) {
#pragma unroll
            for (int i = 0; i < 8; ++i) dzf[i] *= y[i];
            store8<T, kVec>(reinterpret_cast<T*>(a.dz) + c_cur.b * a.dz_bs + c_cur.d * a.dz_ds, t0, L, dzf);
        }
        dD_loc = warp_sum(dD_loc);
        dbias_loc = warp_sum(dbias_loc);
        if (lane == 0) {
            if (a.dD) atomicAdd(a.dD + c_cur.d, dD_loc);
            if (a.ddelta_bias) atomicAdd(a.ddelta_bias + c_cur.d, dbias_loc);
        }
        if (lane < N) atomicAdd(a.dA + c_cur.d * N + lane, dA_acc);
        sc = sc_next;
    }
    // ---- the CTA's dB / dC partial sums -> global (fp32, 128-bit reductions)
    __syncthreads();
    {
        const RowCoord c0 = row_coord(a, rows_seq, 0);
        const int64_t bc_base = ((int64_t)c0.b * a.ngroups + c0.g) * N;
        for (int idx = threadIdx.x; idx < N * kSlots; idx += blockDim.x) {
            const int n = idx / kSlots, s = idx - n * kSlots;
            const int l = s >> 1, half = s & 1;
            const int t = unit * kUnit + l * kVecElems + half * 4;
            const float4 vb = tdB[n * kSlots + half * 32 + l];
            const float4 vc = tdC[n * kSlots + half * 32 + l];
            float* pb = a.dB + (bc_base + n) * L + t;
            float* pc = a.dC + (bc_base + n) * L + t;
            if (kVec) {
                if (t < L) {
                    atomicAdd(reinterpret_cast<float4*>(pb), vb);
                    atomicAdd(reinterpret_cast<float4*>(pc), vc);
                }
            } else {
                const float eb[4] = {vb.x, vb.y, vb.z, vb.w};
                const float ec[4] = {vc.x, vc.y, vc.z, vc.w};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (t + q < L) {
                        atomicAdd(pb + q, eb[q]);
                        atomicAdd(pc + q, ec[q]);
                    }
                }
            }
        }
    }
}

}  // namespace vv
