/*
 * vivim_b200.h -- C ABI of libvivim_b200.so: the B200 (sm_100a) kernels behind Vivim's
 * Temporal-Mamba hot path (causal depthwise conv1d fwd/bwd + fused selective scan fwd/bwd).
 *
 * Drop-in boundary.  The reference binds this path through two pybind11 torch extensions that take
 * at::Tensor (reference: causal-conv1d/csrc/causal_conv1d.cpp:329-333 `causal_conv1d_fwd/_bwd`,
 * mamba/csrc/selective_scan/selective_scan.cpp:494-497 `fwd/bwd`).  This library replaces what sits
 * *below* that binding: raw device pointers + sizes + element strides + a dtype code + a
 * cudaStream_t, no torch types, no allocation, no synchronisation, never throws.  The Python
 * packages `causal_conv1d` / `mamba_ssm` (ctypes) own contiguity fixes, output allocation,
 * zero-initialisation of accumulators and dtype casts, exactly where the reference's .cpp shims
 * do them (causal_conv1d.cpp:130-268, selective_scan.cpp:226-492).  See INTEGRATION.md.
 *
 * Conventions
 *   - every function returns 0 on success, a positive cudaError_t from the launch, or a negative
 *     VV_ERR_* code for an argument the kernels do not serve; vv_last_error() (thread local)
 *     describes the most recent failure.
 *   - strides are in ELEMENTS; the innermost (sequence) stride of every (.., L) tensor must be 1
 *     (same rule as the reference: selective_scan.cpp:262-267, causal_conv1d.cpp:155).
 *   - `stream` is a cudaStream_t passed as void*; the kernels run on the current device.
 *   - accumulated outputs (marked +=) must be zero-initialised by the caller; they are float32.
 */
#ifndef VIVIM_B200_H
#define VIVIM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VV_VERSION 100 /* 0.1.0 */

/* element type of the streamed (B,D,L)/(B,G,N,L) tensors */
enum { VV_F32 = 0, VV_F16 = 1, VV_BF16 = 2 };

enum {
    VV_OK = 0,
    VV_ERR_BAD_ARG = -1,      /* null pointer / non-positive size / bad dtype */
    VV_ERR_UNSUPPORTED = -2,  /* valid for the reference, not served by these kernels */
    VV_ERR_ALIGN = -3         /* pointer not aligned to its element size */
};

/* Sequence positions per checkpoint segment: scan states are checkpointed / carried every 64
 * positions; one segment x 16 channels is the work unit of a CTA in every scan kernel. */
#define VV_SCAN_SEGMENT 64

int vv_version(void);
const char *vv_last_error(void);

/* number of checkpoint segments for a sequence of length L: ceil(L / VV_SCAN_SEGMENT) */
int vv_scan_num_segments(int seqlen);

/* ------------------------------------------------------------------ causal depthwise conv1d
 * Replaces causal_conv1d_cuda.causal_conv1d_fwd / _bwd, channel-first layout only
 * (causal-conv1d/csrc/causal_conv1d.cpp:130-189, 191-268; kernels causal_conv1d_fwd.cu:39-130,
 * causal_conv1d_bwd.cu:46-240).  out[b,d,t] = act(bias[d] + sum_k w[d,k] x[b,d,t-(K-1)+k]).
 */
typedef struct {
    const void *x;      /* (B,D,L)  io_dtype, strides (x_bs, x_ds, 1) */
    const void *weight; /* (D,K)    w_dtype, contiguous */
    const void *bias;   /* (D)      w_dtype, or NULL */
    void *out;          /* fwd: (B,D,L) io_dtype, strides (out_bs, out_ds, 1) */
    const void *dout;   /* bwd: (B,D,L) io_dtype, strides (dout_bs, dout_ds, 1) */
    void *dx;           /* bwd: (B,D,L) io_dtype, strides (dx_bs, dx_ds, 1); may be a view */
    float *dweight;     /* bwd: (D,K) float32  += */
    float *dbias;       /* bwd: (D)   float32  +=, or NULL */
    int32_t batch, dim, seqlen, width; /* width K in {2,3,4} */
    int64_t x_bs, x_ds, out_bs, out_ds, dout_bs, dout_ds, dx_bs, dx_ds;
    int32_t io_dtype, w_dtype; /* VV_F32 / VV_F16 / VV_BF16 */
    int32_t silu;              /* 0: identity, 1: SiLU */
} vv_conv1d_args;

int vv_conv1d_fwd(const vv_conv1d_args *a, void *stream);
int vv_conv1d_bwd(const vv_conv1d_args *a, void *stream);

/* ------------------------------------------------------------------ selective scan
 * Replaces selective_scan_cuda.fwd / .bwd for real A and input-dependent B and C
 * (mamba/csrc/selective_scan/selective_scan.cpp:226-336, 338-492; kernels
 * selective_scan_fwd_kernel.cuh:67-303, selective_scan_bwd_kernel.cuh:75-489):
 *   dt = softplus?(delta + delta_bias);  h_t = exp(dt A) h_{t-1} + dt B_t u_t;
 *   y_t = <C_t, h_t> + D u_t;  out = y;  out_z = y * silu(z).
 *
 * Workspaces (float32, caller-allocated, contents need not be initialised), S = vv_scan_num_segments(L):
 *   agg  : 2 * B*D*S*N floats   per-segment scan aggregates (decay product, local state)
 *   chk  : B*D*S*N floats       fwd: state entering each segment (saved for bwd)
 *                               bwd: the same tensor, read
 *   radj : B*D*S*N floats       bwd only: adjoint entering each segment from the right, already multiplied by the
 *                               decay of the first position to its right (e_t = a_t r_t)
 * The reference's `x` intermediate (B,D,n_chunks,2N) (selective_scan.cpp:307-313) is replaced by chk.
 */
typedef struct {
    const void *u;          /* (B,D,L) io_dtype */
    const void *delta;      /* (B,D,L) io_dtype */
    const float *A;         /* (D,N)   float32, strides (A_ds, A_ns) */
    const void *Bm;         /* (B,G,N,L) io_dtype, strides (B_bs, B_gs, B_ns, 1) */
    const void *Cm;         /* (B,G,N,L) io_dtype, strides (C_bs, C_gs, C_ns, 1) */
    const float *D;         /* (D) float32 or NULL */
    const void *z;          /* (B,D,L) io_dtype or NULL */
    const float *delta_bias;/* (D) float32 or NULL */
    void *out;              /* fwd: (B,D,L) io_dtype, pre-gate y; may be NULL when z != NULL */
    void *out_z;            /* fwd: (B,D,L) io_dtype, y*silu(z); required iff z != NULL */
    float *last_state;      /* fwd: (B,D,N) float32 contiguous, or NULL */
    float *agg, *chk, *radj;/* workspaces, see above (radj: bwd only) */
    /* backward only */
    const void *dout;       /* (B,D,L) io_dtype: grad of out_z (z != NULL) or of out */
    void *du, *ddelta;      /* (B,D,L) io_dtype */
    void *dz;               /* (B,D,L) io_dtype, required iff z != NULL; may be a view */
    float *dA;              /* (D,N) float32 contiguous  += */
    float *dB, *dC;         /* (B,G,N,L) float32 contiguous  += */
    float *dD;              /* (D) float32 += , required iff D != NULL */
    float *ddelta_bias;     /* (D) float32 += , required iff delta_bias != NULL */
    int32_t batch, dim, seqlen, dstate, ngroups; /* dstate N <= 256, G divides D */
    int64_t u_bs, u_ds, delta_bs, delta_ds, z_bs, z_ds, out_bs, out_ds, outz_bs, outz_ds;
    int64_t A_ds, A_ns, B_bs, B_gs, B_ns, C_bs, C_gs, C_ns;
    int64_t dout_bs, dout_ds, du_bs, du_ds, ddelta_bs, ddelta_ds, dz_bs, dz_ds;
    int32_t io_dtype;       /* VV_F32 / VV_F16 / VV_BF16 */
    int32_t delta_softplus; /* 0 / 1 */
    /* bwd, optional: dB / dC additionally in the I/O dtype (B,G,N,L contiguous) -- the cast the reference's shim performs
     * with `.to(B.dtype)` (selective_scan.cpp:488), as a fourth kernel chained to the backward with programmatic
     * dependent launch instead of a separate framework op.  NULL = fp32 dB / dC only. */
    void *dB_io, *dC_io;
    int32_t zero_accumulators; /* bwd: 1 = vv_scan_bwd zero-fills dA, dB, dC, dD, ddelta_bias itself (in its first kernel,
                                  no separate memset launch); 0 = the caller has zeroed them (reference convention,
                                  selective_scan.cpp:460-466) */
} vv_scan_args;

int vv_scan_fwd(const vv_scan_args *a, void *stream);
int vv_scan_bwd(const vv_scan_args *a, void *stream);

/* ------------------------------------------------------------------ depthwise 3x3x3 conv over (frame, y, x)
 * Replaces what the reference's DWConv module (modeling/vivim.py:57-68: tokens -> transpose -> nn.Conv3d(C, C, 3, 1, 1,
 * groups=C) -> flatten -> transpose) asks of cuDNN, directly on the token layout: x, out, dout, dx are
 * (B, frames, H, W, C) contiguous, channels innermost (= the (B, N, C) token tensor, N = frames*H*W).
 *   out[b,t,y,x,c] = bias[c] + sum_{dt,dy,dx in 0..2} weight[(dt*3+dy)*3+dx, c] * in[b, t+dt-1, y+dy-1, x+dx-1, c]
 * weight is the Conv3d parameter (C,1,3,3,3) as float32, TAP-MAJOR: (27,C) = param.view(C,27).t(); zero padding.
 */
typedef struct {
    const void *x;        /* (B,T,H,W,C) io_dtype */
    const float *weight;  /* (27,C) float32, tap-major */
    const float *bias;    /* (C) float32 or NULL */
    void *out;            /* fwd: (B,T,H,W,C) io_dtype */
    const void *dout;     /* bwd: (B,T,H,W,C) io_dtype */
    void *dx;             /* bwd: (B,T,H,W,C) io_dtype, or NULL to skip the input gradient */
    float *dweight;       /* bwd: (27,C) float32 +=, or NULL to skip the parameter gradients */
    float *dbias;         /* bwd: (C) float32 +=, or NULL */
    int32_t batch, frames, height, width, channels;
    int32_t io_dtype;     /* VV_F32 / VV_F16 / VV_BF16 */
} vv_dwconv3d_args;

int vv_dwconv3d_fwd(const vv_dwconv3d_args *a, void *stream);
int vv_dwconv3d_bwd(const vv_dwconv3d_args *a, void *stream);

/* number of kernel launches the last successful call on this thread enqueued (for bench.py) */
int vv_last_launch_count(void);

/* Measurement aid (bench.py, ncu): restrict which of the three scan passes subsequent vv_scan_fwd /
 * vv_scan_bwd calls on this thread launch.  bit0: segment aggregates, bit1: carry fold, bit2: main
 * kernel, bit3: the dB/dC cast of vv_scan_bwd.  Default 15 (all).  Returns the previous mask.  Workspaces must hold valid data from an
 * earlier full call when a pass is skipped. */
int vv_scan_set_pass_mask(int mask);

#ifdef __cplusplus
}
#endif
#endif /* VIVIM_B200_H */
