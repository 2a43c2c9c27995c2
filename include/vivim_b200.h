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

#define VV_VERSION 200 /* 0.2.0 */

/* element type of the streamed (B,D,L)/(B,G,N,L) tensors */
enum { VV_F32 = 0, VV_F16 = 1, VV_BF16 = 2 };

enum {
    VV_OK = 0,
    VV_ERR_BAD_ARG = -1,      /* null pointer / non-positive size / bad dtype */
    VV_ERR_UNSUPPORTED = -2,  /* valid for the reference, not served by these kernels */
    VV_ERR_ALIGN = -3         /* pointer not aligned to its element size */
};

/* Traversal order of a direction (Mamba.forward v3, mamba/mamba_ssm/modules/mamba_simple.py:217-264).  Tensors always
 * stay in MEMORY order; a direction only changes the order in which the recurrence / the causal taps visit the tokens:
 *   VV_DIR_FWD     j -> j                                   (mamba_simple.py:217-229)
 *   VV_DIR_REV     j -> L-1-j                               (xz.flip([-1]), :230-242)
 *   VV_DIR_FRAMES  j -> (j % nframes) * (L/nframes) + j / nframes
 *                  tokens are (frame, pixel) in memory and are visited pixel-major: the
 *                  chunk(nframes) / stack(-1) / flatten(-2) copy of :243-247 */
enum { VV_DIR_FWD = 0, VV_DIR_REV = 1, VV_DIR_FRAMES = 2 };
#define VV_MAX_DIRS 4

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

/* ------------------------------------------------------------------ causal conv1d of several directions, one launch
 * Replaces the per-direction causal_conv1d_fwd / _bwd calls of Mamba.forward v3 (mamba_simple.py:217-260 ->
 * selective_scan_interface.py:177, 239, 281-283) together with the xz.flip / frame-interleave copies that feed them:
 * ONE read of x produces the conv output of every direction, in memory order;
 *   out[b, k*D + d, m(j)] = act(bias[k,d] + sum_i w[k,d,i] * x[b, d, m_k(j - (K-1) + i)])      m_k: traversal of direction k
 * and the backward sums the input gradient of all directions into one dx.
 */
typedef struct {
    const void *x;       /* (B,D,L) io_dtype, strides (x_bs, x_ds, 1) */
    const float *weight; /* (ndirs,D,K) float32 contiguous */
    const float *bias;   /* (ndirs,D) float32 contiguous, or NULL */
    void *out;           /* fwd: (B,ndirs*D,L) io_dtype, strides (out_bs, out_ds, 1) */
    const void *dout;    /* bwd: (B,ndirs*D,L) io_dtype, strides (dout_bs, dout_ds, 1) */
    void *dx;            /* bwd: (B,D,L) io_dtype, strides (dx_bs, dx_ds, 1): sum over the directions; may be a view */
    float *dweight;      /* bwd: (ndirs,D,K) float32 += */
    float *dbias;        /* bwd: (ndirs,D) float32 +=, or NULL */
    int32_t batch, dim, seqlen, width; /* dim = D (channels of x); width K in {2,3,4} */
    int32_t ndirs;                     /* 1..VV_MAX_DIRS */
    int32_t dir_mode[VV_MAX_DIRS];     /* VV_DIR_* per direction */
    int32_t nframes;                   /* VV_DIR_FRAMES: frames per clip (divides seqlen); else ignored */
    int64_t x_bs, x_ds, out_bs, out_ds, dout_bs, dout_ds, dx_bs, dx_ds;
    int32_t io_dtype;                  /* VV_F32 / VV_F16 / VV_BF16 */
    int32_t silu;                      /* 0: identity, 1: SiLU */
} vv_conv1d_dirs_args;

int vv_conv1d_dirs_fwd(const vv_conv1d_dirs_args *a, void *stream);
int vv_conv1d_dirs_bwd(const vv_conv1d_dirs_args *a, void *stream);

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
    /* ---- directions (all zero = one left-to-right scan, the reference op) --------------------------------------
     * ndirs > 1: the `dim` channels are ndirs direction blocks of dim/ndirs channels (ngroups % ndirs == 0, the groups
     * are split evenly over the blocks); block k visits the tokens in order dir_mode[k].  This is Mamba.forward v3's
     * three mamba_inner_fn_no_out_proj calls (mamba_simple.py:217-260) as ONE launch over channel-concatenated
     * parameters, without the flip / interleave copies.  agg / chk / radj and the fp32 dB / dC accumulators are indexed
     * by TRAVERSAL position; every I/O tensor (u, delta, z, B, C, out*, dout, du, ddelta, dz, dB_io, dC_io) by MEMORY
     * position. */
    int32_t ndirs;                  /* 0 or 1: single direction dir_mode[0] */
    int32_t dir_mode[VV_MAX_DIRS];  /* VV_DIR_* */
    int32_t nframes;                /* VV_DIR_FRAMES: frames per clip (divides seqlen) */
    /* B / C sequence stride.  0 or 1: (B,G,N,L) layout, unit sequence stride (B_ns = state stride).  > 1: position-major,
     * e.g. straight out of x_proj's GEMM output x_dbl (B*L, R+2N) (selective_scan_interface.py:181-207 transposes
     * it instead): B_ns = 1, B_ls = R+2N, B_gs / B_bs accordingly. */
    int64_t B_ls, C_ls;
    /* dB_io / dC_io strides (elements); all zero = (B,G,N,L) contiguous.  Position-major = written straight into the
     * dx_dbl slices that selective_scan_interface.py:255-271 fills through two transposes. */
    int64_t dBio_bs, dBio_gs, dBio_ns, dBio_ls, dCio_bs, dCio_gs, dCio_ns, dCio_ls;
    /* > 0: z and dout have only gate_rows channel rows, shared by the direction blocks (row = d % gate_rows): the
     * three directions of v3 are gated by the same z and receive the same upstream gradient. */
    int32_t gate_rows;
    /* Measurement aid (bench.py, ncu): which passes to launch.  0 = all.  bit0: segment aggregates, bit1: carry fold,
     * bit2: main kernel, bit3: dB/dC cast.  Skipped passes need valid workspaces from an earlier full call. */
    int32_t pass_mask;
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

/* ------------------------------------------------------------------ LayerNorm over the channels of token tensors
 * The glue around the Mamba path in a Temporal Mamba block (modeling/vivim.py:153-157: norm1 / norm2 are nn.LayerNorm
 * over (B, tokens, C)): y = (x - mean) * rstd * weight + bias per row, eps inside the square root.  x is (rows, C) with
 * unit channel stride; `out` may be written in another dtype than x (e.g. bf16 for the GEMM that follows under
 * autocast).  C <= 512.  Backward: dx (dtype of x), dweight / dbias accumulated in fp32 (+=, zeroed by the caller).
 */
typedef struct {
    const void *x;        /* (rows, C) io_dtype, row stride x_rs */
    const float *weight;  /* (C) float32 or NULL (= 1) */
    const float *bias;    /* (C) float32 or NULL (= 0) */
    void *out;            /* fwd: (rows, C) out_dtype, row stride out_rs */
    float *mean, *rstd;   /* (rows) float32: written by fwd, read by bwd */
    const void *dout;     /* bwd: (rows, C) out_dtype, row stride dout_rs */
    void *dx;             /* bwd: (rows, C) io_dtype, row stride dx_rs, or NULL */
    float *dweight;       /* bwd: (C) float32 +=, or NULL */
    float *dbias;         /* bwd: (C) float32 +=, or NULL */
    int64_t rows;
    int32_t channels;
    int64_t x_rs, out_rs, dout_rs, dx_rs;
    int32_t io_dtype, out_dtype; /* VV_F32 / VV_F16 / VV_BF16; out_dtype == io_dtype, or io_dtype == VV_F32 */
    float eps;
} vv_layernorm_args;

int vv_layernorm_fwd(const vv_layernorm_args *a, void *stream);
int vv_layernorm_bwd(const vv_layernorm_args *a, void *stream);

/* number of kernel launches the last successful call on this thread enqueued (for bench.py) */
int vv_last_launch_count(void);

/* Test aid: 1 = every kernel takes its element-wise I/O path even where 128-bit accesses are legal (the two paths
 * must agree bit for bit, tests/test_scan_gpu.py).  Initialised once from the environment variable
 * VV_FORCE_SCALAR_IO; returns the previous value.  Process-wide. */
int vv_debug_force_scalar_io(int on);

#ifdef __cplusplus
}
#endif
#endif /* VIVIM_B200_H */
