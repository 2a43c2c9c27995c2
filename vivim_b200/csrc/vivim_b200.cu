// vivim_b200.cu -- C ABI (include/vivim_b200.h) and launch logic for the sm_100a kernels.
// Argument checks mirror the reference shims' TORCH_CHECKs
// (causal-conv1d/csrc/causal_conv1d.cpp:130-268, mamba/csrc/selective_scan/selective_scan.cpp:226-492).
#include <cuda.h>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <type_traits>

#include "../../include/vivim_b200.h"
#include "conv1d.cuh"
#include "conv1d_dirs.cuh"
#include "dwconv3d.cuh"
#include "layernorm.cuh"
#include "scan_bwd.cuh"
#include "scan_seq.cuh"

namespace {

thread_local char g_err[512] = "";
thread_local int g_launches = 0;

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int elem_size(int dtype) { return dtype == VV_F32 ? 4 : 2; }
bool valid_dtype(int dtype) { return dtype == VV_F32 || dtype == VV_F16 || dtype == VV_BF16; }

// 128-bit access is legal for a (.., L) tensor when its base and every row start are 16-byte aligned
bool vec_ok(const void* p, int es, std::initializer_list<int64_t> strides) {
    if (p == nullptr) return true;
    if (reinterpret_cast<uintptr_t>(p) % 16 != 0) return false;
    for (int64_t s : strides)
        if ((s * es) % 16 != 0) return false;
    return true;
}
bool elem_aligned(const void* p, int es) { return p == nullptr || reinterpret_cast<uintptr_t>(p) % es == 0; }

int check_launch(const char* what) {
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail((int)e, "%s: %s", what, cudaGetErrorString(e));
    ++g_launches;
    return VV_OK;
}

template <typename K>
int set_smem(K kernel, size_t bytes) {
    if (bytes <= 48 * 1024) return VV_OK;
    const cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return fail((int)e, "cudaFuncSetAttribute(%zu B): %s", bytes, cudaGetErrorString(e));
    return VV_OK;
}

int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return (v && *v) ? atoi(v) : dflt;
}

// process-wide test switch, read from the environment once (vv_debug_force_scalar_io changes it)
int& force_scalar_io() {
    static int on = env_int("VV_FORCE_SCALAR_IO", 0);
    return on;
}

// SM count of the current device (queried once per device)
int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        cached[dev] = (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) ? n : 148;
    }
    return cached[dev];
}

// ---------------------------------------------------------------- conv1d dispatch
template <typename T, bool kBwd>
int launch_conv_t(const vv_conv1d_args& a, bool vec, cudaStream_t st) {
    const dim3 grid((unsigned)((int64_t)a.batch * a.dim), (unsigned)((a.seqlen + vv::kConvTile - 1) / vv::kConvTile));
    const dim3 block(vv::kConvThreads);
#define VV_CONV_LAUNCH(SILU, VEC)                                                    \
    do {                                                                             \
        if (kBwd) vv::conv1d_bwd_kernel<T, SILU, VEC><<<grid, block, 0, st>>>(a);    \
        else vv::conv1d_fwd_kernel<T, SILU, VEC><<<grid, block, 0, st>>>(a);         \
    } while (0)
    if (a.silu) { if (vec) VV_CONV_LAUNCH(true, true); else VV_CONV_LAUNCH(true, false); }
    else        { if (vec) VV_CONV_LAUNCH(false, true); else VV_CONV_LAUNCH(false, false); }
#undef VV_CONV_LAUNCH
    return check_launch(kBwd ? "conv1d_bwd_kernel" : "conv1d_fwd_kernel");
}

template <bool kBwd>
int launch_conv(const vv_conv1d_args* a, void* stream) {
    g_launches = 0;
    if (!a) return fail(VV_ERR_BAD_ARG, "conv1d: null args");
    if (!a->x || !a->weight) return fail(VV_ERR_BAD_ARG, "conv1d: x and weight are required");
    if (!kBwd && !a->out) return fail(VV_ERR_BAD_ARG, "conv1d_fwd: out is required");
    if (kBwd && (!a->dout || !a->dx || !a->dweight)) return fail(VV_ERR_BAD_ARG, "conv1d_bwd: dout, dx, dweight are required");
    if (kBwd && a->bias && !a->dbias) return fail(VV_ERR_BAD_ARG, "conv1d_bwd: dbias is required when bias is given");
    if (a->batch <= 0 || a->dim <= 0 || a->seqlen <= 0) return fail(VV_ERR_BAD_ARG, "conv1d: sizes must be positive");
    if (!valid_dtype(a->io_dtype) || !valid_dtype(a->w_dtype)) return fail(VV_ERR_BAD_ARG, "conv1d: dtype must be fp32, fp16 or bf16");
    if (a->width < 2 || a->width > 4) return fail(VV_ERR_UNSUPPORTED, "causal_conv1d only supports width between 2 and 4");
    const int es = elem_size(a->io_dtype);
    if (!elem_aligned(a->x, es) || !elem_aligned(a->out, es) || !elem_aligned(a->dout, es) || !elem_aligned(a->dx, es))
        return fail(VV_ERR_ALIGN, "conv1d: tensor not aligned to its element size");
    bool vec = a->seqlen % 8 == 0 && vec_ok(a->x, es, {a->x_bs, a->x_ds});
    if (kBwd) vec = vec && vec_ok(a->dout, es, {a->dout_bs, a->dout_ds}) && vec_ok(a->dx, es, {a->dx_bs, a->dx_ds});
    else vec = vec && vec_ok(a->out, es, {a->out_bs, a->out_ds});
    if (force_scalar_io()) vec = false;
    if ((int64_t)a->batch * a->dim > 2147483647ll || (a->seqlen + vv::kConvTile - 1) / vv::kConvTile > 65535)
        return fail(VV_ERR_UNSUPPORTED, "conv1d: batch * dim > 2^31 - 1 or seqlen > 65535 * %d", vv::kConvTile);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    switch (a->io_dtype) {
        case VV_F32: return launch_conv_t<float, kBwd>(*a, vec, st);
        case VV_F16: return launch_conv_t<__half, kBwd>(*a, vec, st);
        default: return launch_conv_t<__nv_bfloat16, kBwd>(*a, vec, st);
    }
}

// ---------------------------------------------------------------- conv1d of several directions
template <typename T, bool kBwd>
int launch_conv_dirs_t(const vv_conv1d_dirs_args& a, bool vec, int pt, cudaStream_t st) {
    const int nf = a.nframes > 0 ? a.nframes : 1, hw = a.seqlen / nf;
    const dim3 grid((unsigned)((int64_t)a.batch * a.dim), (unsigned)((hw + pt - 1) / pt));
    const size_t tile = (size_t)nf * (pt + 2 * vv::kDirsHalo) * sizeof(float);
    const size_t smem = kBwd ? (a.ndirs + 1) * tile + (vv::kDirsThreads / 32) * VV_MAX_DIRS * 5 * sizeof(float) : tile;
    int rc;
#define VV_DIRS_LAUNCH(SILU, VEC)                                                                                     \
    do {                                                                                                              \
        if (kBwd) {                                                                                                   \
            if ((rc = set_smem(vv::conv1d_dirs_bwd_kernel<T, SILU, VEC>, smem)) != VV_OK) return rc;                  \
            vv::conv1d_dirs_bwd_kernel<T, SILU, VEC><<<grid, vv::kDirsThreads, smem, st>>>(a, pt);                    \
        } else {                                                                                                      \
            if ((rc = set_smem(vv::conv1d_dirs_fwd_kernel<T, SILU, VEC>, smem)) != VV_OK) return rc;                  \
            vv::conv1d_dirs_fwd_kernel<T, SILU, VEC><<<grid, vv::kDirsThreads, smem, st>>>(a, pt);                    \
        }                                                                                                             \
    } while (0)
    if (a.silu) { if (vec) VV_DIRS_LAUNCH(true, true); else VV_DIRS_LAUNCH(true, false); }
    else        { if (vec) VV_DIRS_LAUNCH(false, true); else VV_DIRS_LAUNCH(false, false); }
#undef VV_DIRS_LAUNCH
    return check_launch(kBwd ? "conv1d_dirs_bwd_kernel" : "conv1d_dirs_fwd_kernel");
}

template <bool kBwd>
int launch_conv_dirs(const vv_conv1d_dirs_args* a, void* stream) {
    g_launches = 0;
    if (!a) return fail(VV_ERR_BAD_ARG, "conv1d_dirs: null args");
    if (!a->x || !a->weight) return fail(VV_ERR_BAD_ARG, "conv1d_dirs: x and weight are required");
    if (!kBwd && !a->out) return fail(VV_ERR_BAD_ARG, "conv1d_dirs_fwd: out is required");
    if (kBwd && (!a->dout || !a->dx || !a->dweight)) return fail(VV_ERR_BAD_ARG, "conv1d_dirs_bwd: dout, dx, dweight are required");
    if (kBwd && a->bias && !a->dbias) return fail(VV_ERR_BAD_ARG, "conv1d_dirs_bwd: dbias is required when bias is given");
    if (a->batch <= 0 || a->dim <= 0 || a->seqlen <= 0) return fail(VV_ERR_BAD_ARG, "conv1d_dirs: sizes must be positive");
    if (!valid_dtype(a->io_dtype)) return fail(VV_ERR_BAD_ARG, "conv1d_dirs: dtype must be fp32, fp16 or bf16");
    if (a->width < 2 || a->width > 4) return fail(VV_ERR_UNSUPPORTED, "causal_conv1d only supports width between 2 and 4");
    if (a->ndirs < 1 || a->ndirs > VV_MAX_DIRS) return fail(VV_ERR_BAD_ARG, "conv1d_dirs: ndirs must be in [1, %d]", VV_MAX_DIRS);
    bool frames = false;
    for (int k = 0; k < a->ndirs; ++k) {
        const int m = a->dir_mode[k];
        if (m != VV_DIR_FWD && m != VV_DIR_REV && m != VV_DIR_FRAMES) return fail(VV_ERR_BAD_ARG, "conv1d_dirs: bad dir_mode[%d]", k);
        frames = frames || m == VV_DIR_FRAMES;
    }
    vv_conv1d_dirs_args b = *a;
    if (!frames) b.nframes = 1;       // without a frame-interleaved direction the row is one run of L tokens
    if (b.nframes <= 0 || b.seqlen % b.nframes != 0)
        return fail(VV_ERR_BAD_ARG, "conv1d_dirs: VV_DIR_FRAMES needs nframes > 0 dividing seqlen (nframes %d, seqlen %d)", b.nframes, b.seqlen);
    if (b.nframes > vv::kDirsMaxFrames) return fail(VV_ERR_UNSUPPORTED, "conv1d_dirs: nframes > %d", vv::kDirsMaxFrames);
    const int es = elem_size(a->io_dtype);
    if (!elem_aligned(a->x, es) || !elem_aligned(a->out, es) || !elem_aligned(a->dout, es) || !elem_aligned(a->dx, es))
        return fail(VV_ERR_ALIGN, "conv1d_dirs: tensor not aligned to its element size");
    const int nf = b.nframes, hw = b.seqlen / nf;
    // pixels per CTA: one 4-pixel quad per thread and frame (two without frames), shrunk until the tiles fit 64 KB
    static const int pt_env = env_int("VV_DIRS_PT", 0);     // development knob: pixels per CTA
    int pt = pt_env > 0 ? pt_env / 128 * 128 : (nf >= 3 ? 512 : 1024);
    if (pt < 128) pt = 128;
    while (pt > 128 && (size_t)(kBwd ? a->ndirs + 1 : 1) * nf * (pt + 2 * vv::kDirsHalo) * sizeof(float) > 64 * 1024) pt -= 128;
    if ((int64_t)a->batch * a->dim > 2147483647ll || (hw + pt - 1) / pt > 65535)
        return fail(VV_ERR_UNSUPPORTED, "conv1d_dirs: grid too large");
    bool vec = hw % 8 == 0 && vec_ok(a->x, es, {a->x_bs, a->x_ds});
    if (kBwd) vec = vec && vec_ok(a->dout, es, {a->dout_bs, a->dout_ds});
    if (force_scalar_io()) vec = false;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    switch (a->io_dtype) {
        case VV_F32: return launch_conv_dirs_t<float, kBwd>(b, vec, pt, st);
        case VV_F16: return launch_conv_dirs_t<__half, kBwd>(b, vec, pt, st);
        default: return launch_conv_dirs_t<__nv_bfloat16, kBwd>(b, vec, pt, st);
    }
}

// ---------------------------------------------------------------- LayerNorm dispatch
template <typename TI, typename TO, int V, bool kBwd>
int launch_ln_k(const vv_layernorm_args& a, int K, cudaStream_t st) {
    const int groups = (a.channels + V - 1) / V;               // V-element groups per row
    const int lpr = (V == 4 && groups <= 8) ? 8 : (V == 4 && groups <= 16) ? 16 : 32;
    const int rows_per_cta = vv::kLnWarps * (32 / lpr);
    const int64_t want = (a.rows + rows_per_cta - 1) / rows_per_cta;
    // forward: one row per (sub-)warp and trip; backward: few enough CTAs that the per-CTA channel atomics stay cheap
    const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)sm_count() * (kBwd ? 4 : 16)));
#define VV_LN_LAUNCH(KK, LPR)                                                                              \
    do {                                                                                                   \
        if (kBwd) vv::layernorm_bwd_kernel<TI, TO, V, KK, LPR><<<grid, vv::kLnThreads, 0, st>>>(a);        \
        else vv::layernorm_fwd_kernel<TI, TO, V, KK, LPR><<<grid, vv::kLnThreads, 0, st>>>(a);             \
    } while (0)
    if constexpr (V == 4) {
        if (lpr == 8) VV_LN_LAUNCH(1, 8);
        else if (lpr == 16) VV_LN_LAUNCH(1, 16);
        else switch (K) {
            case 1: VV_LN_LAUNCH(1, 32); break;
            case 2: VV_LN_LAUNCH(2, 32); break;
            case 3: VV_LN_LAUNCH(3, 32); break;
            default: VV_LN_LAUNCH(4, 32); break;
        }
    } else {
        if (K <= 4) VV_LN_LAUNCH(4, 32);
        else if (K <= 8) VV_LN_LAUNCH(8, 32);
        else VV_LN_LAUNCH(16, 32);
    }
#undef VV_LN_LAUNCH
    return check_launch(kBwd ? "layernorm_bwd_kernel" : "layernorm_fwd_kernel");
}

template <typename TI, typename TO, bool kBwd>
int launch_ln_t(const vv_layernorm_args& a, cudaStream_t st) {
    const int ei = sizeof(TI), eo = sizeof(TO);
    auto ok4 = [](const void* p, int64_t rs, int es) {
        return p == nullptr || (reinterpret_cast<uintptr_t>(p) % (4 * es) == 0 && (rs * es) % (4 * es) == 0);
    };
    bool vec = a.channels % 4 == 0 && ok4(a.x, a.x_rs, ei) && !force_scalar_io();
    if (kBwd) vec = vec && ok4(a.dout, a.dout_rs, eo) && ok4(a.dx, a.dx_rs, ei);
    else vec = vec && ok4(a.out, a.out_rs, eo);
    if (vec) return launch_ln_k<TI, TO, 4, kBwd>(a, (a.channels / 4 + 31) / 32, st);
    return launch_ln_k<TI, TO, 1, kBwd>(a, (a.channels + 31) / 32, st);
}

template <bool kBwd>
int launch_ln(const vv_layernorm_args* a, void* stream) {
    g_launches = 0;
    if (!a) return fail(VV_ERR_BAD_ARG, "layernorm: null args");
    if (!a->x || !a->mean || !a->rstd) return fail(VV_ERR_BAD_ARG, "layernorm: x, mean and rstd are required");
    if (!kBwd && !a->out) return fail(VV_ERR_BAD_ARG, "layernorm_fwd: out is required");
    if (kBwd && !a->dout) return fail(VV_ERR_BAD_ARG, "layernorm_bwd: dout is required");
    if (a->rows <= 0 || a->channels <= 0) return fail(VV_ERR_BAD_ARG, "layernorm: sizes must be positive");
    if (a->channels > 512) return fail(VV_ERR_UNSUPPORTED, "layernorm: this build serves channels <= 512 (got %d)", a->channels);
    if (!valid_dtype(a->io_dtype) || !valid_dtype(a->out_dtype)) return fail(VV_ERR_BAD_ARG, "layernorm: dtype must be fp32, fp16 or bf16");
    if (a->out_dtype != a->io_dtype && a->io_dtype != VV_F32)
        return fail(VV_ERR_UNSUPPORTED, "layernorm: out_dtype must equal io_dtype unless io_dtype is float32");
    if (!elem_aligned(a->x, elem_size(a->io_dtype)) || !elem_aligned(a->dx, elem_size(a->io_dtype)) ||
        !elem_aligned(a->out, elem_size(a->out_dtype)) || !elem_aligned(a->dout, elem_size(a->out_dtype)))
        return fail(VV_ERR_ALIGN, "layernorm: tensor not aligned to its element size");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (a->io_dtype == VV_F32) {
        switch (a->out_dtype) {
            case VV_F32: return launch_ln_t<float, float, kBwd>(*a, st);
            case VV_F16: return launch_ln_t<float, __half, kBwd>(*a, st);
            default: return launch_ln_t<float, __nv_bfloat16, kBwd>(*a, st);
        }
    }
    if (a->io_dtype == VV_F16) return launch_ln_t<__half, __half, kBwd>(*a, st);
    return launch_ln_t<__nv_bfloat16, __nv_bfloat16, kBwd>(*a, st);
}

// ---------------------------------------------------------------- depthwise conv3d dispatch
template <typename T, bool kBwd>
int launch_dw_t(const vv_dwconv3d_args& a, cudaStream_t st) {
    const vv::DwGeom g{a.batch, a.frames, a.height, a.width, a.channels};
    const size_t pair_bytes = 2 * sizeof(T);
    auto aligned = [&](const void* p) { return p == nullptr || reinterpret_cast<uintptr_t>(p) % pair_bytes == 0; };
    const bool pair = a.channels % 2 == 0 && aligned(a.x) && aligned(a.out) && aligned(a.dout) && aligned(a.dx) &&
                      reinterpret_cast<uintptr_t>(a.weight) % 8 == 0 && reinterpret_cast<uintptr_t>(a.bias) % 8 == 0 &&
                      !force_scalar_io();
    const int cp = (a.channels + 1) / 2, xt = (a.width + vv::kDwX - 1) / vv::kDwX;
    const int64_t ncols = (int64_t)a.batch * a.height * xt;
    const unsigned blocks = (unsigned)((ncols * cp + vv::kDwThreads - 1) / vv::kDwThreads);
    int rc = VV_OK;
#define VV_DW_STENCIL(MIRROR, IN, BIAS, OUT)                                                                          \
    do {                                                                                                              \
        if (pair) vv::dwconv3d_kernel<T, true, MIRROR><<<blocks, vv::kDwThreads, 0, st>>>(                            \
            reinterpret_cast<const T*>(IN), a.weight, BIAS, reinterpret_cast<T*>(OUT), g);                            \
        else vv::dwconv3d_kernel<T, false, MIRROR><<<blocks, vv::kDwThreads, 0, st>>>(                                \
            reinterpret_cast<const T*>(IN), a.weight, BIAS, reinterpret_cast<T*>(OUT), g);                            \
    } while (0)
    if (!kBwd) {
        VV_DW_STENCIL(false, a.x, a.bias, a.out);
        return check_launch("dwconv3d_kernel<fwd>");
    }
    if (a.dx) {
        VV_DW_STENCIL(true, a.dout, nullptr, a.dx);
        if ((rc = check_launch("dwconv3d_kernel<dgrad>")) != VV_OK) return rc;
    }
#undef VV_DW_STENCIL
    if (a.dweight) {
        const unsigned gx = (unsigned)((a.channels + 63) / 64);
        int64_t gy = (8 * sm_count() + gx - 1) / gx;                // 2 resident CTAs of 256 threads per SM (128 registers), 4 waves
        gy = std::max<int64_t>(1, std::min<int64_t>(gy, (ncols + vv::kDwCols - 1) / vv::kDwCols));
        const dim3 grid(gx, (unsigned)gy), block(32, vv::kDwCols);
        if (pair) vv::dwconv3d_wgrad_kernel<T, true><<<grid, block, 0, st>>>(
            reinterpret_cast<const T*>(a.x), reinterpret_cast<const T*>(a.dout), a.dweight, a.dbias, g);
        else vv::dwconv3d_wgrad_kernel<T, false><<<grid, block, 0, st>>>(
            reinterpret_cast<const T*>(a.x), reinterpret_cast<const T*>(a.dout), a.dweight, a.dbias, g);
        if ((rc = check_launch("dwconv3d_wgrad_kernel")) != VV_OK) return rc;
    }
    return VV_OK;
}

template <bool kBwd>
int launch_dw(const vv_dwconv3d_args* a, void* stream) {
    g_launches = 0;
    if (!a) return fail(VV_ERR_BAD_ARG, "dwconv3d: null args");
    if (!a->weight) return fail(VV_ERR_BAD_ARG, "dwconv3d: weight is required");
    if (!kBwd && (!a->x || !a->out)) return fail(VV_ERR_BAD_ARG, "dwconv3d_fwd: x and out are required");
    if (kBwd && !a->dout) return fail(VV_ERR_BAD_ARG, "dwconv3d_bwd: dout is required");
    if (kBwd && a->dweight && !a->x) return fail(VV_ERR_BAD_ARG, "dwconv3d_bwd: x is required for the weight gradient");
    if (kBwd && a->dbias && !a->dweight) return fail(VV_ERR_BAD_ARG, "dwconv3d_bwd: dbias needs dweight");
    if (a->batch <= 0 || a->frames <= 0 || a->height <= 0 || a->width <= 0 || a->channels <= 0)
        return fail(VV_ERR_BAD_ARG, "dwconv3d: sizes must be positive");
    if ((int64_t)a->batch * a->height * a->width > (1ll << 30))
        return fail(VV_ERR_UNSUPPORTED, "dwconv3d: more than 2^30 (batch, y, x) columns");
    if (!valid_dtype(a->io_dtype)) return fail(VV_ERR_BAD_ARG, "dwconv3d: dtype must be fp32, fp16 or bf16");
    const int es = elem_size(a->io_dtype);
    if (!elem_aligned(a->x, es) || !elem_aligned(a->out, es) || !elem_aligned(a->dout, es) || !elem_aligned(a->dx, es))
        return fail(VV_ERR_ALIGN, "dwconv3d: tensor not aligned to its element size");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    switch (a->io_dtype) {
        case VV_F32: return launch_dw_t<float, kBwd>(*a, st);
        case VV_F16: return launch_dw_t<__half, kBwd>(*a, st);
        default: return launch_dw_t<__nv_bfloat16, kBwd>(*a, st);
    }
}

// Launch `kernel`; with pdl = true the kernel may start while its predecessor in the stream is still
// draining (programmatic dependent launch): everything before its griddepcontrol.wait -- tile
// fills, softplus pre-pass, shared-memory zeroing -- overlaps the predecessor's tail.
template <typename... KArgs, typename... Args>
void launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl, Args... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

bool use_pdl() {
    static const bool on = env_int("VV_PDL", 1) != 0;
    return on;
}

// ---------------------------------------------------------------- scan dispatch
int check_scan_common(const vv_scan_args* a, bool bwd) {
    if (!a) return fail(VV_ERR_BAD_ARG, "scan: null args");
    if (!a->u || !a->delta || !a->A || !a->Bm || !a->Cm) return fail(VV_ERR_BAD_ARG, "scan: u, delta, A, B, C are required");
    if (!a->agg || !a->chk) return fail(VV_ERR_BAD_ARG, "scan: agg and chk workspaces are required");
    if (a->batch <= 0 || a->dim <= 0 || a->seqlen <= 0 || a->dstate <= 0 || a->ngroups <= 0)
        return fail(VV_ERR_BAD_ARG, "scan: sizes must be positive");
    if (a->dim % a->ngroups != 0) return fail(VV_ERR_BAD_ARG, "scan: ngroups must divide dim");
    if (a->dstate > 256) return fail(VV_ERR_BAD_ARG, "selective_scan only supports state dimension <= 256");
    if (a->dstate > vv::kMaxState)
        return fail(VV_ERR_UNSUPPORTED, "scan: this build serves dstate <= %d (got %d)", vv::kMaxState, a->dstate);
    if (!valid_dtype(a->io_dtype)) return fail(VV_ERR_BAD_ARG, "scan: dtype must be fp32, fp16 or bf16");
    if (a->batch > 65535) return fail(VV_ERR_UNSUPPORTED, "scan: batch > 65535");
    if (!bwd) {
        if (a->z && !a->out_z) return fail(VV_ERR_BAD_ARG, "scan_fwd: out_z is required when z is given");
        if (!a->z && !a->out) return fail(VV_ERR_BAD_ARG, "scan_fwd: out is required when z is absent");
    } else {
        if (!a->dout || !a->du || !a->ddelta || !a->dA || !a->dB || !a->dC || !a->radj)
            return fail(VV_ERR_BAD_ARG, "scan_bwd: dout, du, ddelta, dA, dB, dC, radj are required");
        if (a->z && !a->dz) return fail(VV_ERR_BAD_ARG, "scan_bwd: dz is required when z is given");
        if ((a->dB_io != nullptr) != (a->dC_io != nullptr)) return fail(VV_ERR_BAD_ARG, "scan_bwd: dB_io and dC_io need each other");
        if (a->D && !a->dD) return fail(VV_ERR_BAD_ARG, "scan_bwd: dD is required when D is given");
        if (a->delta_bias && !a->ddelta_bias) return fail(VV_ERR_BAD_ARG, "scan_bwd: ddelta_bias is required when delta_bias is given");
    }
    const int es = elem_size(a->io_dtype);
    const void* ptrs[] = {a->u, a->delta, a->Bm, a->Cm, a->z, a->out, a->out_z, a->dout, a->du, a->ddelta, a->dz,
                          a->dB_io, a->dC_io};
    for (const void* p : ptrs)
        if (!elem_aligned(p, es)) return fail(VV_ERR_ALIGN, "scan: tensor not aligned to its element size");
    // directions
    if (a->ndirs < 0 || a->ndirs > VV_MAX_DIRS) return fail(VV_ERR_BAD_ARG, "scan: ndirs must be in [0, %d]", VV_MAX_DIRS);
    const int ndirs = a->ndirs > 1 ? a->ndirs : 1;
    if (a->ngroups % ndirs != 0 || a->dim % ndirs != 0)
        return fail(VV_ERR_BAD_ARG, "scan: ndirs must divide ngroups and dim");
    for (int k = 0; k < ndirs; ++k) {
        const int m = a->dir_mode[k];
        if (m != VV_DIR_FWD && m != VV_DIR_REV && m != VV_DIR_FRAMES) return fail(VV_ERR_BAD_ARG, "scan: bad dir_mode[%d]", k);
        if (m == VV_DIR_FRAMES && (a->nframes <= 0 || a->seqlen % a->nframes != 0))
            return fail(VV_ERR_BAD_ARG, "scan: VV_DIR_FRAMES needs nframes > 0 dividing seqlen (nframes %d, seqlen %d)",
                        a->nframes, a->seqlen);
    }
    if (a->gate_rows < 0 || a->gate_rows > a->dim) return fail(VV_ERR_BAD_ARG, "scan: gate_rows must be in [0, dim]");
    if (a->B_ls < 0 || a->C_ls < 0) return fail(VV_ERR_BAD_ARG, "scan: negative B / C sequence stride");
    // grid limits: y = channel blocks of all groups, z = batch
    const int64_t dpg = a->dim / a->ngroups;
    if ((int64_t)a->ngroups * ((dpg + vv::kBwdRows - 1) / vv::kBwdRows) > 65535)
        return fail(VV_ERR_UNSUPPORTED, "scan: more than 65535 channel blocks (dim %d, ngroups %d)", a->dim, a->ngroups);
    return VV_OK;
}

bool all_forward(const vv_scan_args& a) {
    const int ndirs = a.ndirs > 1 ? a.ndirs : 1;
    for (int k = 0; k < ndirs; ++k)
        if (a.dir_mode[k] != VV_DIR_FWD) return false;
    return true;
}

// dB_io / dC_io are plain (B,G,N,L) contiguous tensors filled left to right: the flat cast kernel serves them
bool bc_io_flat(const vv_scan_args& a) {
    const int64_t N = a.dstate, L = a.seqlen, G = a.ngroups;
    auto flat = [&](int64_t bs, int64_t gs, int64_t ns, int64_t ls) {
        return (bs == 0 && gs == 0 && ns == 0 && ls == 0) || (bs == G * N * L && gs == N * L && ns == L && ls == 1);
    };
    return all_forward(a) && flat(a.dBio_bs, a.dBio_gs, a.dBio_ns, a.dBio_ls) && flat(a.dCio_bs, a.dCio_gs, a.dCio_ns, a.dCio_ls);
}

bool scan_vec_ok(const vv_scan_args& a, bool bwd) {
    const int es = elem_size(a.io_dtype);
    bool v = a.seqlen % 8 == 0;
    v = v && vec_ok(a.u, es, {a.u_bs, a.u_ds}) && vec_ok(a.delta, es, {a.delta_bs, a.delta_ds});
    // B / C: 128-bit row loads only in the (.., N, L) layout; position-major rows are read in (position, 4 states) items
    // whose alignment is checked per item in the kernel
    if (a.B_ls <= 1) v = v && vec_ok(a.Bm, es, {a.B_bs, a.B_gs, a.B_ns});
    if (a.C_ls <= 1) v = v && vec_ok(a.Cm, es, {a.C_bs, a.C_gs, a.C_ns});
    v = v && vec_ok(a.z, es, {a.z_bs, a.z_ds});
    if (!bwd) {
        v = v && vec_ok(a.out, es, {a.out_bs, a.out_ds}) && vec_ok(a.out_z, es, {a.outz_bs, a.outz_ds});
    } else {
        v = v && vec_ok(a.dout, es, {a.dout_bs, a.dout_ds}) && vec_ok(a.du, es, {a.du_bs, a.du_ds});
        v = v && vec_ok(a.ddelta, es, {a.ddelta_bs, a.ddelta_ds}) && vec_ok(a.dz, es, {a.dz_bs, a.dz_ds});
        v = v && vec_ok(a.dB, 4, {}) && vec_ok(a.dC, 4, {});
    }
    if (force_scalar_io()) v = false;
    return v;
}


// geometry of the segment kernels (scan_seq.cuh): 32 channels x one 64-position segment per CTA
struct SegPlan {
    int NB;        // compile-time state block: 8, 16 or 32
    int segs;
    dim3 grid;
};

int pass_mask(const vv_scan_args& a) { return a.pass_mask ? (a.pass_mask & 15) : 15; }

SegPlan plan_seg(const vv_scan_args& a) {
    SegPlan p;
    p.segs = vv_scan_num_segments(a.seqlen);
    p.NB = a.dstate <= 8 ? 8 : (a.dstate <= 16 ? 16 : 32);
    const int dpg = a.dim / a.ngroups;
    p.grid = dim3(p.segs, a.ngroups * ((dpg + vv::kSegRows - 1) / vv::kSegRows), a.batch);
    return p;
}

// the launch is the plain reference op: one left-to-right direction, (.., N, L) B / C, no shared gate rows
bool plain_scan(const vv_scan_args& a) {
    return all_forward(a) && a.B_ls <= 1 && a.C_ls <= 1 && a.gate_rows == 0;
}

template <typename T, bool kVec, int NB, bool kRev>
int launch_seg_agg(const vv_scan_args& a, const SegPlan& p, cudaStream_t st) {
    const size_t smem = 2 * (size_t)vv::kSegRows * vv::kF32Pitch + (size_t)vv::kSeg * NB * sizeof(float);
    int rc;
    auto kernel = plain_scan(a) ? vv::seg_agg_kernel<T, kVec, NB, kRev, false> : vv::seg_agg_kernel<T, kVec, NB, kRev, true>;
    if ((rc = set_smem(kernel, smem)) != VV_OK) return rc;
    launch_kernel(kernel, p.grid, dim3(vv::kSegThreads), smem, st, use_pdl(), a);
    return check_launch(kRev ? "seg_agg_kernel<rev>" : "seg_agg_kernel<fwd>");
}

template <typename T, bool kVec, int NB>
int launch_seg_fwd(const vv_scan_args& a, const SegPlan& p, cudaStream_t st) {
    const size_t smem = 2 * (size_t)vv::kSegRows * vv::kF32Pitch + 2 * (size_t)vv::kSeg * NB * sizeof(float) +
                        2 * (size_t)vv::kSegRows * vv::SegTile<T>::kPitch;
    int rc;
    auto kernel = plain_scan(a) ? vv::seg_fwd_kernel<T, kVec, NB, false> : vv::seg_fwd_kernel<T, kVec, NB, true>;
    if ((rc = set_smem(kernel, smem)) != VV_OK) return rc;
    launch_kernel(kernel, p.grid, dim3(vv::kSegThreads), smem, st, use_pdl() && (pass_mask(a) & 2), a);
    return check_launch("seg_fwd_kernel");
}

template <bool kRev>
int launch_seg_carry(const vv_scan_args& a, const SegPlan& p, cudaStream_t st) {
    const int64_t rows = (int64_t)a.batch * a.dim;
    // >= 4 segments per thread when the sequence is short, at most 512 / N chunks per row
    const int chunks = std::max(1, std::min(vv::kCarryThreads / a.dstate, (p.segs + 3) / 4));
    const int rows_per_cta = std::max(1, vv::kCarryThreads / (chunks * a.dstate));
    launch_kernel(vv::seg_carry_kernel<kRev>, dim3((unsigned)((rows + rows_per_cta - 1) / rows_per_cta)),
                  dim3(vv::kCarryThreads), 0, st, use_pdl() && (pass_mask(a) & 1), reinterpret_cast<const float2*>(a.agg),
                  kRev ? a.radj : a.chk, kRev ? (float*)nullptr : a.last_state, p.segs, a.dstate, chunks, rows_per_cta, rows);
    return check_launch(kRev ? "seg_carry_kernel<rev>" : "seg_carry_kernel<fwd>");
}

#define VV_NB_SWITCH(NBVAL, ...)                 \
    switch (NBVAL) {                             \
        case 8: { constexpr int NB = 8; __VA_ARGS__; break; }   \
        case 16: { constexpr int NB = 16; __VA_ARGS__; break; } \
        default: { constexpr int NB = 32; __VA_ARGS__; break; } \
    }

// ---- TMA variant of the backward's B / C tile fill (measurement: VV_TMA_BC=1, see profiles/r02_tma.md)
bool tma_bc_enabled() {
    static const bool on = env_int("VV_TMA_BC", 0) != 0;
    return on;
}

// rank-2 tensor maps over B and C viewed as (B*G*N rows, L) with a (64 positions x N rows) box; false = not encodable
// (layout not contiguous in (N, L), strides not multiples of 16 bytes, driver entry point missing): the caller then
// takes the register route.
template <typename T>
bool encode_bc_maps(const vv_scan_args& a, vv::BwdTmaMaps& maps) {
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn encode = [] {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            fn = nullptr;
        return reinterpret_cast<EncodeFn>(fn);
    }();
    if (!encode) return false;
    const int64_t N = a.dstate, L = a.seqlen, G = a.ngroups;
    auto contiguous = [&](int64_t bs, int64_t gs, int64_t ns) { return ns == L && gs == N * L && bs == G * N * L; };
    if (!contiguous(a.B_bs, a.B_gs, a.B_ns) || !contiguous(a.C_bs, a.C_gs, a.C_ns) || (L * (int64_t)sizeof(T)) % 16 != 0) return false;
    static_assert(sizeof(CUtensorMap) <= 128, "CUtensorMap does not fit the kernel parameter slot");
    const cuuint64_t dims[2] = {(cuuint64_t)L, (cuuint64_t)(a.batch * G * N)};
    const cuuint64_t strides[1] = {(cuuint64_t)(L * sizeof(T))};
    const cuuint32_t box[2] = {(cuuint32_t)vv::kSeg, (cuuint32_t)N};
    const cuuint32_t estr[2] = {1, 1};
    const CUtensorMapDataType dt = std::is_same<T, __half>::value ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    const void* ptrs[2] = {a.Bm, a.Cm};
    unsigned char* dst[2] = {maps.B, maps.C};
    for (int i = 0; i < 2; ++i) {
        if (reinterpret_cast<uintptr_t>(ptrs[i]) % 16 != 0) return false;
        CUtensorMap m;
        if (encode(&m, dt, 2, const_cast<void*>(ptrs[i]), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return false;
        memcpy(dst[i], &m, sizeof(m));
    }
    return true;
}

template <typename T, bool kVec>
int scan_fwd_t(const vv_scan_args& a, cudaStream_t st) {
    const SegPlan p = plan_seg(a);
    int rc = VV_OK;
    if (pass_mask(a) & 1) {
        VV_NB_SWITCH(p.NB, rc = (launch_seg_agg<T, kVec, NB, false>(a, p, st)));
        if (rc != VV_OK) return rc;
    }
    if (pass_mask(a) & 2) {
        if ((rc = launch_seg_carry<false>(a, p, st)) != VV_OK) return rc;
    }
    if (pass_mask(a) & 4) {
        VV_NB_SWITCH(p.NB, rc = (launch_seg_fwd<T, kVec, NB>(a, p, st)));
        if (rc != VV_OK) return rc;
    }
    return VV_OK;
}

template <typename T, bool kVec>
int scan_bwd_t(const vv_scan_args& a, cudaStream_t st) {
    const SegPlan sp = plan_seg(a);
    int rc = VV_OK;
    if (pass_mask(a) & 1) {
        VV_NB_SWITCH(sp.NB, rc = (launch_seg_agg<T, kVec, NB, true>(a, sp, st)));
        if (rc != VV_OK) return rc;
    }
    if (pass_mask(a) & 2) {
        if ((rc = launch_seg_carry<true>(a, sp, st)) != VV_OK) return rc;
    }
    if (pass_mask(a) & 4) {
        const int dpg = a.dim / a.ngroups;
        const dim3 grid(sp.segs, a.ngroups * ((dpg + vv::kBwdRows - 1) / vv::kBwdRows), a.batch);
        VV_NB_SWITCH(sp.NB, {
            const size_t smem = vv::bwd_smem_bytes(NB);
            vv::BwdTmaMaps maps;
            memset(&maps, 0, sizeof(maps));
            const bool tma = kVec && sizeof(T) == 2 && plain_scan(a) && tma_bc_enabled() && encode_bc_maps<T>(a, maps);
            auto kernel = tma ? vv::seg_bwd_kernel<T, kVec && sizeof(T) == 2, NB, false, kVec && sizeof(T) == 2>
                        : plain_scan(a) ? vv::seg_bwd_kernel<T, kVec, NB, false, false> : vv::seg_bwd_kernel<T, kVec, NB, true, false>;
            if ((rc = set_smem(kernel, smem)) != VV_OK) return rc;
            launch_kernel(kernel, grid, dim3(vv::kBwdThreads), smem, st, use_pdl() && (pass_mask(a) & 2), a, maps);
        });
        if ((rc = check_launch("seg_bwd_kernel")) != VV_OK) return rc;
    }
    if ((pass_mask(a) & 8) && a.dB_io != nullptr) {
        if (bc_io_flat(a)) {
            const int64_t n = (int64_t)a.batch * a.ngroups * a.dstate * a.seqlen;
            const dim3 cgrid((unsigned)((n + 2047) / 2048), 2);
            launch_kernel(vv::cast_bc_kernel<T>, cgrid, dim3(256), 0, st, use_pdl() && (pass_mask(a) & 4), (const float*)a.dB, (const float*)a.dC,
                          reinterpret_cast<T*>(a.dB_io), reinterpret_cast<T*>(a.dC_io), n);
            if ((rc = check_launch("cast_bc_kernel")) != VV_OK) return rc;
        } else {
            vv_scan_args b = a;
            if (b.dBio_bs == 0 && b.dBio_gs == 0 && b.dBio_ns == 0 && b.dBio_ls == 0) {
                b.dBio_ls = 1; b.dBio_ns = a.seqlen; b.dBio_gs = (int64_t)a.dstate * a.seqlen; b.dBio_bs = b.dBio_gs * a.ngroups;
            }
            if (b.dCio_bs == 0 && b.dCio_gs == 0 && b.dCio_ns == 0 && b.dCio_ls == 0) {
                b.dCio_ls = 1; b.dCio_ns = a.seqlen; b.dCio_gs = (int64_t)a.dstate * a.seqlen; b.dCio_bs = b.dCio_gs * a.ngroups;
            }
            launch_kernel(vv::cast_bc_strided_kernel<T>, dim3(sp.segs, a.ngroups, a.batch), dim3(vv::kCastThreads), 0, st,
                          use_pdl() && (pass_mask(a) & 4), b);
            if ((rc = check_launch("cast_bc_strided_kernel")) != VV_OK) return rc;
        }
    }
    return VV_OK;
}

template <bool kBwd>
int launch_scan(const vv_scan_args* a, void* stream) {
    g_launches = 0;
    const int rc = check_scan_common(a, kBwd);
    if (rc != VV_OK) return rc;
    const bool vec = scan_vec_ok(*a, kBwd);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
#define VV_SCAN_CALL(T)                                                                      \
    do {                                                                                     \
        if (kBwd) return vec ? scan_bwd_t<T, true>(*a, st) : scan_bwd_t<T, false>(*a, st);   \
        return vec ? scan_fwd_t<T, true>(*a, st) : scan_fwd_t<T, false>(*a, st);             \
    } while (0)
    switch (a->io_dtype) {
        case VV_F32: VV_SCAN_CALL(float);
        case VV_F16: VV_SCAN_CALL(__half);
        default: VV_SCAN_CALL(__nv_bfloat16);
    }
#undef VV_SCAN_CALL
}

}  // namespace

extern "C" {

int vv_version(void) { return VV_VERSION; }
const char* vv_last_error(void) { return g_err; }
int vv_last_launch_count(void) { return g_launches; }
int vv_debug_force_scalar_io(int on) {
    const int prev = force_scalar_io();
    force_scalar_io() = on != 0;
    return prev;
}
int vv_scan_num_segments(int seqlen) { return seqlen <= 0 ? 0 : (seqlen + VV_SCAN_SEGMENT - 1) / VV_SCAN_SEGMENT; }

int vv_conv1d_fwd(const vv_conv1d_args* a, void* stream) { return launch_conv<false>(a, stream); }
int vv_conv1d_bwd(const vv_conv1d_args* a, void* stream) { return launch_conv<true>(a, stream); }
int vv_conv1d_dirs_fwd(const vv_conv1d_dirs_args* a, void* stream) { return launch_conv_dirs<false>(a, stream); }
int vv_conv1d_dirs_bwd(const vv_conv1d_dirs_args* a, void* stream) { return launch_conv_dirs<true>(a, stream); }
int vv_dwconv3d_fwd(const vv_dwconv3d_args* a, void* stream) { return launch_dw<false>(a, stream); }
int vv_dwconv3d_bwd(const vv_dwconv3d_args* a, void* stream) { return launch_dw<true>(a, stream); }
int vv_layernorm_fwd(const vv_layernorm_args* a, void* stream) { return launch_ln<false>(a, stream); }
int vv_layernorm_bwd(const vv_layernorm_args* a, void* stream) { return launch_ln<true>(a, stream); }
int vv_scan_fwd(const vv_scan_args* a, void* stream) { return launch_scan<false>(a, stream); }
int vv_scan_bwd(const vv_scan_args* a, void* stream) { return launch_scan<true>(a, stream); }

}  // extern "C"
