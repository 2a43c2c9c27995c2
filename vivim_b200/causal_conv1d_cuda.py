"""Tensor-level entry points with the reference pybind module's names and argument order
(causal-conv1d/csrc/causal_conv1d.cpp:329-333): ``causal_conv1d_fwd``, ``causal_conv1d_bwd``.

What the reference's C++ shim does around its kernels is done here around the C ABI call:
argument checks (causal_conv1d.cpp:136-163, 198-237), output allocation, fp32 zero-initialised
dweight/dbias cast back to the weight dtype (:247-249, :267), device guard and current stream.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib

_DTYPES = {torch.float32: _lib.VV_F32, torch.float16: _lib.VV_F16, torch.bfloat16: _lib.VV_BF16}

LAUNCHES = 0  # kernels enqueued through this module (bench.py reads the total)


def _check(cond, msg):
    if not cond:
        raise RuntimeError(msg)


def _common_checks(x, weight, bias):
    _check(x.is_cuda and weight.is_cuda, "causal_conv1d: x and weight must be CUDA tensors")
    _check(x.dtype in _DTYPES, "causal_conv1d: x must be float32, float16 or bfloat16")
    _check(weight.dtype in _DTYPES, "causal_conv1d: weight must be float32, float16 or bfloat16")
    _check(x.dim() == 3 and weight.dim() == 2, "causal_conv1d: x must be (batch, dim, seqlen), weight (dim, width)")
    batch, dim, seqlen = x.shape
    _check(weight.shape[0] == dim, "causal_conv1d: weight must have shape (dim, width)")
    _check(2 <= weight.shape[1] <= 4, "causal_conv1d only supports width between 2 and 4")
    _check(x.stride(2) == 1 or x.stride(1) == 1, "causal_conv1d: x must be contiguous along seqlen or dim")
    if bias is not None:
        _check(bias.dtype == weight.dtype, "causal_conv1d: bias must have the dtype of weight")
        _check(bias.is_cuda and bias.shape == (dim,), "causal_conv1d: bias must have shape (dim,)")
        _check(bias.stride(-1) == 1, "causal_conv1d: bias must be contiguous")


def _channel_first(t):
    # The channel-last layout (stride(1) == 1) is accepted like in the reference but served by a
    # copy to channel-first: Vivim never produces it (component 2).
    return t if t.stride(2) == 1 else t.contiguous()


def causal_conv1d_fwd(x, weight, bias, silu_activation):
    global LAUNCHES
    _common_checks(x, weight, bias)
    x = _channel_first(x)
    weight = weight.contiguous()
    out = torch.empty_like(x, memory_format=torch.contiguous_format)
    if x.numel() == 0:
        return out
    a = _lib.ConvArgs()
    a.x, a.weight, a.out = x.data_ptr(), weight.data_ptr(), out.data_ptr()
    a.bias = bias.data_ptr() if bias is not None else None
    a.batch, a.dim, a.seqlen = x.shape
    a.width = weight.shape[1]
    a.x_bs, a.x_ds = x.stride(0), x.stride(1)
    a.out_bs, a.out_ds = out.stride(0), out.stride(1)
    a.io_dtype, a.w_dtype = _DTYPES[x.dtype], _DTYPES[weight.dtype]
    a.silu = int(bool(silu_activation))
    with torch.cuda.device(x.device):
        stream = torch.cuda.current_stream().cuda_stream
        _lib.check(_lib.lib().vv_conv1d_fwd(ctypes.byref(a), ctypes.c_void_p(stream)), "vv_conv1d_fwd")
    LAUNCHES += 1
    return out


def causal_conv1d_bwd(x, weight, bias, dout, dx, silu_activation):
    """-> (dx, dweight, dbias); ``dx`` may be a caller-provided view (e.g. half of dxz)."""
    global LAUNCHES
    _common_checks(x, weight, bias)
    _check(dout.is_cuda and dout.shape == x.shape and dout.dtype == x.dtype,
           "causal_conv1d_bwd: dout must match x")
    x = _channel_first(x)
    dout = _channel_first(dout)
    weight = weight.contiguous()
    if dx is None:
        dx = torch.empty_like(x, memory_format=torch.contiguous_format)
    else:
        _check(dx.shape == x.shape and dx.dtype == x.dtype and dx.stride(2) == 1,
               "causal_conv1d_bwd: dx must match x and be contiguous along seqlen")
    dweight = torch.zeros(weight.shape, dtype=torch.float32, device=x.device)
    dbias = torch.zeros(weight.shape[0], dtype=torch.float32, device=x.device) if bias is not None else None
    if x.numel() > 0:
        a = _lib.ConvArgs()
        a.x, a.weight, a.dout, a.dx = x.data_ptr(), weight.data_ptr(), dout.data_ptr(), dx.data_ptr()
        a.bias = bias.data_ptr() if bias is not None else None
        a.dweight = dweight.data_ptr()
        a.dbias = dbias.data_ptr() if dbias is not None else None
        a.batch, a.dim, a.seqlen = x.shape
        a.width = weight.shape[1]
        a.x_bs, a.x_ds = x.stride(0), x.stride(1)
        a.dout_bs, a.dout_ds = dout.stride(0), dout.stride(1)
        a.dx_bs, a.dx_ds = dx.stride(0), dx.stride(1)
        a.io_dtype, a.w_dtype = _DTYPES[x.dtype], _DTYPES[weight.dtype]
        a.silu = int(bool(silu_activation))
        with torch.cuda.device(x.device):
            stream = torch.cuda.current_stream().cuda_stream
            _lib.check(_lib.lib().vv_conv1d_bwd(ctypes.byref(a), ctypes.c_void_p(stream)), "vv_conv1d_bwd")
        LAUNCHES += 1
    return (dx, dweight.to(weight.dtype),
            dbias.to(weight.dtype) if dbias is not None else None)


# ---------------------------------------------------------------------------------------------
# all scan directions of a Temporal Mamba block in one launch (csrc/conv1d_dirs.cuh)
# ---------------------------------------------------------------------------------------------
DIR_MODES = {"fwd": _lib.VV_DIR_FWD, "rev": _lib.VV_DIR_REV, "frames": _lib.VV_DIR_FRAMES}


def dir_codes(dirs):
    """('fwd', 'rev', 'frames') / VV_DIR_* codes -> tuple of codes."""
    codes = tuple(DIR_MODES[d] if isinstance(d, str) else int(d) for d in dirs)
    _check(1 <= len(codes) <= _lib.VV_MAX_DIRS, f"between 1 and {_lib.VV_MAX_DIRS} directions")
    return codes


def _dirs_checks(x, weight, bias, dirs, nframes):
    _check(x.is_cuda and x.dtype in _DTYPES and x.dim() == 3 and x.stride(2) == 1,
           "causal_conv1d_dirs: x must be a CUDA (batch, dim, seqlen) tensor, contiguous along seqlen")
    nd = len(dirs)
    _check(weight.dtype == torch.float32 and weight.dim() == 3 and weight.shape[0] == nd and weight.shape[1] == x.shape[1]
           and weight.is_contiguous(), "causal_conv1d_dirs: weight must be contiguous float32 (ndirs, dim, width)")
    _check(2 <= weight.shape[2] <= 4, "causal_conv1d only supports width between 2 and 4")
    if bias is not None:
        _check(bias.dtype == torch.float32 and bias.shape == weight.shape[:2] and bias.is_contiguous(),
               "causal_conv1d_dirs: bias must be contiguous float32 (ndirs, dim)")
    if _lib.VV_DIR_FRAMES in dirs:
        _check(nframes > 0 and x.shape[2] % nframes == 0, "causal_conv1d_dirs: nframes must divide seqlen")


def _fill_dirs(a, x, weight, bias, dirs, nframes, silu):
    a.x, a.weight = x.data_ptr(), weight.data_ptr()
    a.bias = bias.data_ptr() if bias is not None else None
    a.batch, a.dim, a.seqlen = x.shape
    a.width = weight.shape[2]
    a.ndirs = len(dirs)
    for k, m in enumerate(dirs):
        a.dir_mode[k] = m
    a.nframes = int(nframes)
    a.x_bs, a.x_ds = x.stride(0), x.stride(1)
    a.io_dtype = _DTYPES[x.dtype]
    a.silu = int(bool(silu))


def causal_conv1d_dirs_fwd(x, weight, bias, dirs, nframes, silu_activation=True):
    """x (B, D, L), weight (ndirs, D, K) fp32, bias (ndirs, D) fp32 -> out (B, ndirs*D, L): direction k's causal conv
    (taps along its own traversal order, see include/vivim_b200.h) in channels [k*D, (k+1)*D), memory order."""
    global LAUNCHES
    dirs = dir_codes(dirs)
    _dirs_checks(x, weight, bias, dirs, nframes)
    batch, dim, seqlen = x.shape
    out = torch.empty((batch, len(dirs) * dim, seqlen), dtype=x.dtype, device=x.device)
    if x.numel() == 0:
        return out
    a = _lib.ConvDirsArgs()
    _fill_dirs(a, x, weight, bias, dirs, nframes, silu_activation)
    a.out, a.out_bs, a.out_ds = out.data_ptr(), out.stride(0), out.stride(1)
    with torch.cuda.device(x.device):
        stream = torch.cuda.current_stream().cuda_stream
        _lib.check(_lib.lib().vv_conv1d_dirs_fwd(ctypes.byref(a), ctypes.c_void_p(stream)), "vv_conv1d_dirs_fwd")
    LAUNCHES += 1
    return out


def causal_conv1d_dirs_bwd(x, weight, bias, dout, dx, dirs, nframes, silu_activation=True):
    """dout (B, ndirs*D, L) -> (dx (B, D, L) summed over the directions, dweight (ndirs, D, K), dbias (ndirs, D)), fp32
    parameter gradients.  ``dx`` may be a caller-provided view."""
    global LAUNCHES
    dirs = dir_codes(dirs)
    _dirs_checks(x, weight, bias, dirs, nframes)
    batch, dim, seqlen = x.shape
    _check(dout.shape == (batch, len(dirs) * dim, seqlen) and dout.dtype == x.dtype and dout.stride(2) == 1,
           "causal_conv1d_dirs_bwd: dout must be (batch, ndirs*dim, seqlen) in the dtype of x")
    if dx is None:
        dx = torch.empty_like(x, memory_format=torch.contiguous_format)
    else:
        _check(dx.shape == x.shape and dx.dtype == x.dtype and dx.stride(2) == 1,
               "causal_conv1d_dirs_bwd: dx must match x and be contiguous along seqlen")
    dweight = torch.zeros_like(weight)
    dbias = torch.zeros_like(bias) if bias is not None else None
    if x.numel() > 0:
        a = _lib.ConvDirsArgs()
        _fill_dirs(a, x, weight, bias, dirs, nframes, silu_activation)
        a.dout, a.dout_bs, a.dout_ds = dout.data_ptr(), dout.stride(0), dout.stride(1)
        a.dx, a.dx_bs, a.dx_ds = dx.data_ptr(), dx.stride(0), dx.stride(1)
        a.dweight = dweight.data_ptr()
        a.dbias = dbias.data_ptr() if dbias is not None else None
        with torch.cuda.device(x.device):
            stream = torch.cuda.current_stream().cuda_stream
            _lib.check(_lib.lib().vv_conv1d_dirs_bwd(ctypes.byref(a), ctypes.c_void_p(stream)), "vv_conv1d_dirs_bwd")
        LAUNCHES += 1
    return dx, dweight, dbias
