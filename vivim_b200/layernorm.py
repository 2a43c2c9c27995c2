"""LayerNorm over the channels of token tensors on the sm_100a kernels (csrc/layernorm.cuh): the glue around the Mamba
path in a Temporal Mamba block (reference: modeling/vivim.py:153-157, ``norm1`` / ``norm2`` = nn.LayerNorm).

``TokenLayerNorm`` is an ``nn.LayerNorm`` (same parameters, same state-dict keys) whose forward / backward run one
kernel each; under autocast it can hand the consumer GEMM its input directly in the autocast dtype
(``autocast_output=True``), which removes the fp32 -> bf16 cast pass over the activations.  ``use_token_layernorm(model)``
swaps the class of every eligible nn.LayerNorm of a model in place.  No fallback for CUDA tensors of an eligible shape;
CPU tensors and channel counts the kernels do not serve (> 512) go to ``F.layer_norm``.
"""
from __future__ import annotations

import ctypes

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib

_DTYPES = {torch.float32: _lib.VV_F32, torch.float16: _lib.VV_F16, torch.bfloat16: _lib.VV_BF16}
MAX_CHANNELS = 512


def _call(fn_name, a, device):
    with torch.cuda.device(device):
        stream = torch.cuda.current_stream().cuda_stream
        _lib.check(getattr(_lib.lib(), fn_name)(ctypes.byref(a), ctypes.c_void_p(stream)), fn_name)


class LayerNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, eps, out_dtype):
        C = x.shape[-1]
        x2 = x.reshape(-1, C)
        if x2.stride(-1) != 1:
            x2 = x2.contiguous()
        rows = x2.shape[0]
        out = torch.empty((rows, C), dtype=out_dtype, device=x.device)
        mean = torch.empty(rows, dtype=torch.float32, device=x.device)
        rstd = torch.empty(rows, dtype=torch.float32, device=x.device)
        w = weight.float().contiguous() if weight is not None else None
        b = bias.float().contiguous() if bias is not None else None
        if rows > 0:
            a = _lib.LayerNormArgs()
            a.x, a.out, a.mean, a.rstd = x2.data_ptr(), out.data_ptr(), mean.data_ptr(), rstd.data_ptr()
            a.weight = w.data_ptr() if w is not None else None
            a.bias = b.data_ptr() if b is not None else None
            a.rows, a.channels = rows, C
            a.x_rs, a.out_rs = x2.stride(0), out.stride(0)
            a.io_dtype, a.out_dtype, a.eps = _DTYPES[x2.dtype], _DTYPES[out_dtype], eps
            _call("vv_layernorm_fwd", a, x.device)
        ctx.save_for_backward(x2, w, mean, rstd)
        ctx.has_w, ctx.has_b, ctx.eps, ctx.shape = weight is not None, bias is not None, eps, x.shape
        ctx.w_dtype = weight.dtype if weight is not None else None
        ctx.b_dtype = bias.dtype if bias is not None else None
        return out.view(*x.shape[:-1], C)

    @staticmethod
    def backward(ctx, dout):
        x2, w, mean, rstd = ctx.saved_tensors
        rows, C = x2.shape
        g = dout.reshape(rows, C)
        if g.stride(-1) != 1:
            g = g.contiguous()
        need_dx = ctx.needs_input_grad[0]
        dx = torch.empty_like(x2) if need_dx else None
        dwb = torch.zeros(2, C, dtype=torch.float32, device=x2.device)
        if rows > 0:
            a = _lib.LayerNormArgs()
            a.x, a.mean, a.rstd, a.dout = x2.data_ptr(), mean.data_ptr(), rstd.data_ptr(), g.data_ptr()
            a.weight = w.data_ptr() if w is not None else None
            a.dx = dx.data_ptr() if dx is not None else None
            a.dweight = dwb[0].data_ptr() if ctx.has_w else None
            a.dbias = dwb[1].data_ptr() if ctx.has_b else None
            a.rows, a.channels = rows, C
            a.x_rs, a.dout_rs = x2.stride(0), g.stride(0)
            a.dx_rs = dx.stride(0) if dx is not None else 0
            a.io_dtype, a.out_dtype, a.eps = _DTYPES[x2.dtype], _DTYPES[g.dtype], ctx.eps
            _call("vv_layernorm_bwd", a, x2.device)
        return (dx.view(ctx.shape) if dx is not None else None,
                dwb[0].to(ctx.w_dtype) if ctx.has_w else None, dwb[1].to(ctx.b_dtype) if ctx.has_b else None, None, None)


def layer_norm_tokens(x, weight, bias, eps=1e-5, out_dtype=None):
    """LayerNorm over the last axis of x (.., C), C <= 512, on the sm_100a kernels.  ``out_dtype``: dtype of the result
    (default: dtype of x; only float32 inputs may change dtype)."""
    return LayerNormFn.apply(x, weight, bias, float(eps), out_dtype or x.dtype)


def eligible(x, normalized_shape):
    return (x.is_cuda and len(normalized_shape) == 1 and x.shape[-1] == normalized_shape[0]
            and normalized_shape[0] <= MAX_CHANNELS and x.dtype in _DTYPES and x.numel() > 0)


class TokenLayerNorm(nn.LayerNorm):
    """Drop-in nn.LayerNorm (identical parameters / state dict) on the sm_100a kernels."""
    autocast_output = False   # True: under autocast the result is produced in the autocast dtype

    def forward(self, x):
        if not eligible(x, self.normalized_shape):
            return F.layer_norm(x, self.normalized_shape, self.weight, self.bias, self.eps)
        # autocast runs layer_norm in float32 (it is on autocast's fp32 list): same here
        amp = torch.is_autocast_enabled("cuda")
        xin = x.float() if (amp and x.dtype != torch.float32) else x
        out_dtype = torch.get_autocast_dtype("cuda") if (amp and self.autocast_output) else xin.dtype
        with torch.autocast("cuda", enabled=False):
            return layer_norm_tokens(xin, self.weight, self.bias, self.eps, out_dtype)


def use_token_layernorm(model: nn.Module, autocast_output: bool = False) -> int:
    """Re-class every nn.LayerNorm over <= 512 channels of ``model`` as TokenLayerNorm, in place (parameters, buffers and
    state-dict keys are untouched).  Returns the number of modules switched."""
    n = 0
    for m in model.modules():
        if type(m) is nn.LayerNorm and len(m.normalized_shape) == 1 and m.normalized_shape[0] <= MAX_CHANNELS:
            m.__class__ = TokenLayerNorm
            m.autocast_output = autocast_output
            n += 1
    return n
