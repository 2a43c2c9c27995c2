"""Depthwise 3x3x3 convolution over (frame, y, x) on the token layout -- the ``DWConv`` inside the Mlp of every
Temporal Mamba block (reference: modeling/vivim.py:57-68, ``tokens -> transpose -> nn.Conv3d(C, C, 3, 1, 1,
groups=C) -> flatten -> transpose``), served by the sm_100a kernels of ``csrc/dwconv3d.cuh`` through the C ABI
(``vv_dwconv3d_fwd`` / ``vv_dwconv3d_bwd``).  No transposes, no cuDNN, no fallback.

    y = dwconv3d_tokens(tokens, conv.weight, conv.bias, frames, height, width)     # tokens (B, frames*H*W, C)

``weight`` is the Conv3d parameter, shape (C, 1, 3, 3, 3); ``bias`` (C) or None.  Parameters are used in fp32
(under autocast too: the accumulation is fp32 either way); the output has the dtype of ``tokens``.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib

_DTYPES = {torch.float32: _lib.VV_F32, torch.float16: _lib.VV_F16, torch.bfloat16: _lib.VV_BF16}
LAUNCHES = 0


def _args(tokens, weight27, bias, frames, height, width):
    a = _lib.DwConv3dArgs()
    a.weight = weight27.data_ptr()
    a.bias = bias.data_ptr() if bias is not None else None
    a.batch, a.frames, a.height, a.width, a.channels = tokens.shape[0], frames, height, width, tokens.shape[2]
    a.io_dtype = _DTYPES[tokens.dtype]
    return a


def _call(fn_name, a, device):
    with torch.cuda.device(device):
        stream = torch.cuda.current_stream().cuda_stream
        _lib.check(getattr(_lib.lib(), fn_name)(ctypes.byref(a), ctypes.c_void_p(stream)), fn_name)


def _check_inputs(tokens, weight, bias, frames, height, width):
    if not (tokens.is_cuda and weight.is_cuda):
        raise RuntimeError("dwconv3d_tokens: tensors must be on a CUDA device (there is no CPU path)")
    if tokens.dtype not in _DTYPES:
        raise RuntimeError("dwconv3d_tokens: tokens must be float32, float16 or bfloat16")
    if tokens.dim() != 3 or tokens.shape[1] != frames * height * width:
        raise RuntimeError("dwconv3d_tokens: tokens must be (batch, frames*height*width, channels)")
    c = tokens.shape[2]
    if tuple(weight.shape) != (c, 1, 3, 3, 3):
        raise RuntimeError("dwconv3d_tokens: weight must be (channels, 1, 3, 3, 3)")
    if bias is not None and tuple(bias.shape) != (c,):
        raise RuntimeError("dwconv3d_tokens: bias must be (channels,)")


class _DwConv3dTokens(torch.autograd.Function):
    @staticmethod
    def forward(ctx, tokens, weight, bias, frames, height, width):
        global LAUNCHES
        _check_inputs(tokens, weight, bias, frames, height, width)
        x = tokens.contiguous()
        w27 = weight.detach().float().reshape(weight.shape[0], 27).t().contiguous()     # tap-major (27, C)
        b32 = bias.detach().float().contiguous() if bias is not None else None
        out = torch.empty_like(x)
        if x.numel():
            a = _args(x, w27, b32, frames, height, width)
            a.x, a.out = x.data_ptr(), out.data_ptr()
            _call("vv_dwconv3d_fwd", a, x.device)
            LAUNCHES += 1
        ctx.save_for_backward(x, w27)
        ctx.geom = (frames, height, width)
        ctx.has_bias = bias is not None
        ctx.param_dtypes = (weight.dtype, bias.dtype if bias is not None else None)
        return out

    @staticmethod
    def backward(ctx, dout):
        global LAUNCHES
        x, w27 = ctx.saved_tensors
        frames, height, width = ctx.geom
        dout = dout.contiguous()
        need_x, need_w, need_b = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.has_bias and ctx.needs_input_grad[2]
        dx = torch.empty_like(x) if need_x else None
        # accumulated in fp32 zeros, cast back to the parameter dtype (the convention of the scan / conv1d shims)
        dw = torch.zeros(w27.shape, dtype=torch.float32, device=x.device) if (need_w or need_b) else None
        db = torch.zeros(w27.shape[1], dtype=torch.float32, device=x.device) if need_b else None
        if x.numel() and (need_x or need_w or need_b):
            a = _args(x, w27, None, frames, height, width)
            a.x, a.dout = x.data_ptr(), dout.data_ptr()
            a.dx = dx.data_ptr() if dx is not None else None
            a.dweight = dw.data_ptr() if dw is not None else None
            a.dbias = db.data_ptr() if db is not None else None
            _call("vv_dwconv3d_bwd", a, x.device)
            LAUNCHES += (1 if need_x else 0) + (1 if dw is not None else 0)
        wd, bd = ctx.param_dtypes
        return (dx, dw.t().reshape(-1, 1, 3, 3, 3).to(wd) if need_w else None, db.to(bd) if need_b else None,
                None, None, None)


def dwconv3d_tokens(tokens, weight, bias, frames, height, width):
    return _DwConv3dTokens.apply(tokens, weight, bias, frames, height, width)
