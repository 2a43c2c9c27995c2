"""Drop-in for ``causal_conv1d.causal_conv1d_interface`` of the reference
(causal-conv1d/causal_conv1d/causal_conv1d_interface.py): same public names, argument meaning and
error behaviour; the work is done by the sm_100a kernels in libvivim_b200.so."""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import causal_conv1d_cuda

_ACTIVATIONS = (None, "silu", "swish")


def _check_activation(activation):
    if activation not in _ACTIVATIONS:
        raise NotImplementedError("activation must be None, silu, or swish")
    return activation is not None


class CausalConv1dFn(torch.autograd.Function):
    """reference: causal_conv1d_interface.py:10-34"""

    @staticmethod
    def forward(ctx, x, weight, bias=None, activation=None):
        ctx.silu = _check_activation(activation)
        if x.stride(2) != 1 and x.stride(1) != 1:
            x = x.contiguous()
        if bias is not None:
            bias = bias.contiguous()
        ctx.save_for_backward(x, weight, bias)
        return causal_conv1d_cuda.causal_conv1d_fwd(x, weight, bias, ctx.silu)

    @staticmethod
    def backward(ctx, dout):
        x, weight, bias = ctx.saved_tensors
        if dout.stride(2) != 1 and dout.stride(1) != 1:
            dout = dout.contiguous()
        dx, dweight, dbias = causal_conv1d_cuda.causal_conv1d_bwd(x, weight, bias, dout, None, ctx.silu)
        return dx, dweight, dbias, None


def causal_conv1d_fn(x, weight, bias=None, activation=None):
    """x (batch, dim, seqlen); weight (dim, width); bias (dim,); activation None | "silu" | "swish".
    Returns (batch, dim, seqlen).  reference: causal_conv1d_interface.py:37-46"""
    return CausalConv1dFn.apply(x, weight, bias, activation)


def causal_conv1d_ref(x, weight, bias=None, activation=None):
    """Pure-torch statement of the op's semantics (reference: causal_conv1d_interface.py:49-65).
    Kept importable for API parity; never called by the product path."""
    silu = _check_activation(activation)
    seqlen = x.shape[-1]
    dim, width = weight.shape
    y = F.conv1d(x.to(weight.dtype), weight[:, None], bias, padding=width - 1, groups=dim)[..., :seqlen]
    return (F.silu(y) if silu else y).to(x.dtype)


def causal_conv1d_update(x, conv_state, weight, bias=None, activation=None):
    """Single-token rolling-window update used for autoregressive decoding
    (reference: causal_conv1d_interface.py:68-82).  Not on Vivim's path (SURVEY.md section 2,
    component 2); composed from torch ops on the tensors' own device."""
    silu = _check_activation(activation)
    conv_state.copy_(torch.roll(conv_state, shifts=-1, dims=-1))
    conv_state[:, :, -1] = x
    y = torch.sum(conv_state.to(weight.dtype) * weight, dim=-1)
    if bias is not None:
        y = y + bias
    return (F.silu(y) if silu else y).to(x.dtype)


causal_conv1d_update_ref = causal_conv1d_update
