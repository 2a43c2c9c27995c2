"""CUDA-graph capture of a Temporal-Mamba block (forward + backward).

At Vivim's shapes one ``Mamba(bimamba_type="v3")`` step is ~60 kernel launches of 5-60 us each; run eagerly it is
bound by Python and launch overhead (3.2 ms per fwd+bwd at every stage shape on B200, of which 1.1 ms is GPU
time).  Everything on this path is capture-safe -- the C-ABI launches go to the current stream, never allocate
and never synchronise, and the autograd Functions only call torch ops and those launches -- so the whole
block can be replayed as two CUDA graphs (``torch.cuda.make_graphed_callables``).

    block = MambaLayer(...).cuda()
    block = graph_module(block, (sample_input,), autocast_dtype=torch.bfloat16)
    y = block(x); y.backward(g)          # graph replays, same numerics as eager (bit-identical)

Constraints are those of CUDA graphs: static shapes and dtypes (one graphed callable per input shape), no
data-dependent control flow, call it with tensors that ``require_grad`` like the sample did.
"""
from __future__ import annotations

import torch
import torch.nn as nn


class _Autocast(nn.Module):
    def __init__(self, inner: nn.Module, dtype):
        super().__init__()
        self.inner = inner
        self.dtype = dtype

    def forward(self, *args):
        # cache_enabled=False: the autocast weight-cast cache must not outlive the capture
        with torch.autocast("cuda", dtype=self.dtype, cache_enabled=False):
            return self.inner(*args)


def graph_module(module: nn.Module, sample_args, autocast_dtype=None, num_warmup_iters: int = 3):
    """Return a callable with ``module``'s signature whose forward and backward are CUDA-graph replays.

    ``sample_args``: tuple of CUDA tensors of the shapes / dtypes / requires_grad the callable will be used with.
    ``autocast_dtype``: run the module under ``torch.autocast("cuda", dtype=...)`` inside the graph.
    """
    if not torch.cuda.is_available():
        raise RuntimeError("graph_module needs a CUDA device (the hot path has no CPU fallback)")
    if not isinstance(sample_args, (tuple, list)):
        sample_args = (sample_args,)
    target = _Autocast(module, autocast_dtype) if autocast_dtype is not None else module
    return torch.cuda.make_graphed_callables(target, tuple(sample_args), num_warmup_iters=num_warmup_iters)
