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


def bucket_ranges(numels, limit):
    """Cut a flat buffer holding tensors of ``numels`` elements (in order) into contiguous buckets of at least ``limit``
    elements (the last one may be smaller), never splitting a tensor.  Returns (list of (lo, hi) element ranges, bucket
    index of every tensor)."""
    buckets, owner = [], []
    lo = off = 0
    for n in numels:
        owner.append(len(buckets))
        off += n
        if off - lo >= limit:
            buckets.append((lo, off))
            lo = off
    if off > lo:
        buckets.append((lo, off))
    return buckets, owner


class TrainStepGraph:
    """Forward + backward of a whole network as ONE CUDA graph, for fixed-shape training steps.

    Every gradient is a view into one flat fp32 buffer (``flat_grad``), so data-parallel training needs a single
    ``dist.all_reduce(step.flat_grad, op=dist.ReduceOp.AVG)`` per iteration, followed by a (fused, capturable)
    optimizer step -- three host-side calls per iteration instead of ~3000 launches.  Vivim (batch 3, 256x256,
    clip 5) on B200: 94 ms eager -> 49 ms per step, and 7.49x on 8 GPUs where per-rank launch overhead had limited
    DDP to 6.0x (profiles/r01_vivim_step.md).

        step = TrainStepGraph(model, loss_fn, (clip,), (target,), autocast_dtype=torch.bfloat16,
                              grad_allreduce=(lambda t: dist.all_reduce(t, op=dist.ReduceOp.AVG)) if world > 1 else None)
        opt = torch.optim.AdamW(step.params, lr=1e-4, fused=True, capturable=True)
        for clip_batch, target_batch in loader:
            loss = step(clip_batch, target_batch)          # copies into the static inputs, replays (gradients arrive averaged)
            opt.step()

    The model must be capture-safe (static shapes, no host synchronisation in forward/backward); everything in this
    repo is.  ``loss_fn(output, *targets)`` must return a scalar tensor.
    """

    def __init__(self, model: nn.Module, loss_fn, inputs, targets=(), autocast_dtype=None, warmup_iters: int = 3,
                 grad_allreduce=None, bucket_mb: float = 32.0):
        """``grad_allreduce``: optional ``fn(tensor)`` that all-reduces a slice of ``flat_grad`` in place (e.g.
        ``lambda t: dist.all_reduce(t, op=dist.ReduceOp.AVG)``).  When given, the gradient buffer is cut into buckets of
        about ``bucket_mb`` MB in parameter order and each bucket is reduced INSIDE the graph, on a side stream, as soon
        as the backward has produced its last gradient -- the collective of the late layers overlaps the backward of the
        early ones (what DDP's bucketing does, but captured: no per-step Python, no hooks at replay time).  Without it the
        caller reduces ``flat_grad`` after the replay."""
        if not torch.cuda.is_available():
            raise RuntimeError("TrainStepGraph needs a CUDA device")
        self.model, self.loss_fn, self.autocast_dtype = model, loss_fn, autocast_dtype
        self.grad_allreduce = grad_allreduce
        self.inputs = tuple(t.clone() for t in inputs)
        self.targets = tuple(t.clone() for t in targets)
        self.params = [p for p in model.parameters() if p.requires_grad]
        device = self.params[0].device
        self.flat_grad = torch.zeros(sum(p.numel() for p in self.params), device=device, dtype=torch.float32)
        off = 0
        for p in self.params:
            if p.dtype != torch.float32:
                raise RuntimeError("TrainStepGraph keeps gradients in one fp32 buffer: parameters must be float32")
            p.grad = self.flat_grad[off:off + p.numel()].view_as(p)
            off += p.numel()
        self.loss = torch.zeros((), device=device)
        self._buckets, self._in_step = [], False
        if grad_allreduce is not None:
            self.comm_stream = torch.cuda.Stream(device)
            self._buckets, owner = bucket_ranges([p.numel() for p in self.params], int(bucket_mb * 1e6 / 4))
            self._bucket_params = [owner.count(b) for b in range(len(self._buckets))]
            self._hooks = [p.register_post_accumulate_grad_hook(lambda _p, b=b: self._grad_ready(b))
                           for p, b in zip(self.params, owner)]
        side = torch.cuda.Stream(device)
        side.wait_stream(torch.cuda.current_stream(device))
        with torch.cuda.stream(side):
            for _ in range(warmup_iters):
                self._fwd_bwd()
        torch.cuda.current_stream(device).wait_stream(side)
        torch.cuda.synchronize(device)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._fwd_bwd()

    def _grad_ready(self, b):
        if not self._in_step:
            return
        self._pending[b] -= 1
        if self._pending[b] == 0:
            self._reduce_bucket(b)

    def _reduce_bucket(self, b):
        lo, hi = self._buckets[b]
        self.comm_stream.wait_stream(torch.cuda.current_stream())     # fork: the bucket's gradients are complete
        with torch.cuda.stream(self.comm_stream):
            self.grad_allreduce(self.flat_grad[lo:hi])
        self._pending[b] = -1

    def _fwd_bwd(self):
        self.flat_grad.zero_()
        if self._buckets:
            self._pending = list(self._bucket_params)
            self._in_step = True
        if self.autocast_dtype is not None:
            with torch.autocast("cuda", dtype=self.autocast_dtype, cache_enabled=False):
                out = self.model(*self.inputs)
        else:
            out = self.model(*self.inputs)
        loss = self.loss_fn(out, *self.targets)
        loss.backward()
        if self._buckets:
            self._in_step = False
            for b in range(len(self._buckets) - 1, -1, -1):            # parameters that received no gradient this step
                if self._pending[b] >= 0:
                    self._reduce_bucket(b)
            torch.cuda.current_stream().wait_stream(self.comm_stream)    # join
        self.loss.copy_(loss.detach())

    def close(self):
        """Release the captured graph (and the gradient hooks).  Call this before ``dist.destroy_process_group()`` when
        the graph contains collectives: a live CUDA graph keeps its NCCL nodes alive and the communicator teardown waits
        on them."""
        for h in getattr(self, "_hooks", []):
            h.remove()
        self._hooks = []
        torch.cuda.synchronize()
        if getattr(self, "graph", None) is not None:
            self.graph.reset()
            self.graph = None

    def __call__(self, *batch):
        """Copy ``batch`` (inputs followed by targets; omit to reuse the captured tensors) into the static buffers and
        replay.  Returns the (device-resident, overwritten on the next call) loss."""
        if batch:
            static = self.inputs + self.targets
            if len(batch) != len(static):
                raise ValueError(f"expected {len(static)} tensors, got {len(batch)}")
            for dst, src in zip(static, batch):
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.loss


class InferenceGraph:
    """The whole forward of a network as ONE CUDA graph (eval mode, no_grad, static shapes).

        infer = InferenceGraph(model.eval(), (clip,), autocast_dtype=torch.bfloat16)
        logits = infer(next_clip)        # copies into the static input, replays; the result is overwritten by the next call

    Vivim (4 clips of 5x256x256) on B200: 28.6 ms eager (launch bound) -> 16.1 ms, i.e. 140 -> 248 clips/s per GPU and
    1983 clips/s on eight (profiles/r01_vivim_step.md).
    """

    def __init__(self, model: nn.Module, inputs, autocast_dtype=None, warmup_iters: int = 3):
        if not torch.cuda.is_available():
            raise RuntimeError("InferenceGraph needs a CUDA device")
        self.model, self.autocast_dtype = model, autocast_dtype
        self.inputs = tuple(t.clone() for t in inputs)
        device = self.inputs[0].device
        side = torch.cuda.Stream(device)
        side.wait_stream(torch.cuda.current_stream(device))
        with torch.cuda.stream(side):
            for _ in range(warmup_iters):
                self._forward()
        torch.cuda.current_stream(device).wait_stream(side)
        torch.cuda.synchronize(device)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.output = self._forward()

    def _forward(self):
        with torch.no_grad():
            if self.autocast_dtype is not None:
                with torch.autocast("cuda", dtype=self.autocast_dtype, cache_enabled=False):
                    return self.model(*self.inputs)
            return self.model(*self.inputs)

    def __call__(self, *inputs):
        if inputs:
            if len(inputs) != len(self.inputs):
                raise ValueError(f"expected {len(self.inputs)} tensors, got {len(inputs)}")
            for dst, src in zip(self.inputs, inputs):
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.output
