#!/usr/bin/env python
"""Mamba(bimamba_type="v3") -- the core of a Temporal Mamba block -- forward + backward at the four Vivim stage shapes:
the direction-fused route (one conv launch + one scan launch chain, vivim_b200/mamba_block.py) against the reference's
data flow (three mamba_inner_fn_no_out_proj calls on flipped / interleaved copies, mamba_simple.py:217-264) on the same
kernels; eager and replayed as CUDA graphs; GPU time and launches per step from the torch profiler.

    python scripts/bench_mamba_block.py [--batch B] [--iters N]     -> markdown table on stdout
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from mamba_ssm import Mamba  # noqa: E402
from vivim_b200.graphed import graph_module  # noqa: E402

STAGES = {1: (64, 5 * 64 * 64), 2: (128, 5 * 32 * 32), 3: (320, 5 * 16 * 16), 4: (512, 5 * 8 * 8)}
ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=1)
ap.add_argument("--iters", type=int, default=50)
args = ap.parse_args()
dev = torch.device("cuda", 0)


def timed(fn, iters):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


print(f"| stage (d_model, L), batch {args.batch} | route | eager ms | graphed ms | GPU ms / step | launches / step |")
print("|---|---|---|---|---|---|")
for stage, (d_model, L) in STAGES.items():
    for fused in (False, True):
        torch.manual_seed(0)
        m = Mamba(d_model=d_model, d_state=16, d_conv=4, expand=2, bimamba_type="v3", nframes=5).to(dev)
        m.fuse_directions = fused
        x = torch.randn(args.batch, L, d_model, device=dev, requires_grad=True)
        gy = torch.randn(args.batch, L, d_model, device=dev)

        def step():
            with torch.autocast("cuda", dtype=torch.bfloat16):
                y = m(x)
            y.backward(gy.to(y.dtype))

        eager = timed(step, args.iters)
        try:
            with profile(activities=[ProfilerActivity.CUDA]) as prof:
                step()
                torch.cuda.synchronize()
            ev = [e for e in prof.key_averages() if e.device_time_total > 0]
            gpu_ms, launches = sum(e.device_time_total for e in ev) / 1e3, sum(e.count for e in ev)
        except Exception:   # noqa: BLE001
            gpu_ms, launches = float("nan"), -1
        gm = graph_module(m, (torch.randn_like(x).requires_grad_(),), autocast_dtype=torch.bfloat16)

        def gstep():
            y = gm(x)
            y.backward(gy.to(y.dtype))

        graphed = timed(gstep, args.iters)
        print(f"| {stage} ({d_model}, {L}) | {'fused directions' if fused else 'three calls on copies'} | {eager:.3f} | "
              f"{graphed:.3f} | {gpu_ms:.3f} | {launches} |", flush=True)
        del gm, m
