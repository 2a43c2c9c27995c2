#!/usr/bin/env python
"""Time mamba_ssm.Mamba (bimamba_type="v3", Vivim's Temporal Mamba block core) fwd+bwd at the four Vivim stage
shapes and print the per-kernel breakdown of one step (torch profiler).

    python scripts/bench_mamba_block.py [--batch B] [--stage 1..4] [--graph]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from mamba_ssm import Mamba  # noqa: E402

STAGES = {1: (64, 5 * 64 * 64), 2: (128, 5 * 32 * 32), 3: (320, 5 * 16 * 16), 4: (512, 5 * 8 * 8)}

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=1)
ap.add_argument("--stage", type=int, default=1)
ap.add_argument("--iters", type=int, default=50)
ap.add_argument("--profile", action="store_true")
ap.add_argument("--graph", action="store_true")
args = ap.parse_args()

dev = torch.device("cuda", 0)
d_model, L = STAGES[args.stage]
torch.manual_seed(0)
m = Mamba(d_model=d_model, d_state=16, d_conv=4, expand=2, bimamba_type="v3", nframes=5).to(dev)
x = torch.randn(args.batch, L, d_model, device=dev, requires_grad=True)
gy = torch.randn(args.batch, L, d_model, device=dev)


def step():
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y = m(x)
    y.backward(gy.to(y.dtype))
    return y


for _ in range(5):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.iters):
    step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / args.iters
print(f"stage {args.stage} d_model {d_model} L {L} batch {args.batch}: {ms:.3f} ms / fwd+bwd (eager)")

if args.graph:
    # whole block fwd+bwd as two CUDA graphs: the ~60 launches of a step are replayed without Python / launch overhead
    from vivim_b200.graphed import graph_module
    x.grad = None
    y_ref = step().detach().float()
    gx_ref = x.grad.clone()
    x.grad = None
    xs = torch.randn(args.batch, L, d_model, device=dev, requires_grad=True)
    gm = graph_module(m, (xs,), autocast_dtype=torch.bfloat16)

    def gstep():
        y = gm(x)
        y.backward(gy.to(y.dtype))
        return y

    y2 = gstep().detach().float()
    torch.cuda.synchronize()
    print("graphed vs eager: max |dy|", (y2 - y_ref).abs().max().item(), " max |dgx|", (x.grad - gx_ref).abs().max().item())
    for _ in range(5):
        gstep()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(args.iters):
        gstep()
    e1.record()
    torch.cuda.synchronize()
    print(f"stage {args.stage} batch {args.batch}: {e0.elapsed_time(e1) / args.iters:.3f} ms / fwd+bwd (CUDA graphs)")

if args.profile:
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(3):
            step()
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=70))
