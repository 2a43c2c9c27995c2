#!/usr/bin/env python
"""Vivim clips/s on B200 (BASELINE.json configs[2] and [3]): synthetic clips (clip_length 5, 3x256x256, 3 classes),
random-init weights, bf16 autocast; one process per GPU, DDP (NCCL) gradient all-reduce in training.

    python scripts/bench_vivim.py --mode train --batch 3 [--steps 20] [--graph]
    python scripts/bench_vivim.py --mode infer --batch 4
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/bench_vivim.py --mode train

--graph replays every Temporal Mamba block (fwd + bwd) as CUDA graphs (vivim_b200.graphed.graph_module).
Prints one JSON line on rank 0: clips/s over all ranks (weak scaling: per-GPU batch fixed), max-over-ranks time.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from vivim_b200.temporal_model import RecallFocusedLoss, Vivim  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--mode", choices=("train", "infer"), default="train")
ap.add_argument("--batch", type=int, default=None)
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--warmup", type=int, default=5)
ap.add_argument("--image", type=int, default=256)
ap.add_argument("--graph", action="store_true")
ap.add_argument("--full-graph", action="store_true",
                help="training: capture forward + backward of the WHOLE network in one CUDA graph; gradients live in one "
                     "flat buffer that is all-reduced (NCCL, average) with a single call, then a fused AdamW step")
ap.add_argument("--bucket-mb", type=int, default=25)
ap.add_argument("--bf16-allreduce", action="store_true")
args = ap.parse_args()
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
batch = args.batch or (3 if args.mode == "train" else 4)
frames = 5

torch.manual_seed(0)
model = Vivim(out_chans=3).to(dev)
for name, p in model.named_parameters():        # unused by the forward, as in the reference
    if "downsample_layers.layer_norm" in name or "decoder.classifier" in name:
        p.requires_grad_(False)
clip = torch.randn(batch, frames, 3, args.image, args.image, device=dev)
target = torch.randint(0, 3, (batch * frames, args.image, args.image), device=dev)

if args.graph:
    from vivim_b200.graphed import graph_module
    size = args.image // 4
    for stage in model.encoder.stages:
        for seq in stage:
            blk = seq[0]
            sample = torch.randn(batch, blk.norm1.normalized_shape[0], frames, size, size, device=dev,
                                 requires_grad=args.mode == "train")
            if args.mode == "infer":
                blk.eval()
            seq[0] = graph_module(blk, (sample,), autocast_dtype=torch.bfloat16)
        size //= 2

if args.mode == "train" and args.full_graph:
    from vivim_b200.graphed import TrainStepGraph
    model.train()
    tsg = TrainStepGraph(model, RecallFocusedLoss().to(dev), (clip,), (target,), autocast_dtype=torch.bfloat16)
    opt = torch.optim.AdamW(tsg.params, lr=1e-4, weight_decay=1e-2, fused=True, capturable=True)

    def step():
        loss = tsg()
        if world > 1:
            dist.all_reduce(tsg.flat_grad, op=dist.ReduceOp.AVG)
        opt.step()
        return loss
elif args.mode == "train":
    model.train()
    net = model
    if world > 1:
        # gradients live directly in the all-reduce buckets (no copy), and travel as bf16 (the DDP all-reduce is the
        # only collective of this workload: north_star)
        from torch.distributed.algorithms.ddp_comm_hooks import default_hooks
        net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local], gradient_as_bucket_view=True,
                                                        bucket_cap_mb=args.bucket_mb)
        if args.bf16_allreduce:
            net.register_comm_hook(None, default_hooks.bf16_compress_hook)
    opt = torch.optim.AdamW([p for p in model.parameters() if p.requires_grad], lr=1e-4, weight_decay=1e-2)
    loss_fn = RecallFocusedLoss().to(dev)

    def step():
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            logits = net(clip)
        loss = loss_fn(logits, target)
        loss.backward()
        opt.step()
        return loss
elif args.full_graph:
    # inference: the whole forward as one CUDA graph (static input / output buffers)
    from vivim_b200.graphed import InferenceGraph
    infer = InferenceGraph(model.eval(), (clip,), autocast_dtype=torch.bfloat16)

    def step():
        return infer().float().mean()
else:
    model.eval()

    def step():
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            return model(clip).float().mean()

first = None
for _ in range(args.warmup):
    out = step()
    first = float(out) if first is None else first
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.steps):
    out = step()
e1.record()
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
t = torch.tensor([e0.elapsed_time(e1) / 1e3], device=dev)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps({"metric": f"vivim_{args.mode}_clips_per_s", "value": world * batch * args.steps / t.item(),
                      "unit": "clips/s", "n_gpus": world, "steps": args.steps, "ms_per_step": t.item() / args.steps * 1e3,
                      "scaling": "weak", "dtype": "bf16", "data": "synthetic",
                      "config": {"workload": f"Vivim multiclass {args.mode}, image {args.image}, clip_length {frames}, "
                                             f"batch {batch} per GPU, 3 classes, random init",
                                 "graphed_mamba_blocks": bool(args.graph), "whole_step_graph": bool(args.full_graph)},
                      "first_value": first, "last_value": float(out)}), flush=True)
if world > 1:
    dist.destroy_process_group()
