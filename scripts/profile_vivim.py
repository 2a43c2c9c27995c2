#!/usr/bin/env python
"""torch-profiler kernel breakdown of one Vivim training step (batch 3, image 256, bf16 autocast, recall_focused_loss,
AdamW): GPU time by kernel and by category.  python scripts/profile_vivim.py [--infer] > profiles/rNN_vivim_step.md"""
import collections
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from vivim_b200.temporal_model import RecallFocusedLoss, Vivim  # noqa: E402

infer = "--infer" in sys.argv
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = Vivim(out_chans=3).to(dev)
batch = 4 if infer else 3
clip = torch.randn(batch, 5, 3, 256, 256, device=dev)
target = torch.randint(0, 3, (batch * 5, 256, 256), device=dev)
loss_fn = RecallFocusedLoss().to(dev)
if "--torch-layernorm" not in sys.argv:     # what bench.py does: the SegFormer stages' nn.LayerNorm on the same kernels
    from vivim_b200.layernorm import use_token_layernorm
    use_token_layernorm(model)
if infer:
    model.eval()
else:
    model.train()
    for name, p in model.named_parameters():
        if "downsample_layers.layer_norm" in name or "decoder.classifier" in name:
            p.requires_grad_(False)
    opt = torch.optim.AdamW([p for p in model.parameters() if p.requires_grad], lr=1e-4, weight_decay=1e-2, fused=True)


def step():
    if infer:
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            return model(clip)
    opt.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        loss = loss_fn(model(clip), target)
    loss.backward()
    opt.step()


CATS = [("vivim_b200 scan", r"seg_|cast_bc"), ("vivim_b200 conv1d", r"conv1d_"), ("vivim_b200 dwconv3d", r"dwconv3d"),
        ("GEMM", r"gemm|nvjet|cutlass|cublas|sm100_|sm90_|xmma|gemv|splitK|wgrad2d"), ("attention", r"flash|fmha|attention|sdpa|softmax"),
        ("layer_norm", r"layer_norm|LayerNorm|layernorm"), ("batch_norm", r"batch_norm|bn_"), ("cudnn conv", r"cudnn|convolve|conv2d|implicit"),
        ("copy / cat / transpose", r"copy|Copy|CatArray|transpose|permute|gather|index|scatter"), ("reduce", r"reduce|Reduce|sum"),
        ("upsample", r"upsample|interp|bilinear"), ("dropout / rng", r"dropout|philox|bernoulli|rand"), ("optimizer", r"adam|multi_tensor|foreach"),
        ("elementwise", r"elementwise|vectorized|unrolled|gelu|silu|sigmoid|mul|add")]


def category(name):
    for cat, pat in CATS:
        if re.search(pat, name):
            return cat
    return "other"


for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
rows = [(e.key, e.count, e.device_time_total) for e in prof.key_averages() if e.device_time_total > 0]
total = sum(r[2] for r in rows)
by_cat = collections.Counter()
n_cat = collections.Counter()
for name, n, us in rows:
    by_cat[category(name)] += us
    n_cat[category(name)] += n
kind = "inference (batch 4)" if infer else "training step (batch 3)"
print(f"## Vivim {kind}, eager, one B200: {total / 1e3:.1f} ms of GPU time, {sum(r[1] for r in rows)} launches\n")
print("| category | ms | share | launches |\n|---|---|---|---|")
for cat, us in by_cat.most_common():
    print(f"| {cat} | {us / 1e3:.2f} | {100 * us / total:.1f}% | {n_cat[cat]} |")
print("\n| kernel | calls | ms | share |\n|---|---|---|---|")
for name, n, us in sorted(rows, key=lambda r: -r[2])[:40]:
    print(f"| `{name[:110]}` | {n} | {us / 1e3:.2f} | {100 * us / total:.1f}% |")
