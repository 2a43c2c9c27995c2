#!/usr/bin/env python
"""torch-profiler kernel breakdown of one Vivim training step (batch 3, image 256, bf16 autocast)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from vivim_b200.temporal_model import Vivim  # noqa: E402

dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = Vivim(out_chans=3).to(dev).train()
clip = torch.randn(3, 5, 3, 256, 256, device=dev)
target = torch.randint(0, 3, (15, 256, 256), device=dev)
opt = torch.optim.AdamW(model.parameters(), lr=1e-4)


def step():
    opt.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        loss = torch.nn.functional.cross_entropy(model(clip).float(), target)
    loss.backward()
    opt.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=30, max_name_column_width=90))
