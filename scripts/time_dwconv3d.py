#!/usr/bin/env python
"""dwconv3d forward+backward time at the four Vivim stage shapes (batch 3, bf16): python scripts/time_dwconv3d.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from vivim_b200.dwconv3d import dwconv3d_tokens  # noqa: E402

dev = torch.device("cuda", 0)
out = []
for C, hw in ((256, 64), (512, 32), (1280, 16), (2048, 8)):
    x = torch.randn(3, 5 * hw * hw, C, device=dev, dtype=torch.bfloat16, requires_grad=True)
    w = torch.randn(C, 1, 3, 3, 3, device=dev, requires_grad=True)
    b = torch.randn(C, device=dev, requires_grad=True)
    go = torch.randn_like(x)
    fn = lambda: torch.autograd.grad(dwconv3d_tokens(x, w, b, 5, hw, hw), (x, w, b), go)  # noqa: E731
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        fn()
    e1.record()
    torch.cuda.synchronize()
    out.append(f"C={C} {hw}x{hw}: {e0.elapsed_time(e1) / 50 * 1e3:.1f} us")
print("  ".join(out))
