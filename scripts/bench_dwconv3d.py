#!/usr/bin/env python
"""dwconv3d fwd / dgrad / wgrad times at the Vivim stage shapes (batch 3, bf16) vs torch's Conv3d (cuDNN)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from vivim_b200.dwconv3d import dwconv3d_tokens  # noqa: E402

dev = torch.device("cuda", 0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 3


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


for C, hw in ((256, 64), (512, 32), (1280, 16), (2048, 8)):
    x = torch.randn(B, 5 * hw * hw, C, device=dev, dtype=torch.bfloat16, requires_grad=True)
    w = torch.randn(C, 1, 3, 3, 3, device=dev, requires_grad=True)
    b = torch.randn(C, device=dev, requires_grad=True)
    go = torch.randn_like(x)
    y = dwconv3d_tokens(x, w, b, 5, hw, hw)
    t_f = timeit(lambda: dwconv3d_tokens(x, w, b, 5, hw, hw))
    t_fb = timeit(lambda: torch.autograd.grad(dwconv3d_tokens(x, w, b, 5, hw, hw), (x, w, b), go))
    nbytes = x.numel() * 2
    conv = torch.nn.Conv3d(C, C, 3, 1, 1, groups=C).to(dev).to(torch.bfloat16)
    xv = x.detach().transpose(1, 2).reshape(B, C, 5, hw, hw).contiguous().requires_grad_()
    gv = go.transpose(1, 2).reshape(B, C, 5, hw, hw).contiguous()
    t_cf = timeit(lambda: conv(xv), 5)
    t_cfb = timeit(lambda: torch.autograd.grad(conv(xv), (xv, conv.weight, conv.bias), gv), 3)
    print(f"C={C:5d} {hw}x{hw}x5 B={B}: ours fwd {t_f:7.1f} us ({2 * nbytes / t_f / 1e3:6.0f} GB/s)  fwd+bwd {t_fb:8.1f} us"
          f" ({7 * nbytes / t_fb / 1e3:6.0f} GB/s on 7 tensors) | cuDNN fwd {t_cf:9.1f} us  fwd+bwd {t_cfb:10.1f} us")
