#!/usr/bin/env python
"""Launch every kernel of libvivim_b200.so a few times at the Vivim stage-1 shape (for ncu):

    ncu --set full --clock-control none --import-source on -k regex:"vv::" -s <warm-up launches> -c <n> -o prof \\
        python scripts/profile_kernels.py

Order per round: conv1d fwd/bwd (one direction), conv1d_dirs fwd/bwd (3 directions), scan fwd (agg, carry, main) + bwd
(agg, carry, main, cast) for one direction in the (B,G,N,L) layout, the same for the 3-direction block with B / C from
x_dbl rows, layernorm fwd/bwd (61 440 x 64, fp32 -> bf16), dwconv3d fwd/bwd (B = 3, C = 256, 5 x 64 x 64).
26 launches per round; the first rounds are warm-up."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from vivim_b200 import _lib, build  # noqa: E402
from vivim_b200 import causal_conv1d_cuda as ccc  # noqa: E402
from vivim_b200.dwconv3d import dwconv3d_tokens  # noqa: E402
from vivim_b200.layernorm import layer_norm_tokens  # noqa: E402

rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 3
build.build()
lib = _lib.lib()
dev = torch.device("cuda", 0)
bf = torch.bfloat16
D, L = bench.D_INNER, bench.SEQLEN
xz = torch.randn(1, 2 * D, L, device=dev, dtype=bf)
w1, b1 = torch.randn(D, 4, device=dev), torch.randn(D, device=dev)
w3, b3 = torch.randn(3, D, 4, device=dev), torch.randn(3, D, device=dev)
g1 = torch.randn(1, D, L, device=dev, dtype=bf)
g3 = torch.randn(1, 3 * D, L, device=dev, dtype=bf)
dxz = torch.empty_like(xz)
s1 = bench.ScanSet(1, dev, seed=1, ndirs=1)
s3 = bench.ScanSet(1, dev, seed=2, ndirs=3)
xln = torch.randn(3 * L, 64, device=dev, requires_grad=True)
wln, bln = torch.ones(64, device=dev, requires_grad=True), torch.zeros(64, device=dev, requires_grad=True)
gln = torch.randn(3 * L, 64, device=dev, dtype=bf)
xdw = torch.randn(3, L, 256, device=dev, dtype=bf, requires_grad=True)
wdw = (0.2 * torch.randn(256, 1, 3, 3, 3, device=dev)).requires_grad_()
bdw = torch.randn(256, device=dev, requires_grad=True)
gdw = torch.randn(3, L, 256, device=dev, dtype=bf)
stream = torch.cuda.current_stream().cuda_stream
for _ in range(rounds):
    ccc.causal_conv1d_fwd(xz[:, :D], w1, b1, True)
    ccc.causal_conv1d_bwd(xz[:, :D], w1, b1, g1, dxz[:, :D], True)
    ccc.causal_conv1d_dirs_fwd(xz[:, :D], w3, b3, bench.DIRS, bench.NFRAMES, True)
    ccc.causal_conv1d_dirs_bwd(xz[:, :D], w3, b3, g3, dxz[:, :D], bench.DIRS, bench.NFRAMES, True)
    bench.launch_step(s1, lib, stream)
    bench.launch_step(s3, lib, stream)
    y = layer_norm_tokens(xln, wln, bln, 1e-5, bf)
    y.backward(gln)
    z = dwconv3d_tokens(xdw, wdw, bdw, 5, 64, 64)
    z.backward(gdw)
    torch.cuda.synchronize()
print("ok")
