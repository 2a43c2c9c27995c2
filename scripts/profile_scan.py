#!/usr/bin/env python
"""Launch the scan kernels a few times at the stage-1 shape (for ncu / sanitizer runs).

    python scripts/profile_scan.py [--clips C] [--iters N] [--conv]
"""
import argparse
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402
from vivim_b200 import _lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--clips", type=int, default=1)
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--conv", action="store_true")
args = ap.parse_args()

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
lib = _lib.lib()
s = bench.ScanSet(args.clips, dev, seed=0)
stream = torch.cuda.current_stream().cuda_stream
for _ in range(args.iters):
    bench.launch_step(s, lib, stream)
if args.conv:
    bench.bench_conv(args.clips, dev, torch)
torch.cuda.synchronize()
print("done")
