#!/usr/bin/env python
"""BASELINE.json configs[4]: selective-scan fwd+bwd over L = 4K .. 74K tokens (clip_length 3/5/8 x image 256/384,
plus 4096 / 5120), d_inner 128, d_state 16, bf16, B = 1 -- CUDA-graph replay, inputs rotated through > L2.

    python scripts/sweep_scan.py            # prints a markdown table
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from vivim_b200 import _lib  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
lib = _lib.lib()
peak, _ = bench.measured_peak()
print("| L (tokens) | us / fwd+bwd | algorithmic GB/s | % of measured HBM peak | us per 1K tokens |\n|---|---|---|---|---|")
for L in (4096, 5120, 12288, 20480, 27648, 32768, 46080, 73728):
    bench.SEQLEN = L
    probe = bench.ScanSet(1, dev, seed=L)
    n_sets = max(2, -(-2 * bench.L2_BYTES // probe.input_bytes()) + 1)
    sets = [probe] + [bench.ScanSet(1, dev, seed=L + i) for i in range(1, n_sets)]
    side = torch.cuda.Stream(dev)
    with torch.cuda.stream(side):
        for s in sets:
            bench.launch_step(s, lib, side.cuda_stream)
    side.synchronize()
    graphs = []
    for s in sets:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            bench.launch_step(s, lib, torch.cuda.current_stream().cuda_stream)
        graphs.append(g)
    torch.cuda.synchronize()
    for i in range(20):
        graphs[i % n_sets].replay()
    torch.cuda.synchronize()
    steps = 300
    t = bench.time_events(lambda i: graphs[i % n_sets].replay(), steps, torch) / steps
    fb, bb = bench.algo_bytes(1, L)
    gbs = (fb + bb) / t / 1e9
    print(f"| {L} | {t * 1e6:.1f} | {gbs:.0f} | {100 * gbs / peak:.1f} | {t * 1e6 / L * 1024:.2f} |")
    del sets, graphs, probe
    torch.cuda.empty_cache()
