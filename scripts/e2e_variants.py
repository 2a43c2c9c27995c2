#!/usr/bin/env python
"""Variants of bench.py's end-to-end pipeline (host buffers in / out every step): buffer-set depth, copies split into
chunks, host submission order.  python scripts/e2e_variants.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from vivim_b200 import _lib  # noqa: E402

dev = torch.device("cuda", 0)
lib = _lib.lib()


def run(depth, chunks, early_h2d, steps=60, kernels=True):
    sets = [bench.ScanSet(1, dev, seed=500 + k, ndirs=1, flat=True) for k in range(depth)]
    n_in, n_out = sets[0].flat_in.numel(), sets[0].flat_out.numel()
    pin_in = torch.empty(n_in, dtype=torch.uint8, pin_memory=True)
    pin_in.copy_(sets[0].flat_in)
    host_out = [torch.empty(n_out, dtype=torch.uint8, pin_memory=True) for _ in range(depth)]
    s_in, s_cmp, s_out = (torch.cuda.Stream(dev) for _ in range(3))
    ev_in = [torch.cuda.Event() for _ in range(depth)]
    ev_cmp = [torch.cuda.Event() for _ in range(depth)]
    ev_out = [torch.cuda.Event() for _ in range(depth)]

    def pieces(n):
        step = -(-n // chunks // 256) * 256
        return [(o, min(o + step, n)) for o in range(0, n, step)]

    def h2d(i):
        k = i % depth
        with torch.cuda.stream(s_in):
            s_in.wait_event(ev_cmp[k])
            for lo, hi in pieces(n_in):
                sets[k].flat_in[lo:hi].copy_(pin_in[lo:hi], non_blocking=True)
            ev_in[k].record(s_in)

    def rest(i):
        k = i % depth
        with torch.cuda.stream(s_cmp):
            s_cmp.wait_event(ev_in[k])
            s_cmp.wait_event(ev_out[k])
            if kernels:
                bench.launch_step(sets[k], lib, s_cmp.cuda_stream)
            ev_cmp[k].record(s_cmp)
        with torch.cuda.stream(s_out):
            s_out.wait_event(ev_cmp[k])
            for lo, hi in pieces(n_out):
                host_out[k][lo:hi].copy_(sets[k].flat_out[lo:hi], non_blocking=True)
            ev_out[k].record(s_out)

    def loop(n):
        if early_h2d:
            # the copy-in of step i + depth - 1 is submitted before the kernels of step i
            for j in range(min(depth - 1, n)):
                h2d(j)
            for i in range(n):
                if i + depth - 1 < n:
                    h2d(i + depth - 1)
                rest(i)
        else:
            for i in range(n):
                h2d(i)
                rest(i)

    loop(2 * depth)
    torch.cuda.synchronize()
    out = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for st in (s_in, s_cmp, s_out):
            st.wait_event(e0)
        loop(steps)
        cur = torch.cuda.current_stream()
        for st in (s_in, s_cmp, s_out):
            done = torch.cuda.Event()
            done.record(st)
            cur.wait_event(done)
        e1.record()
        torch.cuda.synchronize()
        out.append(e0.elapsed_time(e1) / steps)
    return sorted(out)[1]


for depth, chunks, early in [(2, 1, False), (2, 1, True), (3, 1, False), (3, 1, True), (2, 4, False), (3, 4, True), (4, 1, True),
                             (2, 1, False)]:
    print(f"depth {depth} chunks {chunks} early_h2d {early}: {run(depth, chunks, early):.4f} ms/step   "
          f"(copies only: {run(depth, chunks, early, kernels=False):.4f})", flush=True)
