#!/usr/bin/env python
"""What each kernel of the fwd+bwd launch chain costs INSIDE the chain (graph replay, rotating input sets, as bench.py
times the headline): the step with one pass left out (results are wrong then; the timing is what is asked for).
python scripts/chain_ablation.py"""
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
_, n_sets = bench.workload_config(1, 1)
sets = [bench.ScanSet(1, dev, seed=i, ndirs=1) for i in range(n_sets)]
bench.graphs_for(sets, lib, dev, torch)[0].replay()          # a full pass first: the workspaces hold valid data
torch.cuda.synchronize()
NAMES = {1: "aggregate", 2: "carry", 4: "main", 8: "cast"}
for label, mask in [("all passes", 15), ("all passes", 15)] + [("without " + NAMES[b], 15 & ~b) for b in (1, 2, 4, 8)]:
    for s in sets:
        s.args.pass_mask = mask
    graphs = bench.graphs_for(sets, lib, dev, torch)
    for i in range(50):
        graphs[i % n_sets].replay()
    torch.cuda.synchronize()
    t = bench.time_events(lambda i: graphs[i % n_sets].replay(), 1000, torch) / 1000
    print(f"{label:20s} (forward and backward): {t * 1e6:7.2f} us per step", flush=True)
