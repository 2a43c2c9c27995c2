#!/usr/bin/env python
"""Timeline of the end-to-end (host buffers in / out) pipeline of bench.py: device-side start / duration of every copy
and kernel of a few steps, from the torch profiler.  python scripts/trace_e2e.py > gpurun_out/e2e_trace.json"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import bench  # noqa: E402
from vivim_b200 import _lib  # noqa: E402

dev = torch.device("cuda", 0)
lib = _lib.lib()
bench.bench_e2e(1, 1, 10, 1, dev, torch, None, lib)          # warm
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    res = bench.bench_e2e(1, 1, 12, 1, dev, torch, None, lib)
    torch.cuda.synchronize()
ev = [{"name": e.name[:60], "ts": e.time_range.start, "dur": e.time_range.end - e.time_range.start,
       "stream": getattr(e, "stream", None) if hasattr(e, "stream") else None}
      for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e["ts"])
json.dump({"result": res, "events": ev}, sys.stdout)
