#!/usr/bin/env python
"""Print the interesting keys of bench.py JSON lines:  python scripts/show_bench.py gpurun_out/*.log"""
import json
import sys

for path in sys.argv[1:]:
    for line in open(path):
        if not line.startswith("{"):
            continue
        d = json.loads(line)
        if "metric" not in d:
            print(f"== {path}: n_gpus {d.get('n_gpus')}  " + "  ".join(f"{k} {d[k]:.2f}" for k in d if k.endswith(("_per_s", "_per_step"))))
            continue
        print(f"== {path}: {d.get('config', {}).get('directions')} value {d.get('value', 0):.1f} GB/s  step {d.get('ms_per_step', 0) * 1e3:.1f} us"
              f"  per-scan {d.get('us_per_direction_scan', 0):.1f} us  e2e {d.get('e2e', {}).get('value', 0):.1f}")
        if "kernel_us" in d:
            print("   kernels: " + "  ".join(f"{k} {v:.1f}" for k, v in d["kernel_us"].items()))
        if "conv1d" in d:
            print("   conv:    " + "  ".join(f"{k} {v:.1f}" for k, v in d["conv1d"].items() if k.endswith("_us")))
        for k in ("single_direction", "ref_cuda_us", "vivim_train_clips_per_s", "vivim_train_ms_per_step",
                  "vivim_infer_clips_per_s", "vivim_infer_ms_per_step"):
            if k in d:
                print(f"   {k}: {d[k]}")
