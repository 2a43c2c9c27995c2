#!/usr/bin/env python
"""BASELINE.json configs[0]: the UNMODIFIED reference Vivim (modeling/vivim.py, SegFormer-b3 config-initialised,
mamba_ssm.Mamba backed by the reference's own selective_scan_ref + causal_conv1d_ref) forward on CPU:
image 256, clip_length 5, batch 1, 3 classes, fp32, eval / no_grad.  Build container only (needs /root/reference).

    python tests/golden/time_reference_vivim.py [runs]
"""
import os
import sys
import time

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden_vivim as mgv  # noqa: E402

mgv.DEPTHS = [3, 4, 18, 3]          # the real SegFormer-b3 depths


def main(runs=3):
    import importlib.util
    import types
    from importlib.machinery import ModuleSpec
    mg = mgv.mg
    cci, ssi, ms = mg.load_reference()
    sys.modules["mamba_ssm"].Mamba = ms.Mamba
    import transformers
    from transformers import SegformerForSemanticSegmentation  # noqa: F401
    from vivim_b200.temporal_model import segformer
    layers = types.ModuleType("timm.models.layers")

    class DropPath(torch.nn.Module):
        def __init__(self, p=0.0):
            super().__init__()

        def forward(self, x):
            return x

    layers.DropPath, layers.to_2tuple, layers.trunc_normal_ = DropPath, (lambda v: (v, v)), torch.nn.init.trunc_normal_
    for name in ("timm", "timm.models"):
        sys.modules[name] = types.ModuleType(name)
    sys.modules["timm.models.layers"] = layers
    for name in ("timm", "timm.models", "timm.models.layers"):
        sys.modules[name].__spec__ = ModuleSpec(name, None)
    transformers.SegformerForSemanticSegmentation.from_pretrained = staticmethod(lambda *a, **k: segformer())
    spec = importlib.util.spec_from_file_location("ref_vivim", os.path.join(mg.REF, "modeling", "vivim.py"))
    ref_vivim = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_vivim)
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count() or 1)
    model = ref_vivim.Vivim(out_chans=3).eval()
    clip = torch.randn(1, 5, 3, 256, 256)
    times = []
    with torch.no_grad():
        model(clip)
        for _ in range(runs):
            t0 = time.perf_counter()
            model(clip)
            times.append(time.perf_counter() - t0)
    times.sort()
    print(f"reference Vivim forward on CPU ({torch.get_num_threads()} threads): median {times[len(times) // 2]:.2f} s / clip "
          f"({1 / times[len(times) // 2]:.3f} clips/s), runs {[round(t, 2) for t in times]}")


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 3)
