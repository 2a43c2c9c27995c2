#!/usr/bin/env python
"""Golden vectors for the whole Vivim network from the REAL reference model (build container only).

    python tests/golden/make_golden_vivim.py      # writes tests/golden/vivim_model.npz

Runs, unmodified, modeling/vivim.py of the reference on CPU: ``Vivim(out_chans=3).eval()`` on one seeded clip
(1, 5, 3, 64, 64), with

* ``mamba_ssm.Mamba`` = the reference's own mamba_simple.Mamba whose CUDA entry points are re-bound to the
  reference's ``*_ref`` functions (see make_golden.load_reference),
* ``timm.models.layers`` stubbed (timm is not installed; DropPath is inactive in eval mode and ``trunc_normal_``
  is torch.nn.init's, the same algorithm),
* ``SegformerForSemanticSegmentation.from_pretrained`` replaced by a config-initialised SegFormer-b3 with depths
  (1, 1, 1, 1) (no network; the shallow backbone keeps the CPU run short -- the Temporal Mamba stages are full).

The model is built under ``torch.manual_seed(0)``; the fixture stores the input, the logits and, for every entry of
the state dict, (sum, sum of squares) in float64, so that the test can check that this repo's restatement built
under the same seed has bit-identical parameters without committing 90 MB of weights.
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import make_golden as mg  # noqa: E402

SEED, SHAPE, DEPTHS = 0, (1, 5, 3, 64, 64), [1, 1, 1, 1]


def main():
    cci, ssi, ms = mg.load_reference()
    sys.modules["mamba_ssm"].Mamba = ms.Mamba

    import transformers
    from transformers import SegformerForSemanticSegmentation  # noqa: F401  (resolve the lazy import before timm is stubbed)
    from vivim_b200.temporal_model import segformer
    segformer(depths=DEPTHS)
    layers = types.ModuleType("timm.models.layers")

    class DropPath(torch.nn.Module):
        def __init__(self, p=0.0):
            super().__init__()
            self.p = p

        def forward(self, x):
            assert not self.training
            return x

    layers.DropPath, layers.to_2tuple, layers.trunc_normal_ = DropPath, (lambda v: (v, v)), torch.nn.init.trunc_normal_
    from importlib.machinery import ModuleSpec
    for name in ("timm", "timm.models"):
        sys.modules[name] = types.ModuleType(name)
    sys.modules["timm.models.layers"] = layers
    for name in ("timm", "timm.models", "timm.models.layers"):
        sys.modules[name].__spec__ = ModuleSpec(name, None)

    transformers.SegformerForSemanticSegmentation.from_pretrained = staticmethod(lambda *a, **k: segformer(depths=DEPTHS))

    spec = importlib.util.spec_from_file_location("ref_vivim", os.path.join(mg.REF, "modeling", "vivim.py"))
    ref_vivim = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_vivim)

    torch.manual_seed(SEED)
    model = ref_vivim.Vivim(out_chans=3).eval()
    g = torch.Generator().manual_seed(SEED + 1)
    clip = torch.randn(*SHAPE, generator=g)
    with torch.no_grad():
        logits = model(clip)
    sd = model.state_dict()
    keys = list(sd.keys())
    stats = np.array([[float(v.double().sum()), float((v.double() ** 2).sum())] for v in sd.values()])
    mg.save("vivim_model", clip=mg.npy(clip), logits=mg.npy(logits), keys=np.array(keys), stats=stats,
            depths=np.array(DEPTHS), seed=np.array(SEED))
    print("logits", tuple(logits.shape), "params", sum(v.numel() for v in sd.values()))

    # the reference Mamba(v3) at the REAL stage-1 size (d_model 64, L = 5*64*64 tokens): the fixture keeps the seeds and
    # every 61st output token (parameters are reproduced from the seed: the constructors draw identically)
    torch.manual_seed(SEED + 5)
    m = ms.Mamba(d_model=64, d_state=16, d_conv=4, expand=2, bimamba_type="v3", nframes=5).eval()
    gx = torch.Generator().manual_seed(SEED + 6)
    xs = torch.randn(1, 5 * 64 * 64, 64, generator=gx)
    with torch.no_grad():
        ys = m(xs)
    mg.save("mamba_stage1_full", y_sub=mg.npy(ys[:, ::61]), stride=np.array(61), seed_model=np.array(SEED + 5),
            seed_input=np.array(SEED + 6), y_absmax=np.array(float(ys.abs().max())))
    print("stage-1 Mamba forward", tuple(ys.shape), "absmax", float(ys.abs().max()))

    # the reference's own DWConv module (modeling/vivim.py:57-68), forward + autograd gradients, on a tiny token tensor
    torch.manual_seed(SEED + 2)
    dw = ref_vivim.DWConv(dim=12)
    nf, H, W = 3, 5, 6
    x = torch.randn(2, nf * H * W, 12, requires_grad=True)
    y = dw(x, nf, H, W)
    go = torch.randn_like(y)
    y.backward(go)
    mg.save("vivim_dwconv", x=mg.npy(x), weight=mg.npy(dw.dwconv.weight), bias=mg.npy(dw.dwconv.bias), dout=mg.npy(go),
            out=mg.npy(y), dx=mg.npy(x.grad), dweight=mg.npy(dw.dwconv.weight.grad), dbias=mg.npy(dw.dwconv.bias.grad),
            geom=np.array([nf, H, W]))


if __name__ == "__main__":
    main()
