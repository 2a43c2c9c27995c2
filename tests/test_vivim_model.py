"""The whole-network caller of the hot path (vivim_b200/temporal_model.py) against the REAL reference model
(modeling/vivim.py run on CPU by tests/golden/make_golden_vivim.py):

* CPU: built under the same torch seed, every entry of the state dict has the reference's key, and bit-identical
  contents (compared through float64 sum / sum of squares per tensor) -- i.e. a reference checkpoint loads
  unchanged and nothing in the construction order / initialisers drifted;
* GPU: the eval-mode logits of one clip, with every Temporal Mamba block running on the sm_100a kernels, match
  the reference's CPU logits (selective_scan_ref + causal_conv1d_ref inside the reference Mamba).
"""
import numpy as np
import pytest
import torch

from conftest import golden, rel_err


def _build(g):
    from vivim_b200.temporal_model import Vivim, segformer
    torch.manual_seed(int(g["seed"]))
    return Vivim(out_chans=3, backbone=segformer(depths=[int(v) for v in g["depths"]])).eval()


def test_vivim_parameters_match_reference_constructor():
    g = golden("vivim_model")
    sd = _build(g).state_dict()
    assert list(sd.keys()) == [str(k) for k in g["keys"]]
    stats = np.array([[float(v.double().sum()), float((v.double() ** 2).sum())] for v in sd.values()])
    np.testing.assert_array_equal(stats, g["stats"])


@pytest.mark.gpu
def test_vivim_logits_match_reference_model(cuda_device):
    g = golden("vivim_model")
    model = _build(g).cuda()
    with torch.no_grad():
        logits = model(torch.from_numpy(g["clip"]).cuda())
    assert logits.shape == g["logits"].shape
    err = rel_err(logits.float().cpu().numpy(), g["logits"])
    print(f"vivim logits max rel err {err:.2e}")
    assert err < 2e-3      # fp32 end to end; cuBLAS/cuDNN vs CPU GEMM/conv ordering dominates


@pytest.mark.gpu
def test_vivim_training_step_runs(cuda_device):
    """One bf16-autocast training step through the whole network: finite loss, a gradient for every Temporal Mamba
    parameter (the SegFormer stage norms and classifier are unused, as in the reference)."""
    g = golden("vivim_model")
    model = _build(g).cuda().train()
    clip = torch.from_numpy(g["clip"]).cuda()
    target = torch.randint(0, 3, (clip.shape[0] * clip.shape[1],) + clip.shape[-2:], device="cuda")
    with torch.autocast("cuda", dtype=torch.bfloat16):
        loss = torch.nn.functional.cross_entropy(model(clip).float(), target)
    loss.backward()
    assert torch.isfinite(loss)
    for name, p in model.named_parameters():
        if ".stages." in name:
            assert p.grad is not None and torch.isfinite(p.grad).all(), name


@pytest.mark.gpu
def test_whole_step_graph_matches_eager_gradients(cuda_device):
    """vivim_b200.graphed.TrainStepGraph on one Temporal Mamba block (eval-free: no dropout inside): the flat
    gradient buffer after a replay equals the eager gradients, and a second replay with a new batch tracks it."""
    from vivim_b200.graphed import TrainStepGraph
    from vivim_b200.temporal_model import TemporalMambaBlock
    torch.manual_seed(0)
    blk = TemporalMambaBlock(32).cuda().train()
    loss_fn = lambda y, t: torch.nn.functional.mse_loss(y.float(), t)   # noqa: E731
    xs = [torch.randn(2, 32, 5, 8, 8, device="cuda") for _ in range(2)]
    ts = [torch.randn(2, 32, 5, 8, 8, device="cuda") for _ in range(2)]

    def eager(x, t):
        for p in blk.parameters():
            p.grad = None
        with torch.autocast("cuda", dtype=torch.bfloat16):
            loss = loss_fn(blk(x), t)
        loss.backward()
        return float(loss), torch.cat([p.grad.flatten() for p in blk.parameters()]).clone()

    want = [eager(x, t) for x, t in zip(xs, ts)]
    step = TrainStepGraph(blk, loss_fn, (xs[0],), (ts[0],), autocast_dtype=torch.bfloat16)
    for (x, t), (loss_ref, grad_ref) in zip(zip(xs, ts), want):
        loss = float(step(x, t))
        assert abs(loss - loss_ref) <= 1e-3 * abs(loss_ref)
        assert rel_err(step.flat_grad.cpu().numpy(), grad_ref.cpu().numpy()) < 1e-2


@pytest.mark.gpu
def test_inference_graph_matches_eager(cuda_device):
    """vivim_b200.graphed.InferenceGraph: the whole-network forward replayed as one CUDA graph equals the eager
    forward bit for bit, also after the static input has been replaced."""
    from vivim_b200.graphed import InferenceGraph
    g = golden("vivim_model")
    model = _build(g).cuda().eval()
    clips = [torch.from_numpy(g["clip"]).cuda(), torch.randn(*g["clip"].shape, device="cuda")]
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        want = [model(c).float().cpu().numpy() for c in clips]
    infer = InferenceGraph(model, (clips[0],), autocast_dtype=torch.bfloat16)
    for c, w in zip(clips, want):
        assert np.array_equal(infer(c).float().cpu().numpy(), w)
