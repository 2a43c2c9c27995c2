"""Depthwise 3x3x3 conv on the token layout (vv_dwconv3d_fwd / _bwd) against what the reference's DWConv module
computes (modeling/vivim.py:57-68: transpose -> nn.Conv3d(groups=C) -> transpose), evaluated by torch on the CPU in
fp64 from the same inputs."""
import numpy as np
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu

# (batch, frames, H, W, C): Vivim stage shapes (narrow), ragged W, odd C, single frame
SHAPES = [(1, 5, 16, 16, 64), (2, 5, 8, 8, 256), (1, 5, 7, 9, 40), (2, 3, 5, 6, 12), (1, 1, 4, 4, 8), (1, 2, 3, 5, 7)]


def _reference(x, w, b, geom, dout):
    bt, f, h, wd, c = geom
    xr = x.double().requires_grad_()
    wr, br = w.double().requires_grad_(), b.double().requires_grad_()
    vol = xr.transpose(1, 2).reshape(bt, c, f, h, wd)
    y = torch.nn.functional.conv3d(vol, wr, br, stride=1, padding=1, groups=c).flatten(2).transpose(1, 2)
    y.backward(dout.double())
    return y.detach(), xr.grad, wr.grad, br.grad


@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "x".join(map(str, s)))
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16], ids=["fp32", "bf16", "fp16"])
def test_dwconv3d_matches_conv3d(cuda_device, shape, dtype):
    from vivim_b200.dwconv3d import dwconv3d_tokens
    bt, f, h, wd, c = shape
    g = torch.Generator().manual_seed(sum(shape))
    x = torch.randn(bt, f * h * wd, c, generator=g).to(dtype)
    w = torch.randn(c, 1, 3, 3, 3, generator=g) * 0.3
    b = torch.randn(c, generator=g)
    dout = torch.randn(bt, f * h * wd, c, generator=g).to(dtype)
    y_ref, dx_ref, dw_ref, db_ref = _reference(x.float(), w, b, shape, dout.float())

    xd = x.cuda().requires_grad_()
    wd_, bd = w.cuda().requires_grad_(), b.cuda().requires_grad_()
    y = dwconv3d_tokens(xd, wd_, bd, f, h, wd)
    y.backward(dout.cuda())
    tol = 1e-5 if dtype == torch.float32 else 2e-2
    for name, got, want in (("out", y, y_ref), ("dx", xd.grad, dx_ref), ("dw", wd_.grad, dw_ref), ("db", bd.grad, db_ref)):
        e = rel_err(got.detach().float().cpu().numpy(), want.numpy())
        assert e < tol, f"{name}: {e:.2e}"


def test_dwconv3d_matches_reference_module_golden(cuda_device):
    """fixture produced by the reference's own DWConv module (tests/golden/make_golden_vivim.py)."""
    from conftest import golden
    from vivim_b200.dwconv3d import dwconv3d_tokens
    g = golden("vivim_dwconv")
    nf, h, w = (int(v) for v in g["geom"])
    x = torch.from_numpy(g["x"]).cuda().requires_grad_()
    wt = torch.from_numpy(g["weight"]).cuda().requires_grad_()
    b = torch.from_numpy(g["bias"]).cuda().requires_grad_()
    y = dwconv3d_tokens(x, wt, b, nf, h, w)
    y.backward(torch.from_numpy(g["dout"]).cuda())
    for name, got in (("out", y), ("dx", x.grad), ("dweight", wt.grad), ("dbias", b.grad)):
        assert rel_err(got.detach().cpu().numpy(), g[name]) < 1e-5, name


def test_dwconv3d_without_bias_and_partial_grads(cuda_device):
    from vivim_b200.dwconv3d import dwconv3d_tokens
    torch.manual_seed(0)
    x = torch.randn(1, 5 * 6 * 6, 16, device="cuda")
    w = torch.randn(16, 1, 3, 3, 3, device="cuda", requires_grad=True)
    y = dwconv3d_tokens(x, w, None, 5, 6, 6)            # x does not require grad: only the weight gradient runs
    y.sum().backward()
    vol = x.transpose(1, 2).reshape(1, 16, 5, 6, 6)
    wr = w.detach().clone().requires_grad_()
    torch.nn.functional.conv3d(vol, wr, None, padding=1, groups=16).sum().backward()
    assert rel_err(w.grad.cpu().numpy(), wr.grad.cpu().numpy()) < 1e-4


def test_dwconv3d_rejects_bad_arguments(cuda_device):
    from vivim_b200.dwconv3d import dwconv3d_tokens
    x = torch.randn(1, 10, 8, device="cuda")
    w = torch.randn(8, 1, 3, 3, 3, device="cuda")
    with pytest.raises(RuntimeError):
        dwconv3d_tokens(x, w, None, 5, 2, 2)            # 5*2*2 != 10
    with pytest.raises(RuntimeError):
        dwconv3d_tokens(x.cpu(), w.cpu(), None, 5, 2, 1)   # no CPU path
    with pytest.raises(RuntimeError):
        dwconv3d_tokens(x, w[:4], None, 5, 2, 1)
