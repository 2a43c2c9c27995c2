"""Parity of the CUDA causal conv1d (out, dx, dweight, dbias) with the oracle / golden vectors.
Grid mirrors causal-conv1d/tests/test_causal_conv1d.py:14-75 (channel slice of a wider tensor,
widths 2/3/4, odd sequence lengths), with dim reduced so the CPU oracle stays fast."""
import numpy as np
import pytest
import torch

import oracle
from conftest import golden, golden_names
from gpu_util import TOL, TOL_W, compare, dev, host, quantize

pytestmark = pytest.mark.gpu


def run_cuda(x_wide, lo, hi, w, b, dout, silu, dtype, w_dtype=torch.float32):
    from causal_conv1d import causal_conv1d_fn
    xw = dev(x_wide, dtype)
    x = xw[:, lo:hi, :].requires_grad_()           # non-contiguous batch stride, like the reference test
    wt = dev(w, w_dtype, grad=True)
    bt = dev(b, w_dtype, grad=True) if b is not None else None
    out = causal_conv1d_fn(x, wt, bt, activation="silu" if silu else None)
    out.backward(dev(dout, dtype))
    res = dict(out=host(out), dx=host(x.grad), dw=host(wt.grad))
    if bt is not None:
        res["db"] = host(bt.grad)
    return res


def run_oracle(x, w, b, dout, silu):
    out = oracle.conv1d_fwd(x, w, b, silu)
    dx, dw, db = oracle.conv1d_bwd(x, w, b, dout, silu)
    return dict(out=out, dx=dx, dw=dw, db=db if b is not None else None)


@pytest.mark.parametrize("name", golden_names("conv_"))
def test_conv_matches_reference_golden(cuda_device, name):
    g = golden(name)
    silu = "silu1" in name
    got = run_cuda(g["x"], 0, g["x"].shape[1], g["w"], g.get("bias"), g["dout"], silu, torch.float32)
    compare(got, {k: g[k] for k in ("out", "dx", "dw", "db") if k in g}, TOL[torch.float32], label=name)


@pytest.mark.parametrize("seqlen", [8, 16, 64, 151, 256, 372, 1024, 1134, 2048, 4096])
@pytest.mark.parametrize("width", [2, 3, 4])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16], ids=["fp32", "bf16", "fp16"])
def test_conv_matches_oracle(cuda_device, seqlen, width, dtype):
    rng = np.random.default_rng(seqlen * 7 + width)
    batch, dim, lo = 2, 72, 40                                     # slice 40:112 of 120 channels
    for silu, has_bias in ((True, True), (False, False)):
        xw = quantize(rng.standard_normal((batch, 120, seqlen)).astype(np.float32), dtype)
        w = rng.standard_normal((dim, width)).astype(np.float32)
        b = rng.standard_normal(dim).astype(np.float32) if has_bias else None
        dout = quantize(rng.standard_normal((batch, dim, seqlen)).astype(np.float32), dtype)
        got = run_cuda(xw, lo, lo + dim, w, b, dout, silu, dtype)
        want = run_oracle(xw[:, lo:lo + dim], w, b, dout, silu)
        compare(got, want, TOL[dtype], TOL_W[dtype], label=f"L={seqlen} K={width} {dtype} silu={silu}")


def test_conv_vivim_stage1_full_size(cuda_device):
    """x is the first half of xz (1, 256, 20480) bf16, weight (128, 4) fp32."""
    rng = np.random.default_rng(0)
    xz = quantize(rng.standard_normal((1, 256, 20480)).astype(np.float32), torch.bfloat16)
    w = (0.5 * rng.standard_normal((128, 4))).astype(np.float32)
    b = rng.standard_normal(128).astype(np.float32)
    dout = quantize(rng.standard_normal((1, 128, 20480)).astype(np.float32), torch.bfloat16)
    got = run_cuda(xz, 0, 128, w, b, dout, True, torch.bfloat16)
    compare(got, run_oracle(xz[:, :128], w, b, dout, True), TOL[torch.bfloat16], TOL_W[torch.bfloat16],
            label="stage1 conv")


def test_conv_16bit_weights_and_channel_last(cuda_device):
    from causal_conv1d import causal_conv1d_fn
    rng = np.random.default_rng(3)
    x = quantize(rng.standard_normal((2, 24, 300)).astype(np.float32), torch.bfloat16)
    w = quantize(rng.standard_normal((24, 4)).astype(np.float32), torch.bfloat16)
    b = quantize(rng.standard_normal(24).astype(np.float32), torch.bfloat16)
    dout = quantize(rng.standard_normal((2, 24, 300)).astype(np.float32), torch.bfloat16)
    got = run_cuda(x, 0, 24, w, b, dout, True, torch.bfloat16, w_dtype=torch.bfloat16)
    compare(got, run_oracle(x, w, b, dout, True), 2e-2, 2e-2, label="bf16 weights")
    # channel-last input (stride(1) == 1) is accepted, as in the reference
    xt = dev(x, torch.bfloat16).transpose(1, 2).contiguous().transpose(1, 2)
    assert xt.stride(1) == 1
    out = causal_conv1d_fn(xt, dev(w, torch.float32), dev(b, torch.float32), "silu")
    assert np.abs(host(out) - oracle.conv1d_fwd(x, w, b, True)).max() < 5e-2


def test_conv_errors_match_reference(cuda_device):
    from causal_conv1d import causal_conv1d_fn
    x = torch.randn(1, 4, 16, device="cuda")
    with pytest.raises(NotImplementedError, match="activation must be None, silu, or swish"):
        causal_conv1d_fn(x, torch.randn(4, 4, device="cuda"), None, "relu")
    with pytest.raises(RuntimeError, match="width between 2 and 4"):
        causal_conv1d_fn(x, torch.randn(4, 5, device="cuda"))


def test_conv_is_deterministic(cuda_device):
    """causal-conv1d/tests/test_causal_conv1d.py:117-173 (race detector): out and dx bitwise equal
    across repeats, dweight/dbias within 1e-4 (they are fp32 atomics)."""
    from causal_conv1d import causal_conv1d_fn
    torch.manual_seed(0)
    x = torch.randn(2, 256, 2048, device="cuda", dtype=torch.bfloat16)[:, 64:192].requires_grad_()
    w = torch.randn(128, 4, device="cuda", requires_grad=True)
    b = torch.randn(128, device="cuda", requires_grad=True)
    out0 = causal_conv1d_fn(x, w, b, "silu")
    g = torch.randn_like(out0)
    dx0, dw0, db0 = torch.autograd.grad(out0, (x, w, b), g)
    for _ in range(200):
        out = causal_conv1d_fn(x, w, b, "silu")
        dx, dw, db = torch.autograd.grad(out, (x, w, b), g)
        assert torch.equal(out, out0) and torch.equal(dx, dx0)
        assert torch.allclose(dw, dw0, atol=1e-4, rtol=1e-4) and torch.allclose(db, db0, atol=1e-4, rtol=1e-4)
