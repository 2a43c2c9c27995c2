"""Shared helpers of the GPU parity tests: run the CUDA path through the public Python API (which
goes through the C ABI) and the oracle on the same inputs, report max relative errors."""
import numpy as np
import torch

import oracle
from conftest import rel_err

# BASELINE.json north_star: max relative error <= 1e-3 for fp32, <= 2e-2 for bf16 inputs with fp32
# state (fp16 sits between; the reference's own fp16 tolerances are 3e-3/5e-3).
TOL = {torch.float32: 1e-3, torch.float16: 5e-3, torch.bfloat16: 2e-2}
# weight gradients are fp32 sums over B*L terms of rounded 16-bit products
TOL_W = {torch.float32: 1e-3, torch.float16: 5e-3, torch.bfloat16: 2e-2}


def quantize(a, dtype):
    """numpy fp32 -> values representable in `dtype` (still numpy fp32)."""
    return torch.from_numpy(np.ascontiguousarray(a)).to(dtype).float().numpy()


def dev(a, dtype, device="cuda", grad=False):
    if a is None:
        return None
    t = torch.from_numpy(np.ascontiguousarray(a)).to(device=device, dtype=dtype)
    return t.requires_grad_() if grad else t


def host(t):
    return None if t is None else t.detach().float().cpu().numpy()


def make_scan_inputs(batch, dim, seqlen, dstate, groups=1, dtype=torch.float32, seed=0,
                     has_D=True, has_z=True, has_bias=True, vivim_init=False):
    """Input recipe of mamba/tests/ops/test_selective_scan.py:53-96 (vivim_init: the module's own
    parameter init, mamba_simple.py:99-117)."""
    g = np.random.default_rng(seed)
    f = lambda *s: g.standard_normal(s).astype(np.float32)  # noqa: E731
    if vivim_init:
        A = -np.tile(np.arange(1, dstate + 1, dtype=np.float32), (dim, 1))
        dt0 = np.exp(g.random(dim) * (np.log(0.1) - np.log(1e-3)) + np.log(1e-3)).clip(min=1e-4)
        bias = (dt0 + np.log(-np.expm1(-dt0))).astype(np.float32)
        delta = 0.5 * f(batch, dim, seqlen)
        D = np.ones(dim, np.float32)
    else:
        A = (-0.5 * g.random((dim, dstate))).astype(np.float32)
        bias = (0.5 * g.random(dim)).astype(np.float32)
        delta = (0.5 * g.random((batch, dim, seqlen))).astype(np.float32)
        D = f(dim)
    shape_bc = (batch, dstate, seqlen) if groups == 1 else (batch, groups, dstate, seqlen)
    d = dict(u=quantize(f(batch, dim, seqlen), dtype), delta=quantize(delta, dtype), A=A,
             B=quantize(f(*shape_bc), dtype), C=quantize(f(*shape_bc), dtype),
             D=D if has_D else None, z=quantize(f(batch, dim, seqlen), dtype) if has_z else None,
             delta_bias=bias if has_bias else None, dout=quantize(f(batch, dim, seqlen), dtype))
    return d


def run_scan_cuda(d, dtype, softplus=True, return_last_state=True):
    from mamba_ssm.ops.selective_scan_interface import selective_scan_fn
    t = {k: dev(d[k], dtype if k in ("u", "delta", "B", "C", "z") else torch.float32, grad=True)
         for k in ("u", "delta", "A", "B", "C", "D", "z", "delta_bias")}
    out, last = selective_scan_fn(t["u"], t["delta"], t["A"], t["B"], t["C"], t["D"], z=t["z"],
                                  delta_bias=t["delta_bias"], delta_softplus=softplus,
                                  return_last_state=True)
    out.backward(dev(d["dout"], dtype))
    res = dict(out=host(out), last_state=host(last), du=host(t["u"].grad), ddelta=host(t["delta"].grad),
               dA=host(t["A"].grad), dB=host(t["B"].grad), dC=host(t["C"].grad))
    if t["D"] is not None:
        res["dD"] = host(t["D"].grad)
    if t["z"] is not None:
        res["dz"] = host(t["z"].grad)
    if t["delta_bias"] is not None:
        res["ddelta_bias"] = host(t["delta_bias"].grad)
    return res


def run_scan_oracle(d, softplus=True):
    f = oracle.scan_fwd(d["u"], d["delta"], d["A"], d["B"], d["C"], d["D"], d["z"], d["delta_bias"], softplus)
    b = oracle.scan_bwd(d["u"], d["delta"], d["A"], d["B"], d["C"], d["D"], d["z"], d["delta_bias"],
                        d["dout"], softplus)
    res = dict(out=f["out_z"] if d["z"] is not None else f["out"], last_state=f["last_state"])
    for k, v in b.items():
        if v is not None:
            res[k] = v.reshape(d["B"].shape) if k in ("dB", "dC") else v
    return res


def compare(got, want, tol, tol_w=None, keys=None, label=""):
    """Assert max relative error per tensor; prints a table so a failing GPU run is diagnosable."""
    tol_w = tol if tol_w is None else tol_w
    weights = {"dA", "dD", "ddelta_bias", "dw", "db", "last_state"}
    rows, bad = [], []
    for k in (keys or want.keys()):
        if k not in want or want[k] is None:
            continue
        e = rel_err(got[k], want[k])
        lim = tol_w if k in weights else tol
        rows.append(f"{k}={e:.2e}")
        if not (e <= lim) or not np.isfinite(got[k]).all():
            bad.append(f"{k}: rel_err {e:.3e} > {lim:.1e}")
    print(f"[{label}] " + " ".join(rows))
    assert not bad, f"{label}: " + "; ".join(bad)
