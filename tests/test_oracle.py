"""The oracle itself is pinned here: C restatement and torch port vs golden vectors that the
REAL Python reference produced (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

import oracle
from oracle import torch_port
from conftest import golden, golden_names, rel_err

TOL = 2e-5  # fp32 reference vs fp64-accumulating restatement


@pytest.mark.parametrize("name", golden_names("scan_"))
def test_c_scan_matches_reference(name):
    g = golden(name)
    sp = "delta_bias" in g  # every golden with a bias was generated with softplus on
    f = oracle.scan_fwd(g["u"], g["delta"], g["A"], g["B"], g["C"], g.get("D"), g.get("z"),
                        g.get("delta_bias"), delta_softplus=sp)
    key = "out_z" if "z" in g else "out"
    assert rel_err(f[key], g["out"]) < TOL
    assert rel_err(f["last_state"], g["last_state"]) < TOL
    b = oracle.scan_bwd(g["u"], g["delta"], g["A"], g["B"], g["C"], g.get("D"), g.get("z"),
                        g.get("delta_bias"), g["dout"], delta_softplus=sp)
    for k in ("du", "ddelta", "dA", "dB", "dC", "dD", "dz", "ddelta_bias"):
        if k in g:
            assert rel_err(b[k].reshape(g[k].shape), g[k]) < 5 * TOL, k


@pytest.mark.parametrize("name", golden_names("conv_"))
def test_c_conv_matches_reference(name):
    g = golden(name)
    silu = "silu1" in name
    out = oracle.conv1d_fwd(g["x"], g["w"], g.get("bias"), silu)
    assert rel_err(out, g["out"]) < TOL
    dx, dw, db = oracle.conv1d_bwd(g["x"], g["w"], g.get("bias"), g["dout"], silu)
    assert rel_err(dx, g["dx"]) < TOL
    assert rel_err(dw, g["dw"]) < TOL
    if "db" in g:
        assert rel_err(db, g["db"]) < TOL


@pytest.mark.parametrize("name", ["scan_L128_G1", "scan_L256_G2", "scan_vivim_L320", "scan_plain_L96"])
def test_torch_port_scan_matches_reference(name):
    g = {k: torch.from_numpy(v) for k, v in golden(name).items()}
    leaves = {k: g[k].clone().requires_grad_() for k in ("u", "delta", "A", "B", "C", "D", "z", "delta_bias") if k in g}
    out, last = torch_port.selective_scan_port(
        leaves["u"], leaves["delta"], leaves["A"], leaves["B"], leaves["C"], leaves.get("D"),
        leaves.get("z"), leaves.get("delta_bias"), delta_softplus="delta_bias" in g,
        return_last_state=True)
    assert rel_err(out.detach(), g["out"]) < TOL
    assert rel_err(last.detach(), g["last_state"]) < TOL
    out.backward(g["dout"])
    for k, gk in (("u", "du"), ("delta", "ddelta"), ("A", "dA"), ("B", "dB"), ("C", "dC"),
                  ("D", "dD"), ("z", "dz"), ("delta_bias", "ddelta_bias")):
        if gk in g:
            assert rel_err(leaves[k].grad, g[gk]) < 5 * TOL, gk


def test_torch_port_conv_matches_reference():
    for name in golden_names("conv_L151"):
        g = {k: torch.from_numpy(v) for k, v in golden(name).items()}
        out = torch_port.causal_conv1d_port(g["x"], g["w"], g.get("bias"),
                                            "silu" if "silu1" in name else None)
        assert rel_err(out, g["out"]) < TOL


def test_torch_port_inner_matches_reference():
    g = {k: torch.from_numpy(v) for k, v in golden("inner_noproj").items()}
    y = torch_port.mamba_inner_port(g["xz"], g["conv_w"], g["conv_b"], g["x_proj_w"], g["dt_proj_w"],
                                    g["A"], g["D"], g["dt_bias"])
    assert rel_err(y, g["out"]) < TOL


def test_oracle_edge_cases():
    # L shorter than the conv width, and a single-step scan
    x = np.arange(2 * 3 * 2, dtype=np.float32).reshape(2, 3, 2)
    w = np.ones((3, 4), dtype=np.float32)
    out = oracle.conv1d_fwd(x, w, None, False)
    assert np.allclose(out[..., 0], x[..., 0]) and np.allclose(out[..., 1], x[..., 0] + x[..., 1])
    u = np.ones((1, 2, 1), np.float32)
    f = oracle.scan_fwd(u, u, -np.ones((2, 3), np.float32), np.ones((1, 3, 1), np.float32),
                        np.ones((1, 3, 1), np.float32))
    assert np.allclose(f["out"], 3.0)  # h = delta*B*u = 1 per state, y = sum_n C h = 3


def test_oracle_threads_reported():
    assert oracle.num_threads() >= 1
