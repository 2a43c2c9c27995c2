"""Host logic of the direction-fused Temporal Mamba block (vivim_b200/mamba_block.py) on the CPU.

The two kernel entry points are replaced by the CPU oracle (oracle/dirs.py), so what is checked is the algebra that
the autograd Function owns -- the batched x_proj / dt_proj GEMMs on strided views of x_dbl, the mean over directions
folded into out_proj, the sum of the directions' dz folded into in_proj's backward, every parameter gradient -- against
torch autograd through the reference's own data flow (oracle/torch_port.py::mamba_v3_port, a restatement of
mamba_simple.py:204-264 with the flip / interleave copies)."""
import numpy as np
import pytest
import torch

from oracle import dirs as odirs
from oracle.torch_port import mamba_v3_port


def _np(t):
    return None if t is None else t.detach().float().cpu().numpy()


@pytest.fixture
def cpu_kernels(monkeypatch):
    from vivim_b200 import causal_conv1d_cuda as ccc
    from vivim_b200 import selective_scan_cuda as ssc

    def conv_fwd(x, w, b, dirs, nframes, silu=True):
        return torch.from_numpy(odirs.conv1d_dirs_fwd(_np(x), _np(w), _np(b), dirs, nframes, silu)).to(x.dtype)

    def conv_bwd(x, w, b, dout, dx, dirs, nframes, silu=True):
        g, dw, db = odirs.conv1d_dirs_bwd(_np(x), _np(w), _np(b), _np(dout), dirs, nframes, silu)
        dx.copy_(torch.from_numpy(g))
        return dx, torch.from_numpy(dw), torch.from_numpy(db)

    def scan_fwd(u, delta, A, B, C, D, z, bias, softplus, want_out=True, dirs=None, nframes=0):
        r = odirs.scan_dirs_fwd(_np(u), _np(delta), _np(A), _np(B), _np(C), _np(D), _np(z), _np(bias), softplus, dirs, nframes)
        return [None, torch.zeros(1), None, torch.from_numpy(r["out_z"]).to(u.dtype)]

    def scan_bwd(u, delta, A, B, C, D, z, bias, dout, chk, dz, softplus, dirs=None, nframes=0, dBC_out=None):
        r = odirs.scan_dirs_bwd(_np(u), _np(delta), _np(A), _np(B), _np(C), _np(D), _np(z), _np(bias), _np(dout),
                                softplus, dirs, nframes)
        t = {k: (None if v is None else torch.from_numpy(v)) for k, v in r.items()}
        dz.copy_(t["dz"])
        dBC_out[0].copy_(t["dB"])
        dBC_out[1].copy_(t["dC"])
        return [t["du"], t["ddelta"], t["dA"], dBC_out[0], dBC_out[1], t["dD"], t["ddelta_bias"], dz]

    monkeypatch.setattr(ccc, "causal_conv1d_dirs_fwd", conv_fwd)
    monkeypatch.setattr(ccc, "causal_conv1d_dirs_bwd", conv_bwd)
    monkeypatch.setattr(ssc, "fwd", scan_fwd)
    monkeypatch.setattr(ssc, "bwd", scan_bwd)


@pytest.mark.parametrize("kind,batch,bias", [("v3", 2, False), ("v3", 1, True), ("v2", 2, True)])
def test_fused_block_matches_reference_flow(cpu_kernels, kind, batch, bias):
    from vivim_b200.mamba_block import mamba_dirs_fn
    from vivim_b200.mamba_simple import _DIRECTIONS, _TRAVERSAL, Mamba
    torch.manual_seed(0)
    nf, hw = 5, 6
    m = Mamba(d_model=8, d_state=4, d_conv=4, expand=2, bimamba_type=kind, nframes=nf, bias=bias).double().float()
    with torch.no_grad():                      # make every parameter count
        for n, p in m.named_parameters():
            if n.startswith(("D", "A")):
                p.add_(0.1 * torch.randn_like(p))
    hidden = torch.randn(batch, nf * hw, 8, requires_grad=True)
    gout = torch.randn(batch, nf * hw, 8)

    if kind == "v3":
        ref = mamba_v3_port(m, hidden)
    else:                                      # v2: forward + flipped, summed (mamba_simple.py:265-293)
        from oracle.torch_port import mamba_inner_port
        xz = (m.in_proj.weight @ hidden.reshape(-1, 8).t()).reshape(-1, batch, nf * hw).transpose(0, 1)
        if m.in_proj.bias is not None:
            xz = xz + m.in_proj.bias[:, None]
        run = lambda inp, s: mamba_inner_port(  # noqa: E731
            inp, getattr(m, "conv1d" + s).weight, getattr(m, "conv1d" + s).bias, getattr(m, "x_proj" + s).weight,
            getattr(m, "dt_proj" + s).weight, -torch.exp(getattr(m, "A" + s + "_log")), getattr(m, "D" + s),
            getattr(m, "dt_proj" + s).bias)
        y = run(xz, "") + run(xz.flip([-1]), "_b").flip([-1])
        ref = torch.nn.functional.linear(y.transpose(1, 2), m.out_proj.weight, m.out_proj.bias)
    ref.backward(gout)
    want = {n: p.grad.clone() for n, p in m.named_parameters()}
    want_h = hidden.grad.clone()
    m.zero_grad()
    hidden.grad = None

    sfxs = _DIRECTIONS[kind]
    got = mamba_dirs_fn(hidden, m.in_proj.weight, m.in_proj.bias, m.out_proj.weight, m.out_proj.bias,
                        [m._direction_params(s) for s in sfxs], tuple(_TRAVERSAL[s] for s in sfxs), nf,
                        scale=1.0 / 3.0 if kind == "v3" else 1.0)
    got.backward(gout)
    assert torch.allclose(got, ref, rtol=1e-4, atol=1e-5), (got - ref).abs().max()
    assert torch.allclose(hidden.grad, want_h, rtol=1e-4, atol=1e-5), (hidden.grad - want_h).abs().max()
    for n, p in m.named_parameters():
        assert p.grad is not None, n
        err = (p.grad - want[n]).abs().max() / want[n].abs().max().clamp_min(1e-6)
        assert err < 2e-4, (n, float(err))


def test_traversal_is_a_permutation_and_matches_the_reference_copies():
    L, nf = 30, 5
    x = torch.arange(L)
    assert np.array_equal(odirs.traversal("rev", L, nf), x.flip([-1]).numpy())
    inter = torch.stack(x.chunk(nf, dim=-1), dim=-1).flatten(-2)      # mamba_simple.py:245-247
    assert np.array_equal(odirs.traversal("frames", L, nf), inter.numpy())
    for mode in ("fwd", "rev", "frames"):
        assert sorted(odirs.traversal(mode, L, nf)) == list(range(L))
