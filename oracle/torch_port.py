"""Torch restatement of the reference's CPU path.  TEST / BASELINE INFRASTRUCTURE ONLY.

These functions restate, with the same tensor-op structure (and therefore the same CPU cost
profile and fp32 rounding behaviour), the reference's own pure-torch oracles:

* ``causal_conv1d_port``   <- causal-conv1d/causal_conv1d/causal_conv1d_interface.py:49-65
* ``selective_scan_port``  <- mamba/mamba_ssm/ops/selective_scan_interface.py:86-152
* ``mamba_inner_port``     <- composition order of mamba_inner_ref (…:636-670) with the two CUDA
  calls (:646, :669) replaced by the ports, and the x_dbl slicing of
  MambaInnerFnNoOutProj.forward (:181-207)
* ``mamba_v3_port``        <- Mamba.forward, bimamba_type="v3" (mamba_ssm/modules/mamba_simple.py:204-264)

They are what ``bench.py --impl reference`` and ``cpu_baseline`` time on the host cores
(``kind: "port"``), and a second checker in the tests.  Pinned against the real reference by
``tests/test_oracle.py`` (golden vectors from ``tests/golden/make_golden.py``).
Only real-valued A and input-dependent ("variable") or constant B/C are restated.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def causal_conv1d_port(x, weight, bias=None, activation=None):
    """x (B,D,L), weight (D,K), bias (D) -> (B,D,L) in x.dtype; compute in weight.dtype."""
    if activation not in (None, "silu", "swish"):
        raise NotImplementedError("activation must be None, silu, or swish")
    in_dtype = x.dtype
    L = x.shape[-1]
    D, K = weight.shape
    y = F.conv1d(x.to(weight.dtype), weight[:, None, :], bias, padding=K - 1, groups=D)[..., :L]
    if activation is not None:
        y = F.silu(y)
    return y.to(in_dtype)


def selective_scan_port(u, delta, A, B, C, D=None, z=None, delta_bias=None,
                        delta_softplus=False, return_last_state=False):
    """Sequential-in-time scan, one einsum per step, fp32 state (real A only)."""
    in_dtype = u.dtype
    u32 = u.float()
    dt = delta.float()
    if delta_bias is not None:
        dt = dt + delta_bias.float()[:, None]
    if delta_softplus:
        dt = F.softplus(dt)
    nb, nd, L = u32.shape
    ns = A.shape[1]
    Bf, Cf = B.float(), C.float()
    var_B, var_C = Bf.dim() >= 3, Cf.dim() >= 3
    if var_B and Bf.dim() == 4:  # (B,G,N,L): broadcast each group over its channels
        Bf = Bf.repeat_interleave(nd // Bf.shape[1], dim=1)
    if var_C and Cf.dim() == 4:
        Cf = Cf.repeat_interleave(nd // Cf.shape[1], dim=1)
    decay = torch.exp(dt[..., None] * A[None, :, None, :])           # (b,d,l,n)
    if not var_B:
        drive = (dt * u32)[..., None] * Bf[None, :, None, :]
    elif Bf.dim() == 3:
        drive = (dt * u32)[..., None] * Bf.transpose(1, 2)[:, None]  # (b,1,l,n)
    else:
        drive = (dt * u32)[..., None] * Bf.transpose(2, 3)
    h = A.new_zeros((nb, nd, ns))
    ys = []
    for t in range(L):
        h = decay[:, :, t] * h + drive[:, :, t]
        if not var_C:
            yt = (h * Cf[None]).sum(-1)
        elif Cf.dim() == 3:
            yt = (h * Cf[:, None, :, t]).sum(-1)
        else:
            yt = (h * Cf[:, :, :, t]).sum(-1)
        ys.append(yt)
    y = torch.stack(ys, dim=2)
    if D is not None:
        y = y + u32 * D[:, None]
    if z is not None:
        y = y * F.silu(z)
    y = y.to(in_dtype)
    return (y, h) if return_last_state else y


def mamba_inner_port(xz, conv1d_weight, conv1d_bias, x_proj_weight, delta_proj_weight,
                     A, D=None, delta_bias=None, out_proj_weight=None, out_proj_bias=None):
    """conv+SiLU -> x_proj -> dt_proj -> selective scan (gated by z) [-> out_proj]."""
    nb, _, L = xz.shape
    rank = delta_proj_weight.shape[1]
    ns = A.shape[-1]
    x, z = xz.chunk(2, dim=1)
    xc = causal_conv1d_port(x, conv1d_weight.squeeze(1), conv1d_bias, "silu")
    x_dbl = F.linear(xc.transpose(1, 2).reshape(nb * L, -1), x_proj_weight)   # (b l, R+2N)
    dt = (delta_proj_weight @ x_dbl[:, :rank].t()).reshape(-1, nb, L).transpose(0, 1)
    Bm = x_dbl[:, rank:rank + ns].reshape(nb, L, ns).transpose(1, 2).contiguous()
    Cm = x_dbl[:, -ns:].reshape(nb, L, ns).transpose(1, 2).contiguous()
    y = selective_scan_port(xc, dt, A, Bm, Cm, D, z=z, delta_bias=delta_bias, delta_softplus=True)
    if out_proj_weight is None:
        return y
    return F.linear(y.transpose(1, 2), out_proj_weight, out_proj_bias)


def mamba_v3_port(m, hidden_states):
    """Forward of a ``Mamba(bimamba_type="v3")`` module ``m`` using only the ports above."""
    nb, L, _ = hidden_states.shape
    xz = (m.in_proj.weight @ hidden_states.reshape(nb * L, -1).t()).reshape(-1, nb, L).transpose(0, 1)
    if m.in_proj.bias is not None:
        xz = xz + m.in_proj.bias.to(xz.dtype)[:, None]

    def run(inp, sfx):
        g = lambda name: getattr(m, name + sfx)  # noqa: E731
        return mamba_inner_port(
            inp, g("conv1d").weight, g("conv1d").bias, g("x_proj").weight, g("dt_proj").weight,
            -torch.exp(getattr(m, "A" + sfx + "_log").float()), getattr(m, "D" + sfx).float(),
            g("dt_proj").bias.float())

    y_f = run(xz, "")
    y_b = run(xz.flip([-1]), "_b").flip([-1])
    nf = m.nframes
    xz_s = xz.reshape(nb, -1, nf, L // nf).transpose(2, 3).reshape(nb, -1, L)  # (t,hw)->(hw,t)
    y_s = run(xz_s, "_s").reshape(nb, -1, L // nf, nf).transpose(2, 3).reshape(nb, -1, L)
    y = (y_f + y_b + y_s).transpose(1, 2) / 3
    return F.linear(y, m.out_proj.weight, m.out_proj.bias)
