#!/usr/bin/env python
"""Generate golden vectors from the REAL Python reference (run in the build container only).

    python tests/golden/make_golden.py            # writes tests/golden/*.npz

The reference lives at /root/reference (read-only, absent on the GPU box), so the vectors are
committed.  What is imported, unmodified, from the reference:

* ``causal_conv1d_ref``  causal-conv1d/causal_conv1d/causal_conv1d_interface.py:49-65
* ``selective_scan_ref`` mamba/mamba_ssm/ops/selective_scan_interface.py:86-152
* ``mamba_inner_ref``    mamba/mamba_ssm/ops/selective_scan_interface.py:636-670
* ``bimamba_inner_ref``  mamba/mamba_ssm/ops/selective_scan_interface.py:673-709
* ``Mamba`` (v3 forward) mamba/mamba_ssm/modules/mamba_simple.py:188-264

The two compiled modules (``causal_conv1d_cuda``, ``selective_scan_cuda``) are stubbed; inside the
loaded reference modules the names ``causal_conv1d_fn`` / ``selective_scan_fn`` are re-bound to the
reference's own ``*_ref`` functions, so every line that runs is reference code on CPU.
Gradients come from torch autograd through the refs (that is how the reference's tests get them).
Input recipes follow mamba/tests/ops/test_selective_scan.py:53-96 and
causal-conv1d/tests/test_causal_conv1d.py:24-54 (seed 0, same distributions; smaller dims).
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def load_reference():
    for name in ("causal_conv1d_cuda", "selective_scan_cuda"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.path.insert(0, os.path.join(REF, "causal-conv1d"))
    import causal_conv1d.causal_conv1d_interface as cci  # the reference package, unmodified

    for pkg in ("mamba_ssm", "mamba_ssm.ops", "mamba_ssm.modules"):
        m = types.ModuleType(pkg)
        m.__path__ = []  # mark as package
        sys.modules[pkg] = m

    def by_path(modname, rel):
        spec = importlib.util.spec_from_file_location(modname, os.path.join(REF, rel))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[modname] = mod
        spec.loader.exec_module(mod)
        return mod

    ssi = by_path("mamba_ssm.ops.selective_scan_interface",
                  "mamba/mamba_ssm/ops/selective_scan_interface.py")
    # re-bind the CUDA-backed entry points to the reference's own CPU refs
    ssi.causal_conv1d_fn = cci.causal_conv1d_ref
    ssi.selective_scan_fn = ssi.selective_scan_ref

    def inner_no_out_proj(xz, conv1d_weight, conv1d_bias, x_proj_weight, delta_proj_weight,
                          A, B=None, C=None, D=None, delta_bias=None, B_proj_bias=None,
                          C_proj_bias=None, delta_softplus=True):
        # mamba_inner_ref with an identity out_proj == the no-out-proj variant, transposed back
        eye = torch.eye(A.shape[0], dtype=xz.dtype)
        y = ssi.mamba_inner_ref(xz, conv1d_weight, conv1d_bias, x_proj_weight, delta_proj_weight,
                                eye, None, A, B, C, D, delta_bias, B_proj_bias, C_proj_bias,
                                delta_softplus)
        return y.transpose(1, 2)

    ssi.mamba_inner_fn_no_out_proj = inner_no_out_proj
    ms = by_path("mamba_ssm.modules.mamba_simple", "mamba/mamba_ssm/modules/mamba_simple.py")
    return cci, ssi, ms


def npy(t):
    return None if t is None else t.detach().float().numpy()


def save(name, **arrays):
    arrays = {k: v for k, v in arrays.items() if v is not None}
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"wrote {path}  ({os.path.getsize(path) / 1024:.1f} KiB)")


def gen_scan(ssi):
    # test_selective_scan.py:53-96 recipe: batch 2, dim 4, dstate 8, everything switched on
    for seqlen, groups in ((128, 1), (256, 2), (512, 1)):
        torch.random.manual_seed(0)
        nb, nd, ns = 2, 4, 8
        A = (-0.5 * torch.rand(nd, ns)).requires_grad_()
        bshape = (nb, ns, seqlen) if groups == 1 else (nb, groups, ns, seqlen)
        B = torch.randn(*bshape, requires_grad=True)
        C = torch.randn(*bshape, requires_grad=True)
        D = torch.randn(nd, requires_grad=True)
        z = torch.randn(nb, nd, seqlen, requires_grad=True)
        delta_bias = (0.5 * torch.rand(nd)).requires_grad_()
        u = torch.randn(nb, nd, seqlen, requires_grad=True)
        delta = (0.5 * torch.rand(nb, nd, seqlen)).requires_grad_()
        out, state = ssi.selective_scan_ref(u, delta, A, B, C, D, z=z, delta_bias=delta_bias,
                                            delta_softplus=True, return_last_state=True)
        g = torch.randn_like(out)
        out.backward(g)
        save(f"scan_L{seqlen}_G{groups}", u=npy(u), delta=npy(delta), A=npy(A), B=npy(B), C=npy(C),
             D=npy(D), z=npy(z), delta_bias=npy(delta_bias), dout=npy(g), out=npy(out),
             last_state=npy(state), du=npy(u.grad), ddelta=npy(delta.grad), dA=npy(A.grad),
             dB=npy(B.grad), dC=npy(C.grad), dD=npy(D.grad), dz=npy(z.grad),
             ddelta_bias=npy(delta_bias.grad))

    # Vivim-shaped case: N=16, S4D-real A, dt-bias init of mamba_simple.py:99-117, ragged L
    torch.random.manual_seed(1)
    nb, nd, ns, seqlen = 1, 8, 16, 320
    A = (-torch.arange(1, ns + 1, dtype=torch.float32)).repeat(nd, 1).requires_grad_()
    dt0 = torch.exp(torch.rand(nd) * (np.log(0.1) - np.log(1e-3)) + np.log(1e-3)).clamp(min=1e-4)
    delta_bias = (dt0 + torch.log(-torch.expm1(-dt0))).requires_grad_()
    B = torch.randn(nb, ns, seqlen, requires_grad=True)
    C = torch.randn(nb, ns, seqlen, requires_grad=True)
    D = torch.ones(nd, requires_grad=True)
    z = torch.randn(nb, nd, seqlen, requires_grad=True)
    u = torch.randn(nb, nd, seqlen, requires_grad=True)
    delta = (0.5 * torch.randn(nb, nd, seqlen)).requires_grad_()
    out, state = ssi.selective_scan_ref(u, delta, A, B, C, D, z=z, delta_bias=delta_bias,
                                        delta_softplus=True, return_last_state=True)
    g = torch.randn_like(out)
    out.backward(g)
    save("scan_vivim_L320", u=npy(u), delta=npy(delta), A=npy(A), B=npy(B), C=npy(C), D=npy(D),
         z=npy(z), delta_bias=npy(delta_bias), dout=npy(g), out=npy(out), last_state=npy(state),
         du=npy(u.grad), ddelta=npy(delta.grad), dA=npy(A.grad), dB=npy(B.grad), dC=npy(C.grad),
         dD=npy(D.grad), dz=npy(z.grad), ddelta_bias=npy(delta_bias.grad))

    # plain variant: no D, no z, no bias, no softplus (the other branch of every optional)
    torch.random.manual_seed(2)
    nb, nd, ns, seqlen = 2, 4, 8, 96
    A = (-0.5 * torch.rand(nd, ns)).requires_grad_()
    B = torch.randn(nb, ns, seqlen, requires_grad=True)
    C = torch.randn(nb, ns, seqlen, requires_grad=True)
    u = torch.randn(nb, nd, seqlen, requires_grad=True)
    delta = (0.5 * torch.rand(nb, nd, seqlen)).requires_grad_()
    out, state = ssi.selective_scan_ref(u, delta, A, B, C, return_last_state=True)
    g = torch.randn_like(out)
    out.backward(g)
    save("scan_plain_L96", u=npy(u), delta=npy(delta), A=npy(A), B=npy(B), C=npy(C), dout=npy(g),
         out=npy(out), last_state=npy(state), du=npy(u.grad), ddelta=npy(delta.grad),
         dA=npy(A.grad), dB=npy(B.grad), dC=npy(C.grad))


def gen_conv(cci):
    # test_causal_conv1d.py:24-54 recipe with a narrower channel slice (12+8+6 instead of 4096+4128+64)
    for seqlen in (8, 151, 372):
        for width in (2, 3, 4):
            for silu in (False, True):
                for has_bias in (False, True):
                    torch.random.manual_seed(0)
                    nb, nd = 2, 8
                    x = torch.randn(nb, 12 + nd + 6, seqlen)[:, 12:12 + nd, :].requires_grad_()
                    w = torch.randn(nd, width, requires_grad=True)
                    b = torch.randn(nd, requires_grad=True) if has_bias else None
                    out = cci.causal_conv1d_ref(x, w, b, activation="silu" if silu else None)
                    g = torch.randn_like(out)
                    out.backward(g)
                    save(f"conv_L{seqlen}_K{width}_silu{int(silu)}_bias{int(has_bias)}",
                         x=npy(x), w=npy(w), bias=npy(b), dout=npy(g), out=npy(out),
                         dx=npy(x.grad), dw=npy(w.grad), db=npy(b.grad) if has_bias else None)


def gen_inner_and_module(ssi, ms):
    # mamba_inner_ref composition (the function Vivim calls has no reference test: SURVEY §4)
    torch.random.manual_seed(3)
    nb, nd, ns, rank, L, K = 2, 32, 16, 2, 160, 4
    xz = torch.randn(nb, 2 * nd, L, requires_grad=True)
    conv_w = (0.5 * torch.randn(nd, 1, K)).requires_grad_()
    conv_b = (0.1 * torch.randn(nd)).requires_grad_()
    x_proj_w = (torch.randn(rank + 2 * ns, nd) / np.sqrt(nd)).requires_grad_()
    dt_proj_w = (torch.randn(nd, rank) / np.sqrt(rank)).requires_grad_()
    A = (-torch.arange(1, ns + 1, dtype=torch.float32)).repeat(nd, 1).requires_grad_()
    D = torch.ones(nd, requires_grad=True)
    dt_bias = (torch.rand(nd) - 4.0).requires_grad_()
    y = ssi.mamba_inner_fn_no_out_proj(xz, conv_w, conv_b, x_proj_w, dt_proj_w, A, None, None, D,
                                       delta_bias=dt_bias, delta_softplus=True)
    g = torch.randn_like(y)
    y.backward(g)
    save("inner_noproj", xz=npy(xz), conv_w=npy(conv_w), conv_b=npy(conv_b), x_proj_w=npy(x_proj_w),
         dt_proj_w=npy(dt_proj_w), A=npy(A), D=npy(D), dt_bias=npy(dt_bias), dout=npy(g), out=npy(y),
         dxz=npy(xz.grad), dconv_w=npy(conv_w.grad), dconv_b=npy(conv_b.grad),
         dx_proj_w=npy(x_proj_w.grad), ddt_proj_w=npy(dt_proj_w.grad), dA=npy(A.grad),
         dD=npy(D.grad), ddt_bias=npy(dt_bias.grad))

    # Mamba module, v3 forward + backward, nframes=5, L = 5 * 8 * 8
    torch.random.manual_seed(4)
    m = ms.Mamba(d_model=16, d_state=16, d_conv=4, expand=2, bimamba_type="v3", nframes=5)
    with torch.no_grad():  # move the per-direction parameters apart so the directions differ
        for p in m.parameters():
            p.add_(0.02 * torch.randn_like(p))
    h = torch.randn(2, 5 * 8 * 8, 16, requires_grad=True)
    y = m(h)
    g = torch.randn_like(y)
    y.backward(g)
    arrays = {"param:" + k: npy(v) for k, v in m.state_dict().items()}
    arrays.update({"grad:" + k: npy(p.grad) for k, p in m.named_parameters()})
    save("mamba_v3_module", hidden=npy(h), dout=npy(g), out=npy(y), dhidden=npy(h.grad), **arrays)


def gen_bimamba(ssi):
    """bimamba_inner_ref (shared conv / projections, forward scan with A + reversed scan with A_b, out_proj).  The
    reference's own test compares bimamba_inner_fn with ITSELF (tests/ops/test_selective_scan.py:314-320), so this
    fixture is the only thing that pins the function."""
    torch.random.manual_seed(5)
    nb, nd, ns, rank, L, K, dm = 2, 24, 16, 3, 200, 4, 12
    xz = torch.randn(nb, 2 * nd, L, requires_grad=True)
    conv_w = (0.5 * torch.randn(nd, 1, K)).requires_grad_()
    conv_b = (0.1 * torch.randn(nd)).requires_grad_()
    x_proj_w = (torch.randn(rank + 2 * ns, nd) / np.sqrt(nd)).requires_grad_()
    dt_proj_w = (torch.randn(nd, rank) / np.sqrt(rank)).requires_grad_()
    out_w = (torch.randn(dm, nd) / np.sqrt(nd)).requires_grad_()
    out_b = (0.1 * torch.randn(dm)).requires_grad_()
    A = (-torch.arange(1, ns + 1, dtype=torch.float32)).repeat(nd, 1).requires_grad_()
    A_b = (-0.5 * torch.rand(nd, ns) - 0.1).requires_grad_()
    D = torch.randn(nd, requires_grad=True)
    dt_bias = (torch.rand(nd) - 4.0).requires_grad_()
    y = ssi.bimamba_inner_ref(xz, conv_w, conv_b, x_proj_w, dt_proj_w, out_w, out_b, A, A_b, None, None, D,
                              delta_bias=dt_bias, delta_softplus=True)
    g = torch.randn_like(y)
    y.backward(g)
    save("bimamba_inner", xz=npy(xz), conv_w=npy(conv_w), conv_b=npy(conv_b), x_proj_w=npy(x_proj_w),
         dt_proj_w=npy(dt_proj_w), out_w=npy(out_w), out_b=npy(out_b), A=npy(A), A_b=npy(A_b), D=npy(D),
         dt_bias=npy(dt_bias), dout=npy(g), out=npy(y), dxz=npy(xz.grad), dconv_w=npy(conv_w.grad),
         dconv_b=npy(conv_b.grad), dx_proj_w=npy(x_proj_w.grad), ddt_proj_w=npy(dt_proj_w.grad),
         dout_w=npy(out_w.grad), dout_b=npy(out_b.grad), dA=npy(A.grad), dA_b=npy(A_b.grad), dD=npy(D.grad),
         ddt_bias=npy(dt_bias.grad))


if __name__ == "__main__":
    cci, ssi, ms = load_reference()
    if "--only-bimamba" not in sys.argv:
        gen_scan(ssi)
        gen_conv(cci)
        gen_inner_and_module(ssi, ms)
    gen_bimamba(ssi)
