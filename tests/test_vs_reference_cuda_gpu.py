"""Same inputs, same GPU: this library's kernels against the UNMODIFIED reference CUDA kernels
(mamba/csrc/selective_scan, causal-conv1d/csrc) compiled for sm_100a into the git-ignored baseline/_ref/ by
baseline/build_ref.py (they travel to the GPU box as built files; nothing here reads /root/reference).

This is the strongest statement of "results identical to the reference's on the same inputs" available on a B200: the
CPU refs pin the semantics (tests/test_oracle.py), these pin the behaviour of the kernels Vivim actually runs -- fast-math
exp2 / softplus, fp32 state, bf16 / fp16 rounding at load / store only.  Skipped when baseline/_ref was not built."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT, rel_err
from gpu_util import TOL, TOL_W, dev, host, make_scan_inputs, quantize

pytestmark = pytest.mark.gpu
REF_DIR = os.path.join(ROOT, "baseline", "_ref")


@pytest.fixture(scope="module")
def ref():
    if not os.path.exists(os.path.join(REF_DIR, "selective_scan_cuda.so")):
        pytest.skip("baseline/_ref not built (python baseline/build_ref.py, needs the reference checkout)")
    sys.path.insert(0, REF_DIR)
    try:
        import causal_conv1d_cuda as rconv
        import selective_scan_cuda as rscan
    except Exception as e:   # noqa: BLE001 -- e.g. built against another torch
        pytest.skip(f"reference extensions do not import here: {e}")
    finally:
        sys.path.remove(REF_DIR)
    return rscan, rconv


@pytest.mark.parametrize("shape,dtype", [
    ((2, 4, 1024, 8, 1), torch.float32), ((2, 4, 4096, 8, 2), torch.float32),           # the reference's own test grid
    ((1, 128, 20480, 16, 1), torch.bfloat16), ((3, 128, 20480, 16, 1), torch.bfloat16),  # Vivim stage 1, B = 1 and 3
    ((1, 128, 20480, 16, 1), torch.float16), ((1, 256, 5120, 16, 1), torch.bfloat16),
    ((2, 640, 1280, 16, 1), torch.bfloat16), ((1, 1024, 320, 16, 1), torch.float32),     # stages 3 and 4
], ids=lambda v: str(v).replace("torch.", "").replace(" ", ""))
def test_scan_equals_reference_cuda_kernels(cuda_device, ref, shape, dtype):
    from vivim_b200 import selective_scan_cuda as ssc
    rscan, _ = ref
    batch, dim, seqlen, dstate, groups = shape
    d = make_scan_inputs(batch, dim, seqlen, dstate, groups, dtype, seed=seqlen + dim, vivim_init=dstate == 16)
    t = {k: dev(d[k], dtype if k in ("u", "delta", "B", "C", "z", "dout") else torch.float32)
         for k in ("u", "delta", "A", "B", "C", "D", "z", "delta_bias", "dout")}
    Bm = t["B"] if t["B"].dim() == 4 else t["B"].unsqueeze(1)
    Cm = t["C"] if t["C"].dim() == 4 else t["C"].unsqueeze(1)
    # reference: selective_scan_interface.py:35-36, 59-62
    out_r, x_r, out_z_r = rscan.fwd(t["u"], t["delta"], t["A"], Bm, Cm, t["D"], t["z"], t["delta_bias"], True)
    du_r, ddelta_r, dA_r, dB_r, dC_r, dD_r, dbias_r, dz_r = rscan.bwd(
        t["u"], t["delta"], t["A"], Bm, Cm, t["D"], t["z"], t["delta_bias"], t["dout"], x_r, out_r, None, True, False)
    _, chk, last, out_z = ssc.fwd(t["u"], t["delta"], t["A"], Bm, Cm, t["D"], t["z"], t["delta_bias"], True, want_out=False)
    du, ddelta, dA, dB, dC, dD, dbias, dz = ssc.bwd(t["u"], t["delta"], t["A"], Bm, Cm, t["D"], t["z"], t["delta_bias"],
                                                    t["dout"], chk, None, True)
    torch.cuda.synchronize()
    pairs = dict(out=(out_z, out_z_r), du=(du, du_r), ddelta=(ddelta, ddelta_r), dz=(dz, dz_r), dB=(dB, dB_r), dC=(dC, dC_r),
                 dA=(dA, dA_r), dD=(dD, dD_r), ddelta_bias=(dbias, dbias_r),
                 last_state=(last, x_r[:, :, -1, 1::2]))                 # selective_scan_interface.py:40
    errs = {k: rel_err(host(a), host(b)) for k, (a, b) in pairs.items()}
    print(f"[vs reference CUDA {shape} {dtype}] " + " ".join(f"{k}={v:.2e}" for k, v in errs.items()))
    weights = {"dA", "dD", "ddelta_bias", "last_state"}
    # both sides round to the I/O dtype independently: allow the two roundings
    bad = {k: v for k, v in errs.items() if not v <= 2 * (TOL_W[dtype] if k in weights else TOL[dtype])}
    assert not bad, bad


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
@pytest.mark.parametrize("width", [2, 4])
def test_conv_equals_reference_cuda_kernels(cuda_device, ref, dtype, width):
    from vivim_b200 import causal_conv1d_cuda as ccc
    _, rconv = ref
    g = np.random.default_rng(width)
    B_, D_, L_ = 2, 128, 20480
    xz = dev(quantize(g.standard_normal((B_, 2 * D_, L_)).astype(np.float32), dtype), dtype)
    x = xz[:, :D_]                                      # strided half, as MambaInnerFn passes it
    w = dev((0.5 * g.standard_normal((D_, width))).astype(np.float32), torch.float32)
    b = dev(g.standard_normal(D_).astype(np.float32), torch.float32)
    dout = dev(quantize(g.standard_normal((B_, D_, L_)).astype(np.float32), dtype), dtype)
    out_r = rconv.causal_conv1d_fwd(x, w, b, True)
    dx_r, dw_r, db_r = rconv.causal_conv1d_bwd(x, w, b, dout, None, True)
    out = ccc.causal_conv1d_fwd(x, w, b, True)
    dx, dw, db = ccc.causal_conv1d_bwd(x, w, b, dout, None, True)
    torch.cuda.synchronize()
    for name, a, r, tol in (("out", out, out_r, TOL[dtype]), ("dx", dx, dx_r, TOL[dtype]), ("dw", dw, dw_r, TOL_W[dtype]),
                            ("db", db, db_r, TOL_W[dtype])):
        e = rel_err(host(a), host(r))
        assert e <= 2 * tol, (name, e)
