"""Direction-fused kernels: causal conv1d of all directions in one launch, the selective scan
with per-direction-block traversal order, position-major B / C read straight from x_dbl, dB / dC written into dx_dbl,
and the fused Mamba(v3) block built on them.  Everything goes Python -> ctypes -> libvivim_b200.so.

Checked against (1) the CPU oracle restating each direction the reference's way -- gather into traversal order, run the
plain op, scatter back (oracle/dirs.py) -- and (2) this library's own single-direction kernels on materialised flipped /
frame-interleaved copies, where the results must be BIT-IDENTICAL (same arithmetic, different addressing)."""
import numpy as np
import pytest
import torch

from conftest import rel_err
from gpu_util import TOL, TOL_W, compare, dev, host, quantize
from oracle import dirs as odirs

pytestmark = pytest.mark.gpu

V3 = ("fwd", "rev", "frames")


def _perm(mode, L, nf, device="cuda"):
    return torch.from_numpy(odirs.traversal(mode, L, nf)).to(device)


# ------------------------------------------------------------------------------------------------ conv
CONV_CASES = [
    # batch, dim, nframes, pixels/frame, width, dirs
    (2, 6, 5, 64, 4, V3),            # vector path (hw % 8 == 0)
    (1, 5, 5, 13, 4, V3),            # ragged frame length: element-wise path, pixel tile larger than the frame
    (2, 4, 3, 200, 3, V3),           # width 3, several pixel tiles per frame (pt = 384 > 200: one tile) ...
    (1, 3, 5, 600, 4, V3),           # ... and more than one tile (pt = 256)
    (1, 8, 1, 2048, 4, ("fwd", "rev")),   # v2: no frames, tiles of 1024
    (2, 4, 2, 24, 2, V3),            # width 2, taps wrap over two pixels (nf = 2 < K - 1)
    (1, 4, 1, 40, 4, ("frames",)),   # degenerate: one frame, frames == fwd
    (1, 16, 5, 4096, 4, V3),         # Vivim stage 1 (a slice of the channels)
    (1, 4, 8, 96, 4, V3),            # clip_length 8 (BASELINE configs[4])
    (2, 3, 16, 40, 4, V3),           # the most frames the kernels stage (pixel tiles shrink to fit shared memory)
]


@pytest.mark.parametrize("case", CONV_CASES, ids=lambda c: "-".join(map(str, c[:5])) + "-" + "".join(d[0] for d in c[5]))
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
@pytest.mark.parametrize("silu,has_bias", [(True, True), (False, False)])
def test_conv_dirs_matches_oracle(cuda_device, case, dtype, silu, has_bias):
    from vivim_b200 import causal_conv1d_cuda as ccc
    batch, dim, nf, hw, K, dirs = case
    L = nf * hw
    g = np.random.default_rng(L + dim)
    x = quantize(g.standard_normal((batch, dim, L)).astype(np.float32), dtype)
    w = g.standard_normal((len(dirs), dim, K)).astype(np.float32)
    b = g.standard_normal((len(dirs), dim)).astype(np.float32) if has_bias else None
    dout = quantize(g.standard_normal((batch, len(dirs) * dim, L)).astype(np.float32), dtype)
    # x as the first half of an xz tensor laid out like in_proj's GEMM output: (B, 2D, L) view of (2D, B, L)
    xz = torch.zeros(2 * dim, batch, L, device="cuda", dtype=dtype).transpose(0, 1)
    xz[:, :dim] = dev(x, dtype)
    xv = xz[:, :dim]
    wt, bt = dev(w, torch.float32), dev(b, torch.float32)
    out = ccc.causal_conv1d_dirs_fwd(xv, wt, bt, dirs, nf, silu)
    dxz = torch.full((batch, 4 * dim, L), 7.0, device="cuda", dtype=dtype)
    dx, dw, db = ccc.causal_conv1d_dirs_bwd(xv, wt, bt, dev(dout, dtype), dxz[:, :dim], dirs, nf, silu)
    want_out = odirs.conv1d_dirs_fwd(x, w, b, dirs, nf, silu)
    want_dx, want_dw, want_db = odirs.conv1d_dirs_bwd(x, w, b, dout, dirs, nf, silu)
    got = dict(out=host(out), dx=host(dx), dw=host(dw))
    want = dict(out=want_out, dx=want_dx, dw=want_dw)
    if has_bias:
        got["db"], want["db"] = host(db), want_db
    compare(got, want, TOL[dtype], TOL_W[dtype], label=f"conv_dirs {case} {dtype}")
    assert torch.all(dxz[:, dim:] == 7.0)            # only the dx rows of the caller's buffer are written


def test_conv_dirs_bit_identical_to_single_direction_kernels(cuda_device):
    """Same taps, same summation order: direction k of the fused launch == conv1d_fwd_kernel on the gathered copy."""
    from vivim_b200 import causal_conv1d_cuda as ccc
    torch.manual_seed(0)
    B_, D_, nf, hw = 2, 24, 5, 256
    L = nf * hw
    x = torch.randn(B_, D_, L, device="cuda", dtype=torch.bfloat16)
    w = torch.randn(3, D_, 4, device="cuda")
    b = torch.randn(3, D_, device="cuda")
    out = ccc.causal_conv1d_dirs_fwd(x, w, b, V3, nf, True)
    for k, mode in enumerate(V3):
        p = _perm(mode, L, nf)
        single = ccc.causal_conv1d_fwd(x[:, :, p].contiguous(), w[k].contiguous(), b[k].contiguous(), True)
        assert torch.equal(out[:, k * D_:(k + 1) * D_][:, :, p], single), mode


# ------------------------------------------------------------------------------------------------ scan
def _scan_inputs(batch, D, L, N, nd, dtype, seed, R=4):
    """Channel-concatenated inputs of nd direction blocks; B / C as column blocks of an x_dbl-like (B, nd, L, R+2N)."""
    g = np.random.default_rng(seed)
    f = lambda *s: g.standard_normal(s).astype(np.float32)  # noqa: E731
    dim = nd * D
    d = dict(u=quantize(f(batch, dim, L), dtype), delta=quantize((0.5 * g.random((batch, dim, L))).astype(np.float32), dtype),
             A=(-0.5 * g.random((dim, N))).astype(np.float32), D=f(dim), delta_bias=(0.5 * g.random(dim)).astype(np.float32),
             z=quantize(f(batch, D, L), dtype), dout=quantize(f(batch, D, L), dtype),
             x_dbl=quantize(f(batch, nd, L, R + 2 * N), dtype))
    d["B"] = np.ascontiguousarray(d["x_dbl"][..., R:R + N].transpose(0, 1, 3, 2))
    d["C"] = np.ascontiguousarray(d["x_dbl"][..., R + N:].transpose(0, 1, 3, 2))
    return d


def _run_scan_dirs(d, dtype, dirs, nf, N, R=4, position_major=True):
    from vivim_b200 import selective_scan_cuda as ssc
    t = {k: dev(d[k], dtype if k in ("u", "delta", "z", "dout", "x_dbl", "B", "C") else torch.float32)
         for k in ("u", "delta", "A", "D", "delta_bias", "z", "dout", "x_dbl", "B", "C")}
    if position_major:
        Bv = t["x_dbl"][..., R:R + N].permute(0, 1, 3, 2)
        Cv = t["x_dbl"][..., R + N:].permute(0, 1, 3, 2)
        dx_dbl = torch.full_like(t["x_dbl"], 7.0)
        dBC = (dx_dbl[..., R:R + N].permute(0, 1, 3, 2), dx_dbl[..., R + N:].permute(0, 1, 3, 2))
    else:
        Bv, Cv, dBC, dx_dbl = t["B"], t["C"], None, None
    _, chk, _, out_z = ssc.fwd(t["u"], t["delta"], t["A"], Bv, Cv, t["D"], t["z"], t["delta_bias"], True,
                               want_out=False, dirs=dirs, nframes=nf)
    du, ddelta, dA, dB, dC, dD, dbias, dz = ssc.bwd(t["u"], t["delta"], t["A"], Bv, Cv, t["D"], t["z"], t["delta_bias"],
                                                    t["dout"], chk, None, True, dirs=dirs, nframes=nf, dBC_out=dBC)
    if dx_dbl is not None:
        assert torch.all(dx_dbl[..., :R] == 7.0)     # the dt columns are not the scan's to write
    return dict(out=host(out_z), du=host(du), ddelta=host(ddelta), dA=host(dA), dB=host(dB), dC=host(dC), dD=host(dD),
                ddelta_bias=host(dbias), dz=host(dz))


SCAN_CASES = [
    # batch, D per direction, nframes, pixels/frame, dstate, dirs
    (2, 8, 5, 64, 16, V3),
    (1, 5, 5, 13, 8, V3),            # ragged: L = 65, element-wise I/O, partly empty warps
    (1, 40, 3, 200, 16, V3),         # two channel blocks per direction in the forward kernels, three in the backward
    (2, 16, 1, 512, 16, ("fwd", "rev")),
    (1, 4, 5, 40, 3, ("frames",)),
    (1, 16, 5, 1024, 16, V3),        # Vivim stage 2 (a slice of the channels)
    (1, 20, 8, 576, 16, V3),         # clip_length 8 (BASELINE configs[4]): the 8-frame instantiation of the run-wise gathers
    (2, 8, 7, 64, 16, V3),           # 7 frames: the same instantiation with its last frame slot empty
    (1, 8, 6, 104, 8, ("frames", "frames")),
    (1, 8, 12, 64, 16, V3),          # more frames than the run-wise gathers serve: element-wise route
]


@pytest.mark.parametrize("case", SCAN_CASES, ids=lambda c: "-".join(map(str, c[:5])) + "-" + "".join(d[0] for d in c[5]))
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
def test_scan_dirs_matches_oracle(cuda_device, case, dtype):
    batch, D, nf, hw, N, dirs = case
    L = nf * hw
    d = _scan_inputs(batch, D, L, N, len(dirs), dtype, seed=L + D)
    got = _run_scan_dirs(d, dtype, dirs, nf, N)
    f = odirs.scan_dirs_fwd(d["u"], d["delta"], d["A"], d["B"], d["C"], d["D"], d["z"], d["delta_bias"], True, dirs, nf)
    want = odirs.scan_dirs_bwd(d["u"], d["delta"], d["A"], d["B"], d["C"], d["D"], d["z"], d["delta_bias"], d["dout"],
                               True, dirs, nf)
    want["out"] = f["out_z"]
    compare(got, want, TOL[dtype], TOL_W[dtype], label=f"scan_dirs {case} {dtype}")


def test_scan_dirs_bit_identical_to_separate_scans(cuda_device):
    """One launch over three direction blocks == three launches of the plain op on gathered copies: the same
    arithmetic in the same order, so every non-atomic output is bit-identical."""
    from vivim_b200 import selective_scan_cuda as ssc
    dtype, nf, hw, D, N = torch.bfloat16, 5, 256, 32, 16
    L = nf * hw
    d = _scan_inputs(2, D, L, N, 3, dtype, seed=7)
    fused = _run_scan_dirs(d, dtype, V3, nf, N)
    for k, mode in enumerate(V3):
        p = _perm(mode, L, nf)
        ch = slice(k * D, (k + 1) * D)
        t = lambda name, sl=slice(None): dev(d[name], dtype if name in ("u", "delta", "z", "dout", "B", "C") else torch.float32)[sl]  # noqa: E731
        u, delta = t("u")[:, ch][:, :, p].contiguous(), t("delta")[:, ch][:, :, p].contiguous()
        z, dout = t("z")[:, :, p].contiguous(), t("dout")[:, :, p].contiguous()
        Bm, Cm = t("B")[:, k:k + 1][..., p].contiguous(), t("C")[:, k:k + 1][..., p].contiguous()
        A, Dv, bias = t("A")[ch].contiguous(), t("D")[ch].contiguous(), t("delta_bias")[ch].contiguous()
        _, chk, _, out_z = ssc.fwd(u, delta, A, Bm, Cm, Dv, z, bias, True, want_out=False)
        du, ddelta, dA, dB, dC, dD, dbias, dz = ssc.bwd(u, delta, A, Bm, Cm, Dv, z, bias, dout, chk, None, True)
        inv = torch.empty_like(p)
        inv[p] = torch.arange(L, device="cuda")
        for name, got in (("out", out_z), ("du", du), ("ddelta", ddelta), ("dz", dz)):
            assert np.array_equal(fused[name][:, ch], host(got[:, :, inv])), (mode, name)
        assert rel_err(fused["dB"][:, k:k + 1], host(dB[..., inv])) < 1e-2    # fp32 atomics, then bf16 rounding


def test_scan_position_major_equals_state_major(cuda_device):
    """B / C read from x_dbl rows == B / C as (B,G,N,L) tensors: bit-identical (forward direction, one group)."""
    dtype, L, D, N = torch.bfloat16, 1280, 48, 16
    d = _scan_inputs(2, D, L, N, 1, dtype, seed=3)
    a = _run_scan_dirs(d, dtype, ("fwd",), 1, N, position_major=True)
    b = _run_scan_dirs(d, dtype, ("fwd",), 1, N, position_major=False)
    for k in ("out", "du", "ddelta", "dz"):
        assert np.array_equal(a[k], b[k]), k
    for k in ("dB", "dC", "dA"):
        assert rel_err(a[k], b[k]) < 1e-2, k


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
def test_scan_dirs_vivim_stage1_block(cuda_device, dtype):
    """One Temporal Mamba block of Vivim's stage 1 (BASELINE configs[1]): 3 directions x 128 channels x 20480 tokens,
    one launch, against the O(L) oracle."""
    nf, hw, D, N = 5, 4096, 128, 16
    d = _scan_inputs(1, D, nf * hw, N, 3, dtype, seed=1)
    got = _run_scan_dirs(d, dtype, V3, nf, N)
    f = odirs.scan_dirs_fwd(d["u"], d["delta"], d["A"], d["B"], d["C"], d["D"], d["z"], d["delta_bias"], True, V3, nf)
    want = odirs.scan_dirs_bwd(d["u"], d["delta"], d["A"], d["B"], d["C"], d["D"], d["z"], d["delta_bias"], d["dout"],
                               True, V3, nf)
    want["out"] = f["out_z"]
    compare(got, want, TOL[dtype], TOL_W[dtype], label=f"stage-1 block {dtype}")


# ------------------------------------------------------------------------------------------------ the fused block
def _module_run(m, x, gy, autocast):
    x = x.clone().requires_grad_()
    for p in m.parameters():
        p.grad = None
    if autocast:
        with torch.autocast("cuda", dtype=torch.bfloat16):
            y = m(x)
    else:
        y = m(x)
    y.float().backward(gy)
    torch.cuda.synchronize()
    return host(y), host(x.grad), {n: host(p.grad) for n, p in m.named_parameters()}


@pytest.mark.parametrize("kind,d_model,nf,hw,batch,autocast", [
    ("v3", 16, 5, 64, 2, False), ("v3", 64, 5, 1024, 1, True), ("v3", 128, 5, 256, 3, True),
    ("v2", 16, 5, 64, 2, False), ("v3", 24, 5, 13, 2, False)])
def test_fused_block_equals_three_call_route(cuda_device, kind, d_model, nf, hw, batch, autocast):
    """Mamba(v3): directions fused into one conv + one scan launch vs the reference's data flow (three
    mamba_inner_fn_no_out_proj calls on flipped / interleaved copies, mamba_simple.py:217-264) run on this library's
    single-direction kernels.  The conv / scan arithmetic is identical; the GEMMs are batched differently."""
    from mamba_ssm import Mamba
    torch.manual_seed(0)
    m = Mamba(d_model=d_model, d_state=16, d_conv=4, expand=2, bimamba_type=kind, nframes=nf).cuda()
    with torch.no_grad():
        for n, p in m.named_parameters():
            if n.startswith("D") or "A" in n.split(".")[0]:
                p.add_(0.1 * torch.randn_like(p))
    x = torch.randn(batch, nf * hw, d_model, device="cuda")
    gy = torch.randn(batch, nf * hw, d_model, device="cuda")
    m.fuse_directions = True
    y1, dx1, g1 = _module_run(m, x, gy, autocast)
    m.fuse_directions = False
    y2, dx2, g2 = _module_run(m, x, gy, autocast)
    tol = 3e-2 if autocast else 2e-3
    assert rel_err(y1, y2) < tol, rel_err(y1, y2)
    assert rel_err(dx1, dx2) < tol, rel_err(dx1, dx2)
    bad = {n: rel_err(g1[n], g2[n]) for n in g1 if not rel_err(g1[n], g2[n]) < tol}
    assert not bad, bad


def test_fused_block_launch_count(cuda_device):
    """What f1 / f2 are for: kernels launched per Mamba(v3) forward + backward, fused vs three-call route."""
    from mamba_ssm import Mamba
    from torch.profiler import ProfilerActivity, profile
    torch.manual_seed(0)
    m = Mamba(d_model=64, d_state=16, bimamba_type="v3", nframes=5).cuda()
    x = torch.randn(1, 5 * 256, 64, device="cuda", requires_grad=True)
    counts = {}
    for fused in (True, False):
        m.fuse_directions = fused
        for _ in range(2):
            m(x).sum().backward()
        torch.cuda.synchronize()
        try:
            with profile(activities=[ProfilerActivity.CUDA]) as prof:
                m(x).sum().backward()
                torch.cuda.synchronize()
            counts[fused] = sum(e.count for e in prof.key_averages() if e.device_type.name == "CUDA")
        except Exception as e:   # no CUPTI on the box: nothing to count
            pytest.skip(f"torch profiler unavailable: {e}")
    print(f"[launches per Mamba(v3) fwd+bwd] fused {counts[True]}  three-call {counts[False]}")
    assert 0 < counts[True] < counts[False]


@pytest.mark.parametrize("shared_gate", [True, False])
def test_selective_scan_dirs_fn_autograd(cuda_device, shared_gate):
    """The public multi-direction op (selective_scan_dirs_fn) through autograd == one selective_scan_fn per direction on
    gathered copies, summed gradients for the shared gate."""
    from mamba_ssm.ops.selective_scan_interface import selective_scan_fn
    from vivim_b200.selective_scan_interface import selective_scan_dirs_fn
    torch.manual_seed(1)
    B_, D, nf, hw, N = 2, 16, 5, 64, 16
    L = nf * hw
    mk = lambda *s: torch.randn(*s, device="cuda").requires_grad_()  # noqa: E731
    u, delta = mk(B_, 3 * D, L), (0.5 * torch.rand(B_, 3 * D, L, device="cuda")).requires_grad_()
    A = (-0.5 * torch.rand(3 * D, N, device="cuda")).requires_grad_()
    Bm, Cm = mk(B_, 3, N, L), mk(B_, 3, N, L)
    Dv, bias = mk(3 * D), (0.5 * torch.rand(3 * D, device="cuda")).requires_grad_()
    z = mk(B_, D if shared_gate else 3 * D, L)
    g = torch.randn(B_, 3 * D, L, device="cuda")
    leaves = (u, delta, A, Bm, Cm, Dv, z, bias)
    out = selective_scan_dirs_fn(u, delta, A, Bm, Cm, Dv, z=z, delta_bias=bias, delta_softplus=True, dirs=V3, nframes=nf)
    got = [host(out)] + [host(t) for t in torch.autograd.grad(out, leaves, g)]
    outs = []
    for k, mode in enumerate(V3):
        p = _perm(mode, L, nf)
        ch = slice(k * D, (k + 1) * D)
        zk = (z if shared_gate else z[:, ch])[:, :, p]
        o = selective_scan_fn(u[:, ch][:, :, p], delta[:, ch][:, :, p], A[ch], Bm[:, k][:, :, p], Cm[:, k][:, :, p], Dv[ch],
                              z=zk, delta_bias=bias[ch], delta_softplus=True)
        full = torch.zeros(B_, D, L, device="cuda")
        full[:, :, p] = o
        outs.append(full)
    ref = torch.cat(outs, dim=1)
    want = [host(ref)] + [host(t) for t in torch.autograd.grad(ref, leaves, g)]
    for name, a, b in zip(("out", "du", "ddelta", "dA", "dB", "dC", "dD", "dz", "dbias"), got, want):
        assert rel_err(a, b) < 1e-4, (name, rel_err(a, b))
