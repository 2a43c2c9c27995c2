"""Tensor-level entry points with the reference pybind module's names
(mamba/csrc/selective_scan/selective_scan.cpp:494-497): ``fwd`` and ``bwd``.

Differences from the reference, all on the private side of the autograd Functions:
* the scan intermediate returned by ``fwd`` is the checkpoint tensor ``chk`` (B, D, segments, N) of
  states entering each 64-position segment, not the reference's (B, D, n_chunks, 2N) pairs
  (selective_scan.cpp:307-313); ``fwd`` therefore also returns ``last_state`` explicitly;
* ``bwd`` recomputes the forward states from ``chk`` and never reads a saved ``out``.
Served: real fp32 ``A``, input-dependent B and C ((B,N,L) or (B,G,N,L)), dstate <= 32.
Constant (D,N) B/C are expanded to per-channel groups (slow, correct); complex ``A`` raises.

Beyond the reference op (keyword arguments, all optional):
* ``dirs`` / ``nframes``: the channels are ``len(dirs)`` direction blocks, block k visits the tokens in order
  ``dirs[k]`` ('fwd', 'rev', 'frames'): Mamba.forward v3's three scans (mamba_simple.py:217-260) as one launch, with all
  tensors left in memory order;
* B / C may be any strided (B,G,N,L) VIEW -- in particular a permuted view of x_proj's output x_dbl (B*L, R+2N), which the
  reference transposes into (B,1,N,L) (selective_scan_interface.py:187-207); ``bwd(..., dBC_out=(dB_view, dC_view))``
  writes dB / dC straight into such views (e.g. column blocks of dx_dbl, :255-271);
* ``z`` (and ``dout``) may have fewer channel rows than u: row = d % z.shape[1] (one gate shared by the direction blocks).
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib
from .causal_conv1d_cuda import dir_codes

_DTYPES = {torch.float32: _lib.VV_F32, torch.float16: _lib.VV_F16, torch.bfloat16: _lib.VV_BF16}

LAUNCHES = 0  # kernels enqueued through this module (bench.py reads the total)


def _check(cond, msg):
    if not cond:
        raise RuntimeError(msg)


def num_segments(seqlen: int) -> int:
    return (seqlen + _lib.VV_SCAN_SEGMENT - 1) // _lib.VV_SCAN_SEGMENT


_SCRATCH = {}


def _scratch(name, numel, device):
    """fp32 scratch (agg / radj) cached per (device, stream, size): contents never outlive one call, and calls on one
    stream are ordered, so consecutive scans of the same shape reuse one buffer instead of allocating 5 MB each."""
    key = (name, device.index, torch.cuda.current_stream(device).cuda_stream, numel)
    t = _SCRATCH.get(key)
    if t is None:
        if len(_SCRATCH) > 64:
            _SCRATCH.clear()
        t = _SCRATCH[key] = torch.empty(numel, dtype=torch.float32, device=device)
    return t


def _checks(u, delta, A, B, C, D, z, delta_bias):
    _check(u.is_cuda, "selective_scan: tensors must be on a CUDA device")
    _check(u.dtype in _DTYPES, "selective_scan: u must be float32, float16 or bfloat16")
    if A.is_complex():
        raise NotImplementedError("selective_scan: complex A is not served by the B200 kernels "
                                  "(Vivim's A is real fp32, mamba_simple.py:212)")
    _check(A.dtype == torch.float32, "selective_scan: A must be float32")
    _check(delta.dtype == u.dtype and delta.shape == u.shape, "selective_scan: delta must match u")
    _check(u.dim() == 3, "selective_scan: u must be (batch, dim, seqlen)")
    batch, dim, seqlen = u.shape
    dstate = A.shape[1]
    _check(A.shape == (dim, dstate), "selective_scan: A must be (dim, dstate)")
    _check(dstate <= 256, "selective_scan only supports state dimension <= 256")
    _check(u.stride(-1) == 1 and delta.stride(-1) == 1, "selective_scan: u and delta must be contiguous along seqlen")
    for name, M in (("B", B), ("C", C)):
        _check(M.dim() == 4, f"selective_scan: {name} must be (batch, groups, dstate, seqlen) here")
        _check(M.dtype == u.dtype, f"selective_scan: variable {name} must have the dtype of u")
        _check(M.shape[0] == batch and M.shape[2] == dstate and M.shape[3] == seqlen,
               f"selective_scan: {name} has the wrong shape")
        _check(M.stride(-1) == 1 or M.stride(2) == 1,
               f"selective_scan: {name} must be contiguous along seqlen or along dstate (a view of x_dbl)")
        _check(dim % M.shape[1] == 0, f"selective_scan: groups of {name} must divide dim")
    _check(B.shape[1] == C.shape[1], "selective_scan: B and C must have the same number of groups")
    if D is not None:
        _check(D.dtype == torch.float32 and D.shape == (dim,) and D.stride(-1) == 1,
               "selective_scan: D must be a contiguous float32 (dim,) tensor")
    if delta_bias is not None:
        _check(delta_bias.dtype == torch.float32 and delta_bias.shape == (dim,) and delta_bias.stride(-1) == 1,
               "selective_scan: delta_bias must be a contiguous float32 (dim,) tensor")
    if z is not None:
        _check(z.dtype == u.dtype and z.dim() == 3 and z.shape[0] == batch and z.shape[2] == seqlen
               and dim % z.shape[1] == 0 and z.stride(-1) == 1,
               "selective_scan: z must be (batch, dim or a divisor of dim, seqlen), contiguous along seqlen")


def _fill_common(a, u, delta, A, B, C, D, z, delta_bias, delta_softplus, dirs=None, nframes=0):
    a.u, a.delta, a.A, a.Bm, a.Cm = u.data_ptr(), delta.data_ptr(), A.data_ptr(), B.data_ptr(), C.data_ptr()
    a.D = D.data_ptr() if D is not None else None
    a.z = z.data_ptr() if z is not None else None
    a.delta_bias = delta_bias.data_ptr() if delta_bias is not None else None
    a.batch, a.dim, a.seqlen = u.shape
    a.dstate = A.shape[1]
    a.ngroups = B.shape[1]
    a.u_bs, a.u_ds = u.stride(0), u.stride(1)
    a.delta_bs, a.delta_ds = delta.stride(0), delta.stride(1)
    if z is not None:
        a.z_bs, a.z_ds = z.stride(0), z.stride(1)
    a.A_ds, a.A_ns = A.stride(0), A.stride(1)
    a.B_bs, a.B_gs, a.B_ns, a.B_ls = B.stride(0), B.stride(1), B.stride(2), B.stride(3)
    a.C_bs, a.C_gs, a.C_ns, a.C_ls = C.stride(0), C.stride(1), C.stride(2), C.stride(3)
    a.io_dtype = _DTYPES[u.dtype]
    a.delta_softplus = int(bool(delta_softplus))
    if dirs is not None:
        codes = dir_codes(dirs)
        _check(a.dim % len(codes) == 0 and a.ngroups % len(codes) == 0,
               "selective_scan: the number of directions must divide dim and the groups of B / C")
        a.ndirs = len(codes)
        for k, m in enumerate(codes):
            a.dir_mode[k] = m
        a.nframes = int(nframes)
    if z is not None and z.shape[1] != u.shape[1]:
        a.gate_rows = z.shape[1]


def fwd(u, delta, A, B, C, D, z, delta_bias, delta_softplus, want_out=True, dirs=None, nframes=0):
    """-> [out, chk, last_state] (+ [out_z] when z is given).  ``out`` is None when
    ``want_out`` is False and z is given (the pre-gate y is not needed by the backward)."""
    global LAUNCHES
    _checks(u, delta, A, B, C, D, z, delta_bias)
    batch, dim, seqlen = u.shape
    dstate = A.shape[1]
    U = num_segments(seqlen)
    dev = u.device
    need_out = want_out or z is None
    out = torch.empty_like(u, memory_format=torch.contiguous_format) if need_out else None
    out_z = torch.empty_like(u, memory_format=torch.contiguous_format) if z is not None else None
    chk = torch.empty((batch, dim, U, dstate), dtype=torch.float32, device=dev)
    last_state = torch.empty((batch, dim, dstate), dtype=torch.float32, device=dev)
    if u.numel() > 0:
        agg = _scratch("agg", batch * dim * U * dstate * 2, dev)
        a = _lib.ScanArgs()
        _fill_common(a, u, delta, A, B, C, D, z, delta_bias, delta_softplus, dirs, nframes)
        if out is not None:
            a.out, a.out_bs, a.out_ds = out.data_ptr(), out.stride(0), out.stride(1)
        if out_z is not None:
            a.out_z, a.outz_bs, a.outz_ds = out_z.data_ptr(), out_z.stride(0), out_z.stride(1)
        a.last_state, a.agg, a.chk = last_state.data_ptr(), agg.data_ptr(), chk.data_ptr()
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream().cuda_stream
            _lib.check(_lib.lib().vv_scan_fwd(ctypes.byref(a), ctypes.c_void_p(stream)), "vv_scan_fwd")
        LAUNCHES += 3
    res = [out, chk, last_state]
    if z is not None:
        res.append(out_z)
    return res


def bwd(u, delta, A, B, C, D, z, delta_bias, dout, chk, dz, delta_softplus, dirs=None, nframes=0, dBC_out=None):
    """-> [du, ddelta, dA, dB, dC, dD, ddelta_bias] (+ [dz] when z is given).
    dB/dC come back in the dtype of B/C (fp32 accumulation, then cast, as selective_scan.cpp:461-488).
    ``dz`` may be a caller-provided view (e.g. half of dxz)."""
    global LAUNCHES
    _checks(u, delta, A, B, C, D, z, delta_bias)
    batch, dim, seqlen = u.shape
    gate_rows = z.shape[1] if z is not None else dim
    _check(dout.shape == (batch, gate_rows, seqlen) and dout.dtype == u.dtype and dout.stride(-1) == 1,
           "selective_scan bwd: dout must match u (or z, when z has shared rows) and be contiguous along seqlen")
    dstate = A.shape[1]
    U = num_segments(seqlen)
    dev = u.device
    _check(chk.shape == (batch, dim, U, dstate) and chk.dtype == torch.float32 and chk.is_contiguous(),
           "selective_scan bwd: bad checkpoint tensor")
    du = torch.empty_like(u, memory_format=torch.contiguous_format)
    ddelta = torch.empty_like(delta, memory_format=torch.contiguous_format)
    # all fp32 accumulators live in one buffer [dB | dC | dA | dD | ddelta_bias]; vv_scan_bwd zero-fills it in its
    # first kernel (zero_accumulators = 1), so there is no fill launch (the reference's shim calls torch::zeros five
    # times: selective_scan.cpp:460-466)
    n_bc = B.numel()
    acc = (torch.empty if u.numel() > 0 else torch.zeros)(2 * n_bc + dim * dstate + 2 * dim, dtype=torch.float32, device=dev)
    dBC = acc[:2 * n_bc].view(2, *B.shape)
    dB, dC = dBC[0], dBC[1]
    dA = acc[2 * n_bc:2 * n_bc + dim * dstate].view(dim, dstate)
    dD = acc[2 * n_bc + dim * dstate:2 * n_bc + dim * dstate + dim] if D is not None else None
    ddelta_bias = acc[2 * n_bc + dim * dstate + dim:] if delta_bias is not None else None
    if z is not None:
        if dz is None:
            dz = torch.empty_like(u, memory_format=torch.contiguous_format)   # one gradient row per scanned channel
        else:
            _check(dz.shape == u.shape and dz.dtype == z.dtype and dz.stride(-1) == 1,
                   "selective_scan bwd: dz must match u and be contiguous along seqlen")
    if dBC_out is not None:
        for M in dBC_out:
            _check(M.shape == B.shape and M.dtype == B.dtype and M.device == B.device,
                   "selective_scan bwd: dBC_out must be two views of the shape and dtype of B")
    if u.numel() > 0:
        agg = _scratch("agg", batch * dim * U * dstate * 2, dev)
        radj = _scratch("radj", batch * dim * U * dstate, dev)
        a = _lib.ScanArgs()
        _fill_common(a, u, delta, A, B, C, D, z, delta_bias, delta_softplus, dirs, nframes)
        a.agg, a.chk, a.radj = agg.data_ptr(), chk.data_ptr(), radj.data_ptr()
        a.dout, a.dout_bs, a.dout_ds = dout.data_ptr(), dout.stride(0), dout.stride(1)
        a.du, a.du_bs, a.du_ds = du.data_ptr(), du.stride(0), du.stride(1)
        a.ddelta, a.ddelta_bs, a.ddelta_ds = ddelta.data_ptr(), ddelta.stride(0), ddelta.stride(1)
        if z is not None:
            a.dz, a.dz_bs, a.dz_ds = dz.data_ptr(), dz.stride(0), dz.stride(1)
        a.dA, a.dB, a.dC = dA.data_ptr(), dB.data_ptr(), dC.data_ptr()
        a.dD = dD.data_ptr() if dD is not None else None
        a.ddelta_bias = ddelta_bias.data_ptr() if ddelta_bias is not None else None
        a.zero_accumulators = 1
        traversal = dirs is not None and any(m != _lib.VV_DIR_FWD for m in dir_codes(dirs))
        if dBC_out is not None:        # written by the cast kernel, any strides, memory order
            dBC_io = dBC_out
            a.dB_io, a.dC_io = dBC_io[0].data_ptr(), dBC_io[1].data_ptr()
            a.dBio_bs, a.dBio_gs, a.dBio_ns, a.dBio_ls = dBC_io[0].stride()
            a.dCio_bs, a.dCio_gs, a.dCio_ns, a.dCio_ls = dBC_io[1].stride()
        elif B.dtype != torch.float32 or traversal:
            # dB / dC in the dtype of B / C (and in memory order): converted by a kernel chained to the backward
            dBC_io = torch.empty((2, *B.shape), dtype=B.dtype, device=dev)
            a.dB_io, a.dC_io = dBC_io[0].data_ptr(), dBC_io[1].data_ptr()
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream().cuda_stream
            _lib.check(_lib.lib().vv_scan_bwd(ctypes.byref(a), ctypes.c_void_p(stream)), "vv_scan_bwd")
        LAUNCHES += 3
    # fp32 accumulation, then one cast (selective_scan.cpp:461-462, 488)
    dBC = dBC_io if (u.numel() > 0 and (B.dtype != torch.float32 or dBC_out is not None or traversal)) else dBC.to(B.dtype)
    res = [du, ddelta, dA, dBC[0], dBC[1], dD, ddelta_bias]
    if z is not None:
        res.append(dz)
    return res
