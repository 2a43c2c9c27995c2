"""Drop-in for ``mamba_ssm.ops.selective_scan_interface`` of the reference
(mamba/mamba_ssm/ops/selective_scan_interface.py): ``selective_scan_fn``, ``mamba_inner_fn``,
``mamba_inner_fn_no_out_proj``, ``bimamba_inner_fn`` and the ``*_ref`` statements, with the same
signatures.  The conv and scan run in the sm_100a kernels of libvivim_b200.so; the x_proj / dt_proj /
out_proj GEMMs stay torch (cuBLAS) calls, as in the reference.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import causal_conv1d_cuda, selective_scan_cuda
from .causal_conv1d_interface import causal_conv1d_fn

_fwd_amp = torch.amp.custom_fwd(device_type="cuda")
_bwd_amp = torch.amp.custom_bwd(device_type="cuda")


def _last_contig(t):
    return t if t is None or t.stride(-1) == 1 else t.contiguous()


def _as_groups(M, u, A, name):
    """B / C as (batch, groups, dstate, seqlen).  Returns (tensor, kind) with kind in
    'bnl' (was 3-D variable), 'bgnl' (4-D variable) or 'dn' (constant (dim, dstate))."""
    if A.is_complex():
        raise NotImplementedError("complex A is not served by the B200 kernels")
    if M.dim() == 4:
        return _last_contig(M), "bgnl"
    if M.dim() == 3:
        return _last_contig(M).unsqueeze(1), "bnl"
    if M.dim() == 2:
        # constant B/C (dim, dstate): one group per channel, broadcast over batch and time.
        # Correct but slow; Vivim always uses input-dependent B and C.
        batch, dim, seqlen = u.shape
        return M.to(u.dtype)[None, :, :, None].expand(batch, dim, M.shape[1], seqlen).contiguous(), "dn"
    raise RuntimeError(f"selective_scan: {name} must have 2, 3 or 4 dimensions")


def _ungroup_grad(dM, kind):
    if kind == "bnl":
        return dM.squeeze(1)
    if kind == "dn":
        return dM.float().sum(dim=(0, 3))   # autograd casts to the parameter's dtype
    return dM


class SelectiveScanFn(torch.autograd.Function):
    """reference: selective_scan_interface.py:14-74"""

    @staticmethod
    def forward(ctx, u, delta, A, B, C, D=None, z=None, delta_bias=None, delta_softplus=False,
                return_last_state=False):
        u, delta, z = _last_contig(u), _last_contig(delta), _last_contig(z)
        if D is not None:
            D = D.contiguous()
        Bg, ctx.kind_B = _as_groups(B, u, A, "B")
        Cg, ctx.kind_C = _as_groups(C, u, A, "C")
        out, chk, last_state, *rest = selective_scan_cuda.fwd(
            u, delta, A, Bg, Cg, D, z, delta_bias, delta_softplus, want_out=z is None)
        ctx.delta_softplus = delta_softplus
        ctx.has_z = z is not None
        ctx.save_for_backward(u, delta, A, Bg, Cg, D, z, delta_bias, chk)
        result = rest[0] if ctx.has_z else out
        if return_last_state:
            ctx.mark_non_differentiable(last_state)
            return result, last_state
        return result

    @staticmethod
    def backward(ctx, dout, *unused):
        u, delta, A, Bg, Cg, D, z, delta_bias, chk = ctx.saved_tensors
        dout = _last_contig(dout)
        du, ddelta, dA, dB, dC, dD, ddelta_bias, *rest = selective_scan_cuda.bwd(
            u, delta, A, Bg, Cg, D, z, delta_bias, dout, chk, None, ctx.delta_softplus)
        dz = rest[0] if ctx.has_z else None
        return (du, ddelta, dA, _ungroup_grad(dB, ctx.kind_B), _ungroup_grad(dC, ctx.kind_C),
                dD, dz, ddelta_bias, None, None)


def selective_scan_fn(u, delta, A, B, C, D=None, z=None, delta_bias=None, delta_softplus=False,
                      return_last_state=False):
    """If ``return_last_state`` is True returns (out, last_state) with last_state (batch, dim, dstate);
    the gradient of last_state is not propagated.  reference: selective_scan_interface.py:77-83"""
    return SelectiveScanFn.apply(u, delta, A, B, C, D, z, delta_bias, delta_softplus, return_last_state)


class SelectiveScanDirsFn(torch.autograd.Function):
    """The scans of all directions of a Temporal Mamba block as one op (one launch chain instead of
    ``len(dirs)`` calls of SelectiveScanFn on flipped / frame-interleaved copies, mamba_simple.py:217-260).

    u, delta (B, nd*D, L): nd direction blocks of D channels; A (nd*D, N), D, delta_bias (nd*D,); B, C (B, nd*G, N, L),
    any strides (e.g. permuted views of x_proj's output); z (B, D, L): ONE gate shared by the blocks (or (B, nd*D, L)).
    Block k visits the tokens in order dirs[k] in ('fwd', 'rev', 'frames'); every tensor stays in memory order.
    Gradient of z: summed over the blocks when z is shared."""

    @staticmethod
    def forward(ctx, u, delta, A, B, C, D, z, delta_bias, delta_softplus, dirs, nframes):
        u, delta, z = _last_contig(u), _last_contig(delta), _last_contig(z)
        out, chk, _, *rest = selective_scan_cuda.fwd(u, delta, A, B, C, D, z, delta_bias, delta_softplus,
                                                     want_out=z is None, dirs=dirs, nframes=nframes)
        ctx.delta_softplus, ctx.dirs, ctx.nframes = delta_softplus, dirs, nframes
        ctx.save_for_backward(u, delta, A, B, C, D, z, delta_bias, chk)
        return rest[0] if z is not None else out

    @staticmethod
    def backward(ctx, dout):
        u, delta, A, B, C, D, z, delta_bias, chk = ctx.saved_tensors
        nd = len(ctx.dirs)
        shared = z is not None and z.shape[1] != u.shape[1]
        if shared:
            # the blocks are gated by the same z: the op's output is (B, nd*D, L), so each block has its own upstream
            # gradient rows -- run the backward with full-width gate tensors
            z_full = z.repeat(1, nd, 1)
        else:
            z_full = z
        du, ddelta, dA, dB, dC, dD, ddelta_bias, *rest = selective_scan_cuda.bwd(
            u, delta, A, B, C, D, z_full, delta_bias, _last_contig(dout), chk, None, ctx.delta_softplus,
            dirs=ctx.dirs, nframes=ctx.nframes)
        dz = rest[0] if z is not None else None
        if shared:
            dz = dz.view(dz.shape[0], nd, -1, dz.shape[2]).sum(1).to(z.dtype)
        return du, ddelta, dA, dB, dC, dD, dz, ddelta_bias, None, None, None


def selective_scan_dirs_fn(u, delta, A, B, C, D=None, z=None, delta_bias=None, delta_softplus=False,
                           dirs=("fwd", "rev", "frames"), nframes=5):
    """Multi-direction selective scan, see SelectiveScanDirsFn.  B / C must be 4-D (batch, groups, dstate, seqlen)."""
    return SelectiveScanDirsFn.apply(u, delta, A, B, C, D, z, delta_bias, delta_softplus, tuple(dirs), int(nframes))


def selective_scan_ref(u, delta, A, B, C, D=None, z=None, delta_bias=None, delta_softplus=False,
                       return_last_state=False):
    """Pure-torch statement of the op's semantics for real A (reference:
    selective_scan_interface.py:86-152).  Kept importable for API parity; never called by the
    product path.  Sequential in time; O(L) time, fp32 state."""
    if A.is_complex():
        raise NotImplementedError("selective_scan_ref here covers real A only")
    dt = delta.float()
    if delta_bias is not None:
        dt = dt + delta_bias.float()[:, None]
    if delta_softplus:
        dt = F.softplus(dt)
    uf = u.float()
    batch, dim, seqlen = uf.shape

    def per_channel(M):  # -> (batch, dim, dstate, seqlen)
        M = M.float()
        if M.dim() == 2:
            return M[None, :, :, None].expand(batch, -1, -1, seqlen)
        if M.dim() == 3:
            return M[:, None].expand(-1, dim, -1, -1)
        return M.repeat_interleave(dim // M.shape[1], dim=1)

    Bf, Cf = per_channel(B), per_channel(C)
    h = uf.new_zeros(batch, dim, A.shape[1])
    ys = []
    for t in range(seqlen):
        h = torch.exp(dt[:, :, t, None] * A) * h + (dt[:, :, t] * uf[:, :, t])[..., None] * Bf[..., t]
        ys.append((h * Cf[..., t]).sum(-1))
    y = torch.stack(ys, dim=2)
    if D is not None:
        y = y + uf * D[:, None]
    if z is not None:
        y = y * F.silu(z.float())
    y = y.to(u.dtype)
    return (y, h) if return_last_state else y


# ---------------------------------------------------------------------------------------------
# fused inner block: conv1d+SiLU -> x_proj -> dt_proj -> selective scan (gated by z) [-> out_proj]
# ---------------------------------------------------------------------------------------------

def _xdbl_layout(rank, dstate):
    """Column layout of x_dbl inside the fused inner functions: [dt (rank) | zeros | B (dstate) | C (dstate) | zeros] with the
    dt block and the whole row padded to multiples of 8 elements.  The reference packs rank + 2*dstate columns (36 at
    Vivim's stage 1), which leaves every GEMM that touches x_dbl / dx_dbl on 4-byte aligned operands (cuBLAS then picks its
    slow `align2` kernels: 370 us of a 1.15 ms block, profiles/r02_block.md).  -> (Rp, R2p): offset of B, padded width."""
    Rp = -(-rank // 8) * 8
    return Rp, -(-(Rp + 2 * dstate) // 8) * 8


def _pad_proj_weights(x_proj_weight, delta_proj_weight, dstate):
    """x_proj (rank+2N, d) -> (R2p, d), dt_proj (d, rank) -> (d, Rp): zero rows / columns in the padding."""
    rank = delta_proj_weight.shape[1]
    Rp, R2p = _xdbl_layout(rank, dstate)
    if Rp == rank and R2p == rank + 2 * dstate:
        return x_proj_weight, delta_proj_weight
    xw = x_proj_weight.new_zeros(R2p, x_proj_weight.shape[1])
    xw[:rank] = x_proj_weight[:rank]
    xw[Rp:Rp + 2 * dstate] = x_proj_weight[rank:]
    dw = delta_proj_weight.new_zeros(delta_proj_weight.shape[0], Rp)
    dw[:, :rank] = delta_proj_weight
    return xw, dw


def _project(conv_out, xw, dw, rank, A, B, C, B_proj_bias, C_proj_bias):
    """x_dbl, delta, B, C from the conv output (reference: selective_scan_interface.py:181-210); xw / dw are the padded
    projection weights of _pad_proj_weights."""
    batch, _, L = conv_out.shape
    dstate = A.shape[-1]
    Rp = dw.shape[1]
    x_dbl = F.linear(conv_out.transpose(1, 2).reshape(batch * L, -1), xw)              # (b l, R2p)
    delta = _delta_from(x_dbl, dw, batch, L)                                           # (b, d, l) view

    def pick(M, lo, proj_bias, name):
        if M is not None:   # caller-provided B / C: not input-dependent through x_proj
            return _as_groups(M, conv_out, A, name)
        cols = x_dbl[:, lo:lo + dstate]
        if proj_bias is not None:
            cols = (cols + proj_bias.to(cols.dtype)).reshape(batch, L, dstate).transpose(1, 2).contiguous().unsqueeze(1)
            return cols, "proj"
        # No copy: a (b, 1, n, l) VIEW of x_dbl's columns (state stride 1, sequence stride R2p) -- the kernels read B / C
        # from the rows of the GEMM output where the reference transposes them (selective_scan_interface.py:187-207)
        return cols.unflatten(0, (batch, L)).permute(0, 2, 1).unsqueeze(1), "proj"

    Bm, kind_B = pick(B, Rp, B_proj_bias, "B")
    Cm, kind_C = pick(C, Rp + dstate, C_proj_bias, "C")
    return x_dbl, delta, Bm, Cm, kind_B, kind_C


def _delta_from(x_dbl, dw, batch, L):
    return (dw @ x_dbl[:, :dw.shape[1]].t()).view(-1, batch, L).transpose(0, 1)


class _MambaInner(torch.autograd.Function):
    """One Function behind mamba_inner_fn (with out_proj) and mamba_inner_fn_no_out_proj.
    reference: MambaInnerFnNoOutProj (selective_scan_interface.py:155-289) and MambaInnerFn
    (:292-434).  checkpoint_lvl=1 semantics: conv output and delta are recomputed in backward."""

    @staticmethod
    @_fwd_amp
    def forward(ctx, xz, conv1d_weight, conv1d_bias, x_proj_weight, delta_proj_weight,
                out_proj_weight, out_proj_bias, A, B, C, D, delta_bias, B_proj_bias, C_proj_bias,
                delta_softplus, has_out_proj, checkpoint_lvl):
        assert checkpoint_lvl in (0, 1)
        if A.is_complex():
            raise NotImplementedError("complex A is not served by the B200 kernels")
        if torch.is_autocast_enabled("cuda"):
            amp = torch.get_autocast_dtype("cuda")
            x_proj_weight = x_proj_weight.to(amp)
            delta_proj_weight = delta_proj_weight.to(amp)
            if has_out_proj:
                out_proj_weight = out_proj_weight.to(amp)
                out_proj_bias = out_proj_bias.to(amp) if out_proj_bias is not None else None
        xz = _last_contig(xz)
        conv_w = conv1d_weight.squeeze(1)                      # (d, 1, w) -> (d, w)
        conv_b = conv1d_bias.contiguous() if conv1d_bias is not None else None
        x, z = xz.chunk(2, dim=1)                              # views: batch stride 2*D*L
        conv_out = causal_conv1d_cuda.causal_conv1d_fwd(x, conv_w, conv_b, True)
        rank = delta_proj_weight.shape[1]
        xw, dw = _pad_proj_weights(x_proj_weight, delta_proj_weight, A.shape[-1])
        x_dbl, delta, Bm, Cm, kind_B, kind_C = _project(conv_out, xw, dw, rank, A, B, C, B_proj_bias, C_proj_bias)
        if D is not None:
            D = D.contiguous()
        _, chk, _, out_z = selective_scan_cuda.fwd(
            conv_out, delta, A, Bm, Cm, D, z, delta_bias, delta_softplus, want_out=False)
        ctx.delta_softplus = delta_softplus
        ctx.has_out_proj = has_out_proj
        ctx.kind_B, ctx.kind_C = kind_B, kind_C
        ctx.has_B_bias, ctx.has_C_bias = B_proj_bias is not None, C_proj_bias is not None
        ctx.has_out_bias = out_proj_bias is not None
        ctx.recompute = checkpoint_lvl >= 1
        ctx.rank = rank
        ctx.save_for_backward(xz, conv_w, conv_b, x_dbl, xw, dw,
                              out_proj_weight if has_out_proj else None,
                              None if ctx.recompute else conv_out, None if ctx.recompute else delta,
                              A, Bm, Cm, D, delta_bias, chk, out_z if has_out_proj else None)
        if has_out_proj:
            return F.linear(out_z.transpose(1, 2), out_proj_weight, out_proj_bias)   # (b, l, e)
        return out_z                                                                   # (b, d, l)

    @staticmethod
    @_bwd_amp
    def backward(ctx, dout):
        (xz, conv_w, conv_b, x_dbl, xw, dw, out_proj_weight,
         conv_out, delta, A, Bm, Cm, D, delta_bias, chk, out_z) = ctx.saved_tensors
        batch, two_d, L = xz.shape
        dim = two_d // 2
        rank, Rp = ctx.rank, dw.shape[1]          # x_dbl columns: [dt (rank) | 0 | B | C | 0], see _xdbl_layout
        dstate = A.shape[-1]
        cB, cC, cEnd = Rp, Rp + dstate, Rp + 2 * dstate
        x, z = xz.chunk(2, dim=1)
        if ctx.recompute:
            conv_out = causal_conv1d_cuda.causal_conv1d_fwd(x, conv_w, conv_b, True)
            delta = _delta_from(x_dbl, dw, batch, L)
        dout_proj_weight = dout_proj_bias = None
        if ctx.has_out_proj:
            dflat = dout.reshape(batch * L, -1)                                      # (b l, e)
            dout_y = (dflat @ out_proj_weight).view(batch, L, dim).transpose(1, 2).contiguous()
            dout_proj_weight = dflat.t() @ out_z.transpose(1, 2).reshape(batch * L, dim)
            dout_proj_bias = dflat.sum(0) if ctx.has_out_bias else None
        else:
            dout_y = _last_contig(dout)
        # dx and dz are written next to each other so that no torch.cat is needed
        dxz = torch.empty_like(xz)
        dx, dz = dxz.chunk(2, dim=1)
        dx_dbl = torch.empty_like(x_dbl)
        if dx_dbl.shape[1] != cEnd:
            dx_dbl[:, cEnd:] = 0                   # trailing padding; the dt padding is written (as zeros) by the GEMM below
        both_proj = ctx.kind_B == "proj" and ctx.kind_C == "proj"
        dBC_out = None
        if both_proj:
            # dB / dC land straight in their column blocks of dx_dbl (the cast kernel of vv_scan_bwd writes them there):
            # no rearrange + slice copy (selective_scan_interface.py:255-271)
            cols = lambda lo, hi: dx_dbl[:, lo:hi].unflatten(0, (batch, L)).permute(0, 2, 1).unsqueeze(1)  # noqa: E731
            dBC_out = (cols(cB, cC), cols(cC, cEnd))
        dconv, ddelta, dA, dB, dC, dD, ddelta_bias, dz = selective_scan_cuda.bwd(
            conv_out, delta, A, Bm, Cm, D, z, delta_bias, dout_y, chk, dz, ctx.delta_softplus, dBC_out=dBC_out)
        dB_out = dC_out = dB_proj_bias = dC_proj_bias = None
        if both_proj:
            dB_proj_bias = dx_dbl[:, cB:cC].sum(0) if ctx.has_B_bias else None
            dC_proj_bias = dx_dbl[:, cC:cEnd].sum(0) if ctx.has_C_bias else None
        else:
            if ctx.kind_B == "proj":
                dBf = dB.squeeze(1).transpose(1, 2).reshape(batch * L, dstate)
                dB_proj_bias = dBf.sum(0) if ctx.has_B_bias else None
                dx_dbl[:, cB:cC] = dBf
            else:
                dB_out = _ungroup_grad(dB, ctx.kind_B)
                dx_dbl[:, cB:cC] = 0
            if ctx.kind_C == "proj":
                dCf = dC.squeeze(1).transpose(1, 2).reshape(batch * L, dstate)
                dC_proj_bias = dCf.sum(0) if ctx.has_C_bias else None
                dx_dbl[:, cC:cEnd] = dCf
            else:
                dC_out = _ungroup_grad(dC, ctx.kind_C)
                dx_dbl[:, cC:cEnd] = 0
        ddelta_f = ddelta.transpose(0, 1).reshape(dim, batch * L)                     # (d, b l)
        ddelta_proj_weight = (ddelta_f @ x_dbl[:, :Rp])[:, :rank]
        dx_dbl[:, :Rp] = ddelta_f.t() @ dw                                            # zeros beyond `rank`
        conv_flat = conv_out.transpose(1, 2).reshape(batch * L, dim)                  # (b l, d)
        dxw = dx_dbl.t() @ conv_flat                                                  # (R2p, d)
        dx_proj_weight = dxw if cEnd == rank + 2 * dstate and Rp == rank else torch.cat([dxw[:rank], dxw[cB:cEnd]])
        # dconv (b, d, l) += x_proj_weight^T (d, R2p) @ dx_dbl^T (R2p, l), per batch entry, accumulated in place in the
        # kernels' (b, d, l) layout (no (b, l, d) intermediate and transposing add)
        dconv.baddbmm_(xw.t().unsqueeze(0).expand(batch, -1, -1), dx_dbl.view(batch, L, -1).transpose(1, 2))
        dx, dconv_w, dconv_b = causal_conv1d_cuda.causal_conv1d_bwd(x, conv_w, conv_b, dconv, dx, True)
        return (dxz, dconv_w.unsqueeze(1), dconv_b, dx_proj_weight, ddelta_proj_weight,
                dout_proj_weight, dout_proj_bias, dA, dB_out, dC_out, dD, ddelta_bias,
                dB_proj_bias, dC_proj_bias, None, None, None)


def mamba_inner_fn(xz, conv1d_weight, conv1d_bias, x_proj_weight, delta_proj_weight,
                   out_proj_weight, out_proj_bias, A, B=None, C=None, D=None, delta_bias=None,
                   B_proj_bias=None, C_proj_bias=None, delta_softplus=True):
    """reference: selective_scan_interface.py:606-614.  Returns (batch, seqlen, d_model)."""
    return _MambaInner.apply(xz, conv1d_weight, conv1d_bias, x_proj_weight, delta_proj_weight,
                             out_proj_weight, out_proj_bias, A, B, C, D, delta_bias, B_proj_bias,
                             C_proj_bias, delta_softplus, True, 1)


def mamba_inner_fn_no_out_proj(xz, conv1d_weight, conv1d_bias, x_proj_weight, delta_proj_weight,
                               A, B=None, C=None, D=None, delta_bias=None, B_proj_bias=None,
                               C_proj_bias=None, delta_softplus=True):
    """What Vivim's Mamba(bimamba_type="v3") calls three times per layer
    (reference: selective_scan_interface.py:627-633; caller mamba_simple.py:217-260).
    Returns the gated scan output (batch, d_inner, seqlen)."""
    return _MambaInner.apply(xz, conv1d_weight, conv1d_bias, x_proj_weight, delta_proj_weight,
                             None, None, A, B, C, D, delta_bias, B_proj_bias, C_proj_bias,
                             delta_softplus, False, 1)


def _inner_pre(xz, conv1d_weight, conv1d_bias, x_proj_weight, delta_proj_weight, A, B, C,
               B_proj_bias, C_proj_bias, conv_fn):
    x, z = xz.chunk(2, dim=1)
    xc = conv_fn(x, conv1d_weight.squeeze(1), conv1d_bias, "silu")
    xw, dw = _pad_proj_weights(x_proj_weight, delta_proj_weight, A.shape[-1])
    _, delta, Bm, Cm, _, _ = _project(xc, xw, dw, delta_proj_weight.shape[1], A, B, C, B_proj_bias, C_proj_bias)
    return xc, z, delta, Bm, Cm


def bimamba_inner_fn(xz, conv1d_weight, conv1d_bias, x_proj_weight, delta_proj_weight,
                     out_proj_weight, out_proj_bias, A, A_b, B=None, C=None, D=None, delta_bias=None,
                     B_proj_bias=None, C_proj_bias=None, delta_softplus=True):
    """Shared conv / projections, one scan left-to-right with A and one right-to-left with A_b, summed,
    then out_proj (reference: BiMambaInnerFn, selective_scan_interface.py:437-603, 616-624).
    Exported for API parity (mamba_ssm/__init__.py:3); Vivim never calls it.  Composed from the
    autograd ops above rather than hand-fused."""
    xc, z, delta, Bm, Cm = _inner_pre(xz, conv1d_weight, conv1d_bias, x_proj_weight, delta_proj_weight,
                                      A, B, C, B_proj_bias, C_proj_bias, causal_conv1d_fn)
    y = selective_scan_fn(xc, delta, A, Bm, Cm, D, z=z, delta_bias=delta_bias, delta_softplus=delta_softplus)
    y_b = selective_scan_fn(xc.flip([-1]), delta.flip([-1]), A_b, Bm.flip([-1]), Cm.flip([-1]), D,
                            z=z.flip([-1]), delta_bias=delta_bias, delta_softplus=delta_softplus)
    return F.linear((y + y_b.flip([-1])).transpose(1, 2), out_proj_weight, out_proj_bias)


def mamba_inner_ref(xz, conv1d_weight, conv1d_bias, x_proj_weight, delta_proj_weight,
                    out_proj_weight, out_proj_bias, A, B=None, C=None, D=None, delta_bias=None,
                    B_proj_bias=None, C_proj_bias=None, delta_softplus=True):
    """Unfused composition of the public ops (reference: selective_scan_interface.py:636-670)."""
    xc, z, delta, Bm, Cm = _inner_pre(xz, conv1d_weight, conv1d_bias, x_proj_weight, delta_proj_weight,
                                      A, B, C, B_proj_bias, C_proj_bias, causal_conv1d_fn)
    y = selective_scan_fn(xc, delta, A, Bm, Cm, D, z=z, delta_bias=delta_bias, delta_softplus=True)
    return F.linear(y.transpose(1, 2), out_proj_weight, out_proj_bias)


def bimamba_inner_ref(xz, conv1d_weight, conv1d_bias, x_proj_weight, delta_proj_weight,
                      out_proj_weight, out_proj_bias, A, A_b, B=None, C=None, D=None, delta_bias=None,
                      B_proj_bias=None, C_proj_bias=None, delta_softplus=True):
    """reference: selective_scan_interface.py:673-709 (same composition as bimamba_inner_fn here)."""
    return bimamba_inner_fn(xz, conv1d_weight, conv1d_bias, x_proj_weight, delta_proj_weight,
                            out_proj_weight, out_proj_bias, A, A_b, B, C, D, delta_bias,
                            B_proj_bias, C_proj_bias, True)
