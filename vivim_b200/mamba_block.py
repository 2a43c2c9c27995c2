"""All scan directions of ``Mamba.forward`` (bimamba_type "v2" / "v3") as ONE autograd Function.

Reference: mamba/mamba_ssm/modules/mamba_simple.py:204-264 -- in_proj, then per direction
``mamba_inner_fn_no_out_proj`` (selective_scan_interface.py:155-289) on xz, on ``xz.flip([-1])`` and on the
frame-interleaved copy of xz, then ``out_proj((y + y_b.flip + y_s.permute-back) / 3)``.

Here the directions are addressing modes of the kernels, every tensor stays in
memory order and the three parameter sets are channel-concatenated:

    xz       = in_proj(hidden)                         one GEMM, (B, 2D, L) view of a (2D, B*L) result
    conv_out = conv1d_dirs(x)                          ONE launch: x read once -> (B, nd*D, L)
    x_dbl    = conv_out^T @ x_proj^T                   one batched GEMM -> (B, nd, L, R+2N)
    delta    = dt_proj @ x_dbl[..., :R]^T              one batched GEMM -> (B, nd*D, L)
    out_z    = scan(conv_out, delta, A, B, C, D, z)    ONE launch (3 kernels): B / C are strided VIEWS of x_dbl, z is
                                                       shared by the directions, direction k walks in order dirs[k]
    out      = out_z^T @ [W_out ... W_out]^T * scale   the mean over directions folded into out_proj's K dimension

No flip, no interleave copy, no (b l) n -> b n l transposes, no torch.cat: 13 launches instead of ~45 in the forward.
The backward mirrors it (conv and delta recomputed: checkpoint_lvl = 1 of the reference): dB / dC are written by the scan
straight into column blocks of dx_dbl, the three dz land next to dx in one (B, (1+nd)*D, L) buffer whose sum over
directions is folded into in_proj's backward GEMMs.
"""
from __future__ import annotations

import torch

from . import causal_conv1d_cuda, selective_scan_cuda

_fwd_amp = torch.amp.custom_fwd(device_type="cuda")
_bwd_amp = torch.amp.custom_bwd(device_type="cuda")


def _bmm_shared(x, w):
    """x (B, M, K) @ w (K, N) -> (B, M, N) without folding the batch into M (x may be a transposed view, which a fold
    would have to copy)."""
    return torch.bmm(x, w.unsqueeze(0).expand(x.shape[0], -1, -1))


class MambaDirsFn(torch.autograd.Function):
    """hidden (B, L, E) -> (B, L, E).  Parameters of direction k are element k of the tuples, exactly the reference
    module's tensors (conv1d{sfx}.weight (D,1,K), conv1d{sfx}.bias (D), x_proj{sfx}.weight (R+2N, D),
    dt_proj{sfx}.weight (D, R), A (D, N) = -exp(A_log), D (D), dt_proj{sfx}.bias (D))."""

    @staticmethod
    @_fwd_amp
    def forward(ctx, hidden, in_w, in_b, out_w, out_b, dirs, nframes, scale, nd, *params):
        conv_w, conv_b, xp_w, dt_w, A, Dp, dt_b = (params[i * nd:(i + 1) * nd] for i in range(7))
        has_conv_b = conv_b[0] is not None
        amp = torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled("cuda") else None
        cast = (lambda t: t.to(amp)) if amp is not None else (lambda t: t)
        B_, L, E = hidden.shape
        Dn, R = dt_w[0].shape
        N = A[0].shape[1]
        R2 = R + 2 * N
        hid = cast(hidden).reshape(B_ * L, E)
        in_wc, out_wc = cast(in_w), cast(out_w)
        xz = (in_wc @ hid.t())                                               # (2D, B*L)
        if in_b is not None:
            xz = xz + cast(in_b)[:, None]
        xz = xz.view(2 * Dn, B_, L).transpose(0, 1)                          # (B, 2D, L), strides (L, B*L, 1)
        x, z = xz[:, :Dn], xz[:, Dn:]
        cw = torch.stack([w.reshape(Dn, -1).float() for w in conv_w])        # (nd, D, K)
        cb = torch.stack([b.float() for b in conv_b]) if has_conv_b else None
        # Column layout of x_dbl: [dt (R) | zeros to Rp | B (N) | C (N) | zeros to R2p], Rp and R2p multiples of 8.  With the
        # reference's packed (R + 2N = 36 at stage 1, 52 at stage 3) rows every GEMM that touches x_dbl / dx_dbl or the
        # R-wide dt block runs on 4-byte aligned operands and cuBLAS falls back to its `align2` WMMA kernels (190 + 180 us
        # of a 1.15 ms block at stage 1, profiles/r02_block.md).  The padding meets zero weight rows / columns.
        Rp = -(-R // 8) * 8
        R2p = -(-(Rp + 2 * N) // 8) * 8
        xw = xz.new_zeros(nd, R2p, Dn)                                       # (nd, R2p, D)
        for k, w in enumerate(xp_w):
            xw[k, :R] = w[:R]
            xw[k, Rp:Rp + 2 * N] = w[R:]
        dw = xz.new_zeros(nd, Dn, Rp)                                        # (nd, D, Rp)
        for k, w in enumerate(dt_w):
            dw[k, :, :R] = w
        A_all = torch.cat([a.float() for a in A]).contiguous()               # (nd*D, N)
        D_all = torch.cat([d.float() for d in Dp]).contiguous()
        b_all = torch.cat([b.float() for b in dt_b]).contiguous()
        conv_out = causal_conv1d_cuda.causal_conv1d_dirs_fwd(x, cw, cb, dirs, nframes, True)   # (B, nd*D, L)
        co4 = conv_out.view(B_, nd, Dn, L)
        x_dbl = torch.matmul(co4.transpose(-1, -2), xw.transpose(-1, -2))    # (B, nd, L, R2p)
        delta = torch.matmul(dw, x_dbl[..., :Rp].transpose(-1, -2)).view(B_, nd * Dn, L)
        Bv = x_dbl[..., Rp:Rp + N].permute(0, 1, 3, 2)                       # (B, nd, N, L) views, dstate stride 1
        Cv = x_dbl[..., Rp + N:Rp + 2 * N].permute(0, 1, 3, 2)
        _, chk, _, out_z = selective_scan_cuda.fwd(conv_out, delta, A_all, Bv, Cv, D_all, z, b_all, True,
                                                   want_out=False, dirs=dirs, nframes=nframes)
        w3 = (out_wc * scale).repeat(1, nd)                                  # (E, nd*D): scaled sum over the directions
        out = _bmm_shared(out_z.transpose(1, 2), w3.t())                     # (B, L, E)
        if out_b is not None:
            out = out + cast(out_b)
        ctx.dirs, ctx.nframes, ctx.nd, ctx.scale, ctx.rank = dirs, nframes, nd, scale, R
        ctx.has_in_b, ctx.has_out_b, ctx.has_conv_b = in_b is not None, out_b is not None, has_conv_b
        ctx.save_for_backward(hid, xz, x_dbl, chk, out_z, in_wc, out_wc, cw, cb, xw, dw, A_all, D_all, b_all)
        return out

    @staticmethod
    @_bwd_amp
    def backward(ctx, dout):
        hid, xz, x_dbl, chk, out_z, in_wc, out_wc, cw, cb, xw, dw, A_all, D_all, b_all = ctx.saved_tensors
        dirs, nframes, nd, scale = ctx.dirs, ctx.nframes, ctx.nd, ctx.scale
        B_, two_d, L = xz.shape
        Dn = two_d // 2
        R, Rp = ctx.rank, dw.shape[2]
        N = A_all.shape[1]
        E = hid.shape[1]
        x, z = xz[:, :Dn], xz[:, Dn:]
        dout = dout.to(out_z.dtype).contiguous()                             # (B, L, E)
        # ---- out_proj: every direction receives the same upstream gradient g = scale * dout @ W_out, as (B, D, L)
        g = torch.bmm((out_wc.t() * scale).unsqueeze(0).expand(B_, -1, -1), dout.transpose(1, 2))   # (B, D, L)
        d_out_w = torch.bmm(dout.transpose(1, 2), out_z.transpose(1, 2)).sum(0)                  # (E, nd*D)
        d_out_w = d_out_w.view(E, nd, Dn).sum(1) * scale
        d_out_b = dout.sum((0, 1)) if ctx.has_out_b else None
        # ---- recompute (checkpoint_lvl = 1: selective_scan_interface.py:238-241)
        conv_out = causal_conv1d_cuda.causal_conv1d_dirs_fwd(x, cw, cb, dirs, nframes, True)
        co4 = conv_out.view(B_, nd, Dn, L)
        delta = torch.matmul(dw, x_dbl[..., :Rp].transpose(-1, -2)).view(B_, nd * Dn, L)
        Bv = x_dbl[..., Rp:Rp + N].permute(0, 1, 3, 2)
        Cv = x_dbl[..., Rp + N:Rp + 2 * N].permute(0, 1, 3, 2)
        # ---- scan backward: dz of every direction next to dx; dB / dC straight into dx_dbl
        dxz = torch.empty((B_, (1 + nd) * Dn, L), dtype=xz.dtype, device=xz.device)
        dx_dbl = torch.empty_like(x_dbl)
        if dx_dbl.shape[-1] != Rp + 2 * N:
            dx_dbl[..., Rp + 2 * N:] = 0                                     # trailing padding (the dt padding is written below)
        dBv = dx_dbl[..., Rp:Rp + N].permute(0, 1, 3, 2)
        dCv = dx_dbl[..., Rp + N:Rp + 2 * N].permute(0, 1, 3, 2)
        dconv, ddelta, dA, _, _, dD, ddt_b, _ = selective_scan_cuda.bwd(
            conv_out, delta, A_all, Bv, Cv, D_all, z, b_all, g, chk, dxz[:, Dn:], True,
            dirs=dirs, nframes=nframes, dBC_out=(dBv, dCv))
        # ---- dt_proj / x_proj (selective_scan_interface.py:272-277), batched over the directions
        dd4 = ddelta.view(B_, nd, Dn, L)
        d_dw = torch.matmul(dd4, x_dbl[..., :Rp]).sum(0)                                         # (nd, D, Rp)
        dx_dbl[..., :Rp] = torch.matmul(dd4.transpose(-1, -2), dw)                               # (B, nd, L, Rp): zeros beyond R
        d_xw = torch.matmul(dx_dbl.transpose(-1, -2), co4.transpose(-1, -2)).sum(0)              # (nd, R2p, D)
        d_xw = torch.cat([d_xw[:, :R], d_xw[:, Rp:Rp + 2 * N]], dim=1)                           # (nd, R + 2N, D)
        dconv4 = dconv.view(B_ * nd, Dn, L)
        dconv4.baddbmm_(xw.transpose(-1, -2).unsqueeze(0).expand(B_, -1, -1, -1).reshape(B_ * nd, Dn, -1),
                        dx_dbl.view(B_ * nd, L, -1).transpose(-1, -2))                           # += W_x^T dx_dbl^T
        # ---- conv backward: dx summed over the directions, into the first D rows of dxz
        _, d_cw, d_cb = causal_conv1d_cuda.causal_conv1d_dirs_bwd(x, cw, cb, dconv, dxz[:, :Dn], dirs, nframes, True)
        # ---- in_proj backward with the sum over the directions' dz folded into the GEMMs
        w_ext = torch.cat([in_wc[:Dn]] + [in_wc[Dn:]] * nd)                                      # ((1+nd)*D, E)
        d_hidden = _bmm_shared(dxz.transpose(1, 2), w_ext)                                       # (B, L, E)
        d_w_ext = torch.bmm(dxz, hid.view(B_, L, E)).sum(0)                                      # ((1+nd)*D, E)
        d_in_w = torch.cat([d_w_ext[:Dn], d_w_ext[Dn:].view(nd, Dn, E).sum(0)])
        d_in_b = None
        if ctx.has_in_b:
            s = dxz.float().sum((0, 2))
            d_in_b = torch.cat([s[:Dn], s[Dn:].view(nd, Dn).sum(0)])
        K = cw.shape[2]
        grads = ([d_cw[k].view(Dn, 1, K) for k in range(nd)]
                 + [d_cb[k] if ctx.has_conv_b else None for k in range(nd)]
                 + [d_xw[k] for k in range(nd)] + [d_dw[k, :, :R] for k in range(nd)]
                 + [dA[k * Dn:(k + 1) * Dn] for k in range(nd)] + [dD[k * Dn:(k + 1) * Dn] for k in range(nd)]
                 + [ddt_b[k * Dn:(k + 1) * Dn] for k in range(nd)])
        return (d_hidden, d_in_w, d_in_b, d_out_w, d_out_b, None, None, None, None, *grads)


def mamba_dirs_fn(hidden, in_proj_weight, in_proj_bias, out_proj_weight, out_proj_bias, direction_params, dirs,
                  nframes, scale=None):
    """``direction_params``: one tuple (conv_weight, conv_bias, x_proj_weight, dt_proj_weight, A, D, dt_bias) per
    direction; ``dirs``: matching tuple of 'fwd' / 'rev' / 'frames'; ``scale``: factor on the sum of the directions
    before out_proj (default: 1/len(dirs), the mean of v3 -- mamba_simple.py:264; v2 sums, :293)."""
    scale = 1.0 / len(dirs) if scale is None else float(scale)
    nd = len(dirs)
    flat = [p[i] for i in range(7) for p in direction_params]
    return MambaDirsFn.apply(hidden, in_proj_weight, in_proj_bias, out_proj_weight, out_proj_bias,
                             tuple(causal_conv1d_cuda.dir_codes(dirs)), int(nframes), scale, nd, *flat)
