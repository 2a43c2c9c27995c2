"""CPU statement of the direction-fused ops (csrc/conv1d_dirs.cuh and the `dirs` mode of the scan kernels).
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

A direction is a traversal order of the L tokens of a row (reference: Mamba.forward v3,
mamba/mamba_ssm/modules/mamba_simple.py:217-264):
    fwd     j -> j                                  (:217-229)
    rev     j -> L-1-j                              (xz.flip([-1]), :230-242)
    frames  j -> (j % nf) * (L // nf) + j // nf     (chunk(nf) / stack(-1) / flatten(-2), :243-247)
The fused ops keep every tensor in memory order; here each direction is restated the reference's way -- gather the
tokens into traversal order, run the plain op (oracle.c), scatter the results back -- so the kernels' addressing modes
are checked against explicit copies.
"""
from __future__ import annotations

import numpy as np

from . import conv1d_bwd, conv1d_fwd, scan_bwd, scan_fwd

MODES = {"fwd": 0, "rev": 1, "frames": 2, 0: 0, 1: 1, 2: 2}


def traversal(mode, L, nframes):
    """perm[j] = memory index of the j-th token visited."""
    mode = MODES[mode]
    j = np.arange(L)
    if mode == 0:
        return j
    if mode == 1:
        return L - 1 - j
    assert nframes > 0 and L % nframes == 0
    return (j % nframes) * (L // nframes) + j // nframes


def conv1d_dirs_fwd(x, w, bias, dirs, nframes, silu=True):
    """x (B,D,L), w (nd,D,K), bias (nd,D)|None -> (B, nd*D, L)."""
    x = np.asarray(x, np.float32)
    B, D, L = x.shape
    out = np.empty((B, len(dirs) * D, L), np.float32)
    for k, mode in enumerate(dirs):
        p = traversal(mode, L, nframes)
        out[:, k * D:(k + 1) * D][:, :, p] = conv1d_fwd(x[:, :, p], w[k], None if bias is None else bias[k], silu)
    return out


def conv1d_dirs_bwd(x, w, bias, dout, dirs, nframes, silu=True):
    """-> dx (B,D,L) summed over the directions, dw (nd,D,K), db (nd,D)."""
    x, dout = np.asarray(x, np.float32), np.asarray(dout, np.float32)
    B, D, L = x.shape
    dx = np.zeros((B, D, L), np.float64)
    dw = np.empty(np.shape(w), np.float32)
    db = np.empty((len(dirs), D), np.float32)
    for k, mode in enumerate(dirs):
        p = traversal(mode, L, nframes)
        dxt, dw[k], db[k] = conv1d_bwd(x[:, :, p], w[k], None if bias is None else bias[k],
                                       dout[:, k * D:(k + 1) * D][:, :, p], silu)
        dx[:, :, p] += dxt
    return dx.astype(np.float32), dw, db


def _blocks(dim, groups, dirs):
    nd = len(dirs)
    assert dim % nd == 0 and groups % nd == 0
    return dim // nd, groups // nd


def scan_dirs_fwd(u, delta, A, Bm, Cm, Dv, z, delta_bias, softplus, dirs, nframes):
    """u, delta (B, nd*D, L); Bm, Cm (B, G, N, L) with G % nd == 0; z (B, rows, L) with rows dividing nd*D (shared gate
    rows).  -> dict(out, out_z), memory order."""
    u = np.asarray(u, np.float32)
    B, dim, L = u.shape
    Bm, Cm = np.asarray(Bm, np.float32), np.asarray(Cm, np.float32)
    cb, gb = _blocks(dim, Bm.shape[1], dirs)
    out = np.empty_like(u)
    out_z = np.empty_like(u)
    for k, mode in enumerate(dirs):
        p = traversal(mode, L, nframes)
        ch = np.arange(k * cb, (k + 1) * cb)
        zk = None if z is None else np.asarray(z, np.float32)[:, ch % np.shape(z)[1]][:, :, p]
        r = scan_fwd(u[:, ch][:, :, p], np.asarray(delta, np.float32)[:, ch][:, :, p], A[ch],
                     Bm[:, k * gb:(k + 1) * gb][..., p], Cm[:, k * gb:(k + 1) * gb][..., p],
                     None if Dv is None else Dv[ch], zk, None if delta_bias is None else delta_bias[ch], softplus)
        out[:, k * cb:(k + 1) * cb][:, :, p] = r["out"]
        out_z[:, k * cb:(k + 1) * cb][:, :, p] = r["out_z"]
    return {"out": out, "out_z": out_z}


def scan_dirs_bwd(u, delta, A, Bm, Cm, Dv, z, delta_bias, dout, softplus, dirs, nframes):
    """dout (B, rows, L) shared like z (or (B, nd*D, L)).  -> dict(du, ddelta, dA, dB, dC, dD, ddelta_bias, dz), memory
    order; dz has one row per scanned channel (B, nd*D, L)."""
    u = np.asarray(u, np.float32)
    B, dim, L = u.shape
    Bm, Cm = np.asarray(Bm, np.float32), np.asarray(Cm, np.float32)
    dout = np.asarray(dout, np.float32)
    cb, gb = _blocks(dim, Bm.shape[1], dirs)
    res = {"du": np.empty_like(u), "ddelta": np.empty_like(u), "dA": np.empty(np.shape(A), np.float32),
           "dB": np.empty_like(Bm), "dC": np.empty_like(Cm),
           "dD": None if Dv is None else np.empty(dim, np.float32),
           "ddelta_bias": None if delta_bias is None else np.empty(dim, np.float32),
           "dz": None if z is None else np.empty_like(u)}
    for k, mode in enumerate(dirs):
        p = traversal(mode, L, nframes)
        ch = np.arange(k * cb, (k + 1) * cb)
        zk = None if z is None else np.asarray(z, np.float32)[:, ch % np.shape(z)[1]][:, :, p]
        r = scan_bwd(u[:, ch][:, :, p], np.asarray(delta, np.float32)[:, ch][:, :, p], A[ch],
                     Bm[:, k * gb:(k + 1) * gb][..., p], Cm[:, k * gb:(k + 1) * gb][..., p],
                     None if Dv is None else Dv[ch], zk, None if delta_bias is None else delta_bias[ch],
                     dout[:, ch % dout.shape[1]][:, :, p], softplus)
        for name in ("du", "ddelta", "dz"):
            if res[name] is not None:
                res[name][:, k * cb:(k + 1) * cb][:, :, p] = r[name]
        res["dB"][:, k * gb:(k + 1) * gb][..., p] = r["dB"]
        res["dC"][:, k * gb:(k + 1) * gb][..., p] = r["dC"]
        res["dA"][ch] = r["dA"]
        if res["dD"] is not None:
            res["dD"][ch] = r["dD"]
        if res["ddelta_bias"] is not None:
            res["ddelta_bias"][ch] = r["ddelta_bias"]
    return res
