"""Drop-in for ``mamba_ssm.modules.mamba_simple.Mamba`` of the reference fork
(mamba/mamba_ssm/modules/mamba_simple.py:34-353): same constructor signature, same parameter names,
shapes and initialisation (so Lightning / raw state_dicts of the reference load unchanged), same
forward for ``bimamba_type`` "v3" (Vivim), "v2" and "none".

Out of scope, as in SURVEY.md section 2 component 5: the autoregressive ``step`` / inference cache and
``Block``; ``forward(..., inference_params=...)`` raises NotImplementedError.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from .causal_conv1d_interface import causal_conv1d_fn
from .mamba_block import mamba_dirs_fn
from .selective_scan_interface import mamba_inner_fn, mamba_inner_fn_no_out_proj, selective_scan_fn

_DIRECTIONS = {"none": ("",), "v2": ("", "_b"), "v3": ("", "_b", "_s")}
# traversal order of each parameter suffix (mamba_simple.py:217-260): as is, flipped, frame-interleaved
_TRAVERSAL = {"": "fwd", "_b": "rev", "_s": "frames"}
_MAX_DSTATE = 32   # state block of the sm_100a scan kernels (the reference serves up to 256)


class Mamba(nn.Module):
    def __init__(self, d_model, d_state=16, d_conv=4, expand=2, dt_rank="auto", dt_min=0.001,
                 dt_max=0.1, dt_init="random", dt_scale=1.0, dt_init_floor=1e-4, conv_bias=True,
                 bias=False, use_fast_path=True, layer_idx=None, device=None, dtype=None,
                 bimamba_type="none", nframes=5):
        super().__init__()
        if bimamba_type not in _DIRECTIONS:
            raise ValueError(f"bimamba_type must be one of {sorted(_DIRECTIONS)}")
        if d_state > _MAX_DSTATE:
            raise NotImplementedError(f"the B200 scan kernels serve d_state <= {_MAX_DSTATE} (got {d_state}); "
                                      "the reference kernels serve up to 256 (INTEGRATION.md, 'Limits')")
        kw = {"device": device, "dtype": dtype}
        self.d_model, self.d_state, self.d_conv, self.expand = d_model, d_state, d_conv, expand
        self.d_inner = int(expand * d_model)
        self.dt_rank = math.ceil(d_model / 16) if dt_rank == "auto" else dt_rank
        self.use_fast_path = use_fast_path
        self.layer_idx = layer_idx
        self.bimamba_type = bimamba_type
        self.nframes = nframes
        self.activation = "silu"
        self.act = nn.SiLU()

        # Module creation / init calls in the reference's order (mamba_simple.py:68-186), so that the same
        # torch seed gives the same parameters as the reference constructor.
        self.in_proj = nn.Linear(d_model, 2 * self.d_inner, bias=bias, **kw)
        for i, sfx in enumerate(_DIRECTIONS[bimamba_type]):
            self._make_direction(sfx, conv_bias, device, kw)
            if i == 0:
                self._init_dt(dt_init, dt_scale, dt_min, dt_max, dt_init_floor, kw)

        self.out_proj = nn.Linear(self.d_inner, d_model, bias=bias, **kw)

    def _init_dt(self, dt_init, dt_scale, dt_min, dt_max, dt_init_floor, kw):
        """Only the first direction gets the variance-preserving dt init (mamba_simple.py:88-108); dt_proj_b /
        dt_proj_s keep nn.Linear's default init in the reference fork."""
        std = self.dt_rank ** -0.5 * dt_scale
        if dt_init == "constant":
            nn.init.constant_(self.dt_proj.weight, std)
        elif dt_init == "random":
            nn.init.uniform_(self.dt_proj.weight, -std, std)
        else:
            raise NotImplementedError
        dt = torch.exp(torch.rand(self.d_inner, **kw) * (math.log(dt_max) - math.log(dt_min))
                       + math.log(dt_min)).clamp(min=dt_init_floor)
        with torch.no_grad():
            self.dt_proj.bias.copy_(dt + torch.log(-torch.expm1(-dt)))   # softplus^-1(dt)
        self.dt_proj.bias._no_reinit = True

    def _make_direction(self, sfx, conv_bias, device, kw):
        """Parameters of one scan direction: A{sfx}_log, D{sfx}, conv1d{sfx}, x_proj{sfx}, dt_proj{sfx}."""
        di, ns = self.d_inner, self.d_state
        setattr(self, "conv1d" + sfx, nn.Conv1d(di, di, self.d_conv, groups=di, bias=conv_bias,
                                                 padding=self.d_conv - 1, **kw))
        setattr(self, "x_proj" + sfx, nn.Linear(di, self.dt_rank + 2 * ns, bias=False, **kw))
        setattr(self, "dt_proj" + sfx, nn.Linear(self.dt_rank, di, bias=True, **kw))
        # S4D-real: A = -(1..N) for every channel, kept as log in fp32
        a_log = torch.log(torch.arange(1, ns + 1, dtype=torch.float32, device=device)).repeat(di, 1)
        p = nn.Parameter(a_log.contiguous())
        p._no_weight_decay = True
        setattr(self, "A" + sfx + "_log", p)
        d = nn.Parameter(torch.ones(di, device=device))
        d._no_weight_decay = True
        setattr(self, "D" + sfx, d)

    def _scan_direction(self, xz, sfx):
        """mamba_inner_fn_no_out_proj with the parameter set of one direction -> (b, d_inner, l)."""
        conv, x_proj, dt_proj = (getattr(self, n + sfx) for n in ("conv1d", "x_proj", "dt_proj"))
        A = -torch.exp(getattr(self, "A" + sfx + "_log").float())
        return mamba_inner_fn_no_out_proj(
            xz, conv.weight, conv.bias, x_proj.weight, dt_proj.weight, A, None, None,
            getattr(self, "D" + sfx).float(), delta_bias=dt_proj.bias.float(), delta_softplus=True)

    def forward(self, hidden_states, inference_params=None):
        """hidden_states (B, L, d_model) -> (B, L, d_model)"""
        if inference_params is not None:
            raise NotImplementedError("autoregressive decoding (inference_params) is out of scope")
        batch, seqlen, _ = hidden_states.shape
        sfxs = _DIRECTIONS[self.bimamba_type]
        if self.use_fast_path and self.fuse_directions and len(sfxs) > 1 and hidden_states.is_cuda:
            # every direction in one conv launch and one scan launch, no flip / interleave / transpose copies
            return mamba_dirs_fn(hidden_states, self.in_proj.weight, self.in_proj.bias, self.out_proj.weight,
                                 self.out_proj.bias, [self._direction_params(s) for s in sfxs],
                                 tuple(_TRAVERSAL[s] for s in sfxs), self.nframes,
                                 scale=1.0 / 3.0 if self.bimamba_type == "v3" else 1.0)
        # in_proj and the (b l d) -> (b d l) transpose in one GEMM (mamba_simple.py:204-210)
        xz = (self.in_proj.weight @ hidden_states.reshape(batch * seqlen, -1).t()) \
            .view(-1, batch, seqlen).transpose(0, 1)
        if self.in_proj.bias is not None:
            xz = xz + self.in_proj.bias.to(xz.dtype)[:, None]

        if not self.use_fast_path:
            return self._forward_unfused(xz, seqlen)
        if self.bimamba_type == "none":
            return mamba_inner_fn(
                xz, self.conv1d.weight, self.conv1d.bias, self.x_proj.weight, self.dt_proj.weight,
                self.out_proj.weight, self.out_proj.bias, -torch.exp(self.A_log.float()), None, None,
                self.D.float(), delta_bias=self.dt_proj.bias.float(), delta_softplus=True)

        return self._forward_directions_unfused(xz, batch, seqlen)

    fuse_directions = True   # class-level switch (tests compare the two routes)

    def _direction_params(self, sfx):
        conv, x_proj, dt_proj = (getattr(self, n + sfx) for n in ("conv1d", "x_proj", "dt_proj"))
        return (conv.weight, conv.bias, x_proj.weight, dt_proj.weight,
                -torch.exp(getattr(self, "A" + sfx + "_log").float()), getattr(self, "D" + sfx).float(),
                dt_proj.bias.float())

    def _forward_directions_unfused(self, xz, batch, seqlen):
        """The reference's own data flow (mamba_simple.py:217-264): one mamba_inner_fn_no_out_proj per direction on
        materialised flipped / frame-interleaved copies of xz.  Kept as the comparison route of the fused one."""
        y = self._scan_direction(xz, "")                                   # left to right
        y = y + self._scan_direction(xz.flip([-1]), "_b").flip([-1])      # right to left
        if self.bimamba_type == "v3":
            # spatial-major order: tokens (t, hw) -> (hw, t)  (mamba_simple.py:245-264)
            nf = self.nframes
            xz_s = xz.reshape(batch, -1, nf, seqlen // nf).transpose(2, 3).reshape(batch, -1, seqlen)
            y_s = self._scan_direction(xz_s, "_s")
            y = y + y_s.reshape(batch, -1, seqlen // nf, nf).transpose(2, 3).reshape(batch, -1, seqlen)
            y = y / 3
        return F.linear(y.transpose(1, 2), self.out_proj.weight, self.out_proj.bias)

    def _forward_unfused(self, xz, seqlen):
        """use_fast_path=False: the public ops one by one (mamba_simple.py:311-353)."""
        x, z = xz.chunk(2, dim=1)
        x = causal_conv1d_fn(x, self.conv1d.weight.squeeze(1), self.conv1d.bias, self.activation)
        batch = x.shape[0]
        x_dbl = self.x_proj(x.transpose(1, 2).reshape(batch * seqlen, -1))
        dt, B, C = torch.split(x_dbl, [self.dt_rank, self.d_state, self.d_state], dim=-1)
        dt = (self.dt_proj.weight @ dt.t()).view(-1, batch, seqlen).transpose(0, 1)
        B = B.reshape(batch, seqlen, -1).transpose(1, 2).contiguous()
        C = C.reshape(batch, seqlen, -1).transpose(1, 2).contiguous()
        y = selective_scan_fn(x, dt, -torch.exp(self.A_log.float()), B, C, self.D.float(), z=z,
                              delta_bias=self.dt_proj.bias.float(), delta_softplus=True)
        return self.out_proj(y.transpose(1, 2))
