"""The caller of the hot path: a Vivim-style video segmenter (SegFormer stages interleaved with Temporal Mamba
stages) built on this repo's ``Mamba`` so that the sm_100a kernels can be measured and parity-tested inside the
whole network on a box where the reference checkout does not exist.

It is NOT part of the product (Vivim keeps its own ``modeling/vivim.py`` and simply imports ``mamba_ssm`` from this
repo); it exists for ``scripts/bench_vivim.py`` and ``tests/test_vivim_model*.py``.  Parameter paths, construction
order and initialisers follow the reference (modeling/vivim.py:57-352) so that (1) the same torch seed yields
bit-identical parameters -- checked on CPU against the reference constructor by ``tests/golden/make_golden_vivim.py``
-- and (2) a reference ``state_dict`` loads unchanged:

    encoder.downsample_layers.*             SegFormer encoder (transformers), stage LayerNorms present but unused
    encoder.stages.{s}.{b}.0.norm1|mamba|norm2|mlp.fc1|mlp.dwconv.dwconv|mlp.fc2
    decoder.*                                SegFormer decode head (its classifier is present but unused)
    out                                      1x1 conv to the classes

Not reproduced: ``from_pretrained`` (no network here: the SegFormer-b3 backbone is built from its config) and the
``with_edge`` head.  The coin flip that applies extra feature dropout to each decoder input in training mode
(modeling/vivim.py:310-312: ``if torch.rand(1).item() > 0.5``, a host synchronisation) is drawn ON THE DEVICE here --
same probability, same dropout rate -- so that the training step stays capturable in a CUDA graph.

``RecallFocusedLoss`` restates the training recipe's loss (multiclass_training_folds.py:217-256, 339-361, 363-425).
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from .layernorm import TokenLayerNorm
from .mamba_simple import Mamba

B3 = dict(num_channels=3, num_encoder_blocks=4, depths=[3, 4, 18, 3], sr_ratios=[8, 4, 2, 1],
          hidden_sizes=[64, 128, 320, 512], patch_sizes=[7, 3, 3, 3], strides=[4, 2, 2, 2],
          num_attention_heads=[1, 2, 5, 8], mlp_ratios=[4, 4, 4, 4], decoder_hidden_size=768, num_labels=150)


def segformer(**overrides):
    """SegFormer-b3 (the checkpoint family modeling/vivim.py:265 downloads), randomly initialised from its config."""
    from transformers import SegformerConfig, SegformerForSemanticSegmentation
    return SegformerForSemanticSegmentation(SegformerConfig(**{**B3, **overrides}))


def _reference_init(module):
    """The initialiser the reference applies to every Temporal Mamba block (modeling/vivim.py:133-146): truncated
    normal (std 0.02) for ALL Linear weights -- including the ones inside Mamba, whose dt_proj bias it zeroes --
    and unit LayerNorms."""
    if isinstance(module, nn.Linear):
        nn.init.trunc_normal_(module.weight, std=0.02)
        if module.bias is not None:
            nn.init.zeros_(module.bias)
    elif isinstance(module, nn.LayerNorm):
        nn.init.zeros_(module.bias)
        nn.init.ones_(module.weight)


class _Depthwise3d(nn.Module):
    def __init__(self, channels):
        super().__init__()
        self.dwconv = nn.Conv3d(channels, channels, kernel_size=3, padding=1, groups=channels)

    def forward(self, tokens, frames, height, width):
        if tokens.is_cuda:      # the sm_100a stencil on the token layout: no transposes, no cuDNN
            from .dwconv3d import dwconv3d_tokens
            return dwconv3d_tokens(tokens, self.dwconv.weight, self.dwconv.bias, frames, height, width)
        # CPU tensors only (used as the comparison side in tests): what the reference module does
        b, _, c = tokens.shape
        vol = tokens.transpose(1, 2).reshape(b, c, frames, height, width)
        return self.dwconv(vol).flatten(2).transpose(1, 2)


class _TokenMlp(nn.Module):
    """fc1 -> depthwise 3x3x3 conv over (frame, y, x) -> GELU -> fc2."""

    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.dwconv = _Depthwise3d(hidden)
        self.act = nn.GELU()
        self.fc2 = nn.Linear(hidden, dim)
        self.apply(_reference_init)

    def forward(self, tokens, frames, height, width):
        return self.fc2(self.act(self.dwconv(self.fc1(tokens), frames, height, width)))


def _stochastic_depth(x, rate, training):
    if rate == 0.0 or not training:
        return x
    keep = 1.0 - rate
    mask = torch.empty(x.shape[0], *([1] * (x.ndim - 1)), device=x.device, dtype=x.dtype).bernoulli_(keep)
    return x * (mask / keep)


class TemporalMambaBlock(nn.Module):
    """(B, C, frames, H, W) -> same: x + Mamba_v3(LN(x)), then x + Mlp(LN(x)), on the flattened token axis."""

    def __init__(self, dim, drop_path=0.0, mlp_ratio=4):
        super().__init__()
        # nn.LayerNorm parameters on the sm_100a kernels; under autocast they hand bf16 straight to the GEMMs that follow
        self.norm1 = TokenLayerNorm(dim)
        self.mamba = Mamba(d_model=dim, d_state=16, d_conv=4, expand=2, bimamba_type="v3")
        self.norm2 = TokenLayerNorm(dim)
        self.norm1.autocast_output = self.norm2.autocast_output = True
        self.mlp = _TokenMlp(dim, int(dim * mlp_ratio))
        self.drop_path_rate = float(drop_path)
        self.apply(_reference_init)

    def forward_tokens(self, tokens, frames, height, width):
        """(B, frames*H*W, C) -> same.  The arithmetic of MambaLayer.forward (modeling/vivim.py:147-159) without its
        (B, C, T, H, W) <-> token transposes: the stage keeps the token layout across its blocks."""
        tokens = tokens + _stochastic_depth(self.mamba(self.norm1(tokens)), self.drop_path_rate, self.training)
        return tokens + _stochastic_depth(self.mlp(self.norm2(tokens), frames, height, width),
                                          self.drop_path_rate, self.training)

    def forward(self, x):
        b, c, frames, height, width = x.shape
        tokens = self.forward_tokens(x.flatten(2).transpose(1, 2), frames, height, width)
        return tokens.transpose(1, 2).reshape(b, c, frames, height, width)


class _Encoder(nn.Module):
    def __init__(self, backbone, dims, depths, drop_path_rate):
        super().__init__()
        self.downsample_layers = backbone.segformer.encoder
        rates = torch.linspace(0, drop_path_rate, sum(depths)).tolist()
        # one rate per STAGE, taken from the head of the per-block schedule, as the reference does (vivim.py:186)
        self.stages = nn.ModuleList(
            nn.Sequential(*(nn.Sequential(TemporalMambaBlock(dim, drop_path=rates[s])) for _ in range(depth)))
            for s, (dim, depth) in enumerate(zip(dims, depths)))

    def forward(self, clips):
        n, frames = clips.shape[:2]
        x = clips.flatten(0, 1)
        seg = self.downsample_layers
        features = []
        for embed, blocks, temporal in zip(seg.patch_embeddings, seg.block, self.stages):
            tokens, height, width = embed(x)
            for blk in blocks:
                tokens = blk(tokens, height, width, False)[0]
            # (n*frames, H*W, C) is already the clip's token sequence (frame-major): a view, where the reference goes
            # through (n, C, frames, H, W) and back around every block (vivim.py:213-215, 147-159)
            seq = tokens.reshape(n, frames * height * width, -1)
            for wrapped in temporal:
                seq = wrapped[0].forward_tokens(seq, frames, height, width)
            x = seq.view(n * frames, height, width, -1).permute(0, 3, 1, 2).contiguous()   # (n*frames, C, H, W)
            features.append(x)
        return features


class VivimSegmenter(nn.Module):
    def __init__(self, in_chans=3, out_chans=3, depths=(2, 2, 2, 2), feat_size=(64, 128, 320, 512),
                 drop_path_rate=0.2, dropout_rate=0.3, backbone=None):
        super().__init__()
        backbone = backbone if backbone is not None else segformer()
        self.dropout_rate = dropout_rate
        self.encoder = _Encoder(backbone, list(feat_size), list(depths), drop_path_rate)
        self.decoder = backbone.decode_head
        self.feature_dropout = nn.Dropout2d(dropout_rate)
        self.out = nn.Conv2d(backbone.config.decoder_hidden_size, out_chans, kernel_size=1)

    def _fuse(self, features):
        head = self.decoder
        size = features[0].shape[2:]
        maps = []
        for fmap, proj in zip(features, head.linear_c):
            n, _, height, width = fmap.shape
            up = proj(fmap).transpose(1, 2).reshape(n, -1, height, width)
            up = F.interpolate(up, size=size, mode="bilinear", align_corners=False)
            if self.training:
                # vivim.py:310-312: with probability 1/2 drop features at rate dropout_rate/2 -- coin drawn on the device
                coin = torch.rand((), device=up.device) > 0.5
                up = torch.where(coin, F.dropout(up, p=self.dropout_rate / 2, training=True), up)
            maps.append(up)
        fused = head.activation(head.batch_norm(head.linear_fuse(torch.cat(maps[::-1], dim=1))))
        fused = head.dropout(head.dropout(fused))
        return self.out(self.feature_dropout(fused))

    def forward(self, clips):
        """clips (n, frames, 3, H, W) -> logits (n*frames, classes, H, W)"""
        logits = self._fuse(self.encoder(clips))
        return F.interpolate(logits, size=clips.shape[-2:], mode="bilinear", align_corners=False)


Vivim = VivimSegmenter


class RecallFocusedLoss(nn.Module):
    """``recall_focused_loss`` of the training recipe (multiclass_training_folds.py:339-361):
    0.4 * class-balanced focal loss (alpha = [0.05, 0.475, 0.475], gamma = 2; :363-425) + 0.6 * Tversky loss
    (alpha 0.3, beta 0.7; :217-256).  logits (N, C, H, W), targets (N, H, W) integer labels.  The class weights live in
    a buffer (the reference builds them with torch.tensor(...).to(device) on every call, a host-to-device copy that a
    CUDA-graph capture cannot contain)."""

    def __init__(self, class_weights=(0.05, 0.475, 0.475), gamma=2.0, tversky_alpha=0.3, tversky_beta=0.7, smooth=1e-6):
        super().__init__()
        self.register_buffer("class_weights", torch.tensor(class_weights, dtype=torch.float32))
        self.gamma, self.ta, self.tb, self.smooth = gamma, tversky_alpha, tversky_beta, smooth

    def forward(self, logits, targets):
        C = logits.shape[1]
        probs = F.softmax(logits.float(), dim=1)
        onehot = F.one_hot(targets.long(), num_classes=C).permute(0, 3, 1, 2).float()
        # Tversky, per class and per image, then averaged (:240-256)
        tp = (probs * onehot).sum(dim=(2, 3))
        fp = (probs * (1 - onehot)).sum(dim=(2, 3))
        fn = ((1 - probs) * onehot).sum(dim=(2, 3))
        tversky = (1 - ((tp + self.smooth) / (tp + self.ta * fp + self.tb * fn + self.smooth)).mean(dim=0)).sum() / C
        # class-balanced focal loss (:405-425)
        weight = onehot * (1 - probs) ** self.gamma + (1 - onehot) * probs ** self.gamma
        bce = -onehot * torch.log(probs + 1e-6) - (1 - onehot) * torch.log(1 - probs + 1e-6)
        focal = (self.class_weights.view(1, C, 1, 1) * weight * bce).mean(dim=(0, 2, 3)).sum()
        return 0.4 * focal + 0.6 * tversky

