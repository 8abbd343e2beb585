"""Latent scale / rotation consistency branch of the "ti" finetune configurations.

Training-only: when ``Poser(num_latent_layer=k)`` the spatial-encoder batch is doubled with a copy of the backbone patches
that was "scaled and rotated in latent space" by ``ScaleRotComplexEmbedTransformationGroup.do_sr``, the predictions of that
copy are rotated back and supervised with weight 1e-2 (ref:cs_vit/net/ti_poser.py:442-457, 537-557, 827-837).  Module and
parameter names follow ref:cs_vit/net/latent_transformers.py:248-338 and ref:cs_vit/net/transformer_module.py:84-206 so the
reference's checkpoints load; the computation is re-derived on this repo's kernels:

* the ``sr`` stack is ``EncoderBlock``s on the GEMM / attention / BatchNorm kernels (differentiable through cs_vit/autograd.py);
* the two 3-layer MLPs and the frequency embedders' ``Linear -> GELU -> LayerNorm`` run on ``linear`` / ``gelu`` / ``LayerNormFn``;
* the radial embedding + pairwise 2-D rotation of ``RoPE2DPositionalEncoding`` is a position-dependent elementwise map, kept
  as torch ops like the "trope" encoding.

Behaviour kept verbatim because it is observable in checkpoints and outputs: the angle embedding goes through ``scale_linear``
and the scale embedding through ``angle_linear`` (ref :310-311), and ``truncate(l)`` evaluates ``min(1, max(l, num_layers))``.
"""
from __future__ import annotations

import math
from functools import partial
from typing import List, Optional, Union

import torch
import torch.nn as nn

from .. import autograd as ag
from .. import ops
from .blocks import EncoderBlock, _KernelModule


class RoPE2DPositionalEncoding(nn.Module):
    """Learned radial embedding (``num_point`` samples along the normalised distance from the grid centre, linearly
    interpolated) added to every patch, then each channel pair (2k, 2k+1) rotated by ``theta(p, q) * freq_k`` where theta is the
    polar angle of the patch (ref:cs_vit/net/transformer_module.py:84-158)."""

    def __init__(self, embed_dim: int, num_p: int, num_q: int, num_point: int):
        super().__init__()
        self.embed_dim, self.num_p, self.num_q, self.num_point = embed_dim, num_p, num_q, num_point
        self.embedding = nn.Parameter(torch.randn(num_point, embed_dim))
        self.center_p, self.center_q, self.freq_base = (num_p - 1) / 2, (num_q - 1) / 2, 10000.0
        p, q = torch.meshgrid(torch.arange(num_p), torch.arange(num_q), indexing="ij")
        dp, dq = p.float() - self.center_p, q.float() - self.center_q
        radius = torch.sqrt(dp ** 2 + dq ** 2) / math.sqrt(self.center_p ** 2 + self.center_q ** 2)
        coords = radius.clamp(0.0, 1.0) * (num_point - 1)
        half = embed_dim // 2
        freq = 1.0 / (self.freq_base ** (torch.arange(half).float() / half))
        ang = torch.atan2(dq, dp)[..., None] * freq                                # [p, q, D/2]
        c, s = torch.cos(ang)[..., None], torch.sin(ang)[..., None]
        self.register_buffer("sample_coords", coords)
        self.register_buffer("rot_matrix", torch.cat([c, -s, s, c], dim=-1).view(num_p, num_q, half, 2, 2))
        self.register_buffer("pos_floor", torch.floor(coords).long())
        self.register_buffer("pos_ceil", torch.ceil(coords).long())
        self.register_buffer("alpha", (coords - torch.floor(coords))[..., None])

    def forward(self, patches: torch.Tensor) -> torch.Tensor:
        b = patches.shape[0]
        lo = self.embedding[self.pos_floor.clamp(0, self.num_point - 1)]
        hi = self.embedding[self.pos_ceil.clamp(0, self.num_point - 1)]
        x = patches.view(b, self.num_p, self.num_q, self.embed_dim) + (lo * (1 - self.alpha) + hi * self.alpha)[None]
        x0, x1 = x.reshape(b, self.num_p, self.num_q, -1, 2).unbind(-1)
        cos, sin = self.rot_matrix[..., 0, 0], self.rot_matrix[..., 1, 0]
        out = torch.stack([cos * x0 - sin * x1, sin * x0 + cos * x1], dim=-1)
        return out.reshape(b, self.num_p * self.num_q, self.embed_dim)


def _lin(mod: _KernelModule, x: torch.Tensor, lin: nn.Linear, act: int = ops.ACT_NONE) -> torch.Tensor:
    if mod._grad(x):
        return ag.linear(x, lin.weight, lin.bias, act=act, impl=mod._impl)
    return ops.linear(x, lin.weight.detach().float(), lin.bias.detach().float(), act=act, impl=mod._impl)


class ContinuousAngleEmbedding(_KernelModule):
    """sin / cos features of a scalar at ``num_freq`` learnable frequencies -> Linear -> GELU -> LayerNorm
    (ref:cs_vit/net/transformer_module.py:161-206)."""

    def __init__(self, output_dim: int = 64, num_freq: int = 16, learnable_freq: bool = True, max_angle: float = 2 * math.pi,
                 epsilon: float = 1e-6):
        super().__init__()
        self.output_dim, self.num_freq, self.max_angle, self.epsilon = output_dim, num_freq, max_angle, epsilon
        self.freq_base = nn.Parameter(torch.logspace(0, 1, num_freq, base=10).float(), requires_grad=learnable_freq)
        self.proj = nn.Sequential(nn.Linear(2 * num_freq, output_dim), nn.GELU(), nn.LayerNorm(output_dim))

    def forward(self, angles: torch.Tensor) -> torch.Tensor:
        self._check(angles)
        a = (angles % self.max_angle) / self.max_angle * 2 * math.pi
        scaled = a[..., None] * self.freq_base
        feats = torch.cat([torch.sin(scaled), torch.cos(scaled)], dim=-1).reshape(-1, 2 * self.num_freq).float().contiguous()
        lin, ln = self.proj[0], self.proj[2]
        if self._grad(feats):
            y = ag.LayerNormFn.apply(ag.gelu(ag.linear(feats, lin.weight, lin.bias, impl=self._impl)), ln.weight, ln.bias, ln.eps)
        else:
            y = ops.linear(feats, lin.weight.detach().float(), lin.bias.detach().float(), act=ops.ACT_GELU, impl=self._impl)
            y = ops.layernorm(y, ln.weight.detach().float(), ln.bias.detach().float(), ln.eps)
        return y.view(*angles.shape, self.output_dim)


class ScaleRotComplexEmbedTransformationGroup(_KernelModule):
    def __init__(self, num_layers: int = 1, embed_dim: int = 768, num_heads: int = 12, num_p: int = 16, num_q: int = 16):
        super().__init__()
        self.num_layers, self.truncated, self.embed_dim, self.num_heads = num_layers, num_layers, embed_dim, num_heads
        self.rope2d = RoPE2DPositionalEncoding(embed_dim, num_p, num_q, 32)
        self.sr = nn.Sequential(*[EncoderBlock(dim=embed_dim, num_heads=num_heads) for _ in range(num_layers)])

        def mlp():
            return nn.Sequential(nn.Linear(embed_dim, embed_dim), nn.ReLU(), nn.Linear(embed_dim, embed_dim), nn.ReLU(),
                                 nn.Linear(embed_dim, embed_dim))
        self.scale_embedder = ContinuousAngleEmbedding(output_dim=embed_dim, num_freq=32)
        self.scale_linear = mlp()
        self.angle_embedder = ContinuousAngleEmbedding(output_dim=embed_dim, num_freq=32)
        self.angle_linear = mlp()

    def __repr__(self):
        return f"ImageLatentTransformerGroup(num_layer={self.num_layers}, embed_dim={self.embed_dim}, num_heads={self.num_heads})"

    def truncate(self, l: int) -> None:
        self.truncated = min(1, max(l, self.num_layers))

    def _mlp(self, seq: nn.Sequential, x: torch.Tensor) -> torch.Tensor:
        y = _lin(self, x, seq[0], ops.ACT_RELU)
        y = _lin(self, y, seq[2], ops.ACT_RELU)
        return _lin(self, y, seq[4])

    def do_sr(self, patches: torch.Tensor, scale_ratio: Union[torch.Tensor, List, None],
              angle_rad: Union[torch.Tensor, List, None]) -> torch.Tensor:
        """patches ``[n, p*q, D]`` -> same shape, "scaled by ``scale_ratio[n]`` then rotated by ``angle_rad[n]``" in latent space."""
        self._check(patches)
        n = patches.shape[0]

        def as_tensor(v):
            if v is None:
                return torch.zeros(n, device=patches.device, dtype=patches.dtype)
            return torch.tensor(v, device=patches.device, dtype=patches.dtype) if isinstance(v, list) else v
        angle_rad, scale_ratio = as_tensor(angle_rad), as_tensor(scale_ratio)
        x = self.rope2d(patches)
        angle_embeds = self._mlp(self.scale_linear, self.angle_embedder(angle_rad))      # (sic) ref :310
        scale_embeds = self._mlp(self.angle_linear, self.scale_embedder(scale_ratio))    # (sic) ref :311
        x = scale_embeds[:, None] * x + angle_embeds[:, None]
        for layer in self.sr[: self.truncated]:
            x = layer(x)
        return x

    @staticmethod
    def _unwrap_partial(op: partial):
        return op.keywords["scale_ratio"], op.keywords["angle_rad"]

    def compose(self, first_op: partial, second_op: partial) -> partial:
        s1, r1 = self._unwrap_partial(first_op)
        s2, r2 = self._unwrap_partial(second_op)
        return partial(self.do_sr, scale_ratio=s1 * s2, angle_rad=r1 + r2)

    def get_parameterized_sr(self, scale_ratio: Union[torch.Tensor, List], angle_rad: Union[torch.Tensor, List]) -> partial:
        if isinstance(scale_ratio, list):
            scale_ratio = torch.Tensor(scale_ratio)
        if isinstance(angle_rad, list):
            angle_rad = torch.Tensor(angle_rad)
        return partial(self.do_sr, scale_ratio=scale_ratio, angle_rad=angle_rad)
