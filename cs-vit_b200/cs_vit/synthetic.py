"""Seeded synthetic inputs and random-init backbone directories.

Nothing here exists in the reference: it publishes no benchmark inputs and ships no weights.  The recipe is
the one fixed in SURVEY.md §8(d) so that the CUDA path, the CPU oracle and the committed golden vectors all
see bit-identical tensors: every value comes from a CPU ``torch.Generator`` and is moved afterwards.
"""
from __future__ import annotations

import json
import os
from typing import Dict

import torch

SWIN_VARIANTS = {
    # name: (embed_dim, depths, num_heads)      HF:swin/configuration_swin.py defaults for the rest
    "swin_t": (96, (2, 2, 6, 2), (3, 6, 12, 24)),
    "swin_b": (128, (2, 2, 18, 2), (4, 8, 16, 32)),
    # two-block toy used by fast unit tests (same code paths: shift, mask, merge, 4 stages)
    "swin_xs": (32, (2, 2, 2, 2), (1, 2, 4, 8)),
}


def swin_config_dict(variant: str, image_size: int = 224, window_size: int = 7) -> Dict:
    embed_dim, depths, heads = SWIN_VARIANTS[variant]
    return {
        "architectures": ["SwinModel"],
        "model_type": "swin",
        "image_size": image_size,
        "patch_size": 4,
        "num_channels": 3,
        "embed_dim": embed_dim,
        "depths": list(depths),
        "num_heads": list(heads),
        "window_size": window_size,
        "mlp_ratio": 4.0,
        "qkv_bias": True,
        "hidden_dropout_prob": 0.0,
        "attention_probs_dropout_prob": 0.0,
        "drop_path_rate": 0.0,
        "hidden_act": "gelu",
        "use_absolute_embeddings": False,
        "layer_norm_eps": 1e-5,
        "initializer_range": 0.02,
        "encoder_stride": 32,
        "hidden_size": int(embed_dim * 2 ** (len(depths) - 1)),
        "num_layers": len(depths),
        "out_features": None,
        "out_indices": None,
    }


def _trunc_normal(shape, std, g):
    t = torch.empty(shape)
    torch.nn.init.trunc_normal_(t, mean=0.0, std=std, a=-2 * std, b=2 * std, generator=g)
    return t


def random_swin_state_dict(variant: str, seed: int = 0, window_size: int = 7,
                           bias_table_std: float = 0.02, ln_jitter: float = 0.1) -> Dict[str, torch.Tensor]:
    """Random Swin-v1 weights under the HF ``SwinModel`` key schema (SURVEY.md §8b ``state_dict`` row).

    Follows HF's initialiser (trunc-normal 0.02 weights, zero biases) with three deliberate departures so
    that every term of the kernels is exercised: the relative-position-bias tables are
    randomised instead of zero, biases are small-random instead of zero, and LayerNorm affine parameters are
    jittered around (1, 0).
    """
    embed_dim, depths, heads = SWIN_VARIANTS[variant]
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}

    def linear(prefix, out_f, in_f, bias=True):
        sd[prefix + ".weight"] = _trunc_normal((out_f, in_f), 0.02, g)
        if bias:
            sd[prefix + ".bias"] = torch.randn(out_f, generator=g) * 0.02

    def norm(prefix, dim):
        sd[prefix + ".weight"] = 1.0 + ln_jitter * torch.randn(dim, generator=g)
        sd[prefix + ".bias"] = ln_jitter * torch.randn(dim, generator=g)

    sd["embeddings.patch_embeddings.projection.weight"] = _trunc_normal((embed_dim, 3, 4, 4), 0.02, g)
    sd["embeddings.patch_embeddings.projection.bias"] = torch.randn(embed_dim, generator=g) * 0.02
    norm("embeddings.norm", embed_dim)
    ws = window_size
    coords = torch.stack(torch.meshgrid(torch.arange(ws), torch.arange(ws), indexing="ij")).flatten(1)
    rel = coords[:, :, None] - coords[:, None, :]
    rel_index = ((rel[0] + ws - 1) * (2 * ws - 1) + (rel[1] + ws - 1)).to(torch.int64)  # HF:swin/modeling_swin.py:461-473
    for s, (depth, h) in enumerate(zip(depths, heads)):
        c = embed_dim * 2 ** s
        for i in range(depth):
            p = f"encoder.layers.{s}.blocks.{i}"
            norm(p + ".layernorm_before", c)
            sd[p + ".attention.self.relative_position_bias_table"] = _trunc_normal(((2 * ws - 1) ** 2, h), bias_table_std, g)
            sd[p + ".attention.self.relative_position_index"] = rel_index.clone()
            linear(p + ".attention.self.query", c, c)
            linear(p + ".attention.self.key", c, c)
            linear(p + ".attention.self.value", c, c)
            linear(p + ".attention.output.dense", c, c)
            norm(p + ".layernorm_after", c)
            linear(p + ".intermediate.dense", 4 * c, c)
            linear(p + ".output.dense", c, 4 * c)
        if s < len(depths) - 1:
            linear(f"encoder.layers.{s}.downsample.reduction", 2 * c, 4 * c, bias=False)
            norm(f"encoder.layers.{s}.downsample.norm", 4 * c)
    norm("layernorm", embed_dim * 2 ** (len(depths) - 1))
    return sd


SWINV2_VARIANTS = {
    # name: (embed_dim, depths, num_heads)      HF:swinv2/configuration_swinv2.py; shipped CS-ViT configs use window 16 @ 256
    "swinv2_t": (96, (2, 2, 6, 2), (3, 6, 12, 24)),
    "swinv2_b": (128, (2, 2, 18, 2), (4, 8, 16, 32)),
    "swinv2_xs": (32, (2, 2, 2, 2), (1, 2, 4, 8)),
}


def swinv2_config_dict(variant: str, image_size: int = 256, window_size: int = 16) -> Dict:
    embed_dim, depths, heads = SWINV2_VARIANTS[variant]
    d = {
        "architectures": ["Swinv2Model"],
        "model_type": "swinv2",
        "image_size": image_size,
        "patch_size": 4,
        "num_channels": 3,
        "embed_dim": embed_dim,
        "depths": list(depths),
        "num_heads": list(heads),
        "window_size": window_size,
        "pretrained_window_sizes": [0, 0, 0, 0],
        "mlp_ratio": 4.0,
        "qkv_bias": True,
        "hidden_dropout_prob": 0.0,
        "attention_probs_dropout_prob": 0.0,
        "drop_path_rate": 0.0,
        "hidden_act": "gelu",
        "use_absolute_embeddings": False,
        "layer_norm_eps": 1e-5,
        "initializer_range": 0.02,
        "encoder_stride": 32,
        "hidden_size": int(embed_dim * 2 ** (len(depths) - 1)),
        "num_layers": len(depths),
        "out_features": None,
        "out_indices": None,
    }
    return d


def random_swinv2_state_dict(variant: str, seed: int = 0, ln_jitter: float = 0.1, logit_scale_jitter: float = 0.5) -> Dict[str, torch.Tensor]:
    """Random SwinV2 weights under the HF ``Swinv2Model`` key schema (persistent entries only: the coordinate table and the
    relative-position index are non-persistent buffers, V2:408-410).  Same departures from HF's initialiser as
    ``random_swin_state_dict`` (small random biases, jittered LayerNorm affines), plus per-head jitter on ``logit_scale``
    around HF's ln(10) so that the per-head scale path is exercised."""
    embed_dim, depths, heads = SWINV2_VARIANTS[variant]
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}

    def linear(prefix, out_f, in_f, bias=True, std=0.02):
        sd[prefix + ".weight"] = _trunc_normal((out_f, in_f), std, g)
        if bias:
            sd[prefix + ".bias"] = torch.randn(out_f, generator=g) * 0.02

    def norm(prefix, dim):
        sd[prefix + ".weight"] = 1.0 + ln_jitter * torch.randn(dim, generator=g)
        sd[prefix + ".bias"] = ln_jitter * torch.randn(dim, generator=g)

    sd["embeddings.patch_embeddings.projection.weight"] = _trunc_normal((embed_dim, 3, 4, 4), 0.02, g)
    sd["embeddings.patch_embeddings.projection.bias"] = torch.randn(embed_dim, generator=g) * 0.02
    norm("embeddings.norm", embed_dim)
    for s, (depth, h) in enumerate(zip(depths, heads)):
        c = embed_dim * 2 ** s
        for i in range(depth):
            p = f"encoder.layers.{s}.blocks.{i}"
            a = p + ".attention.self"
            sd[a + ".logit_scale"] = (torch.log(torch.tensor(10.0)) + logit_scale_jitter * torch.randn(h, 1, 1, generator=g))
            linear(a + ".continuous_position_bias_mlp.0", 512, 2, std=0.5)      # wide enough that the bias varies over the window
            sd[a + ".continuous_position_bias_mlp.2.weight"] = _trunc_normal((h, 512), 0.1, g)
            linear(a + ".query", c, c)
            linear(a + ".key", c, c, bias=False)
            linear(a + ".value", c, c)
            linear(p + ".attention.output.dense", c, c)
            norm(p + ".layernorm_before", c)
            linear(p + ".intermediate.dense", 4 * c, c)
            linear(p + ".output.dense", c, 4 * c)
            norm(p + ".layernorm_after", c)
        if s < len(depths) - 1:
            linear(f"encoder.layers.{s}.downsample.reduction", 2 * c, 4 * c, bias=False)
            norm(f"encoder.layers.{s}.downsample.norm", 2 * c)
    norm("layernorm", embed_dim * 2 ** (len(depths) - 1))
    return sd


def make_random_backbone_dir(path: str, variant: str = "swin_t", seed: int = 0, image_size: int = 224, window_size: int = 0) -> str:
    """Write ``config.json`` + ``model.safetensors`` that both this repo and HF ``AutoModel`` can load."""
    from safetensors.torch import save_file

    os.makedirs(path, exist_ok=True)
    v2 = variant in SWINV2_VARIANTS
    with open(os.path.join(path, "config.json"), "w") as f:
        json.dump(swinv2_config_dict(variant, image_size, window_size or 16) if v2 else swin_config_dict(variant, image_size, window_size or 7),
                  f, indent=2)
    sd = random_swinv2_state_dict(variant, seed) if v2 else random_swin_state_dict(variant, seed, window_size or 7)
    sd = {k: v.contiguous() for k, v in sd.items()}
    save_file(sd, os.path.join(path, "model.safetensors"), metadata={"format": "pt"})
    return path


def make_inputs(batch: int, frames: int = 1, image_size: int = 224, seed: int = 0, labels: bool = False):
    """Synthetic batch per SURVEY.md §8(d): DexYCB/HO3D-like intrinsics, U[0,1) crops, 30 fps timestamps."""
    g = torch.Generator().manual_seed(seed)
    B, T, S = batch, frames, image_size
    patches = torch.rand(B, T, 3, S, S, generator=g)
    cx = 150.0 + 300.0 * torch.rand(B, T, generator=g)
    cy = 120.0 + 200.0 * torch.rand(B, T, generator=g)
    half = 60.0 + 60.0 * torch.rand(B, T, generator=g)
    out = {
        "patches": patches,
        "square_bboxes": torch.stack([cx - half, cy - half, cx + half, cy + half], dim=-1),
        "timestamp": (torch.arange(T, dtype=torch.float32) * 33.333)[None].repeat(B, 1),
        "focal": torch.tensor([617.0, 617.0]).expand(B, T, 2).contiguous(),
        "princpt": torch.tensor([312.0, 241.0]).expand(B, T, 2).contiguous(),
    }
    if labels:
        out["joint_cam"] = torch.randn(B, T, 21, 3, generator=g) * 30.0 + torch.tensor([0.0, 0.0, 500.0])
        out["joint_valid"] = torch.ones(B, T, 21)
        out["mano_shape"] = torch.randn(B, T, 10, generator=g) * 0.5
    return out


def randomize_head_(model: torch.nn.Module, seed: int = 1) -> None:
    """Give BatchNorm running stats / affine parameters non-trivial values (SURVEY.md App. A last line).

    Fresh ``BatchNorm1d`` layers carry (mean 0, var 1, γ 1, β 0), which would make the eval-mode
    normalisation an identity and hide errors in the folded scale/shift path.
    """
    g = torch.Generator().manual_seed(seed)
    for m in model.modules():
        if isinstance(m, torch.nn.BatchNorm1d):
            with torch.no_grad():
                m.running_mean.copy_(0.2 * torch.randn(m.num_features, generator=g))
                m.running_var.copy_(0.5 + torch.rand(m.num_features, generator=g))
                m.weight.copy_(1.0 + 0.1 * torch.randn(m.num_features, generator=g))
                m.bias.copy_(0.1 * torch.randn(m.num_features, generator=g))
