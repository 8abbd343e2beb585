"""Generate tests/golden/swinv2_*.npz from the LIVE HuggingFace ``Swinv2Model`` and pin ``oracle/swinv2_restated.py`` to it.

Usage:  python -m oracle.make_swinv2_goldens          (build container; needs only ``transformers``, not /root/reference:
the reference's backbone IS ``transformers``' model, ref:cs_vit/net/ti_poser.py:246)

For every case: random weights from ``cs_vit.synthetic.random_swinv2_state_dict`` (seeded CPU generator) are loaded into an
unmodified ``Swinv2Model`` with ``strict=True``, the model runs on seeded pixels, the restatement is asserted against it
(last_hidden_state and every stage output), and HF's outputs are written as golden vectors.  Integer goldens (window
gather map, shift mask, relative-position index for windows 16 and 8) come from HF's own helpers.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, os.path.join(ROOT, "cs-vit_b200"))
sys.path.insert(0, ROOT)

# name -> (variant, image_size, window, batch, pixel seed)
CASES = {
    "swinv2_xs_w16": ("swinv2_xs", 256, 16, 2, 11),     # windows 16/16/16/8, shift 8 at stages 0-1
    "swinv2_xs_w8": ("swinv2_xs", 256, 8, 2, 12),       # windows 8, shift 4 at stages 0-2
    "swinv2_t_w16": ("swinv2_t", 256, 16, 1, 13),       # the shipped Swin-T-sized configuration
}
STAGE_TOKEN_STRIDE = 7     # stage outputs are stored for every 7th token (keeps the fixtures small)


def pixels(batch: int, size: int, seed: int) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    return torch.randn(batch, 3, size, size, generator=g)


def main() -> None:
    from transformers import Swinv2Config, Swinv2Model
    from transformers.models.swinv2 import modeling_swinv2 as hf

    from cs_vit.synthetic import SWINV2_VARIANTS, random_swinv2_state_dict, swinv2_config_dict
    from oracle import swinv2_restated as v2
    from oracle.swin_restated import window_gather_index

    torch.manual_seed(0)
    manifest = {}
    for name, (variant, size, window, batch, seed) in CASES.items():
        cfg = swinv2_config_dict(variant, size, window)
        model = Swinv2Model(Swinv2Config(**{k: v for k, v in cfg.items() if k not in ("architectures", "model_type")}),
                            add_pooling_layer=False).eval()
        sd = random_swinv2_state_dict(variant, seed=0)
        model.load_state_dict(sd, strict=True)
        x = pixels(batch, size, seed)
        stage_out = []
        hooks = [st.register_forward_hook(lambda m, i, o: stage_out.append(o[1].detach())) for st in model.encoder.layers]
        with torch.no_grad():
            want = model(pixel_values=x).last_hidden_state
        for h in hooks:
            h.remove()
        _, depths, heads = SWINV2_VARIANTS[variant]
        with torch.no_grad():
            got, got_stages = v2.swinv2_forward(x, sd, depths, heads, window=window, return_stages=True)
        err = ((got - want).norm() / want.norm()).item()
        serr = [((a - b).norm() / b.norm()).item() for a, b in zip(got_stages, stage_out)]
        print(f"{name}: restatement vs HF Swinv2Model  last {err:.2e}  stages {['%.1e' % e for e in serr]}")
        assert err < 2e-6 and max(serr) < 2e-6, "oracle/swinv2_restated.py does not reproduce HF Swinv2Model"
        out = {"last_hidden_state": want.numpy()}
        for s, t in enumerate(stage_out):
            out[f"stage{s}"] = t[:, ::STAGE_TOKEN_STRIDE].contiguous().numpy()
        np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **out)
        manifest[name] = {"variant": variant, "image_size": size, "window": window, "batch": batch, "pixel_seed": seed,
                          "weight_seed": 0, "restatement_rel_err": err,
                          "pixels_sum": float(x.double().sum()), "stage_token_stride": STAGE_TOKEN_STRIDE}

    # ---- gradient golden: HF autograd through Swinv2Model under a linear loss <features, R> (the pin for the backward path that
    # round 2 builds; the restatement's autograd is asserted against it here and in the CPU suite) -------------------------
    from oracle.make_train_goldens import projections
    for name, (variant, size, window, batch, seed) in {"train_swinv2_xs_w16_linear": ("swinv2_xs", 256, 16, 2, 21)}.items():
        cfg = swinv2_config_dict(variant, size, window)
        model = Swinv2Model(Swinv2Config(**{k: v for k, v in cfg.items() if k not in ("architectures", "model_type")}),
                            add_pooling_layer=False).train()       # dropout / drop-path are 0: train == eval numerically
        sd = random_swinv2_state_dict(variant, seed=0)
        model.load_state_dict(sd, strict=True)
        x = pixels(batch, size, seed)
        feats = model(pixel_values=x).last_hidden_state
        R = torch.randn(feats.shape, generator=torch.Generator().manual_seed(5))
        loss = (feats * R).sum()
        loss.backward()
        _, depths, heads = SWINV2_VARIANTS[variant]
        leaf = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
        got = v2.swinv2_forward(x, leaf, depths, heads, window=window)
        (got * R).sum().backward()
        gold = {"features": feats.detach().numpy().astype(np.float32), "loss": np.array(loss.item(), dtype=np.float64)}
        names, norms, projs, worst = [], [], [], 0.0
        for pname, p in model.named_parameters():
            g = p.grad.detach().float()
            go = leaf[pname].grad
            err = ((go - g).norm() / g.norm().clamp_min(1e-30)).item()
            worst = max(worst, err)
            names.append(pname)
            norms.append(g.double().norm().item())
            projs.append(projections(g, pname))
            if g.numel() <= 4096:
                gold["grad/" + pname] = g.numpy().astype(np.float32)
        assert worst < 1e-4, f"autograd of oracle/swinv2_restated.py deviates from HF's by {worst:.2e}"
        gold["param_names"] = np.array(names)
        gold["grad_norm"] = np.array(norms, dtype=np.float64)
        gold["grad_proj"] = np.stack(projs).astype(np.float64)
        np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **gold)
        print(f"{name}: {len(names)} parameters, global grad norm {float(np.sqrt((gold['grad_norm'] ** 2).sum())):.4e}, "
              f"restatement autograd vs HF autograd worst parameter {worst:.2e}")
        manifest[name] = {"variant": variant, "image_size": size, "window": window, "batch": batch, "pixel_seed": seed, "weight_seed": 0,
                          "projection_seed": 5, "restatement_grad_rel_err": worst, "pixels_sum": float(x.double().sum()),
                          "stage_token_stride": STAGE_TOKEN_STRIDE, "kind": "gradients"}

    # ---- integer goldens from HF's own functions ------------------------------------------------------------------
    ints = {}
    for H, ws, shift in [(64, 16, 0), (64, 16, 8), (32, 16, 8), (16, 16, 0), (8, 8, 0), (64, 8, 4), (32, 8, 4), (16, 8, 4)]:
        ids = torch.arange(H * H).reshape(1, H, H, 1)
        if shift:
            ids = torch.roll(ids, shifts=(-shift, -shift), dims=(1, 2))
        gather = hf.window_partition(ids, ws).reshape(-1)
        assert torch.equal(gather, window_gather_index(H, H, ws, shift))
        ints[f"gather_H{H}_w{ws}_s{shift}"] = gather.numpy().astype(np.int32)
        if shift:
            layer = hf.Swinv2Layer.__new__(hf.Swinv2Layer)
            layer.window_size, layer.shift_size = ws, shift
            mask = hf.Swinv2Layer.get_attn_mask(layer, H, H, torch.float32)
            # stored as the set of windows that carry a mask plus a packed bitmap (the full fp32 mask of H=64/w16 is 4 MB)
            ints[f"mask_H{H}_w{ws}_s{shift}"] = np.packbits((mask != 0).numpy().reshape(-1))
            assert set(torch.unique(mask).tolist()) <= {0.0, -100.0}
    for ws in (8, 16):
        cfgd = Swinv2Config(embed_dim=32, depths=[1], num_heads=[1], window_size=ws, image_size=ws * 4)
        sa = hf.Swinv2SelfAttention(cfgd, 32, 1, ws, [0, 0])
        ints[f"rel_index_w{ws}"] = sa.relative_position_index.numpy().astype(np.int32)
        ints[f"coords_table_w{ws}"] = sa.relative_coords_table.reshape(-1, 2).numpy()
        assert torch.allclose(v2.relative_coords_table(ws), sa.relative_coords_table.reshape(-1, 2), atol=0, rtol=0)
    np.savez_compressed(os.path.join(GOLDEN, "swinv2_integer_maps.npz"), **ints)
    with open(os.path.join(GOLDEN, "SWINV2_MANIFEST.json"), "w") as f:
        json.dump({"generator": "oracle/make_swinv2_goldens.py", "transformers": __import__("transformers").__version__,
                   "torch": torch.__version__, "cases": manifest}, f, indent=1)
    print("wrote", sorted(os.listdir(GOLDEN)))


if __name__ == "__main__":
    main()
