"""Generate tests/golden/*.npz from the LIVE reference and pin the oracle restatement against it.

Runs only in the build container (needs /root/reference).  Usage:  python -m oracle.make_goldens

Two interpreter passes, because the reference package and the product package are both called ``cs_vit``:

  pass 1 (product)    builds this repo's ``Poser`` on CPU for every case from fixed seeds and saves its
                      ``state_dict`` - the weights both sides will use.  (Construction only; the product has
                      no CPU forward.)
  pass 2 (reference)  imports the unmodified reference ``Poser`` (``oracle/ref_import.py``), loads that
                      ``state_dict`` with ``strict=True`` (which also proves the key schema is identical),
                      runs ``predict_batch`` / the HF backbone / HF's integer helpers on the seeded synthetic
                      inputs, asserts that ``oracle/*_restated.py`` reproduces every output to fp32 round-off,
                      and writes the REFERENCE's outputs as golden vectors.

Nothing in tests/ or bench.py needs /root/reference afterwards: weights and inputs are regenerated from the
same seeds (a checksum in each golden file guards against RNG drift) and outputs are compared to the files.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import subprocess
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

# name -> (variant, Poser kwargs, phase, batch, frames)
CASES = {
    # BASELINE.json configs[0]: Swin-T spatial model, batch 2, defaults of spatial_ih26m_swint_noti ("decoder"+"query")
    "swint_decoder_query_spatial": ("swin_t", dict(spatial_layer_type="decoder", persp_decorate="query"), "spatial", 2, 1),
    # the spenc_addpat variant ("encoder"+"patch") named by configs[1] and shipped as *_swint_spenc_addpat_*
    "swint_encoder_patch_spatial": ("swin_t", dict(spatial_layer_type="encoder", persp_decorate="patch"), "spatial", 2, 1),
    # configs[2]: temporal cross-frame attention, realtime supervision (T collapses to 1, quirk Q7)
    "swint_encoder_patch_realtime": ("swin_t", dict(spatial_layer_type="encoder", persp_decorate="patch",
                                                    temporal_supervision="realtime", temporal_init_method="random"), "inference", 2, 4),
    # "full" temporal supervision (EncoderBlock over frames, absolute PE)
    "swint_decoder_query_full": ("swin_t", dict(spatial_layer_type="decoder", persp_decorate="query",
                                                temporal_supervision="full", temporal_init_method="random"), "inference", 1, 3),
    # sparse perspective embedding
    "swint_encoder_query_sparse": ("swin_t", dict(spatial_layer_type="encoder", persp_decorate="query",
                                                  persp_embed_method="sparse"), "spatial", 2, 1),
    # Swin-B slice of configs[1]
    "swinb_encoder_patch_spatial": ("swin_b", dict(spatial_layer_type="encoder", persp_decorate="patch"), "spatial", 1, 1),
    # the same model on 8 images: the golden the batch-256 test of configs[1] embeds into its benched batch
    "swinb_encoder_patch_spatial_b8": ("swin_b", dict(spatial_layer_type="encoder", persp_decorate="patch"), "spatial", 8, 1),
}
OUT_KEYS = ("joint_cam", "verts_cam", "pose_aa", "shape", "root_transl_norm", "root_transl")


def state_checksum(sd) -> str:
    """Digest of a dict of tensors that is exact and independent of reduction order / thread count:
    float tensors are summed as their int32 bit patterns in int64."""
    h = hashlib.sha256()
    for k in sorted(sd):
        v = sd[k].detach().cpu().contiguous()
        bits = v.float().view(torch.int32) if v.is_floating_point() else v
        h.update(k.encode())
        h.update(str(int(bits.to(torch.int64).sum().item())).encode())
        h.update(str(int((bits.to(torch.int64) & 0xFFFF).mul(3).sum().item())).encode())
    return h.hexdigest()[:16]


# ------------------------------------------------------------------------------------------------ pass 1
def selected_cases(only):
    return {k: v for k, v in CASES.items() if not only or k in only}


def pass_product(workdir: str, only=None) -> None:
    sys.path.insert(0, os.path.join(ROOT, "cs-vit_b200"))
    from cs_vit.net import Poser
    from cs_vit.synthetic import make_random_backbone_dir, randomize_head_
    from cs_vit.utils.mano_standin import SyntheticMANO

    for name, (variant, kw, _phase, _b, _t) in selected_cases(only).items():
        bdir = make_random_backbone_dir(os.path.join(workdir, variant), variant, seed=0)
        torch.manual_seed(0)
        m = Poser(bdir, image_size=224, mano_layer=SyntheticMANO(), **kw)
        randomize_head_(m, seed=1)
        torch.save(m.state_dict(), os.path.join(workdir, name + ".sd.pt"))
        print(f"[product] {name}: {len(m.state_dict())} tensors")


# ------------------------------------------------------------------------------------------------ pass 2
def rel(a, b):
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def pass_reference(workdir: str, only=None) -> None:
    sys.path.insert(0, ROOT)
    from oracle import head_restated as head
    from oracle import swin_restated as swin
    from oracle.ref_import import import_reference, load_product_file

    mano_mod = load_product_file("utils/mano_standin.py", "csvit_mano_standin")
    synth = load_product_file("synthetic.py", "csvit_synthetic")
    ref_poser = import_reference(lambda: mano_mod.SyntheticMANO())
    from transformers.models.swin import modeling_swin as hf

    os.makedirs(GOLDEN, exist_ok=True)
    torch.set_grad_enabled(False)

    # ---- integer goldens straight from HF's own functions -------------------------------------------------
    ints = {}
    for H in (56, 28, 14, 7):
        for shift in (0, 3):
            if H == 7 and shift:
                continue
            ids = torch.arange(H * H).reshape(1, H, H, 1)
            rolled = torch.roll(ids, shifts=(-shift, -shift), dims=(1, 2)) if shift else ids
            gather = hf.window_partition(rolled, 7).reshape(-1)
            ints[f"gather_{H}_{shift}"] = gather.numpy().astype(np.int32)
            assert torch.equal(gather, swin.window_gather_index(H, H, 7, shift))
            # inverse direction: window_reverse + roll(+shift) must land every slot back on its source token
            back = hf.window_reverse(gather.reshape(-1, 7, 7, 1), 7, H, H)
            back = torch.roll(back, shifts=(shift, shift), dims=(1, 2)) if shift else back
            assert torch.equal(back.reshape(-1), torch.arange(H * H))
            if shift:
                cfg = hf.SwinConfig(window_size=7)
                layer = hf.SwinLayer(cfg, dim=32, input_resolution=(H, H), num_heads=1, shift_size=shift)
                mask = layer.get_attn_mask(H, H, torch.float32, torch.device("cpu"))
                assert torch.equal(mask, swin.shift_attention_mask(H, H, 7, shift))
                assert set(mask.unique().tolist()) <= {0.0, -100.0}
                ints[f"mask_{H}_{shift}"] = (mask != 0).numpy().astype(np.uint8)
    att = hf.SwinSelfAttention(hf.SwinConfig(window_size=7), dim=32, num_heads=1, window_size=7)
    rel_index = att.relative_position_index if hasattr(att, "relative_position_index") else att.create_relative_position_index()
    assert torch.equal(rel_index, swin.relative_position_index(7))
    ints["rel_index_7"] = rel_index.numpy().astype(np.int32)
    for H in (56, 28, 14):
        # HF's own SwinPatchMerging with its norm and reduction replaced by identities exposes the concat order
        pm = hf.SwinPatchMerging((H, H), dim=1)
        pm.norm, pm.reduction = torch.nn.Identity(), torch.nn.Identity()
        cat = pm(torch.arange(H * H, dtype=torch.float32).reshape(1, H * H, 1), (H, H)).reshape(-1, 4)
        assert torch.equal(cat.long(), swin.merge_gather_index(H, H))
        ints[f"merge_{H}"] = cat.numpy().astype(np.int32)
    if not only:
        np.savez_compressed(os.path.join(GOLDEN, "integer_maps.npz"), **ints)
    print(f"[reference] integer maps: {len(ints)} arrays, restatement bit-exact vs HF")

    # ---- float goldens ------------------------------------------------------------------------------------
    summary = {}
    if only:      # partial regeneration keeps the other cases' manifest entries (and their files) untouched
        with open(os.path.join(GOLDEN, "MANIFEST.json")) as f:
            summary = json.load(f)["cases"]
    for name, (variant, kw, phase, B, T) in selected_cases(only).items():
        sd = torch.load(os.path.join(workdir, name + ".sd.pt"))
        bdir = os.path.join(workdir, variant)
        m = ref_poser.Poser(backbone=bdir, image_size=224, num_latent_layer=None, **kw)
        m.load_state_dict(sd, strict=True)                       # identical key schema or this raises
        m.phase(ref_poser.Poser.TrainingPhase(phase))
        m.eval()
        inp = synth.make_inputs(B, T, 224, seed=11)
        out = m.predict_batch(inp["patches"].clone(), inp["square_bboxes"].clone(), inp["timestamp"].clone(),
                              inp["focal"].clone(), inp["princpt"].clone())
        flat = inp["patches"].reshape(B * T, 3, 224, 224)
        hf_out = m.backbone(m.image_preprocessor(flat), output_hidden_states=True)
        feats = hf_out.last_hidden_state

        # restatement vs live reference
        embed_dim, depths, heads = synth.SWIN_VARIANTS[variant]
        opt = head.HeadOptions(num_heads=heads[-1], depths=depths, swin_heads=heads, phase=phase,
                               **{k: v for k, v in kw.items() if k != "temporal_init_method"})
        mine = head.predict_batch(inp, sd, opt, mano_mod.SyntheticMANO())
        my_feats = head.backbone_features(flat, sd, opt)
        errs = {"features": rel(my_feats, feats)}
        for k in OUT_KEYS:
            assert mine[k].shape == out[k].shape, (name, k, mine[k].shape, out[k].shape)
            errs[k] = rel(mine[k], out[k])
        # execute_all=False (what the product computes) must give the same numbers (quirk Q2)
        lean = head.predict_batch(inp, sd, opt, mano_mod.SyntheticMANO(), execute_all=False)
        errs["lean_joint_cam"] = rel(lean["joint_cam"], out["joint_cam"])
        worst = max(errs.values())
        print(f"[reference] {name}: restatement vs reference max rel err {worst:.2e}  {errs}")
        assert worst < 2e-4, (name, errs)

        gold = {k: out[k].numpy().astype(np.float32) for k in OUT_KEYS}
        gold["features"] = feats.numpy().astype(np.float32)
        # per-stage hidden states (HF tuple: embeddings, then each stage after its downsample): first 8 tokens
        for i, hsi in enumerate(hf_out.hidden_states):
            gold[f"hidden_{i}"] = hsi[:, :8].numpy().astype(np.float32)
        gold["state_checksum"] = np.array(state_checksum(sd))
        gold["input_checksum"] = np.array(state_checksum(inp))
        np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **gold)
        summary[name] = {"variant": variant, "kwargs": kw, "phase": phase, "batch": B, "frames": T,
                         "restatement_vs_reference_rel_err": errs}
    with open(os.path.join(GOLDEN, "MANIFEST.json"), "w") as f:
        json.dump({"generator": "oracle/make_goldens.py", "torch": torch.__version__,
                   "transformers": __import__("transformers").__version__, "input_seed": 11, "weight_seed": 0,
                   "cases": summary}, f, indent=1)


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--stage", choices=["all", "product", "reference"], default="all")
    ap.add_argument("--workdir", default=None)
    ap.add_argument("--only", nargs="*", default=None, help="regenerate only these cases (the others keep their files)")
    a = ap.parse_args()
    if a.stage == "product":
        pass_product(a.workdir, a.only)
    elif a.stage == "reference":
        pass_reference(a.workdir, a.only)
    else:
        with tempfile.TemporaryDirectory() as wd:
            for stage in ("product", "reference"):
                subprocess.run([sys.executable, "-m", "oracle.make_goldens", "--stage", stage, "--workdir", wd] +
                               (["--only"] + a.only if a.only else []), cwd=ROOT, check=True)


if __name__ == "__main__":
    main()
