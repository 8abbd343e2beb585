"""Golden vectors for the finetune step (BASELINE configs[3]) from the LIVE reference.  TEST INFRASTRUCTURE.

Runs only in the build container (needs /root/reference).  Usage:  python -m oracle.make_train_goldens

Same two-pass scheme as ``oracle/make_goldens.py`` (the reference and the product are both packages called ``cs_vit``):
pass 1 builds this repo's ``Poser`` from fixed seeds and saves its ``state_dict``; pass 2 loads it (``strict=True``) into
the unmodified reference ``Poser``, puts it in the training phase exactly as ``scripts/finetune.py`` does
(ref:scripts/finetune.py:129 ``model.phase(cfg.phase)``: train-mode BatchNorm, trainable subset per
ref:cs_vit/net/ti_poser.py:339-397), runs ``predict_batch`` + ``_criterion`` with autograd on seeded synthetic inputs and
labels, calls ``loss.backward()`` (ref:scripts/finetune.py:224) and records, in fp32 on the CPU:

  * the loss and its five components, the six ``predict_batch`` outputs,
  * for EVERY parameter: whether it received a gradient, the gradient's L2 norm, and 4 projections onto seeded +-1 vectors,
  * the full gradient of every parameter with <= 4096 elements (all norm / bias / bias-table / token parameters),
  * BatchNorm running statistics after the step for the layers that ran in train mode.

``Poser.forward`` itself is not callable here (its ``_vis`` reads image files, quirk Q9); ``predict_batch`` + ``_criterion``
are the differentiable part of it (ref:cs_vit/net/ti_poser.py:815-843).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

# name -> (variant, Poser kwargs, training phase, batch, frames)
TRAIN_CASES = {
    "train_swint_encoder_patch_spatial": ("swin_t", dict(spatial_layer_type="encoder", persp_decorate="patch"), "spatial", 4, 1),
    "train_swint_decoder_query_spatial": ("swin_t", dict(spatial_layer_type="decoder", persp_decorate="query"), "spatial", 4, 1),
    "train_swint_encoder_patch_temporal": ("swin_t", dict(spatial_layer_type="encoder", persp_decorate="patch",
                                                          temporal_supervision="realtime", temporal_init_method="random"),
                                           "temporal", 4, 4),
    # "ti" finetune configuration: latent scale / rotation consistency branch (doubles the spatial-encoder batch)
    "train_swint_encoder_patch_spatial_ti": ("swin_t", dict(spatial_layer_type="encoder", persp_decorate="patch", num_latent_layer=2),
                                             "spatial", 4, 1),
}
LATENT_SEED = 777
OUT_KEYS = ("joint_cam", "verts_cam", "pose_aa", "shape", "root_transl_norm", "root_transl")
FULL_GRAD_MAX = 4096
N_PROJ = 4


def projections(g: torch.Tensor, name: str) -> np.ndarray:
    """<g, s_k> for 4 sign vectors drawn from a generator seeded by the parameter name (reproducible on any device)."""
    seed = int.from_bytes(name.encode()[-8:].rjust(8, b"\0"), "little") % (2 ** 31 - 1)
    gen = torch.Generator().manual_seed(seed)
    signs = torch.randint(0, 2, (N_PROJ, g.numel()), generator=gen, dtype=torch.int8).float() * 2 - 1
    return (signs.double() @ g.reshape(-1).double().cpu()).numpy()


def pass_product(workdir: str) -> None:
    sys.path.insert(0, os.path.join(ROOT, "cs-vit_b200"))
    from cs_vit.net import Poser
    from cs_vit.synthetic import make_random_backbone_dir, randomize_head_
    from cs_vit.utils.mano_standin import SyntheticMANO

    for name, (variant, kw, _phase, _b, _t) in TRAIN_CASES.items():
        bdir = make_random_backbone_dir(os.path.join(workdir, variant), variant, seed=0)
        torch.manual_seed(0)
        m = Poser(bdir, image_size=224, mano_layer=SyntheticMANO(), **kw)
        randomize_head_(m, seed=1)
        torch.save(m.state_dict(), os.path.join(workdir, name + ".sd.pt"))


def pass_reference(workdir: str) -> None:
    sys.path.insert(0, ROOT)
    from oracle.make_goldens import state_checksum
    from oracle.ref_import import import_reference, load_product_file

    mano_mod = load_product_file("utils/mano_standin.py", "csvit_mano_standin")
    synth = load_product_file("synthetic.py", "csvit_synthetic")
    ref_poser = import_reference(lambda: mano_mod.SyntheticMANO())
    summary = {}
    for name, (variant, kw, phase, B, T) in TRAIN_CASES.items():
        sd = torch.load(os.path.join(workdir, name + ".sd.pt"))
        m = ref_poser.Poser(backbone=os.path.join(workdir, variant), image_size=224, **{"num_latent_layer": None, **kw})
        m.load_state_dict(sd, strict=True)
        m.phase(ref_poser.Poser.TrainingPhase(phase))
        batch = synth.make_inputs(B, T, 224, seed=11, labels=True)
        torch.manual_seed(LATENT_SEED)      # the reference draws (scale, angle) from the global RNG inside _decode_pose (:442-447)
        predict = m.predict_batch(batch["patches"].clone(), batch["square_bboxes"].clone(), batch["timestamp"].clone(),
                                  batch["focal"].clone(), batch["princpt"].clone())
        if kw.get("num_latent_layer") is not None:      # as Poser.forward does (ref :823-837)
            loss, parts = m._criterion({k: v[:B] for k, v in predict.items()}, batch)
            loss_trans, _ = m._criterion({k: v[B:] for k, v in predict.items()}, batch)
            loss = loss + 1e-2 * loss_trans
        else:
            loss, parts = m._criterion(predict, batch)
        loss.backward()

        gold = {k: predict[k].detach().numpy().astype(np.float32) for k in OUT_KEYS}
        if kw.get("num_latent_layer") is not None:      # the same draws, for the product's test hook (Poser._latent_override)
            torch.manual_seed(LATENT_SEED)
            gold["latent_scale"] = (torch.randn(B).clamp(-0.3, 0.3) + 1.0).numpy()
            gold["latent_angle"] = (torch.rand(B) * 2 * torch.pi).numpy()
        gold["loss"] = np.array(loss.item(), dtype=np.float64)
        gold["loss_parts"] = np.array([parts[k] for k in ("cam", "rel", "shape", "loss_vel", "loss_accel")], dtype=np.float64)
        names, has_grad, norms, projs = [], [], [], []
        n_full = 0
        for pname, p in m.named_parameters():
            names.append(pname)
            has_grad.append(p.grad is not None)
            if p.grad is None:
                norms.append(0.0)
                projs.append(np.zeros(N_PROJ))
                continue
            g = p.grad.detach().float()
            norms.append(g.double().norm().item())
            projs.append(projections(g, pname))
            if g.numel() <= FULL_GRAD_MAX:
                gold["grad/" + pname] = g.numpy().astype(np.float32)
                n_full += 1
        gold["param_names"] = np.array(names)
        gold["param_has_grad"] = np.array(has_grad)
        gold["grad_norm"] = np.array(norms, dtype=np.float64)
        gold["grad_proj"] = np.stack(projs).astype(np.float64)
        after = m.state_dict()
        for k in after:      # running statistics that moved during the step
            if (k.endswith("running_mean") or k.endswith("running_var")) and not torch.equal(after[k], sd[k]):
                gold["bn/" + k] = after[k].numpy().astype(np.float32)
        gold["state_checksum"] = np.array(state_checksum(sd))
        gold["input_checksum"] = np.array(state_checksum(batch))
        np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **gold)
        total = float(np.sqrt((gold["grad_norm"] ** 2).sum()))
        print(f"[reference] {name}: loss {loss.item():.6f}  params with grad {sum(has_grad)}/{len(names)}  "
              f"global grad norm {total:.4e}  full grads stored {n_full}")
        summary[name] = {"variant": variant, "kwargs": kw, "phase": phase, "batch": B, "frames": T, "loss": loss.item(),
                         "params_with_grad": int(sum(has_grad)), "params": len(names), "global_grad_norm": total}
    # ---- backbone alone under a LINEAR loss <features, R>: the upstream gradient is then identical for every precision mode,
    # which isolates the tensor-core backward kernels from the chaotic (sqrt(d)-multiplied softmax, quirk Q1) head.
    for name, variant, B in (("train_backbone_swint_linear", "swin_t", 4),):
        sd = torch.load(os.path.join(workdir, "train_swint_encoder_patch_spatial.sd.pt"))
        m = ref_poser.Poser(backbone=os.path.join(workdir, variant), image_size=224, num_latent_layer=None,
                            spatial_layer_type="encoder", persp_decorate="patch")
        m.load_state_dict(sd, strict=True)
        m.phase(ref_poser.Poser.TrainingPhase.SPATIAL)
        batch = synth.make_inputs(B, 1, 224, seed=11)
        imgs = batch["patches"].reshape(B, 3, 224, 224)
        feats = m.backbone(m.image_preprocessor(imgs)).last_hidden_state          # ref:cs_vit/net/ti_poser.py:425-426
        R = torch.randn(feats.shape, generator=torch.Generator().manual_seed(5))
        (feats * R).sum().backward()
        gold = {"features": feats.detach().numpy().astype(np.float32), "loss": np.array((feats * R).sum().item(), dtype=np.float64)}
        names, norms, projs = [], [], []
        for pname, p in m.backbone.named_parameters():
            names.append(pname)
            g = p.grad.detach().float()
            norms.append(g.double().norm().item())
            projs.append(projections(g, pname))
            if g.numel() <= FULL_GRAD_MAX:
                gold["grad/" + pname] = g.numpy().astype(np.float32)
        gold["param_names"] = np.array(names)
        gold["param_has_grad"] = np.ones(len(names), dtype=bool)
        gold["grad_norm"] = np.array(norms, dtype=np.float64)
        gold["grad_proj"] = np.stack(projs).astype(np.float64)
        gold["state_checksum"] = np.array(state_checksum(sd))
        gold["input_checksum"] = np.array(state_checksum(batch))
        np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **gold)
        total = float(np.sqrt((gold["grad_norm"] ** 2).sum()))
        print(f"[reference] {name}: {len(names)} backbone parameters, global grad norm {total:.4e}")
        summary[name] = {"variant": variant, "kwargs": dict(spatial_layer_type="encoder", persp_decorate="patch"), "phase": "spatial",
                         "batch": B, "frames": 1, "loss": float(gold["loss"]), "params": len(names), "global_grad_norm": total,
                         "linear_loss_seed": 5}
    # ---- the same, with stochastic depth (drop_path_rate 0.1, HF's default and what the reference trains with): HF's own
    # SwinDropPath modules configured as HF's constructor would (HF:swin/modeling_swin.py:543, 716), with the uniform draws of
    # drop_path() (HF:362 torch.rand) served from a seeded queue and recorded, so that the product can replay the same masks.
    from transformers.models.swin import modeling_swin as hf_swin
    for name, variant, B, rate in (("train_backbone_swint_linear_droppath", "swin_t", 4, 0.1),):
        sd = torch.load(os.path.join(workdir, "train_swint_encoder_patch_spatial.sd.pt"))
        m = ref_poser.Poser(backbone=os.path.join(workdir, variant), image_size=224, num_latent_layer=None,
                            spatial_layer_type="encoder", persp_decorate="patch")
        m.load_state_dict(sd, strict=True)
        m.phase(ref_poser.Poser.TrainingPhase.SPATIAL)
        layers = [blk for stage in m.backbone.encoder.layers for blk in stage.blocks]
        dpr = [x.item() for x in torch.linspace(0, rate, len(layers), device="cpu")]              # HF:716
        for blk, p_drop in zip(layers, dpr):
            blk.drop_path = hf_swin.SwinDropPath(p_drop) if p_drop > 0.0 else torch.nn.Identity()    # HF:543
        m.backbone.train()
        batch = synth.make_inputs(B, 1, 224, seed=11)
        imgs = batch["patches"].reshape(B, 3, 224, 224)
        active = [p_drop for p_drop in dpr if p_drop > 0.0]

        def dropped(seed):      # samples dropped over the whole forward with this seed
            g_ = torch.Generator().manual_seed(seed)
            return sum(int((torch.floor(1 - p_drop + torch.rand(B, generator=g_)) == 0).sum()) for p_drop in active)

        seed = next(s_ for s_ in range(77, 500) if dropped(s_) >= 3)      # first seed that exercises the drop branch a few times
        draws, gen, real_rand = [], torch.Generator().manual_seed(seed), torch.rand

        def seeded_rand(*size, **kw):
            shape = tuple(size[0]) if len(size) == 1 and isinstance(size[0], (tuple, list, torch.Size)) else tuple(size)
            assert shape == (B, 1, 1), shape          # only drop_path draws random numbers in this forward
            u = real_rand(B, generator=gen)
            draws.append(u.clone())
            return u.reshape(shape).to(kw.get("dtype", torch.float32))

        torch.rand = seeded_rand
        try:
            feats = m.backbone(m.image_preprocessor(imgs)).last_hidden_state
        finally:
            torch.rand = real_rand
        assert len(draws) == sum(1 for p_drop in dpr if p_drop > 0.0)
        assert sum(int((torch.floor(1 - p_drop + u) == 0).sum()) for p_drop, u in zip(active, draws)) >= 3
        R = torch.randn(feats.shape, generator=torch.Generator().manual_seed(5))
        (feats * R).sum().backward()
        gold = {"features": feats.detach().numpy().astype(np.float32), "loss": np.array((feats * R).sum().item(), dtype=np.float64),
                "droppath_rand": torch.stack(draws).numpy().astype(np.float32)}
        names, norms, projs = [], [], []
        for pname, p in m.backbone.named_parameters():
            names.append(pname)
            g = p.grad.detach().float()
            norms.append(g.double().norm().item())
            projs.append(projections(g, pname))
            if g.numel() <= FULL_GRAD_MAX:
                gold["grad/" + pname] = g.numpy().astype(np.float32)
        gold["param_names"] = np.array(names)
        gold["param_has_grad"] = np.ones(len(names), dtype=bool)
        gold["grad_norm"] = np.array(norms, dtype=np.float64)
        gold["grad_proj"] = np.stack(projs).astype(np.float64)
        gold["state_checksum"] = np.array(state_checksum(sd))
        gold["input_checksum"] = np.array(state_checksum(batch))
        np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **gold)
        total = float(np.sqrt((gold["grad_norm"] ** 2).sum()))
        print(f"[reference] {name}: {len(names)} backbone parameters, global grad norm {total:.4e}, {len(draws)} drop-path draws")
        summary[name] = {"variant": variant, "kwargs": dict(spatial_layer_type="encoder", persp_decorate="patch"), "phase": "spatial",
                         "batch": B, "frames": 1, "loss": float(gold["loss"]), "params": len(names), "global_grad_norm": total,
                         "linear_loss_seed": 5, "drop_path_rate": rate, "drop_path_seed": seed}
    with open(os.path.join(GOLDEN, "TRAIN_MANIFEST.json"), "w") as f:
        json.dump({"generator": "oracle/make_train_goldens.py", "torch": torch.__version__,
                   "transformers": __import__("transformers").__version__, "input_seed": 11, "weight_seed": 0,
                   "cases": summary}, f, indent=1)


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--stage", choices=["all", "product", "reference"], default="all")
    ap.add_argument("--workdir", default=None)
    a = ap.parse_args()
    if a.stage == "product":
        pass_product(a.workdir)
    elif a.stage == "reference":
        pass_reference(a.workdir)
    else:
        with tempfile.TemporaryDirectory() as wd:
            for stage in ("product", "reference"):
                subprocess.run([sys.executable, "-m", "oracle.make_train_goldens", "--stage", stage, "--workdir", wd],
                               cwd=ROOT, check=True)


if __name__ == "__main__":
    main()
