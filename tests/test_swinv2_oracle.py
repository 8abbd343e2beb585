"""CPU suite for the SwinV2 row: the restatement against the goldens written by the live HF ``Swinv2Model``
(oracle/make_swinv2_goldens.py), the kernels' integer logic at windows 16 / 8 against HF's own maps, and the host-side seam
(state_dict schema, loud failures)."""
import json
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN, rel

with open(os.path.join(GOLDEN, "SWINV2_MANIFEST.json")) as _f:
    _ALL = json.load(_f)["cases"]
V2_CASES = {k: v for k, v in _ALL.items() if v.get("kind") != "gradients"}       # forward goldens
V2_GRAD_CASES = {k: v for k, v in _ALL.items() if v.get("kind") == "gradients"}   # HF-autograd goldens (linear loss)
V2_INTS = dict(np.load(os.path.join(GOLDEN, "swinv2_integer_maps.npz")))


def v2_case(name):
    """(state_dict, pixels, golden dict, case) of a SwinV2 golden case, rebuilt from its seeds."""
    from cs_vit.synthetic import random_swinv2_state_dict
    case = _ALL[name]
    sd = random_swinv2_state_dict(case["variant"], seed=case["weight_seed"])
    g = torch.Generator().manual_seed(case["pixel_seed"])
    px = torch.randn(case["batch"], 3, case["image_size"], case["image_size"], generator=g)
    assert abs(float(px.double().sum()) - case["pixels_sum"]) < 1e-6 * max(1.0, abs(case["pixels_sum"])), "seeded pixels drifted"
    return sd, px, dict(np.load(os.path.join(GOLDEN, name + ".npz"))), case


@pytest.mark.parametrize("name", sorted(V2_CASES))
def test_swinv2_restatement_matches_hf_goldens(name):
    from cs_vit.synthetic import SWINV2_VARIANTS
    from oracle import swinv2_restated as v2

    sd, px, gold, case = v2_case(name)
    _, depths, heads = SWINV2_VARIANTS[case["variant"]]
    with torch.no_grad():
        out, stages = v2.swinv2_forward(px, sd, depths, heads, window=case["window"], return_stages=True)
    assert rel(out, gold["last_hidden_state"]) < 2e-6
    for s, t in enumerate(stages):
        assert rel(t[:, ::case["stage_token_stride"]], gold[f"stage{s}"]) < 2e-6


@pytest.mark.parametrize("name", sorted(V2_GRAD_CASES))
def test_swinv2_restatement_autograd_matches_hf_gradient_goldens(name):
    """Gradients of <features, R> w.r.t. every Swinv2Model parameter, produced by HF's autograd, against torch autograd through
    the restatement: the pin for the SwinV2 backward path (GPU side: test_swinv2_gpu.py::test_swinv2_backward_matches_hf_gradient_golden)."""
    from cs_vit.synthetic import SWINV2_VARIANTS
    from oracle import swinv2_restated as v2
    from oracle.make_train_goldens import projections

    sd, px, gold, case = v2_case(name)
    _, depths, heads = SWINV2_VARIANTS[case["variant"]]
    leaf = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    feats = v2.swinv2_forward(px, leaf, depths, heads, window=case["window"])
    assert rel(feats.detach(), gold["features"]) < 2e-6
    R = torch.randn(feats.shape, generator=torch.Generator().manual_seed(case["projection_seed"]))
    loss = (feats * R).sum()
    assert abs(loss.item() - float(gold["loss"])) < 1e-4 * max(1.0, abs(float(gold["loss"])))
    loss.backward()
    names = [str(n) for n in gold["param_names"]]
    assert set(names) == set(leaf)
    gnorm = float(np.sqrt((gold["grad_norm"] ** 2).sum()))
    for i, n in enumerate(names):
        g = leaf[n].grad
        assert abs(g.double().norm().item() - gold["grad_norm"][i]) < 1e-4 * gold["grad_norm"][i] + 1e-7 * gnorm, n
        assert np.allclose(projections(g, n), gold["grad_proj"][i], rtol=1e-3, atol=1e-6 * gnorm), n
        if "grad/" + n in gold:
            assert rel(g, gold["grad/" + n]) < 1e-4, n


@pytest.mark.parametrize("H,ws,shift", [(64, 16, 0), (64, 16, 8), (32, 16, 8), (16, 16, 0), (8, 8, 0), (64, 8, 4), (32, 8, 4), (16, 8, 4)])
def test_swinv2_integer_maps_kernel_host_code_bit_exact(H, ws, shift):
    from cs_vit import ops
    from oracle import swin_restated as swin

    want = torch.from_numpy(V2_INTS[f"gather_H{H}_w{ws}_s{shift}"])
    assert torch.equal(swin.window_gather_index(H, H, ws, shift).int(), want)
    idx, mask = ops.host_maps(H, H, ws, shift)
    assert torch.equal(idx, want)
    if shift:
        bits = np.packbits((mask != 0).numpy().reshape(-1))
        assert np.array_equal(bits, V2_INTS[f"mask_H{H}_w{ws}_s{shift}"])
        assert set(torch.unique(mask).tolist()) <= {0.0, -100.0}
        assert torch.equal(mask, swin.shift_attention_mask(H, H, ws, shift))
    else:
        assert not mask.any()


@pytest.mark.parametrize("ws", [8, 16])
def test_swinv2_rel_index_and_coords_table(ws):
    from cs_vit import ops
    from cs_vit.net.swinv2_b200 import relative_coords_table
    from oracle import swinv2_restated as v2

    assert torch.equal(ops.host_rel_pos_index(ws), torch.from_numpy(V2_INTS[f"rel_index_w{ws}"]))
    want = torch.from_numpy(V2_INTS[f"coords_table_w{ws}"])
    assert torch.equal(relative_coords_table(ws), want)
    assert torch.equal(v2.relative_coords_table(ws), want)


def test_swinv2_state_dict_schema_matches_hf_and_loads(tmp_path):
    """The product module's keys are exactly Swinv2Model's persistent keys, and an HF-format directory loads."""
    from transformers import Swinv2Config, Swinv2Model

    from cs_vit.net.swinv2_b200 import Swinv2BackboneB200, load_backbone
    from cs_vit.synthetic import make_random_backbone_dir, swinv2_config_dict

    cfg = swinv2_config_dict("swinv2_xs", 256, 16)
    hf = Swinv2Model(Swinv2Config(**{k: v for k, v in cfg.items() if k not in ("architectures", "model_type")}), add_pooling_layer=False)
    d = make_random_backbone_dir(str(tmp_path / "v2"), "swinv2_xs", seed=0, image_size=256, window_size=16)
    ours = load_backbone(d)
    assert isinstance(ours, Swinv2BackboneB200)
    a, b = hf.state_dict(), ours.state_dict()
    assert set(a) == set(b)
    assert all(a[k].shape == b[k].shape for k in a)
    assert [ours.config.stage_geometry(s) for s in range(4)] == [(64, 16, 8), (32, 16, 8), (16, 16, 0), (8, 8, 0)]
    hf.load_state_dict(b, strict=True)


def test_swinv2_no_cpu_fallback(tmp_path):
    from cs_vit.net.swinv2_b200 import load_backbone
    from cs_vit.synthetic import make_random_backbone_dir

    d = make_random_backbone_dir(str(tmp_path / "v2"), "swinv2_xs", seed=0, image_size=256, window_size=16)
    m = load_backbone(d)
    with pytest.raises(RuntimeError, match="CUDA"):
        m.forward_features(torch.zeros(1, 3, 256, 256), normalize=True)
    with open(os.path.join(d, "config.json")) as f:
        cfg = json.load(f)
    cfg["model_type"] = "convnext"
    with open(os.path.join(d, "config.json"), "w") as f:
        json.dump(cfg, f)
    with pytest.raises(NotImplementedError, match="no sm_100a kernels"):
        load_backbone(d)
