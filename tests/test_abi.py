"""CPU suite, part 2: the C-ABI library loads without a GPU and exports every symbol include/csvit.h declares;
host-side contracts of the drop-in surface (state_dict schema, config, error behaviour)."""
import ctypes
import os
import re

import pytest
import torch

from helpers import ROOT, build_product


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "csvit.h")).read()
    return sorted(set(re.findall(r"CSVIT_API\s+[\w\s\*]+?\b(csvit_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from cs_vit import _lib
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/csvit.h but not exported"
    assert set(_lib.SIGNATURES) | {"csvit_last_error"} == set(names), "ctypes table out of sync with the header"
    assert lib.csvit_abi_version() == 11


def test_errors_are_reported_not_thrown():
    from cs_vit import _lib
    lib = _lib.load()
    buf = (ctypes.c_int32 * 64)()
    assert lib.csvit_host_window_index_map(8, 8, 7, 0, buf) != 0          # 8 not divisible by 7
    assert b"not divisible" in lib.csvit_last_error()
    assert lib.csvit_host_window_index_map(14, 14, 7, 9, buf) != 0         # shift out of range
    with pytest.raises(_lib.CsvitError):
        _lib.check(lib.csvit_host_merge_index_map(7, 7, buf))


def test_no_cpu_fallback():
    from cs_vit import ops
    x = torch.randn(4, 8)
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        ops.layernorm(x, torch.ones(8), torch.zeros(8), 1e-5)
    model, inputs, _, _ = build_product("swint_encoder_patch_spatial")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        model.predict_batch(inputs["patches"], inputs["square_bboxes"], inputs["timestamp"], inputs["focal"], inputs["princpt"])


def test_state_dict_schema_matches_reference_dump():
    """Key names/shapes the reference checkpoints carry; the strict load into the real
    reference model is done in oracle/make_goldens.py - here we pin the representative entries."""
    model, _, _, _ = build_product("swint_encoder_patch_realtime")
    sd = model.state_dict()
    expect = {
        "query_token": (3, 768),
        "J_regressor_mano": (21, 778),
        "backbone.embeddings.patch_embeddings.projection.weight": (96, 3, 4, 4),
        "backbone.encoder.layers.2.blocks.5.attention.self.relative_position_bias_table": (169, 12),
        "backbone.encoder.layers.2.blocks.5.attention.self.relative_position_index": (49, 49),
        "backbone.encoder.layers.0.downsample.reduction.weight": (192, 384),
        "backbone.layernorm.weight": (768,),
        "perspective_mlp.proj.weight": (768, 512),
        "perspective_mlp.layer.0.running_mean": (768,),
        "perspective_mlp.layer.9.weight": (768, 768),
        "spatial_encoder.pe_spatial.positions": (512,),
        "spatial_encoder.pe_spatial.pe.weight": (512, 768),
        "spatial_encoder.layers.5.attn.query.weight": (768, 768),
        "spatial_encoder.layers.5.ffn.net.2.weight": (768, 3072),
        "spatial_encoder.layers.5.norm2.num_batches_tracked": (),
        "pose_temporal_encoder.pe_temporal.inv_freq": (384,),
        "root_temporal_encoder.layers.1.cross_atten.output.bias": (768,),
        "shape_temporal_encoder.zero_conv.weight": (768, 768),
        "pose_decoder.0.weight": (96, 768),
        "shape_decoder.0.bias": (10,),
        "root_decoder.0.weight": (3, 768),
    }
    for k, shape in expect.items():
        assert k in sd, k
        assert tuple(sd[k].shape) == shape, (k, tuple(sd[k].shape))
    assert sd["backbone.encoder.layers.0.blocks.0.attention.self.relative_position_index"].dtype == torch.int64


def test_constructor_contract():
    from cs_vit.net import Poser
    from cs_vit.utils.mano_standin import SyntheticMANO
    from helpers import backbone_dir
    d = backbone_dir("swin_t")
    with pytest.raises(AssertionError):
        Poser(d, spatial_layer_type="mlp", mano_layer=SyntheticMANO())
    with pytest.raises(AssertionError):
        Poser(d, num_latent_layer=2, persp_decorate="query", mano_layer=SyntheticMANO())
    m = Poser(d, mano_layer=SyntheticMANO())
    assert (m.hidden_dim, m.num_heads, m.num_p) == (768, 24, 8)
    assert m.training_phase == Poser.TrainingPhase.INFERENCE and not any(p.requires_grad for p in m.parameters())
    m.phase(Poser.TrainingPhase.SPATIAL)
    assert m.backbone.training and not m.pose_temporal_encoder.training
    assert m.query_token.requires_grad and not next(m.pose_temporal_encoder.parameters()).requires_grad
    m.phase(Poser.TrainingPhase("temporal"))
    assert not m.backbone.training and m.root_temporal_encoder.training and not m.query_token.requires_grad


def test_finetune_config_contract():
    from cs_vit.config import FinetuneConfig, default_finetune_cfg
    cfg = FinetuneConfig()
    assert cfg.img_size == 256 and cfg.spatial_layer_type == "decoder" and default_finetune_cfg.lr == 1e-4
    cfg.update({"img_size": 224, "phase": "spatial"})
    assert cfg.img_size == 224
    with pytest.raises(KeyError):
        cfg.update({"not_a_field": 1})
    with pytest.raises(TypeError):
        cfg.update(3)
    assert '"img_size": 224' in cfg.to_json()


def test_latent_branch_schema_and_option_checks():
    """"ti" configurations (num_latent_layer set, ref:cs_vit/net/ti_poser.py:213-214, 255-265): same option assert and the
    reference's parameter / buffer names (the strict load into the real reference model is done in oracle/make_train_goldens.py)."""
    from helpers import backbone_dir
    from cs_vit.net import Poser
    from cs_vit.utils.mano_standin import SyntheticMANO
    with pytest.raises(AssertionError):
        Poser(backbone_dir("swin_t"), image_size=224, mano_layer=SyntheticMANO(), num_latent_layer=2, persp_decorate="query")
    m = Poser(backbone_dir("swin_t"), image_size=224, mano_layer=SyntheticMANO(), num_latent_layer=2, persp_decorate="patch",
              spatial_layer_type="encoder")
    sd = m.state_dict()
    expect = {
        "latent_trans.rope2d.embedding": (32, 768),
        "latent_trans.rope2d.rot_matrix": (7, 7, 384, 2, 2),
        "latent_trans.rope2d.pos_ceil": (7, 7),
        "latent_trans.rope2d.alpha": (7, 7, 1),
        "latent_trans.sr.1.attn.query.weight": (768, 768),
        "latent_trans.sr.0.norm2.running_var": (768,),
        "latent_trans.scale_embedder.freq_base": (32,),
        "latent_trans.scale_embedder.proj.0.weight": (768, 64),
        "latent_trans.angle_embedder.proj.2.bias": (768,),
        "latent_trans.scale_linear.4.weight": (768, 768),
        "latent_trans.angle_linear.0.bias": (768,),
    }
    for k, shape in expect.items():
        assert k in sd and tuple(sd[k].shape) == shape, (k, tuple(sd[k].shape) if k in sd else None)
    # the reference never re-enables the group after the constructor's INFERENCE phase: frozen and in eval mode while finetuning
    m.phase(Poser.TrainingPhase.SPATIAL)
    assert not any(p.requires_grad for p in m.latent_trans.parameters()) and not m.latent_trans.training
