"""Experiment configuration, field-compatible with ref:cs_vit/config/base.py so the reference's
``checkpoints/<exp>/config.json`` files load unchanged (``FinetuneConfig(**json)``, ref:scripts/eval.py:336-340)."""
import json
from dataclasses import asdict, dataclass, field
from typing import Any, Dict, List, Optional, Union


@dataclass
class FinetuneConfig:
    # experiment
    exp: Optional[str] = None
    epoch: int = 30
    # model
    backbone: Optional[str] = None
    num_joints: int = 16
    num_spatial_layer: int = 6
    global_positioning: str = "direct"
    spatial_layer_type: str = "decoder"
    num_temporal_layer: int = 2
    temporal_init_method: str = "zero"
    img_size: int = 256
    expansion_ratio: float = 1.25
    trope_scalar: float = 20.0
    num_latent_layer: Optional[int] = None
    persp_embed_method: str = "dense"
    persp_decorate: str = "query"
    # data
    data: Optional[List[str]] = None
    seq_len: Optional[int] = None
    batch_size: Optional[int] = None
    ih26mseq_root: str = "/data_1/datasets_temp/InterHand2.6M_5fps_batch1"
    ho3d_root: str = "/data_1/datasets_temp/HO3D_v3"
    dexycb_root: str = "/data_1/datasets_temp/dexycb"
    # training
    phase: str = "inference"
    temporal_supervision: str = "full"
    spatial_ckpt: Optional[str] = None
    lr: float = 1e-4
    lr_min: float = 1e-6
    lr_scheduler: Optional[str] = None
    warmup_epoch: int = 1
    cooldown_epoch: int = 10
    # evaluation
    eval_ckpt: Optional[str] = None

    def update(self, other: Union["FinetuneConfig", Dict[str, Any]]) -> None:
        """Merge; unknown keys raise ``KeyError`` as in the reference (ref:cs_vit/config/base.py:50-62)."""
        if isinstance(other, FinetuneConfig):
            other = other.to_dict()
        elif not isinstance(other, dict):
            raise TypeError("can only merge from Config/dict")
        for key, value in other.items():
            if not hasattr(self, key):
                raise KeyError(f"Unexpected key: {key}.")
            setattr(self, key, value)

    def to_dict(self) -> Dict[str, Any]:
        return asdict(self)

    def to_json(self) -> str:
        return json.dumps(self.to_dict(), ensure_ascii=False, indent=4)


default_finetune_cfg = FinetuneConfig()
