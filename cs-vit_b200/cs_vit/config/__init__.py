from .base import FinetuneConfig, default_finetune_cfg

__all__ = ["FinetuneConfig", "default_finetune_cfg"]
