"""``cs_vit.net`` - same import surface the reference scripts use (ref:cs_vit/net/__init__.py, minus the
pre-training models that no launch script reaches)."""
from .poser import Poser
from .lr_scheduler import warmup_scheduler, gen_cosine_scheduler_array

__all__ = ["Poser", "warmup_scheduler", "gen_cosine_scheduler_array"]
