"""Skeleton helpers (ref:cs_vit/utils/joint.py)."""
from typing import List, Tuple

import torch


def mean_connection_length(joints: torch.Tensor, connection: List[Tuple[int, int]]) -> torch.Tensor:
    """Mean bone length over ``connection`` for joints ``(..., J, 3)`` -> ``(...)``   (ref:cs_vit/utils/joint.py:49-70)."""
    a = torch.tensor([i for i, _ in connection], device=joints.device)
    b = torch.tensor([j for _, j in connection], device=joints.device)
    return (joints.index_select(-2, a) - joints.index_select(-2, b)).norm(dim=-1).mean(dim=-1)
