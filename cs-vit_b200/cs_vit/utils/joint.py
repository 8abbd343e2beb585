"""Skeleton helpers (ref:cs_vit/utils/joint.py)."""
from typing import List, Tuple

import torch


_EDGE_CACHE = {}


def _edge_index(connection, device):
    """Index tensors per (skeleton, device), built once: no host->device copy on the hot path (graph-capturable)."""
    key = (connection, str(device))
    if key not in _EDGE_CACHE:
        _EDGE_CACHE[key] = (torch.tensor([i for i, _ in connection], device=device),
                            torch.tensor([j for _, j in connection], device=device))
    return _EDGE_CACHE[key]


def mean_connection_length(joints: torch.Tensor, connection: List[Tuple[int, int]]) -> torch.Tensor:
    """Mean bone length over ``connection`` for joints ``(..., J, 3)`` -> ``(...)``   (ref:cs_vit/utils/joint.py:49-70)."""
    a, b = _edge_index(tuple(connection), joints.device)
    return (joints.index_select(-2, a) - joints.index_select(-2, b)).norm(dim=-1).mean(dim=-1)


def reorder_joints(joints: torch.Tensor, origin, target) -> torch.Tensor:
    """``[..., J, D]`` joints from the ``origin`` naming order to the ``target`` one   (ref:cs_vit/utils/joint.py:8-46: same
    errors for non-lists, different lengths or different name sets)."""
    if not isinstance(origin, (list, tuple)) or not isinstance(target, (list, tuple)):
        raise TypeError("Joint orders must be lists/tuples")
    if len(origin) != len(target):
        raise ValueError("Origin and target joint lists must have same length")
    if set(origin) != set(target):
        raise ValueError("Origin and target joint lists must contain same joints")
    where = {name: i for i, name in enumerate(origin)}
    index = torch.tensor([where[name] for name in target], dtype=torch.int64, device=joints.device)
    return torch.index_select(joints, -2, index)
