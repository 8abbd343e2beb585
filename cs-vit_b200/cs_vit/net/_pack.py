"""Kernel-side views of module parameters (fused / re-typed copies), rebuilt only when a source changes.

The modules keep their parameters in the reference's ``state_dict`` layout so checkpoints load
unchanged; the kernels want Q/K/V stacked into one ``[3C, C]`` matrix, bf16 copies, expanded bias tables and
BatchNorm folded to scale/shift.  ``PackCache`` derives those once and watches ``Tensor._version`` /
``data_ptr`` so an optimizer step, ``load_state_dict`` or ``.to(device)`` invalidates them.
"""
from __future__ import annotations

from typing import Callable, Dict, Sequence, Tuple

import torch


class PackCache:
    def __init__(self) -> None:
        self._store: Dict[str, Tuple[tuple, object]] = {}

    @staticmethod
    def _sig(params: Sequence[torch.Tensor]) -> tuple:
        return tuple((p.data_ptr(), p._version, str(p.device), p.dtype) for p in params)

    def get(self, key: str, params: Sequence[torch.Tensor], build: Callable[[], object]):
        sig = self._sig(params)
        hit = self._store.get(key)
        if hit is None or hit[0] != sig:
            with torch.no_grad():
                hit = (sig, build())
            self._store[key] = hit
        return hit[1]

    def clear(self) -> None:
        self._store.clear()


def fold_batchnorm(bn: torch.nn.BatchNorm1d) -> Tuple[torch.Tensor, torch.Tensor]:
    """Eval-mode ``BatchNorm1d`` as ``y = x * scale + shift``   (ref:cs_vit/net/transformer_module.py:306-307)."""
    scale = bn.weight.float() * torch.rsqrt(bn.running_var.float() + bn.eps)
    shift = bn.bias.float() - bn.running_mean.float() * scale
    return scale.contiguous(), shift.contiguous()
