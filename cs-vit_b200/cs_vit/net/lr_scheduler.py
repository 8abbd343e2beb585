"""Learning-rate schedules imported by the reference scripts (ref:cs_vit/net/lr_scheduler.py).  Host-side."""
import math

import numpy as np
import torch
from torch.optim.lr_scheduler import LambdaLR


def gen_cosine_scheduler_array(base_value, final_value, epochs, niter_per_ep, warmup_epochs=0, start_warmup_value=0):
    """Per-iteration values: linear warm-up then half-cosine to ``final_value``   (ref :9-26)."""
    warm = warmup_epochs * niter_per_ep
    head = np.linspace(start_warmup_value, base_value, warm) if warmup_epochs > 0 else np.array([])
    n = epochs * niter_per_ep - warm
    tail = final_value + 0.5 * (base_value - final_value) * (1 + np.cos(np.pi * np.arange(n) / n))
    out = np.concatenate((head, tail))
    assert len(out) == epochs * niter_per_ep
    return out


def warmup_scheduler(optimizer: torch.optim.Optimizer, max_lr: float, min_lr: float, warmup_epochs: int,
                     annealing_epochs: int, steps_per_epoch: int) -> LambdaLR:
    """Linear warm-up, cosine annealing to ``min_lr``, then constant   (ref :29-60)."""
    assert warmup_epochs >= 0, "warmup_epochs>=0"
    assert annealing_epochs >= 0, "annealing_epochs>=0"
    assert max_lr > min_lr >= 0.0, "max_lr>min_lr>=0.0"
    assert steps_per_epoch > 0
    warm, anneal, floor = warmup_epochs * steps_per_epoch, annealing_epochs * steps_per_epoch, min_lr / max_lr

    def factor(step: int) -> float:
        if step < warm:
            return step / warm
        if step < warm + anneal:
            return floor + (1 - floor) * 0.5 * (1 + math.cos(math.pi * (step - warm) / anneal))
        return floor

    return LambdaLR(optimizer, lr_lambda=factor, last_epoch=-1)
