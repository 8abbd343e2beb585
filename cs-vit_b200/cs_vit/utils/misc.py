"""Host-side helpers the reference scripts import (ref:cs_vit/utils/misc.py).  Not on the GPU hot path."""
from functools import partial
from typing import Any, Dict, List, Union

import torch


def move_to_device(data: Union[Dict[str, Any], List[Any], torch.Tensor], device):
    """Recursive in-place move of tensors in dicts/lists   (ref:cs_vit/utils/misc.py:32-43)."""
    if isinstance(data, dict):
        for v in data.values():
            move_to_device(v, device)
    elif isinstance(data, list):
        for v in data:
            move_to_device(v, device)
    elif isinstance(data, torch.Tensor):
        data.data = data.to(device, non_blocking=False)
    return data


def flatten_dict(d, parent_key="", sep="/"):
    """Yield ``("a/b/c", leaf)`` pairs   (ref:cs_vit/utils/misc.py:46-52)."""
    for k, v in d.items():
        key = f"{parent_key}{sep}{k}" if parent_key else k
        if isinstance(v, dict):
            yield from flatten_dict(v, key, sep)
        else:
            yield key, v


def print_with_prefix(*args, prefix: str = "", **kwargs):
    text = kwargs.pop("sep", " ").join(str(a) for a in args)
    print("\n".join(f"{prefix}{line}" for line in text.split("\n")), **kwargs)


def wrap_prefix_print(prefix: str):
    """ref:cs_vit/utils/misc.py:133-134."""
    return partial(print_with_prefix, prefix=prefix)


def print_grouped_losses(epoch, iteration, total_iters, iter_time, lr, forward_result, print_):
    """Console line with the grouped losses of ``Poser.forward`` (ref:cs_vit/utils/misc.py:137-237), uncoloured."""
    scalars = forward_result["logs"]["scalar"]
    head = (f"Epoch {epoch} [{iteration + 1}/{total_iters}] | iter: {iter_time} | "
            f"ETA: {iter_time * (total_iters - iteration - 1)} | lr: {lr:.4e} | Total: {scalars['total']:.6f}")
    lines = []
    for group, items in scalars.items():
        if not isinstance(items, dict):
            continue
        main = f"{group}: {items[group]:.6f}" if group in items else ""
        rest = ", ".join(f"{k}: {v:.6f}" for k, v in items.items() if k != group)
        lines.append("  * " + main + (f" ({rest})" if rest else ""))
    print_(head + "\n" + "\n".join(lines))
