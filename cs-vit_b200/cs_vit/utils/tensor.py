"""ref:cs_vit/utils/tensor.py."""
import torch


def calculate_gradient_norm(model) -> float:
    """Same quantity the reference logs (ref:cs_vit/utils/tensor.py:4-11: half the summed squared L2 norms),
    computed with one device->host transfer instead of one ``.item()`` per parameter."""
    sq = [p.grad.detach().float().pow(2).sum() for p in model.parameters() if p.grad is not None]
    if not sq:
        return 0.0
    return float(torch.stack(sq).sum().item()) * 0.5
