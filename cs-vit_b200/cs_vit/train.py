"""Finetune-step host logic: the training loop body of ref:scripts/finetune.py:193-289 and its data-parallel gradient
reduction (DDP's bucketed allreduce, ref:scripts/finetune.py:133-135).

``GradReducer`` is the NVLink-era replacement for ``DistributedDataParallel(find_unused_parameters=True)`` on this path:

* gradients live as views into a few large flat fp32 buckets (default 64 MB: on NVSwitch the allreduce cost is launch
  latency + bytes / 900 GB/s, not per-link hops, so few large buckets beat DDP's 25 MB default), so there is no
  flatten / unflatten copy;
* a bucket's ``all_reduce(SUM)`` is launched asynchronously from the post-accumulate-grad hook of its last parameter, i.e.
  overlapped with the rest of the backward pass (buckets are ordered by reverse gradient-ready order, learned on step 1);
* parameters that never receive a gradient (the five discarded "encoder" head layers, quirk Q2; frozen phases) are left
  out of the buckets instead of being searched for on every step (what ``find_unused_parameters=True`` does in the reference);
* ``finish()`` waits for the handles and scales by 1 / world once per bucket.

Works with any ``torch.distributed`` backend: NCCL on the B200 box, gloo in the CPU tests (tests/test_distributed.py).
"""
from __future__ import annotations

import math
from typing import Dict, Iterable, List, Optional

import torch
import torch.distributed as dist


def scaled_lr(base_lr: float, world_size: int, batch_size: int, base_batch: int = 44) -> float:
    """Square-root learning-rate scaling of ref:scripts/finetune.py:138-139."""
    return math.sqrt(world_size * batch_size / base_batch) * base_lr


class GradReducer:
    def __init__(self, params: Iterable[torch.nn.Parameter], bucket_bytes: int = 64 << 20, group=None):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        self.bucket_bytes = bucket_bytes
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self._order: List[int] = []           # parameter indices in gradient-ready order (learned on the first step)
        self._buckets: Optional[List[dict]] = None
        self._bucket_of: Dict[int, int] = {}
        self._handles: List = []
        self._hooks = [p.register_post_accumulate_grad_hook(self._make_hook(i)) for i, p in enumerate(self.params)]

    # ------------------------------------------------------------------------------------------ hooks
    def _make_hook(self, i: int):
        def hook(param):
            if self._buckets is None:
                self._order.append(i)
                return
            b = self._bucket_of.get(i)
            if b is None:
                return                      # did not get a gradient on the first step: not bucketed (reduced in finish())
            bucket = self._buckets[b]
            bucket["pending"] -= 1
            if bucket["pending"] == 0 and self.world > 1:
                self._handles.append(dist.all_reduce(bucket["flat"], op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        return hook

    def _build(self) -> None:
        """After the first backward: bucket the parameters that received gradients, in ready order, and re-point their
        ``.grad`` at views of the flat buffers (the values of this first step are carried over)."""
        buckets, cur, cur_bytes = [], [], 0
        for i in self._order:
            p = self.params[i]
            nbytes = p.numel() * 4
            if cur and (cur_bytes + nbytes > self.bucket_bytes or p.device != self.params[cur[0]].device):
                buckets.append(cur)
                cur, cur_bytes = [], 0
            cur.append(i)
            cur_bytes += nbytes
        if cur:
            buckets.append(cur)
        self._buckets = []
        for b, idxs in enumerate(buckets):
            dev = self.params[idxs[0]].device
            flat = torch.zeros(sum(self.params[i].numel() for i in idxs), dtype=torch.float32, device=dev)
            off = 0
            for i in idxs:
                p = self.params[i]
                view = flat[off:off + p.numel()].view_as(p)
                view.copy_(p.grad)
                p.grad = view
                off += p.numel()
                self._bucket_of[i] = b
            self._buckets.append({"flat": flat, "idxs": idxs, "pending": len(idxs)})

    # ------------------------------------------------------------------------------------------ per-step API
    def zero_grad(self) -> None:
        """Zero the flat buckets (``.grad`` stays a view, so ``optimizer.zero_grad(set_to_none=True)`` must NOT be used)."""
        if self._buckets is None:
            for p in self.params:
                p.grad = None
            return
        for bucket in self._buckets:
            bucket["flat"].zero_()
            bucket["pending"] = len(bucket["idxs"])
        for i, p in enumerate(self.params):
            if i not in self._bucket_of:
                p.grad = None

    def finish(self) -> None:
        """Call after ``loss.backward()``: completes the reduction and leaves averaged gradients in ``param.grad``."""
        first = self._buckets is None
        if first:
            self._build()
            if self.world > 1:
                self._handles = [dist.all_reduce(b["flat"], op=dist.ReduceOp.SUM, group=self.group, async_op=True) for b in self._buckets]
        elif self.world > 1:
            for bucket in self._buckets:     # a bucket whose parameters did not all fire this step is reduced here, late
                if bucket["pending"] > 0:
                    self._handles.append(dist.all_reduce(bucket["flat"], op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        for h in self._handles:
            h.wait()
        self._handles = []
        if self.world > 1:
            for bucket in self._buckets:
                bucket["flat"].mul_(1.0 / self.world)
            stray = [p for i, p in enumerate(self.params) if i not in self._bucket_of and p.grad is not None]
            for p in stray:                   # gradient appeared after step 1 (phase change): reduce it directly
                dist.all_reduce(p.grad, op=dist.ReduceOp.SUM, group=self.group)
                p.grad.mul_(1.0 / self.world)

    def bucket_summary(self) -> List[int]:
        return [b["flat"].numel() * 4 for b in (self._buckets or [])]

    def remove(self) -> None:
        for h in self._hooks:
            h.remove()
        self._hooks = []


def broadcast_parameters(module: torch.nn.Module, src: int = 0, group=None) -> None:
    """DDP-constructor semantics: every rank starts from rank ``src``'s parameters and buffers."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    with torch.no_grad():
        for t in list(module.parameters()) + list(module.buffers()):
            dist.broadcast(t, src=src, group=group)


def finetune_step(model: torch.nn.Module, batch: dict, optimizer: torch.optim.Optimizer, reducer: Optional[GradReducer] = None,
                  max_norm: float = 5.0) -> torch.Tensor:
    """One iteration of ref:scripts/finetune.py:211-227: forward (``Poser.forward`` -> loss), backward, gradient averaging
    across ranks, ``clip_grad_norm_(5.0)``, optimizer step.  Returns the detached loss."""
    if reducer is not None:
        reducer.zero_grad()
    else:
        optimizer.zero_grad(set_to_none=True)
    out = model(batch)
    loss = out["loss"]
    loss.backward()
    if reducer is not None:
        reducer.finish()
    torch.nn.utils.clip_grad_norm_([p for p in model.parameters() if p.grad is not None], max_norm)
    optimizer.step()
    return loss.detach()
