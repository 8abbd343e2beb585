"""Finetune-step host logic: the training loop body of ref:scripts/finetune.py:193-289 and its data-parallel gradient
reduction (DDP's bucketed allreduce, ref:scripts/finetune.py:133-135).

``GradReducer`` is the NVLink-era replacement for ``DistributedDataParallel(find_unused_parameters=True)`` on this path:

* gradients are gathered into a few large flat fp32 buckets (default 64 MB: on NVSwitch the allreduce cost is launch
  latency + bytes / 900 GB/s, not per-link hops, so few large buckets beat DDP's 25 MB default) with one multi-tensor copy
  per bucket, and the optimizer reads them back as views - no per-parameter accumulate or unflatten kernels;
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
            if bucket["pending"] == 0:
                self._flush(bucket)
        return hook

    def _flush(self, bucket: dict) -> None:
        """All gradients of the bucket are ready: gather them into the flat buffer with ONE multi-tensor copy (instead of one
        accumulate kernel per parameter) and start the bucket's allreduce."""
        have = [(v, self.params[i].grad) for v, i in zip(bucket["views"], bucket["idxs"]) if self.params[i].grad is not None]
        if len(have) < len(bucket["idxs"]):
            bucket["flat"].zero_()
        if have:
            torch._foreach_copy_([v for v, _ in have], [g for _, g in have])
        bucket["flushed"] = True
        if self.world > 1:
            self._handles.append(dist.all_reduce(bucket["flat"], op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def _build(self) -> None:
        """After the first backward: bucket the parameters that received gradients, in gradient-ready order."""
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
            views, off = [], 0
            for i in idxs:
                p = self.params[i]
                views.append(flat[off:off + p.numel()].view_as(p))
                off += p.numel()
                self._bucket_of[i] = b
            self._buckets.append({"flat": flat, "idxs": idxs, "views": views, "pending": len(idxs), "flushed": False})

    # ------------------------------------------------------------------------------------------ per-step API
    def zero_grad(self) -> None:
        """Gradients start every step as ``None``: autograd then hands each parameter a fresh tensor (no read-modify-write), which the
        bucket's flush copies into the flat buffer.  (Use this instead of ``optimizer.zero_grad``.)"""
        for p in self.params:
            p.grad = None
        for bucket in self._buckets or []:
            bucket["pending"] = len(bucket["idxs"])
            bucket["flushed"] = False

    def finish(self) -> None:
        """Call after ``loss.backward()``: completes the reduction and leaves the averaged gradients in ``param.grad`` (as views of
        the flat buckets)."""
        if self._buckets is None:
            self._build()
        for bucket in self._buckets:          # first step, or a bucket whose parameters did not all fire this step
            if not bucket["flushed"]:
                self._flush(bucket)
        for h in self._handles:
            h.wait()
        self._handles = []
        for bucket in self._buckets:
            if self.world > 1:
                bucket["flat"].mul_(1.0 / self.world)
            for v, i in zip(bucket["views"], bucket["idxs"]):
                if self.params[i].grad is not None:
                    self.params[i].grad = v
        if self.world > 1:
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


def invalidate_packs(model: torch.nn.Module) -> None:
    """Drop every module's packed (16-bit / stacked) parameter copies.  Needed before EAGER use of a model whose optimizer steps
    ran inside a replayed CUDA graph: replays do not bump ``Tensor._version``, which is what ``PackCache`` watches."""
    for m in model.modules():
        pack = getattr(m, "_pack", None)
        if pack is not None:
            pack.clear()


class GraphedFinetuneStep:
    """The finetune step (zero grads, forward, loss, backward, gradient reduction, clip, optimizer) as ONE CUDA graph.

    At the reference's batch size (32 per GPU) the eager step is bound by the ~2300 host launches it issues (37 ms for 26 ms of
    GPU work on B200).  Everything in the step is stream-ordered device work with static shapes - the kernels of this library,
    torch's fp32 tail, ``clip_grad_norm_`` and a ``capturable`` fused AdamW - so after a few eager warm-up steps (which also
    build the reducer's buckets and the optimizer state) the whole step is captured once and replayed.  Inputs are copied
    into static buffers; the loss and its components are read from static output tensors (no ``.item()`` inside the step:
    ``Poser.loss_tensors``).  The optimizer must have been created with ``capturable=True``.  Call ``invalidate_packs(model)``
    before using the model eagerly again (evaluation between epochs).
    """

    def __init__(self, model: torch.nn.Module, batch: dict, optimizer: torch.optim.Optimizer, reducer: Optional[GradReducer] = None,
                 max_norm: float = 5.0, warmup: int = 3):
        self.model, self.optimizer, self.reducer, self.max_norm = model, optimizer, reducer, max_norm
        self.static = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in batch.items()}
        self.params = [p for p in model.parameters() if p.requires_grad]
        if reducer is None:
            self.reducer = GradReducer(self.params)      # also world size 1: keeps .grad in stable flat buffers for the graph
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(warmup, 2)):
                self._step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss, self.parts = self._step()

    def _step(self):
        self.reducer.zero_grad()
        loss, parts, _ = self.model.loss_tensors(self.static)
        loss.backward()
        self.reducer.finish()
        torch.nn.utils.clip_grad_norm_([p for p in self.params if p.grad is not None], self.max_norm)
        self.optimizer.step()
        return loss.detach(), parts

    def __call__(self, batch: dict) -> torch.Tensor:
        for k, v in batch.items():
            if torch.is_tensor(v) and v.data_ptr() != self.static[k].data_ptr():
                self.static[k].copy_(v, non_blocking=True)
        self.graph.replay()
        return self.loss
