"""Data-parallel plumbing for the hot path: batch sharding and the coalesced eval-result gather.

The reference is single-node DDP only.  In inference every image / clip is independent, so
the forward needs NO collective; what remains is (a) giving each rank its slice of a batch and (b) bringing
the per-sample results back to rank 0.  The reference does (b) with five collectives and a barrier per batch
(``gather_strings`` = all_gather of padded uint8 paths, four ``dist.gather`` calls, ``dist.barrier``;
ref:scripts/eval.py:53-82, 289-292, 315-317).  ``gather_eval_results`` packs everything into ONE
``all_gather_into_tensor`` of a fixed-width row per sample, which is what NVLink/NVSwitch wants: one launch,
one 1-2 KB row per sample, no host sync besides the final copy.

Works on any ``torch.distributed`` backend (NCCL on the B200 box, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

RESULT_FIELDS: Tuple[Tuple[str, int], ...] = (          # ref:scripts/eval.py:289-292 payloads, per sample
    ("joint_cam_gt", 63), ("joint_cam_pred", 63), ("joint_reproj_gt", 42), ("joint_reproj_pred", 42))
PATH_BYTES = 256                                          # fixed-width, zero-padded utf-8


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, near-even ``[lo, hi)`` share of ``n`` items for ``rank`` (first ``n % world`` ranks get one more)."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(batch: Dict[str, object], rank: int, world: int) -> Dict[str, object]:
    """Slice every per-sample entry of a batch dict along dim 0; clips stay whole because the
    frame axis is dim 1.  Lists (``imgs_path``, ``flip``) are sliced too."""
    n = batch["patches"].shape[0]
    lo, hi = shard_range(n, rank, world)
    out = {}
    for k, v in batch.items():
        if isinstance(v, torch.Tensor) and v.dim() > 0 and v.shape[0] == n:
            out[k] = v[lo:hi]
        elif isinstance(v, (list, tuple)) and len(v) == n:
            out[k] = v[lo:hi]
        else:
            out[k] = v
    return out


def _encode_paths(paths: Sequence[str], device) -> torch.Tensor:
    buf = torch.zeros(len(paths), PATH_BYTES, dtype=torch.uint8)
    for i, p in enumerate(paths):
        b = p.encode("utf-8")
        if len(b) > PATH_BYTES:      # the reference pads to the longest path (ref:scripts/eval.py:52-58); never truncate silently
            raise ValueError(f"image path longer than {PATH_BYTES} bytes cannot be gathered: {p!r}")
        buf[i, : len(b)] = torch.tensor(list(b), dtype=torch.uint8)
    return buf.to(device)


def _decode_paths(buf: torch.Tensor) -> List[str]:
    return [bytes(row[row != 0].tolist()).decode("utf-8", "replace") for row in buf.cpu()]


def gather_eval_results(results: Dict[str, torch.Tensor], paths: Sequence[str], max_local: Optional[int] = None,
                        group=None) -> Optional[Tuple[Dict[str, torch.Tensor], List[str]]]:
    """One collective per batch.  ``results[name]``: ``[b_local, J, d]`` fp32 for the four RESULT_FIELDS.

    Every rank contributes ``max_local`` rows (its ``b_local`` valid ones + padding; default: all ranks have the
    same ``b_local``), each row = 210 floats of results + 64 floats carrying the 256 path bytes + 1 validity
    flag.  Rank 0 returns ``(dict of [B_total, J, d] tensors, paths)`` in rank order; other ranks return None.
    """
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    first = results[RESULT_FIELDS[0][0]]
    b_local, device = first.shape[0], first.device
    rows = max_local if max_local is not None else b_local
    width = sum(w for _, w in RESULT_FIELDS) + PATH_BYTES // 4 + 1
    packed = torch.zeros(rows, width, dtype=torch.float32, device=device)
    col = 0
    for name, w in RESULT_FIELDS:
        packed[:b_local, col:col + w] = results[name].reshape(b_local, w).float()
        col += w
    packed[:b_local, col:col + PATH_BYTES // 4] = _encode_paths(paths, device).view(torch.float32)
    packed[:b_local, -1] = 1.0
    if world > 1:
        out = torch.empty(world * rows, width, dtype=torch.float32, device=device)
        dist.all_gather_into_tensor(out, packed, group=group)
    else:
        out = packed
    if rank != 0:
        return None
    out = out[out[:, -1] > 0.5]
    merged, col = {}, 0
    for name, w in RESULT_FIELDS:
        merged[name] = out[:, col:col + w].reshape(out.shape[0], 21, w // 21)
        col += w
    path_bytes = out[:, col:col + PATH_BYTES // 4].contiguous().view(torch.uint8)
    return merged, _decode_paths(path_bytes)
