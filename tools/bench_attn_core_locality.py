"""GPU experiment: is swin_attn_core's bandwidth at the wide stages limited by the row pitch of qkv (a unit reads three 128-byte pieces
of every 3C x 2-byte row)?  Same number of (tile, head pair) units and bytes as the Swin-B stage shapes, but every unit's q | k | v
contiguous (C = 64, heads = 2, images x PAIRS): the bound a head-pair-major column order of the Q/K/V GEMM could reach."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cs-vit_b200"))
import torch
from cs_vit import ops
dt = torch.float16
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return 1e3 * ts[len(ts) // 2]
g = torch.Generator(device="cuda").manual_seed(0)
for stage, (H, heads) in enumerate([(56, 4), (28, 8), (14, 16), (7, 32)]):
    for B, hd in ((256, heads), (256 * heads // 2, 2)):
        C = hd * 32; rows = B * H * H
        table = torch.randn(169, hd, device="cuda", generator=g)
        bias_l2 = ops.pack_rel_bias_log2(table, ops.rel_pos_index(7).long())
        qkv = torch.randn(rows, 3 * C, device="cuda", generator=g).to(dt)
        us = timeit(lambda: ops.swin_attn_core(qkv, bias_l2, B, H, H, hd, 7, 0, token_order=False, q_prescaled=True))
        print(f"stage {stage} H={H:2d} images={B:5d} C={C:4d} (row pitch {6 * C:5d} B): {us:7.1f} us  {rows * C * 8 / us / 1e3:6.0f} GB/s", flush=True)
