"""GPU microbench of the SwinV2 cosine window-attention core on the SwinV2-B w16 @256 batch-256 shapes.
ONLY=s2 REPS=1 runs a single shape once (for ncu captures)."""
import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cs-vit_b200"))
import torch
from cs_vit import ops
B = int(os.environ.get("B", "256")); dt = torch.bfloat16
only = os.environ.get("ONLY"); reps = int(os.environ.get("REPS", "10"))
for s, (hw, c, h, ws) in enumerate([(64, 128, 4, 16), (32, 256, 8, 16), (16, 512, 16, 16), (8, 1024, 32, 8)]):
    if only and only != f"s{s}": continue
    g = torch.Generator(device="cuda").manual_seed(s)
    qkv = torch.randn(B * hw * hw, 3 * c, device="cuda", generator=g).to(dt)
    tab = (16 * torch.sigmoid(torch.randn(h, (2 * ws - 1) ** 2, device="cuda", generator=g))).contiguous()
    scale = torch.full((h,), 10.0, device="cuda")
    for shift in ((0, ws // 2) if hw > ws else (0,)):
        f = lambda: ops.swinv2_window_attention(qkv, tab, scale, B, hw, hw, h, ws, shift)
        for _ in range(3 if reps > 1 else 0): f()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): f()
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / reps
        items = B * (hw // ws) ** 2 * h
        flops = 4.0 * B * hw * hw * ws * ws * c
        print(f"s{s} ws={ws} shift={shift} items={items:7d} {us:8.1f} us  {flops/us/1e6:6.1f} TFLOP/s  {us*1e3/items:7.1f} ns/item  "
              f"{qkv.numel()*2*4/3/us/1e6:5.2f} TB/s", flush=True)
