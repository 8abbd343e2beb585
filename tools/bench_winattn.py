"""GPU microbench of the window-attention core on the Swin-B batch-256 shapes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cs-vit_b200"))
import torch
from cs_vit import ops
B = int(os.environ.get("B", "256")); dt = torch.bfloat16
only = os.environ.get("ONLY")
for s, (hw, c, h) in enumerate([(56, 128, 4), (28, 256, 8), (14, 512, 16), (7, 1024, 32)]):
    if only and only != f"s{s}": continue
    g = torch.Generator(device="cuda").manual_seed(s)
    qkv = torch.randn(B * hw * hw, 3 * c, device="cuda", generator=g).to(dt)
    table = torch.randn(169, h, device="cuda", generator=g)
    for shift, impl in [(sh, im) for sh in ((0, 3) if hw > 7 else (0,)) for im in ("mma.sync", "tcgen05")]:
        bias = ops.expand_rel_bias_mma(table, 7) if impl == "mma.sync" else ops.expand_rel_bias(table, 7)
        f = lambda: ops.window_attention(qkv, bias, B, hw, hw, h, 7, shift)
        for _ in range(3): f()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): f()
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 100
        byts = qkv.numel() * 2 * 4 / 3
        print(f"s{s} {impl:8s} shift={shift} items={B*(hw//7)**2*h:7d} {us:8.1f} us  {byts/us/1e6:6.2f} TB/s  {B*(hw//7)**2*h/us:6.1f} items/us", flush=True)
