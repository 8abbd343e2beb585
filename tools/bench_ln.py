"""GPU microbench of the LayerNorm-gather kernel on the Swin-B batch-256 shapes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cs-vit_b200"))
import torch
from cs_vit import ops
B = 256; dt = torch.bfloat16
for s, (hw, c) in enumerate([(56, 128), (28, 256), (14, 512), (7, 1024)]):
    x = torch.randn(B * hw * hw, c, device="cuda")
    g = torch.ones(c, device="cuda"); b = torch.zeros(c, device="cuda")
    for mode, shift in ((0, 0), (1, 3 if hw > 7 else 0)):
        f = lambda: ops.layernorm(x, g, b, 1e-5, out_dtype=dt, mode=mode, grid=(hw, hw), ws=7, shift=shift)
        for _ in range(3): f()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): f()
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 100
        print(f"s{s} C={c} mode={mode}: {us:7.1f} us  {x.numel() * 6 / us / 1e6:5.2f} TB/s", flush=True)
# reference points: a dtype-converting copy of the same byte volume (torch elementwise kernel) and a plain fp32 copy
for s, (hw, c) in enumerate([(56, 128), (14, 512)]):
    x = torch.randn(B * hw * hw, c, device="cuda"); y = torch.empty_like(x, dtype=dt); z = torch.empty_like(x)
    for name, f, byts in (("cast fp32->bf16", lambda: y.copy_(x), 6), ("copy fp32", lambda: z.copy_(x), 8)):
        for _ in range(3): f()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): f()
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 100
        print(f"s{s} C={c} {name}: {us:7.1f} us  {x.numel() * byts / us / 1e6:5.2f} TB/s", flush=True)
