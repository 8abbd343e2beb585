"""GPU microbench of the two tcgen05 window-attention kernels at the Swin-B stage shapes, batch 256:
csvit_swin_attn_core (qkv -> ctx, 8C B/token) and csvit_swin_attn_fused (fp32 x -> ctx, 6C B/token, C <= 256)."""
import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cs-vit_b200"))
import torch
from cs_vit import ops

B = int(os.environ.get("B", "256"))
dt = torch.float16 if os.environ.get("PREC", "fp16") == "fp16" else torch.bfloat16

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
    C = heads * 32; rows = B * H * H
    table = torch.randn(169, heads, device="cuda", generator=g)
    bias_l2 = ops.pack_rel_bias_log2(table, ops.rel_pos_index(7).long())
    qkv = torch.randn(rows, 3 * C, device="cuda", generator=g).to(dt)
    for shift in ((0, 3) if H > 7 else (0,)):
        for tok in (False, True):
            us = timeit(lambda: ops.swin_attn_core(qkv, bias_l2, B, H, H, heads, 7, shift, token_order=tok, q_prescaled=True))
            print(f"core  stage {stage} H={H:2d} C={C:4d} shift={shift} token_order={int(tok)}: {us:7.1f} us  {rows * C * 8 / us / 1e3:6.0f} GB/s (8C B/token)  "
                  f"{rows * 4 * 49 * C / us / 1e6:5.0f} TFLOP/s", flush=True)
        if C <= 256:
            x = torch.randn(rows, C, device="cuda", generator=g)
            w = [torch.randn(C, C, device="cuda", generator=g) * C ** -0.5 for _ in range(3)]
            b = [0.1 * torch.randn(C, device="cuda", generator=g) for _ in range(3)]
            pk = ops.pack_attn_fused(*w, *b, table, ops.rel_pos_index(7).long(), dt, torch.ones(C, device="cuda"), torch.zeros(C, device="cuda"))
            us = timeit(lambda: ops.swin_attn_fused(x, 1e-5, *pk, B, H, H, heads, 7, shift))
            print(f"fused stage {stage} H={H:2d} C={C:4d} shift={shift}: {us:7.1f} us  {rows * C * 6 / us / 1e3:6.0f} GB/s (6C B/token)  "
                  f"{rows * (6 * C * C + 4 * 49 * C) / us / 1e6:5.0f} TFLOP/s", flush=True)
