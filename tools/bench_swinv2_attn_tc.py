"""GPU microbench: SwinV2-B cosine window attention at batch 256 - the mma.sync kernel vs the tcgen05 kernel (ns per (window, head))."""
import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cs-vit_b200"))
import torch
from cs_vit import ops
B = int(os.environ.get("B", "256"))
def timeit(fn, it=10):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it * 1e3
for dt in (torch.float16, torch.bfloat16):
    for H, heads, shift in ([(32, 8, 8), (16, 16, 0)] if os.environ.get("SHORT") else [(64, 4, 0), (64, 4, 8), (32, 8, 0), (32, 8, 8), (16, 16, 0)]):
        g = torch.Generator(device="cuda").manual_seed(1)
        C = heads * 32; rows = B * H * H
        qkv = torch.randn(rows, 3 * C, device="cuda", generator=g).to(dt)
        tab = (16 * torch.sigmoid(2 * torch.randn(heads, 961, device="cuda", generator=g))).contiguous()
        scale = torch.full((heads,), 10.0, device="cuda")
        qn = qkv.float().view(rows, 3, heads, 32)
        qn[:, 0] = torch.nn.functional.normalize(qn[:, 0], dim=-1) * 14.4
        qn[:, 1] = torch.nn.functional.normalize(qn[:, 1], dim=-1)
        qn = qn.view(rows, 3 * C).to(dt)
        bl = ops.swinv2_bias_log2(tab)
        t_old = timeit(lambda: ops.swinv2_window_attention(qkv, tab, scale, B, H, H, heads, 16, shift, token_order=True))
        t_new = timeit(lambda: ops.swinv2_attn_tc(qn, bl, B, H, H, heads, shift, token_order=True))
        items = rows // 256 * heads
        fl = 4.0 * rows * 256 * C
        print(f"{str(dt)[6:]:9s} H={H} heads={heads} shift={shift}: mma.sync {t_old:7.1f} us ({t_old*1e3/items:5.1f} ns/item, {fl/t_old/1e6:4.0f} TF)  "
              f"tcgen05 {t_new:7.1f} us ({t_new*1e3/items:5.1f} ns/item, {fl/t_new/1e6:4.0f} TF)  x{t_old/t_new:.2f}", flush=True)
