"""One launch of csvit_swin_attn_core (stage-2 shape) and one of csvit_swin_attn_fused (stage-0 shape), batch 256, for ncu."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cs-vit_b200"))
import torch
from cs_vit import ops
B = 256; dt = torch.float16
g = torch.Generator(device="cuda").manual_seed(0)
H, heads = 14, 16; C = heads * 32
table = torch.randn(169, heads, device="cuda", generator=g)
qkv = torch.randn(B * H * H, 3 * C, device="cuda", generator=g).to(dt)
ops.swin_attn_core(qkv, ops.pack_rel_bias_log2(table, ops.rel_pos_index(7).long()), B, H, H, heads, 7, 3, token_order=True, q_prescaled=True)
H, heads = 56, 4; C = heads * 32
table = torch.randn(169, heads, device="cuda", generator=g)
x = torch.randn(B * H * H, C, device="cuda", generator=g)
w = [torch.randn(C, C, device="cuda", generator=g) * C ** -0.5 for _ in range(3)]
b = [0.1 * torch.randn(C, device="cuda", generator=g) for _ in range(3)]
pk = ops.pack_attn_fused(*w, *b, table, ops.rel_pos_index(7).long(), dt, torch.ones(C, device="cuda"), torch.zeros(C, device="cuda"))
ops.swin_attn_fused(x, 1e-5, *pk, B, H, H, heads, 7, 3)
torch.cuda.synchronize()
