"""One launch of csvit_swinv2_attn_tc (SwinV2-B stage-2 shape: 16 x 16 tokens = one window per image, 16 heads, batch 256) and, for
comparison, one of the mma.sync kernel it replaces, for ncu."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cs-vit_b200"))
import torch
from cs_vit import ops
B, H, heads, shift = 256, 16, 16, 0
g = torch.Generator(device="cuda").manual_seed(1)
C = heads * 32; rows = B * H * H
qkv = torch.randn(rows, 3 * C, device="cuda", generator=g).to(torch.float16)
tab = (16 * torch.sigmoid(2 * torch.randn(heads, 961, device="cuda", generator=g))).contiguous()
scale = torch.full((heads,), 10.0, device="cuda")
qn = qkv.float().view(rows, 3, heads, 32)
qn[:, 0] = torch.nn.functional.normalize(qn[:, 0], dim=-1) * 14.4
qn[:, 1] = torch.nn.functional.normalize(qn[:, 1], dim=-1)
qn = qn.view(rows, 3 * C).to(torch.float16)
ops.swinv2_attn_tc(qn, ops.swinv2_bias_log2(tab), B, H, H, heads, shift, token_order=True)
ops.swinv2_window_attention(qkv, tab, scale, B, H, H, heads, 16, shift, token_order=True)
torch.cuda.synchronize()
