"""GPU debugging aid: one traced launch of csvit_swin_attn_core at the Swin-B stage-2 shape (library built with
make EXTRA=-DCSVIT_AC_TRACE_BUILD; CSVIT_AC_TRACE=<file>)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cs-vit_b200"))
import torch
from cs_vit import ops
B, H, heads = 256, 14, 16
g = torch.Generator(device="cuda").manual_seed(0)
C = heads * 32; rows = B * H * H
table = torch.randn(169, heads, device="cuda", generator=g)
bias_l2 = ops.pack_rel_bias_log2(table, ops.rel_pos_index(7).long())
qkv = torch.randn(rows, 3 * C, device="cuda", generator=g).to(torch.float16)
ops.swin_attn_core(qkv, bias_l2, B, H, H, heads, 7, 3, token_order=True, q_prescaled=True)
torch.cuda.synchronize()
print(open(os.environ["CSVIT_AC_TRACE"]).read())
