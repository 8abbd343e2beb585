"""GPU debugging aid: one traced launch of csvit_swinv2_attn_tc (CSVIT_V2_TRACE=<file>): clock64 at the phase boundaries of CTA 0."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cs-vit_b200"))
import torch
from cs_vit import ops
B, H, heads, shift = 256, 16, 16, 0
g = torch.Generator(device="cuda").manual_seed(1)
C = heads * 32; rows = B * H * H
qn = torch.randn(rows, 3 * C, device="cuda", generator=g).to(torch.float16)
bl = ops.swinv2_bias_log2((16 * torch.sigmoid(2 * torch.randn(heads, 961, device="cuda", generator=g))).contiguous())
ops.swinv2_attn_tc(qn, bl, B, H, H, heads, shift, token_order=True)
torch.cuda.synchronize()
print(open(os.environ["CSVIT_V2_TRACE"]).read())
