"""One launch of csvit_swin_attn_fused at a batch-256 Swin-B stage shape (for ncu)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "cs-vit_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
from cs_vit import ops
from test_kernels_gpu import fused_attention_case
H, heads, shift = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
x, gamma, beta, (w, b, bo), ref = fused_attention_case(ops, 256, H, heads, shift, torch.bfloat16, seed=2)
for _ in range(2):
    out = ops.swin_attn_fused(x, 1e-5, w, b, bo, 256, H, H, heads, 7, shift)
torch.cuda.synchronize()
print("ok")
