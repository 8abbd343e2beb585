import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cs-vit_b200"))
import torch
from cs_vit import ops
dt = torch.float16
for (M, N, K, act) in [(50176, 1536, 512, 0), (50176, 2048, 512, 1), (12544, 3072, 1024, 0)]:
    a = torch.randn(M, K, device="cuda").to(dt); w = (torch.randn(N, K, device="cuda") * 0.05).to(dt); b = torch.randn(N, device="cuda")
    for _ in range(3):
        ops.linear(a, w, b, act=ops.ACT_GELU if act else ops.ACT_NONE, out_dtype=dt)
    torch.cuda.synchronize()
