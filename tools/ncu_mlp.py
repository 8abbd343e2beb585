"""One launch of csvit_mlp_fused at the stage-0 and stage-1 shapes of Swin-B, batch 256, for ncu."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cs-vit_b200"))
import torch
from cs_vit import ops
B = 256; dt = torch.float16
for s, (hw, c) in enumerate([(56, 128), (28, 256)]):
    M = B * hw * hw
    g = torch.Generator(device="cuda").manual_seed(s)
    xn = torch.randn(M, c, device="cuda", generator=g).to(dt)
    w1 = (torch.randn(4 * c, c, device="cuda", generator=g) * 0.05).to(dt); b1 = torch.randn(4 * c, device="cuda", generator=g)
    w2 = (torch.randn(c, 4 * c, device="cuda", generator=g) * 0.05).to(dt); b2 = torch.randn(c, device="cuda", generator=g)
    x = torch.randn(M, c, device="cuda", generator=g)
    ops.mlp_fused(xn, w1, b1, w2, b2, x)
torch.cuda.synchronize()
