"""One launch each of the four stage-2 GEMMs of Swin-B at batch 256 (Q/K/V plain 16-bit store, out-proj + residual, fc1 + GELU, fc2 + residual) for ncu."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cs-vit_b200"))
import torch
from cs_vit import ops
M, C = 50176, 512
dt = torch.float16
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(M, C, device="cuda", generator=g)
ctx = torch.randn(M, C, device="cuda", generator=g).to(dt)
wo = (torch.randn(C, C, device="cuda", generator=g) * 0.03).to(dt)
w1 = (torch.randn(4 * C, C, device="cuda", generator=g) * 0.03).to(dt)
w2 = (torch.randn(C, 4 * C, device="cuda", generator=g) * 0.03).to(dt)
b = torch.zeros(4 * C, device="cuda")
wq = (torch.randn(3 * C, C, device="cuda", generator=g) * 0.03).to(dt)
qkv = ops.linear(ctx, wq, b[:3 * C], out_dtype=dt)                # Q/K/V: plain 16-bit TMA store
ops.linear(ctx, wo, b[:C], resid=x, out=x)                       # out-proj + fp32 residual (TMA epilogue)
hid = ops.linear(ctx, w1, b, act=ops.ACT_GELU, out_dtype=dt)     # fc1 + GELU
ops.linear(hid, w2, b[:C], resid=x, out=x)                       # fc2 + fp32 residual
torch.cuda.synchronize()
