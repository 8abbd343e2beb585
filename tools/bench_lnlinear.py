"""GPU microbench: LayerNorm-prologue pair GEMM vs csvit_layernorm + csvit_linear on the Swin-B batch-256 shapes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cs-vit_b200"))
import torch
from cs_vit import ops
B = int(os.environ.get("B", "256")); dt = torch.bfloat16
def timeit(fn, it=10):
    for _ in range(2): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it * 1e3
print(f"{'shape':10s} {'M':>7s} {'N':>5s} {'C':>4s}  fused_us  ln_us  gemm_us  fused_TF")
for s, (hw, c) in enumerate([(56, 128), (28, 256), (14, 512)]):
    M = B * hw * hw
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(M, c, device="cuda", generator=g)
    gamma = torch.ones(c, device="cuda"); beta = torch.zeros(c, device="cuda")
    for name, N, mode, act in (("qkv", 3 * c, 1, 0), ("fc1", 4 * c, 0, 1)):
        w = (torch.randn(N, c, device="cuda", generator=g) * 0.05).to(dt)
        b = torch.randn(N, device="cuda", generator=g)
        kw = dict(mode=mode, grid=(hw, hw), ws=7, shift=3 if mode else 0)
        tf = timeit(lambda: ops.ln_linear(x, gamma, beta, 1e-5, w, b, act=act, **kw))
        tl = timeit(lambda: ops.layernorm(x, gamma, beta, 1e-5, out_dtype=dt, **kw))
        xn = ops.layernorm(x, gamma, beta, 1e-5, out_dtype=dt, **kw)
        tg = timeit(lambda: ops.linear(xn, w, b, act=act, out_dtype=dt))
        print(f"s{s} {name:6s} {M:7d} {N:5d} {c:4d}  {tf:8.1f} {tl:6.1f} {tg:8.1f}  {2.0*M*N*c/tf/1e6:8.0f}", flush=True)
