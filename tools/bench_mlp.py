"""GPU microbench: fused MLP (csvit_mlp_fused) vs fc1 + fc2 GEMMs on the narrow Swin-B stages, batch 256."""
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
for s, (hw, c) in enumerate([(56, 128), (28, 256)]):
    M = B * hw * hw
    g = torch.Generator(device="cuda").manual_seed(s)
    xn = torch.randn(M, c, device="cuda", generator=g).to(dt)
    w1 = (torch.randn(4 * c, c, device="cuda", generator=g) * 0.05).to(dt); b1 = torch.randn(4 * c, device="cuda", generator=g)
    w2 = (torch.randn(c, 4 * c, device="cuda", generator=g) * 0.05).to(dt); b2 = torch.randn(c, device="cuda", generator=g)
    x = torch.randn(M, c, device="cuda", generator=g)
    tf = timeit(lambda: ops.mlp_fused(xn, w1, b1, w2, b2, x))
    def unf():
        h = ops.linear(xn, w1, b1, act=ops.ACT_GELU, out_dtype=dt)
        ops.linear(h, w2, b2, resid=x, out=x)
    tu = timeit(unf)
    fl = 16.0 * M * c * c
    print(f"s{s} M={M} C={c}: fused {tf:7.1f} us ({fl/tf/1e6:5.0f} TFLOP/s)   fc1+fc2 {tu:7.1f} us ({fl/tu/1e6:5.0f} TFLOP/s)", flush=True)
