"""GPU: the GEMM engine vs cuBLAS (torch.matmul) on the Swin-B batch-256 shapes, same box, same run, bf16.
ours/plain = bias + 16-bit store (the like-for-like column: cuBLAS runs no epilogue at all); ours/real = the epilogue the model uses
(qkv: bias + store, proj / fc2: bias + fp32 residual read-modify-write, fc1: bias + GELU)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cs-vit_b200"))
import torch
from cs_vit import ops
B = 256; dt = torch.bfloat16
def timeit(fn, it=20):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it
print(f"{'shape':8s} {'M':>7s} {'N':>5s} {'K':>5s} {'cuBLAS':>8s} {'ours/plain':>11s} {'ratio':>6s} {'ours/real':>10s}   (TFLOP/s)")
for s, (n, c) in enumerate([(3136, 128), (784, 256), (196, 512), (49, 1024)]):
    M = B * n
    for name, N, K, epi in (("qkv", 3 * c, c, "store"), ("proj", c, c, "resid"), ("fc1", 4 * c, c, "gelu"), ("fc2", c, 4 * c, "resid")):
        g = torch.Generator(device="cuda").manual_seed(1)
        a = torch.randn(M, K, device="cuda", generator=g).to(dt)
        w = (torch.randn(N, K, device="cuda", generator=g) * 0.05).to(dt)
        b = torch.randn(N, device="cuda", generator=g)
        x = torch.randn(M, N, device="cuda", generator=g)
        out = torch.empty(M, N, device="cuda", dtype=dt)
        fl = 2.0 * M * N * K / 1e9
        t_cb = timeit(lambda: torch.matmul(a, w.t(), out=out))
        t_pl = timeit(lambda: ops.linear(a, w, b, out_dtype=dt))
        if epi == "store": t_re = t_pl
        elif epi == "gelu": t_re = timeit(lambda: ops.linear(a, w, b, act=ops.ACT_GELU, out_dtype=dt))
        else: t_re = timeit(lambda: ops.linear(a, w, b, resid=x, out=x))
        print(f"s{s} {name:5s} {M:7d} {N:5d} {K:5d} {fl/t_cb:8.0f} {fl/t_pl:11.0f} {t_cb/t_pl:6.2f} {fl/t_re:10.0f}", flush=True)
