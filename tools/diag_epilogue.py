"""GPU diagnostic: what the epilogue costs on the stage-2/3 GEMM shapes (Swin-B, batch 256): same (M, N, K) with a plain 16-bit
store, with GELU, with an fp32 store and with the fp32 residual; and the wave count of the CTA-pair grid (74 pairs, 256 x 256 tiles)."""
import os, sys, math
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cs-vit_b200"))
import torch
from cs_vit import ops
B = 256; dt = torch.float16
def timeit(fn, it=20):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it * 1e3
for name, n, N, K in [("s2 qkv", 196, 1536, 512), ("s2 proj", 196, 512, 512), ("s2 fc1", 196, 2048, 512), ("s2 fc2", 196, 512, 2048),
                      ("s3 qkv", 49, 3072, 1024), ("s3 fc1", 49, 4096, 1024), ("s3 fc2", 49, 1024, 4096)]:
    M = B * n
    g = torch.Generator(device="cuda").manual_seed(1)
    a = torch.randn(M, K, device="cuda", generator=g).to(dt)
    w = (torch.randn(N, K, device="cuda", generator=g) * 0.05).to(dt)
    b = torch.randn(N, device="cuda", generator=g)
    x = torch.randn(M, N, device="cuda", generator=g)
    tiles = math.ceil(M / 256) * math.ceil(N / 256)
    fl = 2.0 * M * N * K
    r = {}
    r["store16"] = timeit(lambda: ops.linear(a, w, b, out_dtype=dt))
    r["gelu16"] = timeit(lambda: ops.linear(a, w, b, act=ops.ACT_GELU, out_dtype=dt))
    r["store32"] = timeit(lambda: ops.linear(a, w, b, out=x))
    r["resid32"] = timeit(lambda: ops.linear(a, w, b, resid=x, out=x))
    print(f"{name:8s} M={M} N={N} K={K} tiles={tiles} waves={tiles/74:.2f} | " +
          " ".join(f"{k} {v:6.1f}us {fl/v/1e6:5.0f}TF" for k, v in r.items()), flush=True)
