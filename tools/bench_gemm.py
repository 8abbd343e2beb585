"""GPU microbench of the GEMM engine on the Swin-B batch-256 shapes (TFLOP/s per epilogue / cluster / store path)."""
import os, sys, itertools
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cs-vit_b200"))
import torch
from cs_vit import ops

B = int(os.environ.get("B", "256"))
dt = torch.bfloat16
shapes = []
for s, (n, c) in enumerate([(3136, 128), (784, 256), (196, 512), (49, 1024)]):
    M = B * n
    shapes += [(f"s{s} qkv", M, 3 * c, c, "store"), (f"s{s} proj", M, c, c, "resid"),
               (f"s{s} fc1", M, 4 * c, c, "gelu"), (f"s{s} fc2", M, c, 4 * c, "resid")]
only = os.environ.get("ONLY")
configs = [(1, 0, 0), (1, 1, 0), (1, 1, 1)]
print(f"{'shape':10s} {'M':>7s} {'N':>5s} {'K':>5s} {'epi':6s} " + " ".join(f"cs{c}/t{t}/p{p}".rjust(10) for c, t, p in configs))
for name, M, N, K, epi in shapes:
    if only and only not in name: continue
    g = torch.Generator(device="cuda").manual_seed(1)
    a = torch.randn(M, K, device="cuda", generator=g).to(dt)
    w = (torch.randn(N, K, device="cuda", generator=g) * 0.05).to(dt)
    b = torch.randn(N, device="cuda", generator=g)
    x = torch.randn(M, N, device="cuda", generator=g) if epi == "resid" else None
    res = []
    ref = None
    for cs, tma, pair in configs:
        ops.set_gemm_tuning(cs, tma, 0, pair)
        def run():
            if epi == "store": return ops.linear(a, w, b, out_dtype=dt)
            if epi == "gelu": return ops.linear(a, w, b, act=ops.ACT_GELU, out_dtype=dt)
            return ops.linear(a, w, b, resid=x, out=x)
        out = run(); torch.cuda.synchronize()
        if epi != "resid":
            if ref is None:
                rr = a[:4096].float() @ w.float().T + b
                ref = torch.nn.functional.gelu(rr) if epi == "gelu" else rr
            err = ((out[:4096].float() - ref).norm() / ref.norm()).item()
            assert err < 6e-3, (name, cs, tma, pair, err)
        for _ in range(2): run()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        it = 10
        e0.record()
        for _ in range(it): run()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / it
        res.append(2.0 * M * N * K / ms / 1e9)
    print(f"{name:10s} {M:7d} {N:5d} {K:5d} {epi:6s} " + " ".join(f"{r:10.0f}" for r in res), flush=True)
ops.set_gemm_tuning()
