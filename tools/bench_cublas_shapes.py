"""GPU: cuBLAS (torch.matmul, bf16, no epilogue) on the Swin-B batch-256 GEMM shapes - the calibration point for tools/bench_gemm.py."""
import torch
B = 256
print(f"{'shape':10s} {'M':>7s} {'N':>5s} {'K':>5s}  cuBLAS TFLOP/s (plain bf16 GEMM, bf16 output, no bias / activation / residual)")
for s, (n, c) in enumerate([(3136, 128), (784, 256), (196, 512), (49, 1024)]):
    M = B * n
    for name, N, K in (("qkv", 3 * c, c), ("proj", c, c), ("fc1", 4 * c, c), ("fc2", c, 4 * c)):
        a = torch.randn(M, K, device="cuda").bfloat16(); w = torch.randn(N, K, device="cuda").bfloat16()
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        f = lambda: torch.matmul(a, w.t(), out=out)
        for _ in range(3): f()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): f()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"s{s} {name:6s} {M:7d} {N:5d} {K:5d}  {2.0 * M * N * K / ms / 1e9:7.0f}", flush=True)
