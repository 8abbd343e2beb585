"""GPU check of the CTA-pair GEMM's 16-bit store paths (gemm_pair_kernel<FMT, 8> plain store, <FMT, 16> GELU at K <= 512) at Swin-B stage-2
sizes: against fp32 torch math on the same 16-bit operands, two launches bit-identical; ragged M (not a multiple of 256) included."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cs-vit_b200"))
import torch
from cs_vit import ops
torch.manual_seed(0)
bad = 0
for dt in (torch.float16, torch.bfloat16):
    for M, N, K, act in [(50176, 1536, 512, 0), (50176, 2048, 512, 1), (40000 + 77, 1024, 512, 0), (38000, 512, 512, 1), (50176, 768, 256, 0)]:
        a = torch.randn(M, K, device="cuda").to(dt)
        w = (torch.randn(N, K, device="cuda") * 0.05).to(dt)
        b = torch.randn(N, device="cuda")
        y = ops.linear(a, w, b, act=ops.ACT_GELU if act else ops.ACT_NONE, out_dtype=dt)
        kern = ops.last_gemm_kernel() if hasattr(ops, "last_gemm_kernel") else -1
        worst = 0.0
        for lo in range(0, M, 8192):
            ref = a[lo:lo + 8192].float() @ w.float().t() + b
            if act:
                ref = torch.nn.functional.gelu(ref)
            err = (y[lo:lo + 8192].float() - ref).abs().max().item() / max(ref.abs().max().item(), 1e-6)
            worst = max(worst, err)
        y2 = ops.linear(a, w, b, act=ops.ACT_GELU if act else ops.ACT_NONE, out_dtype=dt)
        same = torch.equal(y, y2)
        tol = 4e-3 if dt == torch.float16 else 1.6e-2
        ok = worst < tol and same
        bad += 0 if ok else 1
        print(f"{str(dt):15s} M={M} N={N} K={K} act={act} kernel={kern} max rel err {worst:.2e} repeat-identical {same} {'ok' if ok else 'FAIL'}", flush=True)
print("FAILED" if bad else "all ok")
sys.exit(1 if bad else 0)
