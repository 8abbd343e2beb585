"""GPU stress test of csvit_mlp_fused (C = 128: residual epilogue on its own warps, second GEMM2 accumulator): ragged and tiny M, more and
fewer tiles than CTAs, each launch repeated - bit-identical run to run and equal to fp32 torch math within the 16-bit operand rounding."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cs-vit_b200"))
import torch
from cs_vit import ops
bad = 0; n = 0
for c in (128, 256):
    for M in (1, 100, 128, 129, 257, 5000, 148 * 128 - 1, 148 * 128, 148 * 128 + 1, 2 * 148 * 128 + 77, 100352):
        for dt in (torch.float16, torch.bfloat16):
            for rep in range(3):
                g = torch.Generator(device="cuda").manual_seed(M + c + rep)
                xn = torch.randn(M, c, device="cuda", generator=g).to(dt)
                w1 = (torch.randn(4 * c, c, device="cuda", generator=g) * c ** -0.5).to(dt); b1 = 0.1 * torch.randn(4 * c, device="cuda", generator=g)
                w2 = (torch.randn(c, 4 * c, device="cuda", generator=g) * (4 * c) ** -0.5).to(dt); b2 = 0.1 * torch.randn(c, device="cuda", generator=g)
                x0 = torch.randn(M, c, device="cuda", generator=g)
                outs = []
                for _ in range(3):
                    x = x0.clone(); ops.mlp_fused(xn, w1, b1, w2, b2, x); outs.append(x)
                n += 1
                if not (torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])):
                    bad += 1; print("NONDETERMINISTIC", c, M, dt, flush=True)
                h = torch.nn.functional.gelu(xn.float() @ w1.float().T + b1).to(dt).float()
                ref = x0 + h @ w2.float().T + b2
                err = ((outs[0] - ref).norm() / ref.norm()).item()
                if not err < (6e-3 if dt == torch.bfloat16 else 1e-3):
                    bad += 1; print("MISMATCH", c, M, dt, err, flush=True)
torch.cuda.synchronize()
print(f"stress: {n} cases x 3 launches, {bad} failures")
