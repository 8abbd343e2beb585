"""Diagnostics for csvit_swin_attn_fused: error split by head, window parity inside the tile, slot, masked / unmasked windows."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cs-vit_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from cs_vit import ops  # noqa: E402
from test_kernels_gpu import fused_attention_case, rel  # noqa: E402


def run(B, H, heads, shift, dtype, zero_bias=False):
    x, gamma, beta, (w, b, bo), ref = fused_attention_case(ops, B, H, heads, shift, dtype, seed=1, zero_bias=zero_bias)
    out = ops.swin_attn_fused(x, 1e-5, w, b, bo, B, H, H, heads, 7, shift)
    torch.cuda.synchronize()
    C, N = heads * 32, H * H
    print(f"B={B} H={H} heads={heads} shift={shift} {dtype} zero_bias={zero_bias}: rel={rel(out, ref):.3e} finite={bool(torch.isfinite(out.float()).all())}")
    o, r = out.float().view(B, N, heads, 32), ref.view(B, N, heads, 32)
    per_head = [(o[:, :, h] - r[:, :, h]).norm().item() / r[:, :, h].norm().item() for h in range(heads)]
    print("  per head:", " ".join(f"{e:.2e}" for e in per_head))
    idx = ops.window_index_map(H, H, 7, shift).long()
    nW = N // 49
    ow, rw = o[:, idx].reshape(B * nW, 49, -1), r[:, idx].reshape(B * nW, 49, -1)
    per_win = ((ow - rw).flatten(1).norm(dim=1) / rw.flatten(1).norm(dim=1))
    print("  window parity A/B:", f"{per_win[0::2].mean().item():.2e} {per_win[1::2].mean().item():.2e}", " worst window", int(per_win.argmax()),
          f"{per_win.max().item():.2e}", " first 8 windows:", " ".join(f"{e:.1e}" for e in per_win[:8].tolist()))
    per_slot = ((ow - rw).norm(dim=2).mean(0) / rw.norm(dim=2).mean(0))
    print("  per slot (first 10, last 5):", " ".join(f"{e:.1e}" for e in per_slot[:10].tolist()), "...", " ".join(f"{e:.1e}" for e in per_slot[-5:].tolist()))


if __name__ == "__main__":
    for dt in (torch.float16, torch.bfloat16):
        run(2, 14, 4, 0, dt, zero_bias=True)
        run(2, 14, 4, 0, dt)
        run(2, 14, 4, 3, dt)
        run(3, 28, 8, 3, dt)
        run(2, 56, 4, 3, dt)
    # timing at the batch-256 shapes
    for (H, heads) in ((56, 4), (28, 8)):
        for shift in (0, 3):
            x, gamma, beta, (w, b, bo), ref = fused_attention_case(ops, 256, H, heads, shift, torch.bfloat16, seed=2)
            out = torch.empty_like(ref, dtype=torch.bfloat16)
            for _ in range(3):
                ops.swin_attn_fused(x, 1e-5, w, b, bo, 256, H, H, heads, 7, shift, out=out)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                ops.swin_attn_fused(x, 1e-5, w, b, bo, 256, H, H, heads, 7, shift, out=out)
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 100
            C = heads * 32
            print(f"timing H={H} C={C} shift={shift}: {us:.1f} us  {256 * H * H * C * 6 / us / 1e6:.2f} TB/s (6C B/token)  rel={rel(out, ref):.2e}")
