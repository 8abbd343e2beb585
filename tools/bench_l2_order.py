"""GPU experiment: does a consumer that walks its input in the REVERSE of the producer's order find the producer's last-written rows in
the 126 MB L2?  Stage-2 shapes of Swin-B at batch 256: fc2 (in-place residual GEMM, ascending tiles) -> LayerNorm(x) -> QKV GEMM(xn).
Times each kernel of the chain separately with events between the launches.  CSVIT_LN_REVERSE was a switch of the experiment's build of
ln_rows_kernel (block index reversed); the experiment was rejected (profiles/r2_l2_order_experiment.txt) and the switch is not in the library."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cs-vit_b200"))
import torch
from cs_vit import ops
dt = torch.float16
M, C = 256 * 196, 512
g = torch.Generator(device="cuda").manual_seed(1)
hid = torch.randn(M, 4 * C, device="cuda", generator=g).to(dt)
w2 = (torch.randn(C, 4 * C, device="cuda", generator=g) * 0.02).to(dt)
b2 = torch.randn(C, device="cuda", generator=g)
wq = (torch.randn(3 * C, C, device="cuda", generator=g) * 0.05).to(dt)
bq = torch.randn(3 * C, device="cuda", generator=g)
gam, bet = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
x = torch.randn(M, C, device="cuda", generator=g)
it = 20
ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(it)]
for k in range(3 + it):
    e = ev[max(k - 3, 0)]
    e[0].record()
    ops.linear(hid, w2, b2, resid=x, out=x)
    e[1].record()
    xn = ops.layernorm(x, gam, bet, 1e-5, out_dtype=dt)
    e[2].record()
    qkv = ops.linear(xn, wq, bq, out_dtype=dt)
    e[3].record()
torch.cuda.synchronize()
t = [sum(e[i].elapsed_time(e[i + 1]) for e in ev) / it * 1e3 for i in range(3)]
print(f"CSVIT_LN_REVERSE={os.environ.get('CSVIT_LN_REVERSE', '0')}: fc2 {t[0]:.1f} us  LayerNorm {t[1]:.1f} us  QKV {t[2]:.1f} us  (events between launches)")
