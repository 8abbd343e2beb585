"""Timeline of block 0 of csvit_swin_attn_fused (needs a library built with EXTRA=-DCSVIT_FA_TRACE): clock64 stamps of the
issuer (G / S / PV), one softmax warp of each group (D / X / E) and the LayerNorm producer, printed per global head."""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cs-vit_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
from cs_vit import ops, _lib
from test_kernels_gpu import fused_attention_case
H, heads, shift = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
lo, hi = (int(sys.argv[4]), int(sys.argv[5])) if len(sys.argv) > 5 else (24, 48)
x, gamma, beta, (w, b, bo), ref = fused_attention_case(ops, 256, H, heads, shift, torch.bfloat16, seed=2)
for _ in range(3):
    out = ops.swin_attn_fused(x, 1e-5, w, b, bo, 256, H, H, heads, 7, shift)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * (16 * 256))()
lib = _lib.load()
assert lib.csvit_debug_fa_trace(buf) == 0
import numpy as np
t = np.array(buf[:], dtype=np.int64).reshape(16, 256)
t0 = t[0, 0]
names = ["Gwait", "Gdone", "PViss", "Siss", "Scmt", "Dwait", "Dbeg", "Dend", "Xbeg", "Xend", "Ebeg"]
print("gh   " + " ".join(f"{n:>7s}" for n in names))
for gh in range(lo, hi):
    print(f"{gh:3d}  " + " ".join(f"{int(t[k, gh] - t0):7d}" for k in range(11)))
print("gh   Gwait  w_full(kb0) w_full(kb1)  Gdone   (issuer inside G)   TMA issued kb0 / kb1 (producer)")
for gh in range(lo, min(hi, 63)):
    print(f"{gh:3d} {int(t[0, gh] - t0):7d} {int(t[15, 2 * gh] - t0):9d} {int(t[15, 2 * gh + 1] - t0):9d} {int(t[1, gh] - t0):9d}      "
          f"{int(t[15, 128 + 2 * gh] - t0):9d} {int(t[15, 128 + 2 * gh + 1] - t0):9d}")
print("per-head period (G issue):", np.diff(t[0, lo:hi]).tolist())
print("tile LNbeg  LNwaitE LNgotE  LNend")
for ti in range(lo // heads, hi // heads + 1):
    print(f"{ti:3d} " + " ".join(f"{int(t[k, ti] - t0):7d}" for k in (11, 12, 13, 14)))

