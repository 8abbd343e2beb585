"""A/B of two library builds (CSVIT_LIB): window-attention kernel error vs fp64, and golden-case feature / joint errors."""
import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "cs-vit_b200")); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from cs_vit import ops
from helpers import build_product, rel, OUT_KEYS
print("lib:", os.environ.get("CSVIT_LIB", "default"))
for dtype in (torch.float16, torch.bfloat16):
    errs = []
    for seed in range(4):
        g = torch.Generator(device="cuda").manual_seed(seed)
        B, H, heads, ws, L, shift = 4, 28, 8, 7, 49, 3
        C = heads * 32; nW = (H // ws) ** 2
        qkv = (torch.randn(B * H * H, 3 * C, device="cuda", generator=g) * 1.5).to(dtype)
        table = torch.randn(169, heads, device="cuda", generator=g)
        out = ops.window_attention(qkv, ops.expand_rel_bias_mma(table, ws), B, H, H, heads, ws, shift)
        bias = ops.expand_rel_bias(table, ws).double()
        q, k, v = qkv.double().view(B * nW, L, 3, heads, 32).permute(2, 0, 3, 1, 4)
        s = q @ k.transpose(-1, -2) / math.sqrt(32) + bias[None]
        s = (s.view(B, nW, heads, L, L) + ops.shift_mask(H, H, ws, shift).double()[None, :, None]).view(B * nW, heads, L, L)
        ref = (s.softmax(-1) @ v).transpose(1, 2).reshape(B * H * H, C)
        errs.append(((out.double() - ref).norm() / ref.norm()).item())
    print(dtype, "kernel rel err", ["%.3e" % e for e in errs])
for name in ("swint_encoder_patch_spatial", "swinb_encoder_patch_spatial", "swint_encoder_query_sparse"):
    for prec in ("fp16", "bf16"):
        model, inputs, gold, case = build_product(name, prec)
        model = model.cuda()
        feats = model.backbone.forward_features(inputs["patches"].reshape(-1, 3, 224, 224).cuda(), normalize=True)
        dev = {k: v.cuda() for k, v in inputs.items()}
        with torch.no_grad():
            out = model.predict_batch(dev["patches"], dev["square_bboxes"], dev["timestamp"], dev["focal"], dev["princpt"])
        print(name, prec, "features %.3e" % rel(feats, gold["features"]), " ".join(f"{k} {rel(out[k], gold[k]):.2e}" for k in ("joint_cam", "verts_cam", "shape", "root_transl")))
