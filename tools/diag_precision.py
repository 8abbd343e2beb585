"""Diagnostic (GPU): error budget of the bf16 backbone / TF32 head against the reference goldens."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cs-vit_b200")); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from helpers import build_product, manifest, rel, OUT_KEYS
from cs_vit.net.blocks import set_precision

for name in sorted(manifest()["cases"]):
    for bb, hd in (("bf16", "bf16"), ("fp16", "bf16"), ("fp16", "fp32")):
        model, inputs, gold, case = build_product(name, "bf16")
        model = model.cuda()
        set_precision(model, hd); model.precision = hd
        model.backbone.precision = bb
        dev = {k: v.cuda() for k, v in inputs.items()}
        with torch.no_grad():
            f = model.backbone.forward_features(dev["patches"].reshape(-1, 3, 224, 224), normalize=True)
            out = model.predict_batch(dev["patches"], dev["square_bboxes"], dev["timestamp"], dev["focal"], dev["princpt"])
        errs = {k: rel(out[k], gold[k]) for k in OUT_KEYS}
        print(f"{name:34s} backbone={bb} head={'tf32' if hd=='bf16' else 'fp32'} feat={rel(f, gold['features']):.2e} " +
              " ".join(f"{k[:6]}={v:.1e}" for k, v in errs.items()), flush=True)
