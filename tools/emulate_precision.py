"""CPU emulation of tensor-core operand rounding on the oracle: which operand format meets the parity bars?
Rounds every GEMM operand of the BACKBONE (activations, weights, q/k/v, probabilities) to bf16 or fp16, keeps
accumulation / residual / LN / softmax / GELU in fp32, runs the head in exact fp32."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cs-vit_b200")); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, torch.nn.functional as F
from helpers import build_product, head_options, manifest, rel, OUT_KEYS
from cs_vit.utils.mano_standin import SyntheticMANO
from oracle import head_restated as head, swin_restated as swin

orig_linear, orig_matmul = F.linear, torch.Tensor.__matmul__
def patched(dt):
    q = (lambda t: t.to(dt).float()) if dt is not None else (lambda t: t)
    def lin(x, w, b=None): return orig_linear(q(x), q(w), b)
    def mm(a, b): return orig_matmul(q(a), q(b))
    return lin, mm

names = sys.argv[1:] or sorted(manifest()["cases"])
for name in names:
    model, inputs, gold, case = build_product(name)
    sd = {k: v.detach() for k, v in model.state_dict().items()}
    opt = head_options(case)
    flat = inputs["patches"].reshape(-1, 3, 224, 224)
    for label, dt in (("bf16", torch.bfloat16), ("fp16", torch.float16)):
        lin, mm = patched(dt)
        swin.F.linear, torch.Tensor.__matmul__ = lin, mm
        try:
            with torch.no_grad():
                feats = head.backbone_features(flat, sd, opt)
        finally:
            swin.F.linear, torch.Tensor.__matmul__ = orig_linear, orig_matmul
        with torch.no_grad():
            out = head.predict_batch(inputs, sd, opt, SyntheticMANO(), features_fn=lambda x: feats, execute_all=False)
        print(f"{name:32s} operands={label} feat={rel(feats, gold['features']):.2e} " +
              " ".join(f"{k[:7]}={rel(out[k], gold[k]):.1e}" for k in OUT_KEYS), flush=True)
