"""GPU: per-entry-point time breakdown of one Swin-B batch-256 predict_batch step (CUDA events per launch)."""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cs-vit_b200"))
import torch
from cs_vit import ops
from cs_vit.net import Poser
from cs_vit.synthetic import make_inputs, make_random_backbone_dir, randomize_head_
from cs_vit.utils.mano_standin import SyntheticMANO

B = int(os.environ.get("B", "256")); prec = os.environ.get("PREC", "bf16")
VARIANT = os.environ.get("VARIANT", "swin_b"); S = 256 if VARIANT.startswith("swinv2") else 224
bdir = make_random_backbone_dir(os.path.join(tempfile.mkdtemp(), "b"), VARIANT, 0, image_size=S)
torch.manual_seed(0)
m = Poser(bdir, image_size=S, mano_layer=SyntheticMANO(), spatial_layer_type="encoder", persp_decorate="patch", precision=prec)
randomize_head_(m); m.phase(Poser.TrainingPhase.SPATIAL); m.eval(); m = m.cuda()
inp = {k: v.cuda() for k, v in make_inputs(B, 1, S, seed=3).items()}
def step():
    with torch.no_grad():
        return m.predict_batch(inp["patches"], inp["square_bboxes"], inp["timestamp"], inp["focal"], inp["princpt"])
if os.environ.get("FUSE_WIDTHS") is not None:
    m.backbone.fuse_attn_widths = tuple(int(v) for v in os.environ["FUSE_WIDTHS"].split(",") if v)
if os.environ.get("FUSE_ATTN") is not None:
    m.backbone.fuse_attn = os.environ["FUSE_ATTN"] == "1"
for _ in range(3): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): step()
e1.record(); torch.cuda.synchronize()
total = e0.elapsed_time(e1) / 5
print(f"step {total:.3f} ms  -> {B / total * 1e3:.0f} img/s")
acc = 0
for name in ("csvit_linear", "csvit_linear_emit", "csvit_linear_lnfold", "csvit_mlp_fused", "csvit_swin_attn_fused", "csvit_swin_attn_core", "csvit_window_attention", "csvit_window_attention_ex", "csvit_swinv2_window_attention", "csvit_layernorm_post", "csvit_layernorm", "csvit_attention", "csvit_affine_rows", "csvit_patch_im2col"):
    ops.begin_profile(name)
    for _ in range(3): step()
    p = ops.end_profile()
    ms = p["ms"] / 3; acc += ms
    extra = f" {p['flops'] / 3 / ms / 1e9:7.0f} TFLOP/s" if p["flops"] else (f" {p['bytes'] / 3 / ms / 1e6:7.0f} GB/s" if p["bytes"] else "")
    print(f"  {name:26s} {ms:7.3f} ms  {100 * ms / total:5.1f}%  x{p['launches'] // 3}{extra}")
print(f"  {'(torch glue + gaps)':26s} {total - acc:7.3f} ms  {100 * (total - acc) / total:5.1f}%")
