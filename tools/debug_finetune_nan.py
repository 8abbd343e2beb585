"""Locate the first non-finite value in the finetune step (GPU diagnostic)."""
import os, sys, tempfile
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cs-vit_b200"))
from cs_vit.net import Poser
from cs_vit.synthetic import make_inputs, make_random_backbone_dir, randomize_head_
from cs_vit.train import GradReducer, finetune_step
from cs_vit.utils.mano_standin import SyntheticMANO

variant = sys.argv[1] if len(sys.argv) > 1 else "swin_b"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
use_reducer = (sys.argv[3] if len(sys.argv) > 3 else "1") == "1"
dev = torch.device("cuda")
bdir = make_random_backbone_dir(os.path.join(tempfile.mkdtemp(), variant), variant, seed=0)
torch.manual_seed(0)
model = Poser(bdir, image_size=224, mano_layer=SyntheticMANO(), spatial_layer_type="encoder", persp_decorate="patch", precision="bf16")
randomize_head_(model); model.phase(Poser.TrainingPhase.SPATIAL); model = model.to(dev)
trainable = [p for p in model.parameters() if p.requires_grad]
names = {id(p): n for n, p in model.named_parameters()}
reducer = GradReducer(trainable) if use_reducer else None
opt = torch.optim.AdamW(trainable, lr=8e-6, fused=True)
batch = {k: v.to(dev) for k, v in make_inputs(B, 1, 224, seed=100, labels=True).items()}
for step in range(int(os.environ.get('STEPS', '6'))):
    if reducer: reducer.zero_grad()
    else: opt.zero_grad(set_to_none=True)
    out = model(batch); loss = out["loss"]; loss.backward()
    pre = [names[id(p)] for p in trainable if p.grad is not None and not torch.isfinite(p.grad).all()]
    print(f"step {step}: non-finite grads BEFORE finish(): {pre[:6]} ({len(pre)})")
    if reducer: reducer.finish()
    bad = [(names[id(p)], int((~torch.isfinite(p.grad)).sum())) for p in trainable if p.grad is not None and not torch.isfinite(p.grad).all()]
    tot = torch.sqrt(sum((p.grad.double() ** 2).sum() for p in trainable if p.grad is not None)).item()
    big = sorted(((p.grad.abs().max().item(), names[id(p)]) for p in trainable if p.grad is not None), reverse=True)[:4]
    print(f"step {step}: loss {loss.item():.4f} grad norm {tot:.4e} non-finite grads: {bad[:8]} ({len(bad)} params); largest |g|: {big}")
    torch.nn.utils.clip_grad_norm_([p for p in trainable if p.grad is not None], 5.0)
    opt.step()
    badp = [names[id(p)] for p in trainable if not torch.isfinite(p).all()]
    print(f"        params non-finite after step: {badp[:8]} ({len(badp)})")
