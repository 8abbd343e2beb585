"""Which gradient turns non-finite first when post-accumulate-grad hooks are installed (GPU diagnostic)."""
import os, sys, tempfile
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cs-vit_b200"))
from cs_vit.net import Poser
from cs_vit.synthetic import make_inputs, make_random_backbone_dir, randomize_head_
from cs_vit.utils.mano_standin import SyntheticMANO

variant, B = "swin_b", 32
mode = sys.argv[1] if len(sys.argv) > 1 else "check"
dev = torch.device("cuda")
bdir = make_random_backbone_dir(os.path.join(tempfile.mkdtemp(), variant), variant, seed=0)
torch.manual_seed(0)
model = Poser(bdir, image_size=224, mano_layer=SyntheticMANO(), spatial_layer_type="encoder", persp_decorate="patch", precision="bf16")
randomize_head_(model); model.phase(Poser.TrainingPhase.SPATIAL); model = model.to(dev)
names = {id(p): n for n, p in model.named_parameters()}
log = []
def hook(p):
    if mode == "check":
        log.append((names[id(p)], bool(torch.isfinite(p.grad).all()), p.grad.is_contiguous(), p.grad._base is not None))
    else:
        log.append((names[id(p)], None, None, None))
sel = [p for p in model.parameters() if p.requires_grad]
if mode == "headonly": sel = [p for p in sel if not names[id(p)].startswith("backbone.")]
if mode == "bbonly": sel = [p for p in sel if names[id(p)].startswith("backbone.")]
for p in sel: p.register_post_accumulate_grad_hook(hook)
batch = {k: v.to(dev) for k, v in make_inputs(B, 1, 224, seed=100, labels=True).items()}
out = model(batch); out["loss"].backward(); torch.cuda.synchronize()
bad = [names[id(p)] for p in model.parameters() if p.grad is not None and not torch.isfinite(p.grad).all()]
print(mode, "hooks:", len(sel), "fired:", len(log), "non-finite after backward:", len(bad), bad[:5])
if mode == "check":
    for i, e in enumerate(log[:12]): print(i, e)
    firstbad = next((i for i, e in enumerate(log) if not e[1]), None)
    print("first non-finite at hook time:", firstbad, log[firstbad] if firstbad is not None else None)
