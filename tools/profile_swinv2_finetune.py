"""GPU: torch profiler table of one eager SwinV2-B finetune step (batch 32, 256x256): where the differentiable path spends its time."""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cs-vit_b200"))
import torch
from cs_vit.net import Poser
from cs_vit.synthetic import make_inputs, make_random_backbone_dir, randomize_head_
from cs_vit.utils.mano_standin import SyntheticMANO
variant = os.environ.get("VARIANT", "swinv2_b"); S = 256 if variant.startswith("swinv2") else 224
tmp = tempfile.mkdtemp()
bdir = make_random_backbone_dir(os.path.join(tmp, variant), variant, seed=0, image_size=S)
torch.manual_seed(0)
model = Poser(bdir, image_size=S, mano_layer=SyntheticMANO(), spatial_layer_type="encoder", persp_decorate="patch", precision="fp16")
randomize_head_(model); model.phase(Poser.TrainingPhase.SPATIAL); model = model.cuda()
batch = {k: v.cuda() for k, v in make_inputs(32, 1, S, seed=1, labels=True).items()}
def step():
    model.zero_grad(set_to_none=True)
    out = model(batch); out["loss"].backward()
for _ in range(2): step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=28, max_name_column_width=70))
