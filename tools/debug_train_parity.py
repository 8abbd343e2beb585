"""Print per-parameter gradient errors of the product's finetune step against the reference goldens (GPU)."""
import sys, os
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cs-vit_b200")); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
from helpers import build_train_case, grad_projections

name = sys.argv[1]; precision = sys.argv[2] if len(sys.argv) > 2 else "fp32"
model, batch, gold, case = build_train_case(name, precision)
model = model.cuda()
dev = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in batch.items()}
predict = model.predict_batch(dev["patches"], dev["square_bboxes"], dev["timestamp"], dev["focal"], dev["princpt"])
loss, parts = model._criterion(predict, dev)
loss.backward(); torch.cuda.synchronize()
print("loss", loss.item(), float(gold["loss"]))
for k in ("joint_cam", "verts_cam", "shape", "root_transl"):
    ref = torch.from_numpy(gold[k]).double()
    print(k, ((predict[k].detach().double().cpu() - ref).norm() / ref.norm()).item())
names = [str(n) for n in gold["param_names"]]
params = dict(model.named_parameters())
rows = []
for i, n in enumerate(names):
    g = params[n].grad
    if not gold["param_has_grad"][i]:
        if g is not None and float(g.abs().max()) > 0: rows.append((9.9, n, "unexpected grad"))
        continue
    if g is None: rows.append((9.9, n, "MISSING")); continue
    key = "grad/" + n
    if key in gold:
        e = (g.detach().double().cpu() - torch.from_numpy(gold[key]).double()).norm().item(); kind = "full"
    else:
        e = float(np.sqrt(np.mean((grad_projections(g.detach(), n) - gold["grad_proj"][i]) ** 2))); kind = "proj"
    rows.append((e / max(gold["grad_norm"][i], 1e-12), n, f"{kind} norm {g.norm().item():.4e} ref {gold['grad_norm'][i]:.4e}"))
rows = [r for r in rows if not (r[1].endswith("key.bias") or r[1] == "perspective_mlp.proj.bias")]
rows.sort(reverse=True)
for r in rows[:int(os.environ.get("TOPN", "40"))]: print(f"{r[0]:.3e}  {r[1]}  {r[2]}")
print("...")
for r in rows[-5:]: print(f"{r[0]:.3e}  {r[1]}  {r[2]}")
print("median", np.median([r[0] for r in rows]))
