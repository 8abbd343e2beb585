"""Shared test plumbing: rebuild a golden case (weights + inputs from seeds) for the product and the oracle."""
import hashlib
import json
import os
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
OUT_KEYS = ("joint_cam", "verts_cam", "pose_aa", "shape", "root_transl_norm", "root_transl")

_TMP = tempfile.mkdtemp(prefix="csvit_tests_")
_backbone_dirs = {}


def manifest():
    with open(os.path.join(GOLDEN, "MANIFEST.json")) as f:
        return json.load(f)


def backbone_dir(variant: str) -> str:
    from cs_vit.synthetic import make_random_backbone_dir
    if variant not in _backbone_dirs:
        _backbone_dirs[variant] = make_random_backbone_dir(os.path.join(_TMP, variant), variant, seed=0)
    return _backbone_dirs[variant]


def state_checksum(sd) -> str:
    """Digest of a dict of tensors that is exact and independent of reduction order / thread count:
    float tensors are summed as their int32 bit patterns in int64."""
    h = hashlib.sha256()
    for k in sorted(sd):
        v = sd[k].detach().cpu().contiguous()
        bits = v.float().view(torch.int32) if v.is_floating_point() else v
        h.update(k.encode())
        h.update(str(int(bits.to(torch.int64).sum().item())).encode())
        h.update(str(int((bits.to(torch.int64) & 0xFFFF).mul(3).sum().item())).encode())
    return h.hexdigest()[:16]


def build_product(name: str, precision: str = "bf16"):
    """Product ``Poser`` (on CPU; move it yourself) with the golden case's weights, plus inputs and golden outputs."""
    from cs_vit.net import Poser
    from cs_vit.synthetic import make_inputs, randomize_head_
    from cs_vit.utils.mano_standin import SyntheticMANO

    case = manifest()["cases"][name]
    torch.manual_seed(0)
    model = Poser(backbone_dir(case["variant"]), image_size=224, mano_layer=SyntheticMANO(), precision=precision, **case["kwargs"])
    randomize_head_(model, seed=1)
    model.phase(Poser.TrainingPhase(case["phase"]))
    model.eval()
    inputs = make_inputs(case["batch"], case["frames"], 224, seed=11)
    gold = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
    got = state_checksum(model.state_dict())
    want = str(gold["state_checksum"])
    assert got == want, f"seeded weights drifted from the golden run ({got} != {want}): torch RNG changed?"
    assert state_checksum(inputs) == str(gold["input_checksum"]), "seeded inputs drifted from the golden run"
    return model, inputs, gold, case


def head_options(case):
    from cs_vit.synthetic import SWIN_VARIANTS
    from oracle.head_restated import HeadOptions
    _, depths, heads = SWIN_VARIANTS[case["variant"]]
    kw = {k: v for k, v in case["kwargs"].items() if k != "temporal_init_method"}
    return HeadOptions(num_heads=heads[-1], depths=depths, swin_heads=heads, phase=case["phase"], **kw)


def rel(a, b) -> float:
    a = torch.as_tensor(a).double().cpu()
    b = torch.as_tensor(b).double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def train_manifest():
    with open(os.path.join(GOLDEN, "TRAIN_MANIFEST.json")) as f:
        return json.load(f)


def build_train_case(name: str, precision: str = "fp32"):
    """Product ``Poser`` in the golden case's TRAINING phase (as scripts/finetune.py sets it up) + labelled batch + goldens."""
    from cs_vit.net import Poser
    from cs_vit.synthetic import make_inputs, randomize_head_
    from cs_vit.utils.mano_standin import SyntheticMANO

    case = train_manifest()["cases"][name]
    torch.manual_seed(0)
    model = Poser(backbone_dir(case["variant"]), image_size=224, mano_layer=SyntheticMANO(), precision=precision, **case["kwargs"])
    randomize_head_(model, seed=1)
    batch = make_inputs(case["batch"], case["frames"], 224, seed=11, labels="linear_loss_seed" not in case)
    gold = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
    assert state_checksum(model.state_dict()) == str(gold["state_checksum"]), "seeded weights drifted from the golden run"
    assert state_checksum(batch) == str(gold["input_checksum"]), "seeded inputs drifted from the golden run"
    model.phase(Poser.TrainingPhase(case["phase"]))
    return model, batch, gold, case


def grad_projections(g: torch.Tensor, name: str, n_proj: int = 4):
    """Same seeded +-1 projections as oracle/make_train_goldens.py::projections."""
    seed = int.from_bytes(name.encode()[-8:].rjust(8, b"\0"), "little") % (2 ** 31 - 1)
    gen = torch.Generator().manual_seed(seed)
    signs = torch.randint(0, 2, (n_proj, g.numel()), generator=gen, dtype=torch.int8).float() * 2 - 1
    return (signs.double() @ g.reshape(-1).double().cpu()).numpy()
