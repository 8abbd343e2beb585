"""``cs_vit.dataset`` - the import surface ``scripts/eval.py`` / ``scripts/finetune.py`` need
(ref:cs_vit/dataset/__init__.py: ``InterHand26MSeq``, ``HO3D``, ``DexYCB``; ref:scripts/eval.py:19-22, 96-134;
ref:scripts/finetune.py:20, 66-110).

The data sets themselves (h5 readers, OpenCV decode, kornia crops) are outside the hot path this package
rebuilds ("out of scope"), but the scripts cannot even be imported without these names, so
each name is served from one of two sources, chosen by the ``root`` the script passes:

* ``root`` is ``"synthetic"`` or ``"synthetic:<N>"`` (or the directory does not exist and
  ``CSVIT_SYNTHETIC_DATA=1``): a deterministic synthetic data set of N clips with exactly the sample dict of
  ref:cs_vit/dataset/DexYCB.py:229-246 (``patches``, ``square_bboxes``, ``timestamp``, ``focal``, ``princpt``,
  ``joint_cam``, ``joint_img``, ``joint_valid``, ``mano_pose``, ``mano_shape``, ``imgs_path``, ``flip`` ...), so
  the unmodified scripts run end to end on a box that has no data (this is what BASELINE.json benchmarks:
  synthetic hand crops).
* anything else: the reference's own data set class, loaded from the reference checkout named by
  ``CSVIT_REFERENCE_ROOT`` (its ``cs_vit/dataset/<name>.py`` is executed as a submodule of THIS package, so its
  relative imports ``..utils.img`` / ``..utils.geometry`` / ``..constants`` resolve to this package's
  equivalents).  That path needs the reference's own dependencies (h5py, kornia, cv2) and real data.

``collate_fn`` is the reference's (ref:cs_vit/dataset/InterHand26M/InterHand26MSeq.py:21-34): python lists for
``imgs_path`` / ``flip``, ``torch.stack`` for everything else.
"""
from __future__ import annotations

import importlib.util
import os
import sys
from typing import Dict, List, Union

import torch
from torch.utils.data.dataset import Dataset

__all__ = ["InterHand26MSeq", "HO3D", "DexYCB", "SyntheticHandSeq", "collate_fn"]


def collate_fn(batch: List[Dict[str, torch.Tensor]]) -> Dict[str, Union[torch.Tensor, list]]:
    out = {}
    for key in batch[0].keys():
        if key in ("imgs_path", "flip"):
            out[key] = [sample[key] for sample in batch]
        else:
            out[key] = torch.stack([sample[key].contiguous() for sample in batch], dim=0)
    return out


class SyntheticHandSeq(Dataset):
    """N deterministic synthetic clips with the reference data sets' sample layout (all tensors per frame, ``[T, ...]``).

    Sample ``i`` is a pure function of ``(seed, i)``: DexYCB/HO3D-like intrinsics, a hand-sized square box around a
    random image position, U[0,1) crops (the backbone normalises them itself, ref:cs_vit/net/ti_poser.py:239-243),
    camera-space joints in millimetres half a metre from the camera and their pinhole projection."""

    def __init__(self, num_frames: int, length: int = 256, img_size: int = 224, expansion_ratio: float = 1.25,
                 name: str = "synthetic", seed: int = 0):
        self.num_frames, self.length, self.img_size = int(num_frames), int(length), int(img_size)
        self.expansion_ratio, self.name, self.seed = float(expansion_ratio), name, int(seed)

    def __len__(self) -> int:
        return self.length

    def __getitem__(self, ix: int) -> Dict[str, Union[torch.Tensor, list, bool]]:
        if not 0 <= ix < self.length:
            raise IndexError(ix)
        T, S = self.num_frames, self.img_size
        g = torch.Generator().manual_seed(self.seed * 1_000_003 + ix)
        focal = torch.tensor([617.0, 617.0]).expand(T, 2).contiguous()
        princpt = torch.tensor([312.0, 241.0]).expand(T, 2).contiguous()
        joint_cam = torch.randn(T, 21, 3, generator=g) * 30.0 + torch.tensor([0.0, 0.0, 500.0])
        joint_img = focal[:, None] * joint_cam[..., :2] / joint_cam[..., 2:] + princpt[:, None]
        lo, hi = joint_img.min(dim=1).values, joint_img.max(dim=1).values
        bbox_tight = torch.cat([lo, hi], dim=-1)
        centre, side = (lo + hi) / 2, (hi - lo).max(dim=-1, keepdim=True).values * self.expansion_ratio
        square = torch.cat([centre - side / 2, centre + side / 2], dim=-1)
        return {
            "imgs_path": [f"{self.name}/{ix:08d}/{t:04d}.jpg" for t in range(T)],
            "flip": False,
            "rot_rad": torch.zeros(T),
            "patches": torch.rand(T, 3, S, S, generator=g),
            "square_bboxes": square,
            "bbox_tight": bbox_tight,
            "joint_img": joint_img,
            "joint_bbox_img": joint_img - bbox_tight[:, None, :2],
            "joint_cam": joint_cam,
            "joint_valid": torch.ones(T, 21),
            "joint_rel": joint_cam - joint_cam[:, :1],
            "mano_pose": torch.randn(T, 48, generator=g) * 0.2,
            "mano_shape": (torch.randn(10, generator=g) * 0.5)[None].repeat(T, 1),
            "timestamp": torch.arange(T, dtype=torch.float32) * 33.333,
            "focal": focal,
            "princpt": princpt,
        }


def _synthetic_length(root) -> int:
    """N of ``"synthetic:<N>"``; 256 for plain ``"synthetic"``; -1 when ``root`` names real data."""
    root = "" if root is None else str(root)
    if root == "synthetic":
        return 256
    if root.startswith("synthetic:"):
        return int(root.split(":", 1)[1])
    if os.environ.get("CSVIT_SYNTHETIC_DATA") == "1" and not os.path.isdir(root):
        return 256
    return -1


def _load_reference_class(relpath: str, cls_name: str):
    ref_root = os.environ.get("CSVIT_REFERENCE_ROOT")
    path = os.path.join(ref_root, "cs_vit", "dataset", relpath) if ref_root else None
    if not path or not os.path.exists(path):
        raise FileNotFoundError(
            f"cs_vit.dataset.{cls_name}: real data needs the reference's data set code - set CSVIT_REFERENCE_ROOT to a CS-ViT checkout "
            f"(looked for {path}); or pass root='synthetic[:N]' for the synthetic data set")
    sub = relpath[:-3].replace("/", ".")
    name = f"{__name__}._ref.{sub}"
    if name in sys.modules:
        return getattr(sys.modules[name], cls_name)
    # executed as a submodule at the reference's own depth (cs_vit.dataset.X / cs_vit.dataset.InterHand26M.X), so its relative
    # imports (..utils.img, ...constants) land on this package
    depth_pkg = __name__ if "/" not in relpath else f"{__name__}.InterHand26M"
    if depth_pkg != __name__ and depth_pkg not in sys.modules:
        import types
        pkg = types.ModuleType(depth_pkg)
        pkg.__path__ = []
        sys.modules[depth_pkg] = pkg
    spec = importlib.util.spec_from_file_location(f"{depth_pkg}.{os.path.basename(relpath)[:-3]}", path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[spec.name] = mod
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return getattr(mod, cls_name)


def _dataset_name(cls_name: str, relpath: str, doc: str):
    class _Dispatch(Dataset):
        collate_fn = staticmethod(collate_fn)

        def __new__(cls, root=None, num_frames: int = 1, *args, **kwargs):
            n = _synthetic_length(root)
            if n >= 0:
                return SyntheticHandSeq(num_frames, n, kwargs.get("img_size", 224), kwargs.get("expansion_ratio", 1.25),
                                        name=f"synthetic_{cls_name.lower()}")
            return _load_reference_class(relpath, cls_name)(root, num_frames, *args, **kwargs)

    _Dispatch.__name__ = _Dispatch.__qualname__ = cls_name
    _Dispatch.__doc__ = doc
    return _Dispatch


InterHand26MSeq = _dataset_name("InterHand26MSeq", "InterHand26M/InterHand26MSeq.py",
                                "ref:cs_vit/dataset/InterHand26M/InterHand26MSeq.py:20 (root, num_frames, data_split, img_size, expansion_ratio)")
HO3D = _dataset_name("HO3D", "HO3D.py", "ref:cs_vit/dataset/HO3D.py (root, num_frames, data_split, img_size, expansion_ratio)")
DexYCB = _dataset_name("DexYCB", "DexYCB.py", "ref:cs_vit/dataset/DexYCB.py:18 (root, num_frames, protocol, data_split, img_size, expansion_ratio)")
