"""Import the UNMODIFIED reference ``cs_vit`` package from /root/reference.  TEST INFRASTRUCTURE, this container only.

The reference cannot be imported as-is here: ``kornia``, ``peft``, ``smplx`` and ``h5py`` are not installed
(SURVEY.md §8c).  This module registers empty stand-ins for the first three (none of their functions is
reached by ``Poser.predict_batch``), makes ``smplx.create`` return the synthetic MANO layer shared with the
product, imports ``cs_vit.net.ti_poser`` and then removes the ``peft`` stub again because
``transformers.AutoModel`` probes it with ``importlib.util.find_spec``.

Because the reference's package is also called ``cs_vit`` it must not be imported in a process that has
already imported the product package; ``make_goldens.py`` loads the product's pure-host helpers by file path.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("CSVIT_REFERENCE_ROOT", "/root/reference")
_PRODUCT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "cs-vit_b200", "cs_vit")


def load_product_file(relpath: str, name: str):
    """Load a pure-host product module (synthetic inputs, MANO stand-in) without importing ``cs_vit``."""
    spec = importlib.util.spec_from_file_location(name, os.path.join(_PRODUCT, relpath))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def _stub(name: str, **attrs) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def import_reference(mano_factory):
    """Returns the reference's ``cs_vit.net.ti_poser`` module.  ``mano_factory()`` builds the MANO stand-in."""
    if "cs_vit" in sys.modules and not getattr(sys.modules["cs_vit"], "__file__", "").startswith(REFERENCE_ROOT):
        raise RuntimeError("the product `cs_vit` is already imported in this process; run in a fresh interpreter")
    if not os.path.isdir(REFERENCE_ROOT):
        raise FileNotFoundError(f"{REFERENCE_ROOT} not present (goldens are generated in the build container only)")
    k = _stub("kornia")
    k.geometry = _stub("kornia.geometry")
    k.geometry.transform = _stub("kornia.geometry.transform")
    k.augmentation = _stub("kornia.augmentation")
    _stub("peft", LoraConfig=object, get_peft_model=None)
    _stub("peft.peft_model", PeftModel=type("PeftModel", (), {}))
    _stub("smplx", create=lambda *a, **kw: mano_factory())
    if "colorama" not in sys.modules:
        try:
            import colorama  # noqa: F401
        except ImportError:
            _stub("colorama", Fore=types.SimpleNamespace(), Style=types.SimpleNamespace(), init=lambda **kw: None)
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        import cs_vit.net.ti_poser as ref_poser
    finally:
        sys.path.remove(REFERENCE_ROOT)
        sys.modules.pop("peft", None)
        sys.modules.pop("peft.peft_model", None)
    return ref_poser
