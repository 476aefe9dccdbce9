"""Import shims that let the UNMODIFIED reference (/root/reference or a private
copy under baseline/_ref/) import in this image.  Standalone: imports neither mtb200 nor
the kernels' library, so the reference arm of bench.py can use it without loading any
product code.  Users: INTEGRATION.md (binding a reference checkout to the product's
`modules` package), baseline/ref_harness.py (reference arm / CPU baseline),
oracle/gen_golden.py (fixtures, build container), tools/verify_dropin.py.

Why each stub exists:
  matplotlib*            imported at module scope by modules/dynamic_multihead_attention.py:295-297
  src.dataset            star-imported by modules/dynamic_transformer.py:272 and src/utils.py:5;
                         the real one needs fannypack/h5py and a BERT directory
  prettytable            src/utils.py:6
  torchsummary/thop/fvcore.nn   src/train.py:12,24-25 (imported, never called)
  src.models             src/train.py:4 (source file missing upstream)
  transformers.BertModel.from_pretrained   DynamicMULTModel always builds a BertTextEncoder
"""
from __future__ import annotations

import importlib
import os
import sys
import types


def _stub(name: str, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def find_reference() -> str | None:
    here = os.path.dirname(os.path.abspath(__file__))
    for p in (os.environ.get("MTB_REFERENCE_ROOT"), "/root/reference",
              os.path.join(os.path.dirname(here), "baseline", "_ref")):
        if p and os.path.isfile(os.path.join(p, "src", "dynamic_models2.py")):
            return p
    return None


def install(ref_root: str | None = None, use_product_modules: bool = False) -> str:
    """Put the reference on sys.path (optionally with the product's ``modules``
    package shadowing the reference's) and register the stubs.  Returns the root."""
    ref_root = ref_root or find_reference()
    if ref_root is None:
        raise RuntimeError("reference sources not found (need /root/reference or baseline/_ref)")
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.ticker", "matplotlib.cm"):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                _stub(name, cm=None, LinearLocator=None)
    if "matplotlib" in sys.modules and not hasattr(sys.modules["matplotlib"], "cm"):
        sys.modules["matplotlib"].cm = None
    for name in ("prettytable", "torchsummary", "thop"):
        if name not in sys.modules:
            _stub(name, PrettyTable=object, summary=lambda *a, **k: None, profile=lambda *a, **k: None)
    if "fvcore" not in sys.modules:
        _stub("fvcore")
        _stub("fvcore.nn", FlopCountAnalysis=object, parameter_count_table=lambda *a, **k: "")
    try:
        import torchvision  # noqa: F401  (reference imports it at module scope)
    except Exception:
        _stub("torchvision", models=None)
        _stub("torchvision.models")
    import transformers

    class _DummyBert:  # stands in for the text front-end (out of scope)
        @classmethod
        def from_pretrained(cls, *a, **k):
            import torch
            return torch.nn.Identity()

    transformers.BertModel = _DummyBert
    if use_product_modules:
        prod = os.path.dirname(os.path.abspath(__file__))
        sys.path.insert(0, prod)
        sys.path.insert(1, ref_root)
    else:
        sys.path.insert(0, ref_root)
    # src is a namespace package upstream (no __init__.py); stub the two broken members
    _stub("src.dataset", **{n: object for n in ("MOSEI_Datasets", "avMNIST_Datasets", "GentlePush_Datasets",
                                                 "Enrico_Datasets", "EEG2a_Datasets")})
    _stub("src.models")
    return ref_root


def patch_train_module():
    """src/train.py:53 passes verbose= to ReduceLROnPlateau (removed in torch 2.11)."""
    import src.train as T
    from torch.optim.lr_scheduler import ReduceLROnPlateau as _R

    class _Compat(_R):
        def __init__(self, *a, verbose=None, **k):
            super().__init__(*a, **k)

    T.ReduceLROnPlateau = _Compat
    return T
