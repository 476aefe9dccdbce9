"""Checkpoint interchange with the reference.

The reference checkpoints by pickling the WHOLE model object -- ``torch.save(model, path)`` in src/train.py:508-511 -- and
EA.py:264 reads it back with ``torch.load``.  Such a pickle names its classes by module path
(``src.dynamic_models2.DynamicMULTModel``, ``src.models2.ModalityStr``, ``modules.dynamic_transformer...``).  The product's
``modules`` package already has the reference's paths; this file covers the two ``src.*`` modules:

* ``install_reference_aliases()`` registers ``src.dynamic_models2`` / ``src.models2`` as aliases of the product modules
  when no reference checkout is importable, so a reference-written checkpoint loads straight into the product classes
  (``DynamicMULTModel.__setstate__`` back-fills the three attributes the reference does not have);
* ``save_reference_checkpoint(model, path)`` writes a whole-model pickle whose class paths are the REFERENCE's, so the
  unmodified reference (EA.py:264) can ``torch.load`` a model trained here.  A sequence-preserving ``Conv1x1FrontEnd`` is
  exported as the equivalent ``Sequential(Transpose(1, 2), Conv1d(k=1, bias=False))`` (the reference's own, commented-out,
  front-end variant).

``state_dict()`` keys are identical on both sides, so ``load_state_dict`` interchange needs none of this."""
from __future__ import annotations

import contextlib
import copy
import sys
import types

import torch
from torch import nn

_ALIASES = {"src.dynamic_models2": "mtb200.dynamic_models2", "src.models2": "mtb200.models2"}


def install_reference_aliases(force: bool = False) -> bool:
    """Make ``src.dynamic_models2`` / ``src.models2`` resolve to the product modules.  No-op (returns False) when a real
    reference checkout already provides them, unless ``force``."""
    import importlib
    if not force:
        try:
            mod = importlib.import_module("src.dynamic_models2")
            if not getattr(mod, "__mtb200_alias__", False):
                return False
        except Exception:
            pass
    pkg = sys.modules.get("src")
    if pkg is None or force:
        pkg = types.ModuleType("src")
        pkg.__path__ = []
        sys.modules["src"] = pkg
    for alias, target in _ALIASES.items():
        mod = importlib.import_module(target)
        mod.__mtb200_alias__ = True
        sys.modules[alias] = mod
        setattr(pkg, alias.split(".")[1], mod)
    return True


def load_reference_checkpoint(path, map_location=None):
    """``torch.load`` of a whole-model pickle written by the reference's ``torch.save(model)`` (or by
    ``save_reference_checkpoint``) into the product classes."""
    install_reference_aliases()
    return torch.load(path, map_location=map_location, weights_only=False)


@contextlib.contextmanager
def _reference_class_paths():
    """pickle looks classes up as sys.modules[cls.__module__].<qualname>: point the product classes at the alias modules"""
    from . import dynamic_models2 as dm, models2 as m2
    install_reference_aliases(force=True)
    classes = [c for mod in (dm, m2) for c in vars(mod).values()
               if isinstance(c, type) and c.__module__ in (dm.__name__, m2.__name__)]
    saved = [(c, c.__module__) for c in classes]
    try:
        for c in classes:
            c.__module__ = "src.dynamic_models2" if saved[classes.index(c)][1] == dm.__name__ else "src.models2"
        yield
    finally:
        for c, name in saved:
            c.__module__ = name


def save_reference_checkpoint(model, path):
    """Whole-model pickle the unmodified reference can ``torch.load`` (class paths ``src.*`` / ``modules.*`` only)."""
    from .dynamic_models2 import Conv1x1FrontEnd, Transpose
    m = copy.deepcopy(model)                   # __getstate__ drops the plan executor (raw device pointers)
    proj = []
    for fe in m.proj:
        if isinstance(fe, Conv1x1FrontEnd):
            conv = nn.Conv1d(fe.d_in, fe.d, kernel_size=1, bias=False)
            conv.weight = fe.weight
            proj.append(nn.Sequential(Transpose(1, 2), conv))
        else:
            proj.append(fe)
    m.proj = nn.ModuleList(proj)
    m.__dict__.pop("_outside_cache", None)
    for mod in m.modules():                    # the plan executor's access-path mirrors (engine._install_fast_attrs)
        d = mod.__dict__
        for k in [k for k in d if k in mod._modules or k in mod._parameters or k in ("_ll", "_lns")]:
            del d[k]
    with _reference_class_paths():
        torch.save(m, path)
