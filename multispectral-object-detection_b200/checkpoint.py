"""Checkpoint ingestion: the reference's ``attempt_load`` (``models/experimental.py:113-134``)
for two-stream CFT checkpoints, ending in a B200 ``Model``.

The reference saves whole pickled modules (``train.py:850-857``: ``{'model': deepcopy(model).half(), 'ema': ...}``), so
``torch.load`` must be able to resolve ``models.yolo_test.Model``, ``models.common.Conv`` ... by name.  Two cases:

* the reference tree is importable (``models.yolo_test`` on ``sys.path``): the checkpoint unpickles into the reference's
  own classes;
* it is not (the GPU box, a deployment image): ``reference_aliases()`` registers stand-in modules ``models``,
  ``models.common``, ``models.yolo_test``, ``models.experimental`` whose attributes are the drop-in classes of
  ``modules.py`` / ``model.py`` (same names, same sub-module layout), for the duration of the load only.

Either way the unpickled object is walked layer by layer (``model._convert_module``: fresh B200 modules, weights copied
through ``load_state_dict``), BN is optionally folded (``Model.fuse``, the reference's ``.fuse()``) and the result is a
stand-alone ``Model`` with the buffer plan of ``_plan_graph``.  No arithmetic happens here: CPU-side, device independent.
"""
from __future__ import annotations

import contextlib
import importlib
import sys
import types
from typing import Optional

import torch
import torch.nn as nn

from . import model as _model
from . import modules as M
from ._lib import CftError

_ALIAS_NAMES = ("models", "models.common", "models.yolo_test", "models.yolo", "models.experimental")


class _PickledModel(nn.Module):
    """Stand-in for the reference's ``Model`` class while unpickling: only its attribute dictionary is needed."""


def _reference_importable() -> bool:
    try:
        mod = importlib.import_module("models.yolo_test")
        return hasattr(mod, "parse_model") and not getattr(mod, "_cft_alias", False)
    except Exception:
        return False


@contextlib.contextmanager
def reference_aliases():
    """Temporarily provide ``models.*`` modules whose classes are the B200 drop-ins, so that a reference checkpoint
    unpickles without the reference tree."""
    saved = {n: sys.modules.get(n) for n in _ALIAS_NAMES}
    try:
        pkg = types.ModuleType("models")
        pkg.__path__ = []                                   # a package: 'models.common' is a submodule
        pkg._cft_alias = True
        common = types.ModuleType("models.common")
        for name in ("Conv", "Focus", "Bottleneck", "C3", "SPP", "Concat", "Add", "Add2", "GPT", "SelfAttention",
                     "myTransformerBlock"):
            setattr(common, name, getattr(M, name))
        yolo = types.ModuleType("models.yolo_test")
        yolo.Model, yolo.Detect = _PickledModel, M.Detect
        yolo1 = types.ModuleType("models.yolo")             # single-stream file: same Detect class name in old checkpoints
        yolo1.Model, yolo1.Detect = _PickledModel, M.Detect
        exp = types.ModuleType("models.experimental")
        for m in (common, yolo, yolo1, exp):
            m._cft_alias = True
        pkg.common, pkg.yolo_test, pkg.yolo, pkg.experimental = common, yolo, yolo1, exp
        sys.modules.update({"models": pkg, "models.common": common, "models.yolo_test": yolo, "models.yolo": yolo1,
                            "models.experimental": exp})
        yield
    finally:
        for n, mod in saved.items():
            if mod is None:
                sys.modules.pop(n, None)
            else:
                sys.modules[n] = mod


def from_reference_model(ref_model: nn.Module) -> "_model.Model":
    """A stand-alone B200 ``Model`` from a built / unpickled reference two-stream model (or its stand-in)."""
    if not hasattr(ref_model, "model") or not hasattr(ref_model, "save"):
        raise CftError("from_reference_model: not a two-stream Model (needs .model and .save)")
    layers = []
    for m in ref_model.model:
        new = _model._convert_module(m)
        for attr in ("i", "f", "type", "np"):
            setattr(new, attr, getattr(m, attr))
        layers.append(new)
    out = _model.Model.__new__(_model.Model)
    nn.Module.__init__(out)
    out.yaml = getattr(ref_model, "yaml", None)
    out.model = nn.Sequential(*layers)
    out.save = list(ref_model.save)
    det = out.model[-1]
    out.names = list(getattr(ref_model, "names", [str(i) for i in range(getattr(det, "nc", 0))]))
    if isinstance(det, M.Detect):
        out.stride = det.stride
    # BatchNorm eps / momentum (utils/torch_utils.py:144-153 sets 1e-3 / 0.03 in Model.__init__) travel with the pickled
    # modules and are copied by _convert_module
    out._plan = _model._plan_graph(out.model)
    return out


def attempt_load(weights, map_location=None, fuse: bool = True, use_ema: Optional[bool] = None) -> "_model.Model":
    """``models/experimental.py:113-134`` for ONE checkpoint file: ``ckpt['ema' or 'model'].float().fuse().eval()`` as
    a B200 ``Model``.  Ensembles (a list of weights) are outside the two-stream hot path."""
    if isinstance(weights, (list, tuple)):
        if len(weights) != 1:
            raise CftError("attempt_load: model ensembles are outside the CFTx3 hot path (one checkpoint at a time)")
        weights = weights[0]
    ctx = contextlib.nullcontext() if _reference_importable() else reference_aliases()
    with ctx:
        ckpt = torch.load(weights, map_location=map_location or "cpu", weights_only=False)
    if isinstance(ckpt, dict):
        key = "ema" if (ckpt.get("ema") is not None if use_ema is None else use_ema) else "model"
        ref = ckpt.get(key)
        if ref is None:
            raise CftError(f"attempt_load: checkpoint has no '{key}' entry")
    else:
        ref = ckpt
    model = from_reference_model(ref.float())
    if fuse:
        model.fuse()
    return model.eval()
