"""Import shim for the UNMODIFIED reference (build container only; TEST INFRASTRUCTURE).

The reference's ``models/common.py:16-18`` imports matplotlib/seaborn through
``utils.plots``/``utils.metrics``; neither is installed.  Four empty ``sys.modules`` stubs
make it importable.  Nothing under ``/root/reference`` is modified or copied.
The GPU box has no ``/root/reference``: only ``oracle/make_golden.py`` and the ``-m "not gpu"``
cross-check tests (which skip when the tree is absent) use this shim.
"""
import os
import sys
import types

REF_ROOT = os.environ.get("CFT_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "models", "yolo_test.py"))


def import_reference():
    """Returns the reference's ``models.yolo_test`` module."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    os.environ.setdefault("PYTHONDONTWRITEBYTECODE", "1")
    sys.dont_write_bytecode = True
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.colors", "seaborn"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    mpl = sys.modules["matplotlib"]
    mpl.pyplot = sys.modules["matplotlib.pyplot"]
    mpl.colors = sys.modules["matplotlib.colors"]
    mpl.use = lambda *a, **k: None
    mpl.rc = lambda *a, **k: None
    if not hasattr(mpl.colors, "TABLEAU_COLORS"):
        mpl.colors.TABLEAU_COLORS = {"tab:blue": "#1f77b4"}
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import logging
    logging.disable(logging.INFO)
    import models.yolo_test as yt  # noqa: E402
    return yt


def reference_yaml(name: str) -> str:
    return os.path.join(REF_ROOT, "models", "transformer", name + ".yaml")
