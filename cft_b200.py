"""Importable alias of the package directory ``multispectral-object-detection_b200`` (hyphenated
names cannot appear in an ``import`` statement):  ``import cft_b200 as cft; cft.Model(...)``."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("multispectral-object-detection_b200")
sys.modules[__name__] = _pkg
