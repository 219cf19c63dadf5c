"""Importable alias of the package (its directory name carries hyphens): ``import hpusm``."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("humanpose-and-urbanscene-matching_b200")
sys.modules[__name__] = _pkg
