"""Importable alias of the package directory ``w-ofdm-optimization_b200`` (hyphens are not valid
in an ``import`` statement): ``import wofdm_b200`` gives that package."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("w-ofdm-optimization_b200")
sys.modules[__name__] = _pkg
