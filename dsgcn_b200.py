"""Importable alias of the package directory `ds-gcn_b200/` (a hyphen is not valid in an identifier)."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("ds-gcn_b200")
sys.modules[__name__] = _pkg
