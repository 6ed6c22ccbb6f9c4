"""Locates the seldq package from inside the drop-in directory (its directory name is not an
identifier, so the drop-in modules cannot use a plain import statement)."""
import importlib
import os
import sys

_PKG_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_ROOT = os.path.dirname(_PKG_DIR)
if _ROOT not in sys.path:
    sys.path.append(_ROOT)
pkg = importlib.import_module(os.path.basename(_PKG_DIR))
