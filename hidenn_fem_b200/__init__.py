"""Importable name for the product package.

The sources live in ``hidenn-fem_b200/`` (the layout the build contract names); a hyphen
cannot appear in a Python import, so this package extends its search path with that
directory: ``import hidenn_fem_b200.models`` loads ``hidenn-fem_b200/models.py``.
"""
import os as _os

_SRC = _os.path.normpath(_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "..", "hidenn-fem_b200"))
__path__.append(_SRC)
SRC_DIR = _SRC

__all__ = ["SRC_DIR"]
