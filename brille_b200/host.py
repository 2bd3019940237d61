"""Locate brille's own pybind11 module (the host C++ that constructs lattices, zones and grids).

Construction stays brille's: ``brille_b200`` plugs in underneath an existing brille installation.  The module
is looked up in this order: the directory named by ``BRILLE_B200_HOST`` (must contain ``_brille*.so``), an
installed ``brille`` package.
"""
from __future__ import annotations

import importlib
import os
import sys

_host = None


def set_module(module):
    """Use an already imported ``_brille`` module."""
    global _host
    _host = module
    return module


def get():
    global _host
    if _host is not None:
        return _host
    path = os.environ.get("BRILLE_B200_HOST")
    if path:
        if path not in sys.path:
            sys.path.insert(0, path)
        _host = importlib.import_module("_brille")
        return _host
    try:
        _host = importlib.import_module("brille._brille")
    except ImportError as e:
        raise ImportError(
            "brille's host module was not found: install brille or point BRILLE_B200_HOST at the directory "
            "holding _brille*.so"
        ) from e
    return _host
