"""Locate brille's own pybind11 module (the host C++ that constructs lattices, zones and grids).

Construction stays brille's: ``brille_b200`` plugs in underneath an existing brille installation.  The module
is looked up in this order: the directory named by ``BRILLE_B200_HOST`` (must contain ``_brille*.so``), an
installed ``brille`` package, the build of brille's unmodified sources under ``third_party/brille_host``
(``third_party/build_brille_host.sh``; this image has no brille installed).
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
    if "_brille" in sys.modules:  # (already imported from a directory on sys.path)
        _host = sys.modules["_brille"]
        return _host
    try:
        _host = importlib.import_module("brille._brille")
        if not hasattr(_host, "BrillouinZone"):  # (brille_b200's own shim package: it wraps the host module)
            _host = importlib.import_module("_brille")
    except ImportError:
        local = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "third_party", "brille_host")
        if os.path.isdir(local) and any(f.startswith("_brille") and f.endswith(".so") for f in os.listdir(local)):
            if local not in sys.path:
                sys.path.insert(0, local)
            _host = importlib.import_module("_brille")
        else:
            raise ImportError(
                "brille's host module was not found: install brille, point BRILLE_B200_HOST at the directory holding "
                "_brille*.so, or build it from brille's sources with third_party/build_brille_host.sh"
            ) from None
    return _host
