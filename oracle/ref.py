"""Locate the reference build made by oracle/build_ref.sh (oracle/_ref/brille_host).

Test infrastructure only.  ``host()`` returns brille's own pybind11 module ``_brille`` (the UNMODIFIED
reference, compiled from /root/reference) and ``probe()`` the internals probe of oracle/probe.cpp.
"""
from __future__ import annotations

import importlib
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
HOST_DIR = os.path.join(HERE, "_ref", "brille_host")


def available() -> bool:
    return os.path.isdir(HOST_DIR) and any(f.startswith("_brille") and f.endswith(".so") for f in os.listdir(HOST_DIR))


def _load(name):
    if not available():
        raise ImportError(
            "reference build not found under oracle/_ref/brille_host; run oracle/build_ref.sh "
            "(needs /root/reference) or __graft_entry__.build()"
        )
    if HOST_DIR not in sys.path:
        sys.path.insert(0, HOST_DIR)
    return importlib.import_module(name)


def host():
    return _load("_brille")


def probe():
    host()
    return _load("_probe")
