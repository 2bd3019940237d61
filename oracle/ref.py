"""Locate the reference build: brille's own, unmodified host module (third_party/brille_host, built by
third_party/build_brille_host.sh) in its role as parity oracle and CPU baseline, and the probe of oracle/probe.cpp.

Test infrastructure only.  ``host()`` returns brille's pybind11 module ``_brille`` and ``probe()`` the internals probe
(oracle/_ref/_probe*.so, built by oracle/build_ref.sh).
"""
from __future__ import annotations

import importlib
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
HOST_DIR = os.path.join(os.path.dirname(HERE), "third_party", "brille_host")
PROBE_DIR = os.path.join(HERE, "_ref")


def _has(d, stem):
    return os.path.isdir(d) and any(f.startswith(stem) and f.endswith(".so") for f in os.listdir(d))


def available() -> bool:
    return _has(HOST_DIR, "_brille") and _has(PROBE_DIR, "_probe")


def _load(name, d):
    if not available():
        raise ImportError(
            "reference build not found (third_party/brille_host, oracle/_ref); run oracle/build_ref.sh "
            "(needs /root/reference) or __graft_entry__.build()"
        )
    if d not in sys.path:
        sys.path.insert(0, d)
    return importlib.import_module(name)


def host():
    return _load("_brille", HOST_DIR)


def probe():
    host()
    return _load("_probe", PROBE_DIR)
