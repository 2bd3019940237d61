#!/usr/bin/env python3
"""Profiling target: sort_pairs on C3 (12 modes) or C4 (72 modes):  PROF_CONFIG=C3|C4 python profiles/prof_sort_target.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import brille_b200  # noqa: E402
from brille_b200 import _bridge, host, workloads as W  # noqa: E402

b = host.get()
wl = W.c4_p21c_nest(b) if os.environ.get("PROF_CONFIG", "C3") == "C4" else W.c3_p63mmc(b)
plan = _bridge.sort_plan(wl.grid)
g = brille_b200.accelerate(wl.grid)
for _ in range(3):
    row, col = g.sort_pairs(plan["pairs"], plan)
print("done", row.shape)
