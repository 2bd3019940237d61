#!/usr/bin/env python3
"""sort(): the device path next to the reference's OpenMP sort() on the host cores, same grid and data."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import brille_b200  # noqa: E402
from brille_b200 import workloads as W  # noqa: E402
from brille_b200.grid import _bridge  # noqa: E402
from brille_b200 import host as _hostmod  # noqa: E402

host, bridge = _hostmod.get(), _bridge()
for name, wl in (("C3 (12 modes)", W.c3_p63mmc(host)), ("C4 (72 modes, nest)", W.c4_p21c_nest(host))):
    plan = bridge.sort_plan(wl.grid)
    g = brille_b200.accelerate(wl.grid)
    g.sort_pairs(plan["pairs"], plan)  # warm-up at full size (work space, kernel attributes)
    t_dev = 1e9
    for _ in range(3):
        t0 = time.perf_counter()
        row, col = g.sort_pairs(plan["pairs"], plan)
        t_dev = min(t_dev, time.perf_counter() - t0)
    t0 = time.perf_counter()
    g.sort()
    t_all = time.perf_counter() - t0
    t0 = time.perf_counter()
    wl.grid.sort()
    t_ref = time.perf_counter() - t0
    refp = bridge.pair_permutations(wl.grid, plan["pairs"])
    same = (row.astype(np.uint32) == refp[:, 0, :]).all(axis=1).mean()
    print(f"{name}: {len(row)} pairs | device sort_pairs {t_dev*1e3:.1f} ms (host buffers in/out), B200Grid.sort() incl. table rebuild "
          f"{t_all*1e3:.1f} ms | reference sort() on {os.cpu_count()} cores {t_ref*1e3:.1f} ms | x{t_ref/t_dev:.0f} | identical pairs {100*same:.2f} %", flush=True)
    g.close()
