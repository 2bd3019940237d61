"""Wall time of grid.sort() through the drop-in classes (device: cost matrices + assignments on the GPU, brille's permutation
table updated on the host) against brille's own OpenMP sort() on the same object, C3 (12 modes) and C4 (72 modes) sizes."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from brille_b200 import host as H, workloads as W, _accel, _bridge  # noqa: E402

b = H.get()
for name, build in (("C3 trellis, 12 modes", lambda: W.c3_p63mmc(b)), ("C4 nest, 72 modes", lambda: W.c4_p21c_nest(b))):
    wl = build()
    cls = getattr(_accel, type(wl.grid).__name__)
    g = cls(wl.grid)                      # shares the data of the host object
    n_pairs = len(_bridge.sort_plan(g)["pairs"])
    g.ir_interpolate_at(np.zeros((1, 3)))  # device tables up: sort() times the sort, not the first upload
    ref = g.host()
    t = []
    for _ in range(3):
        t0 = time.perf_counter()
        g.sort()
        t.append(time.perf_counter() - t0)
    t0 = time.perf_counter()
    ref.sort()
    t_ref = time.perf_counter() - t0
    pairs = _bridge.sort_plan(ref)["pairs"]
    same = (_bridge.pair_permutations(g, pairs) == _bridge.pair_permutations(ref, pairs)).all(axis=(1, 2)).mean()
    print(f"{name}: {n_pairs} pairs; drop-in sort() {min(t) * 1e3:.1f} ms (runs {[round(x * 1e3, 1) for x in t]}), brille host sort() on "
          f"{os.cpu_count()} cores {t_ref:.2f} s -> {t_ref / min(t):.0f}x; identical permutations for {same * 100:.2f} % of the pairs")
