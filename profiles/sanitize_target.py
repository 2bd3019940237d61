#!/usr/bin/env python3
"""Small run of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool memcheck python profiles/sanitize_target.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import brille_b200  # noqa: E402
from brille_b200 import _bridge, capi, host, workloads as W  # noqa: E402

b = host.get()
rng = np.random.default_rng(0)
for cls in ("BZTrellisQdc", "BZNestQdc", "BZMeshQdc"):
    wl = W.c3_p63mmc(b, density=150, seed=5, cls=cls)
    g = brille_b200.accelerate(wl.grid)
    Q = wl.make_q(60_000, 1)
    Q[:10] = 0.0
    v, w, pr = g.ir_interpolate_at(Q, probe=True)
    v2, w2 = g.ir_interpolate_at(Q)
    assert np.array_equal(v, v2) and np.array_equal(w, w2)
    g.set_option("cell_kernel", 1)
    v3, w3 = g.ir_interpolate_at(Q)
    g.set_option("cell_kernel", 0)
    g.set_option("interp_path", 1)
    g.ir_interpolate_at(Q[:5000])
    g.set_option("interp_path", 0)
    g.set_structure_factor(rng.normal(size=4) + 1j * rng.normal(size=4), positions=rng.uniform(0, 1, (4, 3)),
                           q_transform=np.asarray(_bridge.flatten_bz(wl.bz)["to_xyz"]).reshape(3, 3))
    g.ir_structure_factor(Q)
    g.set_option("sf_fused", 0)
    g.ir_structure_factor(Q[:20000])
    g.set_option("sf_fused", 1)
    h, c = g.ir_powder_sweep((0.2, 6.0), 16, (0.0, 52.0), 32, 3000, seed=1, weight=1)
    g.ir_powder_bin(Q[:30000], (0.2, 6.0), 16, (0.0, 52.0), 32)
    g.moveinto(Q[:1000]); g.ir_moveinto(Q[:1000]); g.ir_moveinto_wedge(Q[:1000]); g.isinside(Q[:1000])
    if cls == "BZTrellisQdc":
        g.sort()
        g.ir_interpolate_at(Q[:30000])
    print(cls, "ok", float(v.sum()), float(h.sum()), flush=True)
    g.close()
wl = W.c4_p21c_nest(b, density=40)
g = brille_b200.accelerate(wl.grid)
plan = _bridge.sort_plan(wl.grid)
g.sort_pairs(plan["pairs"][:300], plan)
g.ir_interpolate_at(wl.make_q(20000, 2))
g.close()
capi.solve_assignments(rng.integers(0, 3, (40, 12, 12)).astype(float))
capi.solve_assignments(rng.integers(0, 3, (6, 72, 72)).astype(float))
print("done")
