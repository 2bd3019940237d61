#!/usr/bin/env python3
"""Profiling target: a few device-resident steps of the bench workload (C3, 1e7 Q) for ncu.

    ncu --set full --clock-control none --import-source on -k regex:k_interp -s 1 -c 1 -o gpurun_out/prof python profiles/prof_target.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import brille_b200  # noqa: E402
from bench import NQ, Q_SEED, build_workload  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
nq = int(float(sys.argv[2])) if len(sys.argv) > 2 else NQ
config = os.environ.get("PROF_CONFIG", "C3")  # C3 (default), C2, C4, C3nest, C3mesh: the workload builders of brille_b200.workloads
if config == "C3":
    wl = build_workload()
else:
    from brille_b200 import host as _host, workloads as _W

    wl = _W.BUILDERS[config](_host.get()) if config in _W.BUILDERS else _W.c3_p63mmc(_host.get(), cls={"C3nest": "BZNestQdc", "C3mesh": "BZMeshQdc"}[config])
grid = brille_b200.accelerate(wl.grid)
dQ = torch.from_numpy(wl.make_q(nq, Q_SEED)).cuda()
vals = torch.empty((nq, wl.modes, 1), dtype=torch.float64, device="cuda")
vecs = torch.empty((nq, wl.modes, wl.n_atoms, 3), dtype=torch.complex128, device="cuda")
if len(sys.argv) > 3 and sys.argv[3] in ("sf", "sf0"):  # the structure-factor consumer: fused finish (sf) / k_structure_factor (sf0)
    rng = np.random.default_rng(5)
    grid.set_structure_factor(rng.normal(size=wl.n_atoms) + 1j * rng.normal(size=wl.n_atoms), positions=rng.uniform(0, 1, (wl.n_atoms, 3)),
                              q_transform=rng.normal(size=(3, 3)))
    sf = torch.empty((nq, wl.modes), dtype=torch.float64, device="cuda")
    for _ in range(steps):
        if sys.argv[3] == "sf":
            grid.ir_structure_factor_device(dQ, vals, sf)
        else:
            grid.ir_structure_factor_device(dQ, vals, sf, scratch=vecs, check=False)
else:
    for _ in range(steps):
        grid.ir_interpolate_at_device(dQ, vals, vecs, check=False)
torch.cuda.synchronize()
print("done", steps, nq)
