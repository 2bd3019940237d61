"""Interpolation-kernel time against the number of points per call (C3): does the size of the output window matter?"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import brille_b200
from bench import Q_SEED, build_workload

wl = build_workload()
g = brille_b200.accelerate(wl.grid)
NMAX = 10_000_000
Q = wl.make_q(NMAX, Q_SEED)
dQ = torch.from_numpy(Q).cuda()
vals = torch.empty((NMAX, wl.modes, 1), dtype=torch.float64, device="cuda")
vecs = torch.empty((NMAX, wl.modes, wl.n_atoms, 3), dtype=torch.complex128, device="cuda")
g.enable_timing(True)
for n in (500_000, 1_000_000, 2_000_000, 4_000_000, 10_000_000):
    ts, tl = [], []
    for _ in range(4):
        g.ir_interpolate_at_device(dQ[:n], vals[:n], vecs[:n], check=False)
        ts.append(g.kernel_ms("interpolate")); tl.append(g.kernel_ms("locate"))
    t = min(ts[1:])
    print(f"n {n:>9}: interpolate {t:.3f} ms = {2424*n/t/1e9:.2f} TB/s   locate {min(tl[1:]):.3f} ms ({min(tl[1:])/n*1e7:.2f} per 1e7)", flush=True)
