"""Staged output stores of the pipelined cell kernel (option staged_stores) against the direct stores: C3, 1e7 Q."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import brille_b200
from bench import NQ, Q_SEED, build_workload

wl = build_workload()
g = brille_b200.accelerate(wl.grid)
dQ = torch.from_numpy(wl.make_q(NQ, Q_SEED)).cuda()
vals = torch.empty((NQ, wl.modes, 1), dtype=torch.float64, device="cuda")
vecs = torch.empty((NQ, wl.modes, wl.n_atoms, 3), dtype=torch.complex128, device="cuda")
g.enable_timing(True)
ref = None
for rep in range(2):
    for flag in (0, 1):
        g.set_option("staged_stores", flag)
        vecs.zero_()
        ts = []
        for _ in range(4):
            g.ir_interpolate_at_device(dQ, vals, vecs, check=False)
            ts.append(g.kernel_ms("interpolate"))
        torch.cuda.synchronize()
        if ref is None:
            ref = vecs.clone()
        print(f"staged_stores {flag}: interpolate {min(ts[1:]):.3f} ms  identical={bool(torch.equal(ref, vecs))}", flush=True)
