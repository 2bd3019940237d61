"""The output stores of the pipelined cell kernel replayed on their own (k_store_replay) against the kernel itself (C3, 1e7 Q)."""
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
g.set_option("replay_stores", 1)
for _ in range(4):
    g.ir_interpolate_at_device(dQ, vals, vecs, check=False)
    r = g.kernel_ms("replay")
    print(f"store replay {r:.3f} ms = {2304e-9 * NQ / r:.2f} TB/s of eigenvector rows", flush=True)
g.set_option("replay_stores", 0)
g.ir_interpolate_at_device(dQ, vals, vecs, check=False)
print(f"real kernel {g.kernel_ms('interpolate'):.3f} ms")
