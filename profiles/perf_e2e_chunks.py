import sys, os, time
import numpy as np
sys.path.insert(0, "/root/repo")
import brille_b200
from bench import build_workload, Q_SEED
wl = build_workload()
g = brille_b200.accelerate(wl.grid)
ne = 2_000_000
Q = wl.make_q(ne, Q_SEED)
hq = brille_b200.PinnedArray((ne, 3), np.float64); hq.array[:] = Q
hv = brille_b200.PinnedArray((ne, wl.modes, 1), np.float64)
hw = brille_b200.PinnedArray((ne, wl.modes, wl.n_atoms, 3), np.complex128)
for hc in (0, 2_000_000, 1_000_000, 500_000, 250_000):
    g.set_option("host_chunk", hc)
    for _ in range(2): g.ir_interpolate_at(hq.array, out=(hv.array, hw.array))
    t0 = time.perf_counter()
    for _ in range(5): g.ir_interpolate_at(hq.array, out=(hv.array, hw.array))
    dt = (time.perf_counter() - t0) / 5
    print(f"host_chunk {hc:>8}: {dt*1e3:7.2f} ms  {ne/dt:.3e} Q/s  D2H {ne*2400/dt/1e9:.1f} GB/s", flush=True)
