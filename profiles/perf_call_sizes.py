"""What a user sees: wall time of ir_interpolate_at(Q) through the host-buffer API against the number of points, with pageable
numpy arrays (the default) and with page-locked ones (pinned=True / out=...), next to the reference on the host cores (C3)."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import brille_b200
from bench import build_workload

wl = build_workload()
g = brille_b200.accelerate(wl.grid)
cores = os.cpu_count()
for n in (1_000, 10_000, 100_000, 1_000_000, 4_000_000):
    Q = wl.make_q(n, 7)
    g.ir_interpolate_at(Q)
    t0 = time.perf_counter(); reps = 3
    for _ in range(reps): v, w = g.ir_interpolate_at(Q)
    t_page = (time.perf_counter() - t0) / reps
    hq = brille_b200.PinnedArray((n, 3), np.float64); hq.array[:] = Q
    hv = brille_b200.PinnedArray((n, wl.modes, 1), np.float64)
    hw = brille_b200.PinnedArray((n, wl.modes, wl.n_atoms, 3), np.complex128)
    g.ir_interpolate_at(hq.array, out=(hv.array, hw.array))
    t0 = time.perf_counter()
    for _ in range(reps): g.ir_interpolate_at(hq.array, out=(hv.array, hw.array))
    t_pin = (time.perf_counter() - t0) / reps
    t_ref = None
    if n <= 100_000:
        t0 = time.perf_counter(); wl.grid.ir_interpolate_at(Q, True, cores); t_ref = time.perf_counter() - t0
    print(f"n {n:>8}: pageable {t_page*1e3:9.2f} ms ({n/t_page:.2e} Q/s)  pinned {t_pin*1e3:9.2f} ms ({n/t_pin:.2e} Q/s)"
          + (f"  reference on {cores} cores {t_ref*1e3:9.1f} ms ({n/t_ref:.2e} Q/s, {t_ref/t_page:.0f}x / {t_ref/t_pin:.0f}x)" if t_ref else ""), flush=True)
    del hq, hv, hw
