"""Structure-factor consumer on C3: device-resident and host-buffer rates against the chunk size of the host pipeline."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import brille_b200
from brille_b200 import workloads as W
from oracle import ref

NQ = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
wl = W.c3_p63mmc(ref.host())
g = brille_b200.accelerate(wl.grid)
rng = np.random.default_rng(5)
g.set_structure_factor(rng.normal(size=wl.n_atoms) + 1j * rng.normal(size=wl.n_atoms), positions=rng.uniform(0, 1, (wl.n_atoms, 3)), q_transform=rng.normal(size=(3, 3)))
Q = wl.make_q(NQ, 3)
hq = brille_b200.PinnedArray((NQ, 3), np.float64); hq.array[:] = Q
hv = brille_b200.PinnedArray((NQ, wl.modes, 1), np.float64)
hs = brille_b200.PinnedArray((NQ, wl.modes), np.float64)
for chunk in [0, 250_000, 500_000, 1_000_000, 2_000_000, 4_000_000]:
    g.set_option("host_chunk", chunk)
    g.ir_structure_factor(hq.array, out=(hv.array, hs.array))
    t0 = time.perf_counter()
    for _ in range(5):
        g.ir_structure_factor(hq.array, out=(hv.array, hs.array))
    dt = (time.perf_counter() - t0) / 5
    print(f"host_chunk {chunk:>8}: {dt*1e3:7.2f} ms  {NQ/dt:.3e} Q/s  D2H {NQ*192/dt/1e9:.1f} GB/s", flush=True)
g.set_option("host_chunk", 0)
dQ = torch.from_numpy(Q).cuda()
vals = torch.empty((NQ, wl.modes, 1), dtype=torch.float64, device="cuda")
sf = torch.empty((NQ, wl.modes), dtype=torch.float64, device="cuda")
out = {}
for fused in (1, 0):
    g.set_option("sf_fused", fused)
    for _ in range(2):
        g.ir_structure_factor_device(dQ, vals, sf)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.ir_structure_factor_device(dQ, vals, sf)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    out[fused] = sf.clone()
    g.enable_timing(True)
    g.ir_structure_factor_device(dQ, vals, sf)
    print("fused=%d: %.2f ms per step (%.3e Q/s); locate %.2f sort %.2f interpolate %.2f consumer %.2f ms" % (
        (fused, ms, NQ / ms * 1e3) + tuple(g.kernel_ms(k) for k in ("locate", "sort", "interpolate", "consumer"))), flush=True)
    g.enable_timing(False)
    g.ir_structure_factor(hq.array, out=(hv.array, hs.array))
    t0 = time.perf_counter()
    for _ in range(3):
        g.ir_structure_factor(hq.array, out=(hv.array, hs.array))
    dt = (time.perf_counter() - t0) / 3
    print(f"fused={fused}: host buffers {dt*1e3:.2f} ms  {NQ/dt:.3e} Q/s", flush=True)
d = (out[1] - out[0]).abs().max().item() / out[0].abs().max().item()
print("fused vs unfused: max rel diff %.2e, bitwise equal: %s" % (d, bool(torch.equal(out[0], out[1]))))
