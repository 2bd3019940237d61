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
dQ = torch.from_numpy(Q).cuda()
vals = torch.empty((NQ, wl.modes, 1), dtype=torch.float64, device="cuda")
sf = torch.empty((NQ, wl.modes), dtype=torch.float64, device="cuda")
scratch = torch.empty((NQ, wl.modes, wl.n_atoms, 3), dtype=torch.complex128, device="cuda")
g.enable_timing(True)
for _ in range(3):
    g.ir_structure_factor_device(dQ, vals, sf, scratch=scratch, check=False)
    print("device: locate %.2f sort %.2f interpolate %.2f consumer %.2f ms" % tuple(g.kernel_ms(k) for k in ("locate", "sort", "interpolate", "consumer")))
print("consumer kernel: %.0f GB/s read" % (NQ * 2304 / g.kernel_ms("consumer") / 1e6))
