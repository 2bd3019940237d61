#!/usr/bin/env python3
"""Device-resident throughput of the BASELINE.json configurations (and the Nest/Mesh variants of C3) on one GPU."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import brille_b200  # noqa: E402
from brille_b200 import workloads as W  # noqa: E402
from brille_b200 import host as _hostmod  # noqa: E402

b = _hostmod.get()
which = sys.argv[1].split(",") if len(sys.argv) > 1 else ["C1", "C2", "C3", "C3nest", "C3mesh", "C4"]


def build(name):
    if name in W.BUILDERS:
        wl = W.BUILDERS[name](b)
        return wl, {"C1": 10_000_000, "C2": 10_000_000, "C3": 10_000_000, "C4": 500_000}[name]
    lat = W.p63mmc_lattice(b)
    bz = b.BrillouinZone(lat)
    g = b.BZNestQdc(bz, bz.ir_polyhedron.volume / 2000, 5) if name == "C3nest" else b.BZMeshQdc(bz, bz.ir_polyhedron.volume / 2000, 3)
    args = W._gamma_fill(g, 12, 4, 3)
    return W.Workload(name, g, bz, 12, 4, W._uniform_q(-3, 3), args), 10_000_000


for name in which:
    wl, nq = build(name)
    grid = brille_b200.accelerate(wl.grid)
    dQ = torch.from_numpy(wl.make_q(nq, 3)).cuda()
    vals, vecs = grid.ir_interpolate_at_device(dQ)
    for _ in range(2):
        grid.ir_interpolate_at_device(dQ, vals, vecs, check=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = 5
    e0.record()
    for _ in range(steps):
        grid.ir_interpolate_at_device(dQ, vals, vecs, check=False)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    grid.enable_timing(True)
    grid.ir_interpolate_at_device(dQ, vals, vecs, check=False)
    t = {k: grid.kernel_ms(k) for k in ("locate", "sort", "interpolate")}
    grid.enable_timing(False)
    bpq = grid.bytes_per_q
    print(f"{name:7s} nQ={nq:.0e} bytes/Q={bpq:6d} step {ms:8.3f} ms  {nq/ms*1e3:.3e} Q/s  {bpq*nq/ms/1e6:7.1f} GB/s algorithmic | "
          + " ".join(f"{k} {v:.3f}" for k, v in t.items()), flush=True)
    if name != "C1":  # the structure-factor consumer behind the path (fused where the atom count allows, else through a scratch)
        import time

        rng = np.random.default_rng(5)
        grid.set_structure_factor(rng.normal(size=wl.n_atoms) + 1j * rng.normal(size=wl.n_atoms), positions=rng.uniform(0, 1, (wl.n_atoms, 3)),
                                  q_transform=rng.normal(size=(3, 3)))
        sf = torch.empty((nq, wl.modes), dtype=torch.float64, device="cuda")
        del vecs
        torch.cuda.empty_cache()
        for _ in range(2):
            grid.ir_structure_factor_device(dQ, vals, sf)
        e0.record()
        for _ in range(steps):
            grid.ir_structure_factor_device(dQ, vals, sf)
        e1.record()
        torch.cuda.synchronize()
        cms = e0.elapsed_time(e1) / steps
        grid.enable_timing(True)
        grid.ir_structure_factor_device(dQ, vals, sf)
        t = {k: grid.kernel_ms(k) for k in ("locate", "sort", "interpolate", "consumer")}
        grid.enable_timing(False)
        hq = brille_b200.PinnedArray((nq, 3), np.float64)
        hq.array[:] = dQ.cpu().numpy()
        hv = brille_b200.PinnedArray(tuple(vals.shape), np.float64)
        hs = brille_b200.PinnedArray((nq, wl.modes), np.float64)
        grid.ir_structure_factor(hq.array, out=(hv.array, hs.array))
        t0 = time.perf_counter()
        for _ in range(3):
            grid.ir_structure_factor(hq.array, out=(hv.array, hs.array))
        dt = (time.perf_counter() - t0) / 3
        print(f"{name:7s} structure factor: device-resident {cms:8.3f} ms  {nq/cms*1e3:.3e} Q/s | " + " ".join(f"{k} {v:.3f}" for k, v in t.items())
              + f" | host buffers {dt*1e3:8.2f} ms  {nq/dt:.3e} Q/s ({16*wl.modes*nq/dt/1e9:.1f} GB/s D2H)", flush=True)
        del sf, hq, hv, hs
    else:
        del vecs
    del vals, dQ
    grid.close()
    torch.cuda.empty_cache()
