#!/usr/bin/env python3
"""Kernel-time sweep over tuning knobs on the bench workload (device resident): prints locate / sort / interpolate ms."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import brille_b200  # noqa: E402
from bench import NQ, Q_SEED, build_workload  # noqa: E402

nq = int(float(sys.argv[1])) if len(sys.argv) > 1 else NQ
chunks = [int(c) for c in sys.argv[2].split(",")] if len(sys.argv) > 2 else [256, 128]
kernels = [int(c) for c in sys.argv[3].split(",")] if len(sys.argv) > 3 else [1, 2]  # cell_kernel option values
tiles = [int(c) for c in sys.argv[4].split(",")] if len(sys.argv) > 4 else [4]  # points per register tile (pipelined kernel)
wl = build_workload()
grid = brille_b200.accelerate(wl.grid)
dQ = torch.from_numpy(wl.make_q(nq, Q_SEED)).cuda()
vals = torch.empty((nq, wl.modes, 1), dtype=torch.float64, device="cuda")
vecs = torch.empty((nq, wl.modes, wl.n_atoms, 3), dtype=torch.complex128, device="cuda")
for chunk, ck, tile in [(c, k, t) for k in kernels for t in (tiles if k == 2 else [4]) for c in chunks]:
    grid.set_option("chunk", chunk)
    grid.set_option("tile", tile)
    grid.set_option("cell_kernel", ck)
    for _ in range(3):
        grid.ir_interpolate_at_device(dQ, vals, vecs, check=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        grid.ir_interpolate_at_device(dQ, vals, vecs, check=False)
    e1.record()
    torch.cuda.synchronize()
    total = e0.elapsed_time(e1) / 10
    grid.enable_timing(True)
    t = {k: [] for k in ("locate", "sort", "interpolate")}
    for _ in range(3):
        grid.ir_interpolate_at_device(dQ, vals, vecs, check=False)
        for k in t:
            t[k].append(grid.kernel_ms(k))
    grid.enable_timing(False)
    print(f"cell_kernel {ck} tile {tile} chunk {chunk}: step {total:.3f} ms ({nq/total/1e3:.3e} Q/s) | " + " ".join(f"{k} {np.mean(v):.3f}" for k, v in t.items()), flush=True)
