#!/usr/bin/env python3
"""A/B timing of library options on one configuration (device resident, per-kernel CUDA-event times), with a bit-identity check
of the outputs between the variants.

    python profiles/perf_ab.py C3 coop_locate=0 coop_locate=1 [nq=1e7] [steps=5]
"""
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
name = sys.argv[1]
variants, nq, steps = [], None, 5
for a in sys.argv[2:]:
    if a.startswith("nq="):
        nq = int(float(a[3:]))
    elif a.startswith("steps="):
        steps = int(a[6:])
    else:
        variants.append(dict((kv.split("=")[0], float(kv.split("=")[1])) for kv in a.split(",") if kv))
if name in W.BUILDERS:
    wl = W.BUILDERS[name](b)
    nq = nq or {"C1": 10_000_000, "C2": 10_000_000, "C3": 10_000_000, "C4": 500_000}[name]
else:
    wl = W.c3_p63mmc(b, cls={"C3nest": "BZNestQdc", "C3mesh": "BZMeshQdc"}[name])
    nq = nq or 10_000_000
grid = brille_b200.accelerate(wl.grid)
dQ = torch.from_numpy(wl.make_q(nq, 3)).cuda()
vals, vecs = grid.ir_interpolate_at_device(dQ)
ref_out = None
for v in variants or [{}]:
    for k, x in v.items():
        grid.set_option(k, x)
    for _ in range(2):
        grid.ir_interpolate_at_device(dQ, vals, vecs, check=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        grid.ir_interpolate_at_device(dQ, vals, vecs, check=False)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    grid.enable_timing(True)
    t = {}
    for _ in range(3):
        grid.ir_interpolate_at_device(dQ, vals, vecs, check=False)
        for k in ("locate", "locate_a", "locate_sort", "locate_b", "sort", "interpolate"):
            t.setdefault(k, []).append(grid.kernel_ms(k))
    grid.enable_timing(False)
    same = ""
    cur = (vals[:: max(1, nq // 200000)].clone(), vecs[:: max(1, nq // 200000)].clone())
    if ref_out is None:
        ref_out = cur
    else:
        same = " bit-identical to the first variant: %s" % (bool(torch.equal(cur[0], ref_out[0]) and torch.equal(cur[1], ref_out[1])))
    import hashlib
    k = max(1, nq // 2000)
    digest = hashlib.sha1(vals[::k].cpu().numpy().tobytes() + vecs[::k].cpu().numpy().tobytes()).hexdigest()[:12]
    same += f" sha1 {digest}"
    print(f"{name} {v}: {ms:.3f} ms/step = {nq / ms / 1e3:.4g} Q/s path {grid.last_path} | " +
          " ".join(f"{k} {np.mean(x):.3f}" for k, x in t.items()) + same, flush=True)
