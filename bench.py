#!/usr/bin/env python3
"""bench.py -- Q-points/s through ir_interpolate_at (eigenvalues + rotated eigenvectors) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[2], the configuration the north-star target is quoted on): P6_3/mmc hexagonal
4-atom cell, BZTrellisQdc hybrid cube/tetrahedron trellis at V_ir/2000, 12 modes, eigenvalues + Gamma-rotated
eigenvectors, 1e7 uniformly random Q in [-3,3)^3 rlu per GPU per step (weak scaling: Q is sharded, tables are
replicated, no collective on the data path).  One step = one pass of the whole path over the batch.

The printed JSON line follows the driver's contract.  `value` is timed with CUDA events with Q and the outputs
resident in HBM; `e2e` goes through the host-buffer C-ABI call (pinned host buffers, H2D of Q and D2H of both
results inside the timed region); `roofline` is the dominant (interpolation) kernel against the measured HBM copy
bandwidth; `cpu_baseline` is the reference's own OpenMP implementation (the unmodified build, third_party/brille_host) on
the host cores.  Beside the contract's keys: `configs` (every other BASELINE configuration, device resident), `consumer` (the
fused structure-factor and powder-average calls) and, on a single-GPU run, `sort` (the device sort() of the C3 grid beside the
reference's OpenMP sort()).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "Q-points/sec interpolated (eigvals+eigvecs)"
UNIT = "Q/s"
NQ = 10_000_000           # Q per GPU per step (BASELINE.json configs[2])
E2E_NQ = 2_000_000        # Q per e2e step (host buffers; the rate is flat in nQ, 24 GB of pinned output is not)
CPU_SAMPLE_NQ = int(os.environ.get("BENCH_CPU_SAMPLE_NQ", "200000"))  # bounded sample for the CPU reference legs
Q_SEED = 3


def build_workload():
    """Host-side construction with brille's own C++ (brille_b200.host: an installed brille, or third_party/brille_host)."""
    from brille_b200 import host
    from brille_b200 import workloads as W

    return W.c3_p63mmc(host.get())


WORKLOAD_TEXT = {
    "C3": "C3: P6_3/mmc 4-atom cell, BZTrellisQdc V_ir/2000 (hybrid cube/tetrahedron), 12 modes eigvals+eigvecs (Gamma), 1e7 uniform random Q per GPU per step",
    "C5": "C5: powder-average sweep (|Q| in U(0.1,10) 1/angstrom, isotropic directions) on the C3 P6_3/mmc 12-mode trellis, 1e7 Q per GPU per step, Q sharded across the GPUs",
}
WORKLOAD = "C3"


def make_q(wl, n, seed):
    """uniform random Q (C3, the default) or the powder-average Q of BASELINE configs[4] (--workload C5)"""
    if WORKLOAD == "C5":
        from brille_b200 import workloads as W
        from brille_b200 import _bridge

        return np.ascontiguousarray(W.powder_q(np.asarray(_bridge.flatten_bz(wl.bz)["to_xyz"]), n, seed))
    return wl.make_q(n, seed)


def workload_config(wl, extra=None):
    cfg = {
        "workload": WORKLOAD_TEXT[WORKLOAD],
        "q_per_gpu_per_step": NQ,
        "modes": wl.modes,
        "atoms": wl.n_atoms,
        "vertices": int(wl.grid.rlu.shape[0]),
        "bytes_per_q": wl.bytes_per_q,
        "sharding": "Q sharded across GPUs, tables replicated, no data-path collective",
        "cache": "inputs+outputs (24.5 GB per step) far larger than the 126 MB L2; the 7.3 MB vertex table is L2-resident by design",
    }
    if extra:
        cfg.update(extra)
    return cfg


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    FIELDS = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "50", "-i", str(self.index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True,
            )
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1]))
                mx.append(float(p[2]))
            except ValueError:
                continue
            for name, v in zip(names, p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {
            "sm_mhz": float(np.median(sm)) if sm else None,
            "sm_max_mhz": float(max(mx)) if mx else None,
            "samples": len(sm),
            "reasons": sorted(reasons),
        }


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def profiled_traffic():
    """dram bytes per launch of the dominant kernel from the committed ncu capture, if one exists."""
    p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p))
        except Exception:
            return None
    return None


# ---------------------------------------------------------------------------------------------------------------------
# the other BASELINE configurations (and the C3 lattice on the Nest / Mesh grid kinds), device resident, reported in the
# `configs` block of the JSON line beside the C3 headline
# ---------------------------------------------------------------------------------------------------------------------
def _config_specs(world):
    """(key, text, builder, points per GPU per step, points per call, Q maker or None)"""
    from brille_b200 import workloads as W
    from brille_b200 import _bridge
    from brille_b200 import host

    b = host.get()

    def powder(wl, n, seed):
        return np.ascontiguousarray(W.powder_q(np.asarray(_bridge.flatten_bz(wl.bz)["to_xyz"]), n, seed))

    specs = [
        ("C5", "BASELINE configs[4]: powder-average sweep (|Q| in U(0.1,10) 1/angstrom, isotropic) on the C3 P6_3/mmc 12-mode trellis, 1e7 Q per GPU per step, Q sharded across the GPUs",
         lambda: W.c3_p63mmc(b), 10_000_000, 10_000_000, powder),
    ]
    if world == 1:
        specs += [
            ("C1", "BASELINE configs[0]: Fd-3m a=4.96 BZTrellisQdc V_ir/2000, one scalar eigenvalue (general kernel), 1e5 Q per step (3.2 MB of traffic: L2 flushed between steps)",
             lambda: W.c1_fd3m_scalar(b), 100_000, 100_000, None),
            ("C2", "BASELINE configs[1]: NaCl-like primitive 2-atom cell BZTrellisQdc V_ir/1000, 6 modes eigvals+eigvecs (Gamma), 1e6 Q per step",
             lambda: W.c2_nacl(b), 1_000_000, 1_000_000, None),
            ("C4", "BASELINE configs[3]: P2_1/c 24-atom cell BZNestQdc V_ir/2000, 72 modes eigvals+eigvecs (Gamma), 1e7 Q per step in calls of 5e5 (41.8 GB of output per call)",
             lambda: W.c4_p21c_nest(b), 10_000_000, 500_000, None),
            ("C3-nest", "the C3 lattice and data on a BZNestQdc V_ir/2000 (tetrahedron tree), 1e7 Q per step",
             lambda: W.c3_p63mmc(b, cls="BZNestQdc"), 10_000_000, 10_000_000, None),
            ("C3-mesh", "the C3 lattice and data on a BZMeshQdc V_ir/2000 (layered tetrahedral meshes), 1e7 Q per step",
             lambda: W.c3_p63mmc(b, cls="BZMeshQdc"), 10_000_000, 10_000_000, None),
        ]
    return specs


def measure_config(torch, dist, brille_b200, spec, local, rank, world, steps, peak):
    """Device-resident Q/s of one configuration: every step is timed with its own CUDA event pair on the launching stream, L2 is
    flushed (256 MB written) between steps, the slowest rank counts."""
    key, text, build, nq, per_call, qmaker = spec
    dev = torch.device("cuda", local)
    wl = build()
    grid = brille_b200.accelerate(wl.grid, device=local)
    Q = qmaker(wl, nq, Q_SEED + 1000 * rank) if qmaker else wl.make_q(nq, Q_SEED + 1000 * rank)
    dQ = torch.from_numpy(np.ascontiguousarray(Q)).to(dev)
    c = min(per_call, nq)
    tv = torch.complex128 if grid._vals_dtype == np.complex128 else torch.float64
    tw = torch.complex128 if grid._vecs_dtype == np.complex128 else torch.float64
    vals = torch.empty((c,) + grid._vals_shape, dtype=tv, device=dev)
    vecs = torch.empty((c,) + grid._vecs_shape, dtype=tw, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream(dev)

    def step(check=False):
        for lo in range(0, nq, c):
            m = min(c, nq - lo)
            grid.ir_interpolate_at_device(dQ[lo:lo + m], vals[:m], vecs[:m], check=check, stream=stream)

    step(check=True)
    l0 = grid.launch_count
    step()
    launches = grid.launch_count - l0
    path = grid.last_path
    times = []
    for _ in range(max(steps, 1)):
        flush.fill_(1)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        step()
        e1.record(stream)
        torch.cuda.synchronize(dev)
        times.append(e0.elapsed_time(e1))
    ms = float(np.mean(times))
    # per-kernel times (CUDA events inside the library, on the launching stream), summed over the calls of one step
    grid.enable_timing(True)
    k = {"locate": 0.0, "sort": 0.0, "interpolate": 0.0}
    for lo in range(0, nq, c):
        m = min(c, nq - lo)
        grid.ir_interpolate_at_device(dQ[lo:lo + m], vals[:m], vecs[:m], check=False, stream=stream)
        for name in k:
            k[name] += max(grid.kernel_ms(name), 0.0)
    grid.enable_timing(False)
    torch.cuda.synchronize(dev)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    bpq = grid.bytes_per_q
    achieved = bpq * nq / (k["interpolate"] * 1e-3) / 1e9 if k["interpolate"] > 0 else None
    names = {4: "k_interp_cell_tma (pipelined cell kernel)", 2: "k_interp_cell (on-the-fly cell kernel)", 8: "k_interp (general kernel)"}
    kernel = next((v for b_, v in names.items() if path & b_), "?")
    # DRAM bytes of the interpolation kernel per step from the committed ncu captures (profiles/roofline_traffic.json), when one
    # exists for this configuration with the same points per launch
    tr = (profiled_traffic() or {}).get("configs", {}).get(key)
    traffic = tr["dram_bytes_per_launch"] * (nq // c) if tr and tr.get("q_per_launch") == c else None
    out = {
        "workload": text, "value": world * nq / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "q_per_gpu_per_step": nq, "q_per_call": c,
        "modes": wl.modes, "atoms": wl.n_atoms, "vertices": int(wl.grid.rlu.shape[0]), "bytes_per_q": bpq, "gpu_launches_per_step": int(launches),
        "two_kernel_location": bool(path & 1),
        "roofline": {"bound": "hbm", "kernel": kernel, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak if achieved else None,
                     "kernel_ms": k["interpolate"], "locate_kernel_ms": k["locate"], "bucket_sort_ms": k["sort"],
                     "path_frac": bpq * nq / (ms * 1e-3) / 1e9 / peak, "traffic": traffic},
    }
    grid.close()
    del dQ, vals, vecs, flush
    torch.cuda.empty_cache()
    return out


def cpu_reference_rate(wl, nq, threads, repeats=1):
    """The reference's own ir_interpolate_at (unmodified brille, third_party/brille_host) on the host cores."""
    Q = make_q(wl, nq, Q_SEED + 100)
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        wl.grid.ir_interpolate_at(Q, True, threads)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return nq / best, best


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    wl = build_workload()
    cores = os.cpu_count() or 1
    nq = CPU_SAMPLE_NQ
    Q = make_q(wl, nq, Q_SEED + 100)
    for _ in range(args.warmup):
        wl.grid.ir_interpolate_at(Q[: nq // 10], True, cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        wl.grid.ir_interpolate_at(Q, True, cores)
    dt = time.perf_counter() - t0
    value = nq * args.steps / dt
    line = {
        "impl": "reference",
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": workload_config(wl, {"reference_sample_q_per_step": nq}),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference",
                         "sample": f"{nq} Q per step of the same workload, brille ir_interpolate_at(Q, True, {cores}), unmodified reference build (third_party/brille_host)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def run_ours(args):
    import torch
    import torch.distributed as dist

    import brille_b200

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- brille_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner to stdout when NCCL_DEBUG is VERSION or WARN; stdout carries the one JSON line, so the
        # communicator is created (init + first collective) with file descriptor 1 pointing at stderr
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    dev = torch.device("cuda", local)

    wl = build_workload()
    grid = brille_b200.accelerate(wl.grid, device=local)
    bpq = grid.bytes_per_q
    assert bpq == wl.bytes_per_q

    # this rank's shard of the (world * NQ)-point job; different points on every rank
    Q = make_q(wl, NQ, Q_SEED + 1000 * rank)
    dQ = torch.from_numpy(Q).to(dev)
    vals = torch.empty((NQ, wl.modes, 1), dtype=torch.float64, device=dev)
    vecs = torch.empty((NQ, wl.modes, wl.n_atoms, 3), dtype=torch.complex128, device=dev)
    stream = torch.cuda.current_stream(dev)

    def step(check=False):
        grid.ir_interpolate_at_device(dQ, vals, vecs, check=check, stream=stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    sampler = ClockSampler(local)  # every rank samples its own GPU, from the warm-up until the per-kernel timing pass is over
    sampler.start()                # (the same steps as the timed region, which alone is shorter than two samples)
    for i in range(max(args.warmup, 3)):
        step(check=(i == 0))  # the first warm-up step also verifies that every Q was placed
    barrier()
    launches0 = grid.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = grid.launch_count - launches0
    per_rank = [torch.zeros(1, dtype=torch.float64, device=dev) for _ in range(world)]
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_gather(per_rank, t)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    else:
        per_rank[0] = t.clone()
    per_rank_ms = [float(x.item()) / args.steps for x in per_rank]
    ms_max = float(t.item())
    value = world * NQ * args.steps / (ms_max * 1e-3)

    # per-kernel durations with CUDA events on the launching stream (separate pass so the events do not perturb `value`)
    grid.enable_timing(True)
    k_loc, k_int, k_sort = [], [], []
    for _ in range(3):
        step()
        k_loc.append(grid.kernel_ms("locate"))
        k_sort.append(grid.kernel_ms("sort"))
        k_int.append(grid.kernel_ms("interpolate"))
    grid.enable_timing(False)
    torch.cuda.synchronize(dev)
    loc_ms, int_ms, sort_ms = float(np.mean(k_loc)), float(np.mean(k_int)), float(np.mean(k_sort))
    clocks = sampler.stop()
    if world > 1:  # slowest GPU's clock and the union of the throttle reasons over the ranks
        all_clocks = [None] * world
        dist.all_gather_object(all_clocks, clocks)
        if rank == 0:
            sm = [c["sm_mhz"] for c in all_clocks if c.get("sm_mhz")]
            clocks = {"sm_mhz": min(sm) if sm else None, "sm_max_mhz": clocks.get("sm_max_mhz"), "samples": sum(c.get("samples", 0) for c in all_clocks),
                      "reasons": sorted(set(r for c in all_clocks for r in c.get("reasons", []))), "per_rank_sm_mhz": [c.get("sm_mhz") for c in all_clocks]}

    # end to end through the host-buffer C-ABI call: pinned host Q in, pinned host results out, every step.  A step is the full
    # batch (1e7 Q per GPU) pushed through a ring of page-locked buffers in calls of `ne` points (24 GB of page-locked output per
    # rank for one call of 1e7 is not what a caller would hold; a consumer drains the ring between the calls)
    ne = E2E_NQ if world <= 2 else E2E_NQ // 2  # (page-locked host memory of all ranks together: 2.4 GB per 1e6 Q)
    calls = (NQ + ne - 1) // ne
    hq = brille_b200.PinnedArray((NQ, 3), np.float64)
    hv = brille_b200.PinnedArray((ne, wl.modes, 1), np.float64)
    hw = brille_b200.PinnedArray((ne, wl.modes, wl.n_atoms, 3), np.complex128)
    hq.array[:] = Q

    def e2e_step():
        for c in range(calls):
            m = min(ne, NQ - c * ne)
            grid.ir_interpolate_at(hq.array[c * ne:c * ne + m], out=(hv.array[:m], hw.array[:m]))

    e2e_step()
    e2e_steps = max(1, min(args.steps, 3))
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * NQ * e2e_steps / float(t.item())
    e2e_d2h_gbs = world * (bpq - 24) * NQ * e2e_steps / float(t.item()) / 1e9
    checksum = float(hv.array[:1000].sum())
    # the default call of the mirror API: plain numpy arrays in, fresh (pageable) numpy arrays out
    npq = 1_000_000
    grid.ir_interpolate_at(Q[:npq])
    barrier()
    t0 = time.perf_counter()
    pv, pw = grid.ir_interpolate_at(Q[:npq])
    e2e_pageable_s = time.perf_counter() - t0
    del pv, pw
    t = torch.tensor([e2e_pageable_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_pageable = world * npq / float(t.item())
    # what the host side of this box can take: device-to-pinned-host copies of the same size on all ranks at once, nothing else
    # running (the GPUs of a box share the host's PCIe root complexes and memory: the end-to-end rate of ir_interpolate_at, 2400
    # bytes per Q, cannot exceed this)
    hdst = torch.from_numpy(hw.array.view(np.uint8).reshape(-1))
    dsrc = torch.empty(hdst.numel(), dtype=torch.uint8, device=dev)
    hdst.copy_(dsrc, non_blocking=True)
    barrier()
    t0 = time.perf_counter()
    for _ in range(5):
        hdst.copy_(dsrc, non_blocking=True)
    torch.cuda.synchronize(dev)
    d2h_s = time.perf_counter() - t0
    t = torch.tensor([d2h_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    d2h_ceiling_gbs = world * 5 * dsrc.numel() / float(t.item()) / 1e9
    del hq, hv, hw, dsrc, hdst  # (page-locked memory back before the next leg takes its own)

    # device-resident consumer: the same path followed by the structure-factor reduction on the device; the
    # eigenvectors never leave HBM, 8*modes bytes per Q of |F|^2 go back instead of 16*modes*3*atoms.  Reported beside the
    # headline numbers, not instead of them (different output).
    consumer = None
    if not args.no_consumer:
        rng = np.random.default_rng(5)
        grid.set_structure_factor(rng.normal(size=wl.n_atoms) + 1j * rng.normal(size=wl.n_atoms), positions=rng.uniform(0, 1, (wl.n_atoms, 3)),
                                  q_transform=rng.normal(size=(3, 3)))
        dsf = torch.empty((NQ, wl.modes), dtype=torch.float64, device=dev)

        def c_step(fused):
            # fused (the default of the API): reduced inside the pipelined cell kernel, the eigenvectors exist nowhere; the call
            # reads its failure counters back, i.e. synchronises.  Unfused: eigenvectors through a scratch + k_structure_factor.
            if fused:
                grid.ir_structure_factor_device(dQ, vals, dsf, stream=stream)
            else:
                grid.ir_structure_factor_device(dQ, vals, dsf, scratch=vecs, check=False, stream=stream)

        c_ms = {}
        for fused in (True, False):
            for _ in range(2):
                c_step(fused)
            barrier()
            e0.record(stream)
            for _ in range(args.steps):
                c_step(fused)
            e1.record(stream)
            barrier()
            c_ms[fused] = e0.elapsed_time(e1) / args.steps
        grid.enable_timing(True)
        c_step(True)
        c_fused_kernel_ms = grid.kernel_ms("interpolate")
        c_step(False)
        c_kernel_ms = grid.kernel_ms("consumer")
        grid.enable_timing(False)
        nc = NQ if world <= 2 else NQ // 2
        hq2 = brille_b200.PinnedArray((nc, 3), np.float64)
        hv2 = brille_b200.PinnedArray((nc, wl.modes, 1), np.float64)
        hs2 = brille_b200.PinnedArray((nc, wl.modes), np.float64)
        hq2.array[:] = Q[:nc]
        grid.ir_structure_factor(hq2.array, out=(hv2.array, hs2.array))
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            grid.ir_structure_factor(hq2.array, out=(hv2.array, hs2.array))
        torch.cuda.synchronize(dev)
        c_e2e_s = (time.perf_counter() - t0) / args.steps
        t = torch.tensor([c_ms[True], c_ms[False], c_e2e_s], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        c_fused_ms, c_unfused_ms, c_e2e_s = float(t[0].item()), float(t[1].item()), float(t[2].item())
        # powder average: points generated on the device (200 |Q| bins x NQ/200 directions per GPU per step), structure factor fused
        # into the cell kernel, histogram accumulated on the device; per step only the histogram (200 x 400 doubles) leaves the GPU and
        # the ranks' partial histograms meet in one small all-reduce
        from brille_b200 import _bridge as _br

        grid.set_structure_factor(rng.normal(size=wl.n_atoms) + 1j * rng.normal(size=wl.n_atoms), positions=rng.uniform(0, 1, (wl.n_atoms, 3)),
                                  q_transform=np.asarray(_br.flatten_bz(wl.bz)["to_xyz"]).reshape(3, 3))
        nqb, nwb = 200, 400
        n_dir = world * (NQ // nqb)
        lo_d, hi_d = rank * (NQ // nqb), (rank + 1) * (NQ // nqb)

        def p_step():
            h, c_ = grid.ir_powder_sweep((0.1, 10.0), nqb, (0.0, 55.0), nwb, n_dir, seed=7, weight=1, dir_range=(lo_d, hi_d))
            if world > 1:
                th = torch.from_numpy(h).to(dev)
                dist.all_reduce(th)
                h = th.cpu().numpy()
            return h, c_

        p_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            ph, pc_ = p_step()
        torch.cuda.synchronize(dev)
        p_s = (time.perf_counter() - t0) / args.steps
        t = torch.tensor([p_s], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        p_s = float(t.item())
        powder = {"value": world * NQ / p_s, "unit": UNIT, "ms_per_step": 1e3 * p_s, "q_per_gpu_per_step": NQ, "h2d_bytes_per_step": 0,
                  "d2h_bytes_per_step": 8 * (nqb * nwb + nqb), "bins": [nqb, nwb], "checksum": float(ph.sum()), "points_per_q_bin": float(pc_[0]),
                  "note": "ir_powder_sweep: Q generated on the device, |F|^2/omega binned on (|Q|, omega) with FP64 atomics; end to end "
                          "through the host call, histogram on the host (all-reduce of the partial histograms for N > 1)"}
        consumer = {
            "what": "ir_structure_factor: the path + |sum_k c_k e^{2 pi i Q.r_k} (TQ . eps_k^*)|^2 per (Q, mode) on the device",
            "powder_average": powder,
            "device_resident": {"value": world * NQ / (c_fused_ms * 1e-3), "unit": UNIT, "ms_per_step": c_fused_ms,
                                "fused_cell_kernel_ms": c_fused_kernel_ms,
                                "note": "reduction fused into the finish of the pipelined cell kernel (k_interp_cell_tma<4,true>): the eigenvectors are never written",
                                "unfused_ms_per_step": c_unfused_ms, "unfused_consumer_kernel_ms": c_kernel_ms},
            "e2e": {"value": world * nc / c_e2e_s, "unit": UNIT, "q_per_step": nc, "h2d_bytes_per_step": 24 * nc,
                    "d2h_bytes_per_step": (8 * wl.modes + 8 * wl.modes) * nc, "checksum": float(hs2.array[:1000].sum())},
        }

    # the other configurations (device resident; see measure_config)
    configs = None
    if not args.no_configs:
        del vals, vecs, dQ
        torch.cuda.empty_cache()
        peak0, _ = measured_peak()
        configs = {}
        for spec in _config_specs(world):
            try:
                configs[spec[0]] = measure_config(torch, dist, brille_b200, spec, local, rank, world, min(args.steps, 5), peak0)
            except Exception as e:  # noqa: BLE001 - a failing side configuration must not take the headline line with it
                configs[spec[0]] = {"workload": spec[1], "error": f"{type(e).__name__}: {e}"}

    # sort() on a fresh grid (the reference's sort() changes the host object), rank 0 of a single-GPU run only:
    # it times the reference's OpenMP sort() beside the device call (a few seconds of host time)
    sort_block = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and not args.no_consumer:
        try:
            sort_block = measure_sort(brille_b200, local)
        except Exception as e:  # noqa: BLE001 - a side measurement must not take the headline line with it
            sort_block = {"error": f"{type(e).__name__}: {e}"}

    line = None
    if rank == 0:
        peak, peak_src = measured_peak()
        achieved = bpq * NQ / (int_ms * 1e-3) / 1e9
        traffic = profiled_traffic()
        cores = os.cpu_count() or 1
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            rate, secs = cpu_reference_rate(wl, CPU_SAMPLE_NQ, cores)
            cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "reference",
                   "sample": f"{CPU_SAMPLE_NQ} Q of the same workload in one call of the unmodified reference (third_party/brille_host) ir_interpolate_at(Q, True, {cores}): {secs:.1f} s"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_max / args.steps, "per_rank_ms_per_step": per_rank_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": workload_config(wl, {"e2e_q_per_step": NQ, "e2e_q_per_call": ne, "parallelism": f"q-shard x{world}"}),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 24 * NQ, "d2h_bytes_per_step": (bpq - 24) * NQ, "q_per_step": NQ,
                    "calls_per_step": calls, "steps": e2e_steps,
                    "note": "b200_ir_interpolate_at, page-locked host buffers: the 1e7 Q of a step go through a ring of pinned output buffers in calls of "
                            f"{ne} points; inside a call chunked H2D / kernels / D2H overlap on two streams",
                    "d2h_gbs_achieved": e2e_d2h_gbs, "d2h_gbs_ceiling": d2h_ceiling_gbs,
                    "ceiling_note": "device-to-pinned-host copies alone, all ranks at once: the host-side bandwidth the GPUs of the box share",
                    "pageable_call": {"value": e2e_pageable, "unit": UNIT, "q_per_call": npq,
                                      "note": "ir_interpolate_at(Q) with plain numpy arrays in and fresh numpy arrays out (page-locked bounce buffers + host threads inside the library)"},
                    "checksum": checksum},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": "k_interp_cell_tma (persistent cell-batched interpolate+rotate, TMA-staged cell records)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "peak_source": peak_src, "algorithmic_bytes_per_q": bpq,
                         "kernel_ms": int_ms, "locate_kernel_ms": loc_ms, "bucket_sort_ms": sort_ms,
                         "path_frac": bpq * NQ / (ms_max / args.steps * 1e-3) / 1e9 / peak,
                         "traffic": (traffic or {}).get("dram_bytes_per_launch") if traffic else None},
            "cpu_baseline": cpu,
            "consumer": consumer,
            "sort": sort_block,
            "configs": configs,
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    grid.close()
    return 0


def measure_sort(brille_b200, device):
    """sort() of the C3 grid -- the mode assignment of every connected vertex pair (interpolatordual.hpp:398-434): cost matrices and
    Jonker-Volgenant assignments on the device through b200_grid_sort_pairs (host buffers in and out) beside the reference's own
    OpenMP sort() on the box's cores."""
    from brille_b200 import _bridge

    wl = build_workload()
    g = brille_b200.accelerate(wl.grid, device=device)
    plan = _bridge.sort_plan(wl.grid)
    pairs = plan["pairs"]
    g.sort_pairs(pairs, plan)  # warm-up: work space
    best = float("inf")
    for _ in range(3):
        t0 = time.perf_counter()
        row, col = g.sort_pairs(pairs, plan)
        best = min(best, time.perf_counter() - t0)
    launches = g.launch_count
    t0 = time.perf_counter()
    wl.grid.sort()
    t_ref = time.perf_counter() - t0
    ref = _bridge.pair_permutations(wl.grid, pairs)
    same = float((row.astype(np.uint32) == ref[:, 0, :]).all(axis=1).mean())
    g.close()
    cores = os.cpu_count() or 1
    return {"what": "sort(): cost matrix + assignment of every connected vertex pair of the C3 grid, b200_grid_sort_pairs (host buffers) "
                    "against the reference's OpenMP sort()",
            "pairs": int(len(pairs)), "modes": int(row.shape[1]), "device_ms": best * 1e3, "pairs_per_s": len(pairs) / best,
            "reference_s": t_ref, "reference_cores": cores, "speedup": t_ref / best, "identical_permutations": same,
            "gpu_launches": int(launches)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-consumer", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the per-configuration block (C1, C2, C4, C5, C3 on Nest / Mesh)")
    ap.add_argument("--workload", default="C3", choices=["C3", "C5"], help="C3 (default, the headline configuration) or the C5 powder-average Q on the same grid")
    args = ap.parse_args()
    global WORKLOAD
    WORKLOAD = args.workload
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
