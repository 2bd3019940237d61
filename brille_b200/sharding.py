"""Q sharding across GPUs / ranks: contiguous row blocks of Q, tables replicated, no data-path
collective.  ``ShardedGrid`` drives several devices from one process; under ``torch.distributed`` every rank owns one
device and calls :func:`shard_bounds` with its rank."""
from __future__ import annotations

import threading

import numpy as np


def shard_bounds(n, rank, world):
    """[lo, hi) of the rows of an ``n``-row Q array owned by ``rank`` of ``world``: contiguous, sizes differ by <= 1."""
    n, rank, world = int(n), int(rank), int(world)
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank/world")
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard(Q, rank, world):
    lo, hi = shard_bounds(len(Q), rank, world)
    return Q[lo:hi]


def gather_rows(local, group=None):
    """All-gather variable-length row blocks in rank order (host-side result assembly; works with gloo and nccl)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    t = torch.as_tensor(local)
    sizes = [torch.zeros(1, dtype=torch.int64, device=t.device) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device), group=group)
    sizes = [int(s.item()) for s in sizes]
    m = max(sizes)
    pad = torch.zeros((m,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[: t.shape[0]] = t
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    return torch.cat([o[:s] for o, s in zip(out, sizes)], dim=0)


class ShardedGrid:
    """The same grid on several GPUs of one process: ``ir_interpolate_at`` splits Q, runs the shards concurrently (one
    host thread per device, the library releases the GIL inside ctypes calls) and writes into one output pair."""

    def __init__(self, host_grid, devices):
        from .grid import B200Grid

        self.grids = [B200Grid(host_grid, device=d) for d in devices]

    def ir_interpolate_at(self, Q, useparallel=False, threads=-1, do_not_move_points=False):
        g0 = self.grids[0]
        Q = g0._check_q(Q)
        g0._check_filled()
        vals = np.empty((len(Q),) + g0._vals_shape, g0._vals_dtype)
        vecs = np.empty((len(Q),) + g0._vecs_shape, g0._vecs_dtype)
        errors = []

        def work(rank, grid):
            lo, hi = shard_bounds(len(Q), rank, len(self.grids))
            try:
                grid.ir_interpolate_at(Q[lo:hi], do_not_move_points=do_not_move_points, out=(vals[lo:hi], vecs[lo:hi]))
            except Exception as e:  # noqa: BLE001 - re-raised below, all-or-nothing like the reference
                errors.append(e)

        ts = [threading.Thread(target=work, args=(r, g)) for r, g in enumerate(self.grids)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        if errors:
            raise errors[0]
        return vals, vecs

    def set_structure_factor(self, *args, **kwargs):
        for g in self.grids:
            g.set_structure_factor(*args, **kwargs)

    def ir_structure_factor(self, Q, do_not_move_points=False):
        """``(vals, sf)`` of :meth:`B200Grid.ir_structure_factor`, the shards reduced on their own devices."""
        g0 = self.grids[0]
        Q = g0._check_q(Q)
        g0._check_filled()
        vals = np.empty((len(Q),) + g0._vals_shape, g0._vals_dtype)
        sf = np.empty((len(Q), int(g0._data_tables.vectors.branches)), np.float64)
        errors = []

        def work(rank, grid):
            lo, hi = shard_bounds(len(Q), rank, len(self.grids))
            try:
                grid.ir_structure_factor(Q[lo:hi], do_not_move_points=do_not_move_points, out=(vals[lo:hi], sf[lo:hi]))
            except Exception as e:  # noqa: BLE001 - re-raised below, all-or-nothing like the reference
                errors.append(e)

        ts = [threading.Thread(target=work, args=(r, g)) for r, g in enumerate(self.grids)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        if errors:
            raise errors[0]
        return vals, sf

    def ir_powder_sweep(self, q_range, n_qbins, w_range, n_wbins, n_dir, seed=0, weight=0):
        """:meth:`B200Grid.ir_powder_sweep` with the direction sequence cut into one contiguous slice per device; the partial
        histograms (a few hundred KB each) are added on the host -- the only exchange of the path."""
        parts = [None] * len(self.grids)
        errors = []

        def work(rank, grid):
            lo, hi = shard_bounds(int(n_dir), rank, len(self.grids))
            try:
                parts[rank] = grid.ir_powder_sweep(q_range, n_qbins, w_range, n_wbins, n_dir, seed=seed, weight=weight, dir_range=(lo, hi))
            except Exception as e:  # noqa: BLE001
                errors.append(e)

        ts = [threading.Thread(target=work, args=(r, g)) for r, g in enumerate(self.grids)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        if errors:
            raise errors[0]
        return sum(p[0] for p in parts), sum(p[1] for p in parts)

    def close(self):
        for g in self.grids:
            g.close()
