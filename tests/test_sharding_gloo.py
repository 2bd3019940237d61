"""CPU, world_size 2, gloo: the multi-rank host logic (shard arithmetic, rank-ordered gather of per-rank results).
The per-rank compute on the GPU box is the CUDA path; here each rank runs the oracle on its shard, which checks that
sharding + gathering reproduces the single-rank result exactly (every Q is independent: SURVEY section 8e)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_bounds_partition():
    from brille_b200.sharding import shard_bounds

    for n in (0, 1, 7, 8, 1000, 10**7 + 3):
        for world in (1, 2, 3, 8):
            b = [shard_bounds(n, r, world) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist

    from brille_b200.sharding import gather_rows, shard
    from helpers import load_golden
    from oracle.oracle import Oracle

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    s, d, _, rest = load_golden("nacl_prim_trellis.npz")
    orc = Oracle(s, d)
    Q = rest["Q"]
    rc, v, w, _ = orc.interpolate_at(shard(Q, rank, world), probe=False)
    assert rc == 0
    V = gather_rows(torch.from_numpy(v)).numpy()
    Wr = gather_rows(torch.from_numpy(np.ascontiguousarray(w.view(np.float64)))).numpy().view(np.complex128)
    # the structure-factor consumer reduces every shard where it was computed; only (n, modes) rows are gathered
    from oracle.consumer import structure_factor

    coef = np.array([1.0 + 0.5j, -0.3 + 2.0j])
    pos = np.array([[0.0, 0.0, 0.0], [0.5, 0.5, 0.5]])
    sf = structure_factor(shard(Q, rank, world), w, coef, positions=pos)
    SF = gather_rows(torch.from_numpy(sf)).numpy()
    # the powder sweep shards the DIRECTION sequence: every rank bins its slice of directions at every |Q| and the partial histograms
    # meet in one all-reduce -- the only exchange of the path
    from brille_b200.sharding import shard_bounds
    from oracle.consumer import powder_histogram, powder_points

    B = np.asarray(s["bz"]["to_xyz"]).reshape(3, 3)
    qr, nqb, wr, nwb, n_dir = (0.3, 3.3), 6, (0.0, 60.0), 10, 40
    lo, hi = shard_bounds(n_dir, rank, world)
    Qp = powder_points(B, qr, nqb, n_dir, seed=9, dir_range=(lo, hi))
    rc, pv, pw, _ = orc.interpolate_at(Qp, probe=False)
    assert rc == 0
    hist, counts = powder_histogram(Qp, B, pv, structure_factor(Qp, pw, coef, positions=pos), qr, nqb, wr, nwb, 1)
    th, tc = torch.from_numpy(hist), torch.from_numpy(counts)
    dist.all_reduce(th)
    dist.all_reduce(tc)
    if rank == 0:
        rc, v1, w1, _ = orc.interpolate_at(Q, probe=False)
        ok = np.array_equal(V, v1) and np.array_equal(Wr, w1) and V.shape[0] == len(Q)
        ok = ok and np.array_equal(SF, structure_factor(Q, w1, coef, positions=pos))
        Qa = powder_points(B, qr, nqb, n_dir, seed=9)
        rc, av, aw, _ = orc.interpolate_at(Qa, probe=False)
        h1, c1 = powder_histogram(Qa, B, av, structure_factor(Qa, aw, coef, positions=pos), qr, nqb, wr, nwb, 1)
        ok = ok and np.array_equal(tc.numpy(), c1) and np.allclose(th.numpy(), h1, rtol=1e-12, atol=0) and c1.sum() == nqb * n_dir
        q.put(bool(ok))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shard_and_gather_matches_single_rank():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok
