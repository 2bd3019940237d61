"""GPU: parity cases added in round 2 -- every BASELINE configuration at its own size, the branches the first round left
untested (Gamma-rotated matrices, points shared by many Nest tetrahedra, wedge rotation of far-away points, refills that
change the row sizes) and the stricter per-(Q, mode) tolerance of helpers.rel_err."""
import numpy as np
import pytest

import brille_b200
from brille_b200 import tables as T
from brille_b200 import workloads as W
from oracle.oracle import Oracle
from helpers import assert_decisions_equal, assert_values_close, rel_err

pytestmark = pytest.mark.gpu


def probe_dict(pr):
    return {"tau": pr.tau, "q_ir": pr.q_ir, "ridx": pr.ridx, "invridx": pr.invridx, "n_vert": pr.n_vert, "vertex": pr.vertex, "weight": pr.weight}


def test_c4_at_its_own_configuration(host, bridge):
    """BASELINE configs[3] as SURVEY 8d states it: P2_1/c, 24 atoms, 72 modes, BZNestQdc at V_ir/2000 (2126 vertices, 3597
    leaves).  1e5 Q in chunks (83.5 KB of output per Q): tau / R / tetrahedron / vertices / weights bit-identical to the oracle,
    values and vectors within 1e-10 per (Q, mode) of the oracle, and of the reference itself on a slice."""
    wl = W.c4_p21c_nest(host)
    assert wl.grid.rlu.shape[0] >= 2000
    g = brille_b200.accelerate(wl.grid)
    orc = Oracle(bridge.flatten(wl.grid), bridge.flatten_data(wl.grid))
    Q = wl.make_q(100_000, 21)
    step = 25_000
    seen = 0
    for lo in range(0, len(Q), step):
        q = Q[lo:lo + step]
        vals, vecs, pr = g.ir_interpolate_at(q, probe=True)
        seen |= g.last_path
        rc, ov, ow, opr = orc.interpolate_at(q)
        assert rc == 0
        assert_decisions_equal(pr, probe_dict(opr), "C4 cuda vs oracle", adaptive_ulps=4)
        assert np.array_equal(pr.tet, opr.tet) and np.array_equal(pr.status, opr.status)
        assert_values_close(vals, ov)
        assert_values_close(vecs, ow)
        if lo == 0:
            rv, rw = wl.grid.ir_interpolate_at(q[:4000], True, 8)
            assert_values_close(vals[:4000], rv)
            assert_values_close(vecs[:4000], rw)
            v2, w2 = g.ir_interpolate_at(q)  # lean outputs of the location (no probe): same bits
            assert np.array_equal(v2, vals) and np.array_equal(w2, vecs)
        del vals, vecs, ov, ow
    assert seen & 4, "the pipelined cell kernel never ran on C4"
    g.close()


def test_c2_at_its_own_configuration(host, bridge):
    """BASELINE configs[1]: NaCl primitive cell, trellis V_ir/1000, 6 modes, 1e6 Q: decisions bit-identical to the oracle on
    every point, values on a 2e5 slice (the oracle is scalar), the reference on 2e4."""
    wl = W.c2_nacl(host)
    g = brille_b200.accelerate(wl.grid)
    orc = Oracle(bridge.flatten(wl.grid), bridge.flatten_data(wl.grid))
    Q = wl.make_q(1_000_000, 2)
    vals, vecs, pr = g.ir_interpolate_at(Q, probe=True)
    assert g.hot_path_taken
    rc, ov, ow, opr = orc.interpolate_at(Q[:200_000])
    assert rc == 0
    assert_decisions_equal(pr, probe_dict(opr), "C2 cuda vs oracle")
    assert_values_close(vals[:200_000], ov)
    assert_values_close(vecs[:200_000], ow)
    rc, opr = orc.moveinto(Q, 1)
    assert np.array_equal(pr.tau, opr.tau) and np.array_equal(pr.ridx, opr.ridx) and np.array_equal(pr.q_ir, opr.q_ir)
    rv, rw = wl.grid.ir_interpolate_at(Q[:20000], True, 8)
    assert_values_close(vals[:20000], rv)
    assert_values_close(vecs[:20000], rw)
    g.close()


def test_c5_full_size_sharded_over_all_devices(host, bridge):
    """BASELINE configs[4] at 1e7 Q (per visible device at most 1e7; all visible devices take a contiguous shard): powder-average
    Q with |Q| up to 10 1/angstrom on the C3 trellis, device resident.  Every point is placed; a 1e5-point sample of every shard
    (every k-th row, read back) has decisions bit-identical to the oracle and values within 1e-10; the eigenvalues are
    invariant under Q -> Q + G on the full set."""
    import torch

    from brille_b200.sharding import shard_bounds

    wl = W.c3_p63mmc(host)
    B = np.asarray(bridge.flatten_bz(wl.bz)["to_xyz"])
    n = 10_000_000
    Q = W.powder_q(B, n, 77)
    orc = Oracle(bridge.flatten(wl.grid), bridge.flatten_data(wl.grid))
    ndev = torch.cuda.device_count()
    for rank in range(ndev):
        lo, hi = shard_bounds(n, rank, ndev)
        g = brille_b200.accelerate(wl.grid, device=rank)
        dev = torch.device("cuda", rank)
        dQ = torch.from_numpy(Q[lo:hi]).to(dev)
        vals, vecs = g.ir_interpolate_at_device(dQ, check=True)  # raises if any point is not placed
        assert g.hot_path_taken
        k = max(1, (hi - lo) // 100_000)
        sel = torch.arange(0, hi - lo, k, device=dev)
        sv, sw = vals[sel].cpu().numpy(), vecs[sel].cpu().numpy()
        qs = Q[lo:hi][::k]
        hv, hw, pr = g.ir_interpolate_at(qs, probe=True)  # the same points through the host-buffer call, with the decisions
        assert np.abs(pr.tau).max() >= 3
        rc, ov, ow, opr = orc.interpolate_at(qs)
        assert rc == 0
        assert_decisions_equal(pr, probe_dict(opr), "C5 cuda vs oracle")
        assert_values_close(sv, ov)
        assert_values_close(sw, ow)
        assert rel_err(hv, ov) <= 1e-10 and rel_err(hw, ow) <= 1e-10
        shift = torch.randint(-3, 4, dQ.shape, device=dev, generator=torch.Generator(device=dev).manual_seed(1)).to(torch.float64)
        del vecs
        v2, w2 = g.ir_interpolate_at_device(dQ + shift, check=True)
        assert float(((v2 - vals).abs() / vals.abs()).max().item()) <= 1e-9
        del v2, w2, vals, dQ
        g.close()
        torch.cuda.empty_cache()


@pytest.mark.parametrize("which,cartesian", [("prim", False), ("conv", False), ("conv", True)])
def test_gamma_rotated_matrices(host, bridge, which, cartesian):
    """RotatesLike::Gamma data with 3x3 matrices (interpolator_gamma.tpp:116-134), real_lattice and angstrom: the general kernel
    against the oracle (which is pinned bit for bit on the reference for these grids) and against the reference itself."""
    wl = W.gamma_matrix_grid(host, which, cartesian=cartesian)
    g = brille_b200.accelerate(wl.grid)
    Q = wl.make_q(20000, 5)
    vals, vecs, pr = g.ir_interpolate_at(Q, probe=True)
    orc = Oracle(bridge.flatten(wl.grid), bridge.flatten_data(wl.grid))
    rc, ov, ow, opr = orc.interpolate_at(Q)
    assert rc == 0
    assert_decisions_equal(pr, probe_dict(opr), "cuda vs oracle")
    assert_values_close(vals, ov)
    assert_values_close(vecs, ow)
    rv, rw = wl.grid.ir_interpolate_at(Q[:4000], True, 4)
    assert_values_close(vals[:4000], rv)
    assert_values_close(vecs[:4000], rw)
    g.close()


@pytest.mark.parametrize("cls,args", [("BZNestQdc", (9,)), ("BZNestQdc", (5,)), ("BZMeshQdc", (3,)), ("BZTrellisQdc", ())])
def test_points_shared_by_many_cells(host, bridge, cls, args):
    """Every grid vertex, every edge midpoint of the first cells and points a few ulp off them: points that many tetrahedra share
    (dozens of containing nodes per level of a Nest -- the first round's descent kept at most 32 pending and gave up beyond; the
    reference's queue is unbounded, nest.hpp:163-222), the 'last all-positive else first' and zero-weight folding rules, compact
    emission.  Decisions against the oracle, values against the reference."""
    lat = W.p63mmc_lattice(host)
    bz = host.BrillouinZone(lat)
    hg = getattr(host, cls)(bz, bz.ir_polyhedron.volume / 400, *args)
    W._gamma_fill(hg, 12, 4, 3)
    g = brille_b200.accelerate(hg)
    v = np.asarray(hg.rlu)
    rng = np.random.default_rng(4)
    i, j = rng.integers(0, len(v), (2, 3000))
    mid = 0.5 * (v[i] + v[j])
    Q = np.vstack([v, mid, v * (1 + 2.3e-16), v + 1e-13 * rng.normal(size=v.shape), np.zeros((1, 3))])
    keep = np.asarray(bz.isinside(Q), dtype=bool)  # (midpoints of non-adjacent vertices may leave the irreducible zone: moved anyway)
    Q = np.vstack([Q, Q[keep][:2000] + rng.integers(-2, 3, (min(2000, int(keep.sum())), 3))])
    orc = Oracle(bridge.flatten(hg), bridge.flatten_data(hg))
    # the reference does not find every such point (a vertex displaced by an ulp can fall between the tolerances of all its
    # neighbours: "N points not found", all-or-nothing) -- the points it cannot place are dropped (the oracle restates the search)
    rc, _, _, opr = orc.interpolate_at(Q)
    n_all = len(Q)
    with pytest.raises(RuntimeError) if rc else __import__("contextlib").nullcontext():
        g.ir_interpolate_at(Q)  # the same verdict for the whole set
    Q = Q[(opr.status & 7) == 0]
    assert len(Q) > 0.9 * n_all
    for no_move in (False,):
        vals, vecs, pr = g.ir_interpolate_at(Q, probe=True)
        rc, ov, ow, opr = orc.interpolate_at(Q)
        assert rc == 0
        assert_decisions_equal(pr, probe_dict(opr), f"{cls} cuda vs oracle", adaptive_ulps=0 if "Trellis" in cls else 4)
        assert np.array_equal(pr.tet, opr.tet) and np.array_equal(pr.status, opr.status)
        assert_values_close(vals, ov)
        assert_values_close(vecs, ow)
        rv, rw = hg.ir_interpolate_at(Q, True, 8)
        assert_values_close(vals, rv)
        assert_values_close(vecs, rw)
    # the same set, many times over, takes the two-kernel location and the cell kernels: same bits as the small call
    big = np.tile(Q, (max(1, 400_000 // len(Q)), 1))
    vb, wb, prb = g.ir_interpolate_at(big, probe=True)
    n = len(Q)
    assert np.array_equal(prb.vertex[:n], pr.vertex) and np.array_equal(prb.weight[:n], pr.weight) and np.array_equal(prb.tet[:n], pr.tet)
    assert rel_err(vb[:n], vals) <= 1e-12 and rel_err(wb[:n], vecs) <= 1e-12
    g.close()


@pytest.mark.parametrize("which", ["C3", "C2", "F23", "P3_timereversal", "Im-3m"])
def test_wedge_rotation_of_far_points(host, bridge, which):
    """ir_moveinto_wedge does not translate (bz_move.cpp:299-356): |Q| up to 1e3 rlu reaches the sign-pattern lookup and the
    certified wedge test, whose rounding bounds must scale with |Q|.  Bit for bit against the reference's own method and the oracle
    (including the far points on wedge planes that no operation places: the reference returns q = 0, operation 0, no error);
    isinside for the same points."""
    wl = W.BUILDERS[which](host) if which in W.BUILDERS else W.zoo_grid(host, which)
    g = brille_b200.accelerate(wl.grid)
    bz = wl.bz
    from test_oracle_vs_reference import far_points

    Q = far_points()
    rots = np.asarray(bridge.flatten_bz(bz)["rotations"]).reshape(-1, 3, 3)
    qw, rw = g.ir_moveinto_wedge(Q)
    rqw, rRw = bz.ir_moveinto_wedge(Q)
    assert np.array_equal(qw, rqw) and np.array_equal(rots[rw], rRw)
    orc = Oracle(bridge.flatten(wl.grid), bridge.flatten_data(wl.grid))
    rc, opr = orc.moveinto(Q, 2)
    assert np.array_equal(qw, opr.q_ir) and np.array_equal(rw, opr.ridx)
    assert np.array_equal(g.isinside(Q), np.asarray(bz.isinside(Q), dtype=bool))
    g.close()


def test_refill_with_other_row_sizes(host, bridge):
    """fill() again with more modes / another span on the same grid (the staging buffers of the host pipeline were
    sized in points of the first fill's rows): same number of Q before and after, host-buffer, pinned and consumer calls."""
    lat = W.p63mmc_lattice(host)
    bz = host.BrillouinZone(lat)
    hg = host.BZTrellisQdc(bz, bz.ir_polyhedron.volume / 300)
    g = brille_b200.accelerate(hg)
    Q = np.random.default_rng(1).uniform(-3, 3, (150_000, 3))
    rng = np.random.default_rng(2)
    for modes, seed in ((3, 1), (12, 2), (6, 3)):
        nv = hg.rlu.shape[0]
        r = np.random.default_rng(seed)
        vals = r.uniform(1, 50, (nv, modes, 1))
        vecs = r.normal(size=(nv, modes, 4, 3)) + 1j * r.normal(size=(nv, modes, 4, 3))
        g.fill(vals, (1, 0, 0, 0, 3), vecs, (0, 12, 0, 2, 3))
        a = g.ir_interpolate_at(Q)
        b = g.ir_interpolate_at(Q, pinned=True)
        assert a[0].shape == (len(Q), modes, 1) and a[1].shape == (len(Q), modes, 4, 3)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
        rv, rw = hg.ir_interpolate_at(Q[:5000], True, 8)
        assert_values_close(a[0][:5000], rv)
        assert_values_close(a[1][:5000], rw)
        g.set_structure_factor(rng.normal(size=4) + 1j * rng.normal(size=4), positions=rng.uniform(0, 1, (4, 3)))
        v1, sf1 = g.ir_structure_factor(Q)
        g.set_option("sf_fused", 0)
        v0, sf0 = g.ir_structure_factor(Q)
        g.set_option("sf_fused", 1)
        assert np.array_equal(v1, a[0]) and float(np.abs(sf1 - sf0).max() / np.abs(sf0).max()) <= 1e-12
    g.close()


def test_device_outputs_are_validated(host):
    import torch

    wl = W.c2_nacl(host, density=100)
    g = brille_b200.accelerate(wl.grid)
    dQ = torch.from_numpy(wl.make_q(1000, 1)).cuda()
    with pytest.raises(RuntimeError, match="vals_out"):
        g.ir_interpolate_at_device(dQ, vals_out=torch.empty((999, 6, 1), dtype=torch.float64, device="cuda"))
    with pytest.raises(RuntimeError, match="vecs_out"):
        g.ir_interpolate_at_device(dQ, vecs_out=torch.empty((1000, 6, 2, 3), dtype=torch.complex64, device="cuda"))
    with pytest.raises(RuntimeError, match="vecs_out"):
        g.ir_interpolate_at_device(dQ, vecs_out=torch.empty((1000, 6, 2, 6), dtype=torch.complex128, device="cuda")[..., ::2])
    g.close()
    s, d, _, rest = __import__("helpers").load_golden("nacl_prim_trellis.npz")
    gf = brille_b200.B200Grid(None, structure=s, data=d)
    with pytest.raises(RuntimeError, match="host grid"):
        gf.sort()
    gf.close()
