"""GPU: the regression cases of the reference's own C++ tests for this path, restated through the Python API and run against
the CUDA library (the reference's test strategy).  Each test names the Catch2 case it follows.

* src/tests/trellis_test.cpp:76-137   "Simple BrillouinZoneTrellis3 interpolation"   (a linear field is reproduced)
* src/tests/nest_test.cpp:45-147       "Simple BrillouinZoneNest3 interpolation"
* src/tests/nest_test.cpp:149-218      "Random BrillouinZoneNest3 interpolation"
* src/tests/trellis_test.cpp:344-349   "... 'P1' hexagonal system must contain Gamma"
* src/tests/trellis_test.cpp:351-373   "PolyNode inclusion rounding error" (quartz, Q = (-0.1 + 10^k, -0.1, 0))
* src/tests/trellis_test.cpp:375-441   "BrillouinZoneTrellis3 inclusion data race error" (La2Zr2O7 from generators, 5000 Q)
* src/tests/brillouinzone_test.cpp:114-155  moveinto / ir_moveinto invariants (Q = q + tau, Q = R^-T q_ir + tau)
"""
import numpy as np
import pytest

import brille_b200
from oracle.oracle import Oracle
from helpers import assert_decisions_equal, assert_values_close

pytestmark = pytest.mark.gpu


def probe_dict(pr):
    return {"tau": pr.tau, "q_ir": pr.q_ir, "ridx": pr.ridx, "invridx": pr.invridx, "n_vert": pr.n_vert, "vertex": pr.vertex, "weight": pr.weight}


def points_in_polyhedron(poly, n, seed, margin=1e-6):
    """Rejection sampling inside a convex brille Polyhedron (Polyhedron::rand_rejection is not bound to Python): Cartesian
    points strictly inside every face plane."""
    rng = np.random.default_rng(seed)
    v = np.asarray(poly.vertices)
    nrm, pts = np.asarray(poly.normals), np.asarray(poly.points)
    lo, hi = v.min(axis=0), v.max(axis=0)
    out = np.empty((0, 3))
    while len(out) < n:
        x = rng.uniform(lo, hi, (4 * n, 3))
        d = np.einsum("fj,nfj->nf", nrm, x[:, None, :] - pts[None, :, :])
        out = np.vstack([out, x[(d < -margin).all(axis=1)]])
    return out[:n]


def nb_lattice(b):
    return b.Lattice((3.2598, 3.2598, 3.2598), (90, 90, 90), "-I 4 2 3")


def fill_with_positions(grid):
    """values: one mode whose 3 elements are the vertex position in 1/angstrom, rotating like a reciprocal-lattice vector:
    linear interpolation reproduces the position of the interpolation point."""
    xyz = np.ascontiguousarray(grid.invA)
    tostore = xyz.reshape(-1, 1, 3)
    grid.fill(tostore, (0, 3, 0, 0, 4), tostore.copy(), (0, 3, 0, 0, 4))
    return xyz


@pytest.mark.parametrize("cls,arg,tol", [("BZTrellisQdd", 0.0001, 2e-10), ("BZNestQdd", 0.01, 2e-14), ("BZMeshQdd", 0.01, 2e-10)])
def test_linear_field_is_reproduced(host, bridge, cls, arg, tol):
    """trellis_test.cpp:76-137, nest_test.cpp:45-218: with the vertex positions stored as one 3-vector mode, the interpolated
    vector at points inside the irreducible polyhedron IS the point (|diff| < 2e-10 trellis, 2e-14 nest in the reference's
    tests), with and without ir_moveinto (tau = 0, identity operation); and for points anywhere the result is q_ir rotated
    back, R^-T x_ir."""
    b = host
    bz = b.BrillouinZone(nb_lattice(b))
    grid = getattr(b, cls)(bz, arg, 5) if cls.startswith("BZNest") else getattr(b, cls)(bz, arg)
    fill_with_positions(grid)
    assert grid.ir_interpolate_at(np.zeros((1, 3)))[0].shape[1] == 1   # branches() == 1, or the vector is not a vector
    g = brille_b200.accelerate(grid)
    fbz = bridge.flatten_bz(bz)
    B, ident = np.asarray(fbz["to_xyz"], dtype=float).reshape(3, 3), int(fbz["identity_index"])
    x = points_in_polyhedron(bz.ir_polyhedron, 20000, 5)
    Q = x @ np.linalg.inv(B).T
    for no_move in (True, False):
        vals, vecs, pr = g.ir_interpolate_at(Q, probe=True, do_not_move_points=no_move)
        assert np.abs(vals.reshape(-1, 3) - x).max() < tol
        assert np.abs(vecs.reshape(-1, 3) - x).max() < tol
        assert not pr.tau.any()
        if not no_move:   # nest_test.cpp:117-123: points of the irreducible polyhedron are not moved
            assert np.abs(pr.q_ir - Q).max() < 1e-12 and (pr.ridx == ident).all()
        rv, rw = grid.ir_interpolate_at(Q[:500], False, 1, no_move)
        assert_values_close(vals[:500], rv)
    # anywhere in reciprocal space: the stored vector rotates back with the point (Q - tau in 1/angstrom)
    Qa = np.random.default_rng(9).uniform(-3, 3, (200_000, 3))
    vals, vecs, pr = g.ir_interpolate_at(Qa, probe=True)
    want = (Qa - pr.tau) @ B.T
    assert np.abs(vals.reshape(-1, 3) - want).max() < 100 * tol
    g.close()


def test_p1_hexagonal_trellis_contains_gamma(host, bridge):
    """trellis_test.cpp:344-349: a 'P 1' hexagonal RECIPROCAL lattice; construction must succeed and the origin must be
    locatable; then the whole path against the oracle on random points."""
    b = host
    lat = b.Lattice((1.154701, 1.154701, 1), (90, 90, 60), "P 1", real_space=False)
    bz = b.BrillouinZone(lat)
    grid = b.BZTrellisQdd(bz, 0.002)
    nv = grid.rlu.shape[0]
    rng = np.random.default_rng(3)
    grid.fill(rng.normal(size=(nv, 2, 1)), (1,), rng.normal(size=(nv, 2, 3)), (0, 3, 0, 0, 4))
    g = brille_b200.accelerate(grid)
    orc = Oracle(bridge.flatten(grid), bridge.flatten_data(grid))
    Q = np.vstack([np.zeros((1, 3)), rng.uniform(-2, 2, (50_000, 3))])
    vals, vecs, pr = g.ir_interpolate_at(Q, probe=True)
    rc, ov, ow, opr = orc.interpolate_at(Q)
    assert rc == 0
    assert_decisions_equal(pr, probe_dict(opr), "P1 hexagonal")
    assert_values_close(vals, ov)
    assert_values_close(vecs, ow)
    rv, rw = grid.ir_interpolate_at(Q[:2000], False, 1)
    assert_values_close(vals[:2000], rv)
    assert_values_close(vecs[:2000], rw)
    g.close()


def quartz_trellis(b):
    lat = b.Lattice((4.85235, 4.85235, 5.350305), (90, 90, 120), 'P 32 2"')
    bz = b.BrillouinZone(lat)
    grid = b.BZTrellisQdd(bz, bz.ir_polyhedron.volume / 2000.0)
    nv = grid.rlu.shape[0]
    return bz, grid, nv


def test_quartz_polynode_inclusion_rounding(host, bridge):
    """trellis_test.cpp:351-373 (issue 61): Q = (-0.1 + 10^k, -0.1, 0), k = -15 .. -1, sits within rounding of a PolyNode
    face in the quartz trellis; the reference must not throw.  Here: neither does the device path, its decisions are those of
    the oracle and the reference bit for bit, also for a dense set of points around the same face."""
    b = host
    bz, grid, nv = quartz_trellis(b)
    rng = np.random.default_rng(61)
    grid.fill(rng.normal(size=(nv, 1)), (1, 0, 0, 0, 4), rng.normal(size=(nv, 1)), (1, 0, 0, 0, 4))
    g = brille_b200.accelerate(grid)
    orc = Oracle(bridge.flatten(grid), bridge.flatten_data(grid))
    Q = np.array([[-0.1 + 10.0 ** k, -0.1, 0.0] for k in range(-15, 0)])
    # the same line at 4000 more offsets in both directions, log-spaced: every rounding regime around the face
    d = 10.0 ** rng.uniform(-16, -1, 4000) * rng.choice([-1.0, 1.0], 4000)
    more = np.stack([-0.1 + d, np.full_like(d, -0.1), np.zeros_like(d)], axis=1)
    Q = np.vstack([Q, more, more[:, [1, 0, 2]], more + [0, 0, 0.5]])
    vals, vecs, pr = g.ir_interpolate_at(Q, probe=True)       # REQUIRE_NOTHROW
    rc, ov, ow, opr = orc.interpolate_at(Q)
    assert rc == 0
    assert_decisions_equal(pr, probe_dict(opr), "quartz face")
    rv, rw = grid.ir_interpolate_at(Q, False, 1)                # the reference itself does not throw either
    assert_values_close(vals, rv)
    assert_values_close(vecs, rw)
    assert_values_close(vals, ov)
    g.close()


def la2zr2o7_lattice(b):
    """trellis_test.cpp:380-404: rhombohedral primitive cell of the pyrochlore, space group given by its generators."""
    W = np.array([[[0, -1, 0], [0, 0, -1], [1, 1, 1]],
                  [[-1, -1, -1], [0, 0, 1], [0, 1, 0]],
                  [[-1, -1, -1], [1, 0, 0], [0, 0, 1]],
                  [[-1, 0, 0], [0, -1, 0], [0, 0, -1]]], dtype=np.int32)
    w = np.array([[0, 0, 0.5], [0.5, 0, 0], [0.5, 0, 0], [0, 0, 0]], dtype=float)
    sym = b.Symmetry(W, w).generate()
    a = 7.583912824346341
    return b.Lattice((a, a, a), (60, 60, 60), sym)


def test_la2zr2o7_from_generators(host, bridge):
    """trellis_test.cpp:375-441 (issue 60): a coarse trellis (V_ir/20) of the 48-operation pyrochlore group built from its
    generators; 5000 Q on the line (1.5, -1.5..2, 4..-3).  The reference's case checks that no thread count makes the
    interpolation throw; here the massively parallel path must not throw and must agree with the one-thread reference and
    the oracle."""
    b = host
    lat = la2zr2o7_lattice(b)
    bz = b.BrillouinZone(lat)
    grid = b.BZTrellisQdd(bz, bz.ir_polyhedron.volume / 20.0)
    nv = grid.rlu.shape[0]
    rng = np.random.default_rng(60)
    grid.fill(rng.normal(size=(nv, 2, 1)), (1, 0, 0, 0, 4), rng.normal(size=(nv, 2, 3)), (0, 3, 0, 0, 4))
    n = 5000
    j = np.arange(n)
    Q = np.stack([np.full(n, 1.5), -1.5 + j * (3.5 / n), 4.0 + j * (-7.0 / n)], axis=1)
    g = brille_b200.accelerate(grid)
    orc = Oracle(bridge.flatten(grid), bridge.flatten_data(grid))
    vals, vecs, pr = g.ir_interpolate_at(Q, probe=True)
    rc, ov, ow, opr = orc.interpolate_at(Q)
    assert rc == 0
    assert_decisions_equal(pr, probe_dict(opr), "La2Zr2O7 line")
    rv, rw = grid.ir_interpolate_at(Q, False, 1)
    assert_values_close(vals, rv)
    assert_values_close(vecs, rw)
    # the special point of trellis_test.cpp:541-632 and its neighbourhood, in a fine trellis (V_ir/1e4)
    fine = b.BZTrellisQdd(bz, bz.ir_polyhedron.volume / 1e4)
    nvf = fine.rlu.shape[0]
    fine.fill(rng.normal(size=(nvf, 1)), (1, 0, 0, 0, 4), rng.normal(size=(nvf, 1)), (1, 0, 0, 0, 4))
    q0 = np.array([0.71056334, 0.87427782, 0.70768662])
    Qs = np.vstack([q0[None], q0 + rng.normal(scale=1e-9, size=(2000, 3)), q0 + rng.normal(scale=1e-3, size=(2000, 3))])
    gf = brille_b200.accelerate(fine)
    of = Oracle(bridge.flatten(fine), bridge.flatten_data(fine))
    vals, vecs, pr = gf.ir_interpolate_at(Qs, probe=True)
    rc, ov, ow, opr = of.interpolate_at(Qs)
    assert rc == 0
    assert_decisions_equal(pr, probe_dict(opr), "La2Zr2O7 special point")
    assert pr.n_vert[0] == 4                                   # trellis_test.cpp:629: the point is found in a PolyNode (a tetrahedron)
    rv, rw = fine.ir_interpolate_at(Qs, False, 1)
    assert_values_close(vals, rv)
    g.close()
    gf.close()


@pytest.mark.parametrize("sg,lengths,angles", [("P 1", (2.87,) * 3, (90, 90, 90)),      # the three of brillouinzone_test.cpp
                                                ("Im-3m", (2.87,) * 3, (90, 90, 90)),
                                                ("Fd-3c", (2.87,) * 3, (90, 90, 90)),
                                                ('P 32 2"', (4.85235, 4.85235, 5.350305), (90, 90, 120))])
def test_moveinto_invariants_at_scale(host, bridge, sg, lengths, angles):
    """brillouinzone_test.cpp:114-155: for random Q, (1) moveinto gives Q = q + tau with q inside the first zone, (2)
    ir_moveinto gives Q = R^-T q_ir + tau with q_ir inside the irreducible zone -- checked on 1e6 points from the device
    probes: identities in exact integer / rounding-level arithmetic, membership with the reference's own isinside on a slice
    (the device isinside on all of them)."""
    b = host
    lat = b.Lattice(lengths, angles, sg)
    bz = b.BrillouinZone(lat)
    grid = b.BZTrellisQdd(bz, bz.ir_polyhedron.volume / 50.0)
    nv = grid.rlu.shape[0]
    grid.fill(np.zeros((nv, 1)), (1,), np.zeros((nv, 1)), (1,))
    g = brille_b200.accelerate(grid)
    fl = bridge.flatten(grid)
    rot = np.asarray(fl["bz"]["rotations"], dtype=np.int64).reshape(-1, 3, 3)
    Q = np.random.default_rng(114).uniform(-5, 5, (1_000_000, 3))
    q, tau = g.moveinto(Q)                               # first zone
    assert np.abs(q + tau - Q).max() < 1e-13
    assert g.isinside(q).all()
    assert np.asarray(bz.isinside(q[:20000])).all()
    q_ir, tau_ir, ridx, invridx = g.ir_moveinto(Q)
    assert np.array_equal(rot[ridx] @ rot[invridx], np.broadcast_to(np.eye(3, dtype=np.int64), (len(Q), 3, 3)))
    # Q = R^T q_ir + tau with the matrices BrillouinZone.ir_moveinto hands to Python (wrap/_bz.cpp:452-461)
    rebuilt = np.einsum("nji,nj->ni", rot[ridx].astype(float), q_ir) + tau_ir
    assert np.abs(rebuilt - Q).max() < 1e-12
    assert g.isinside(q_ir).all()
    hq, htau, hr, hinv = bz.ir_moveinto(Q[:20000])
    assert np.array_equal(htau, tau_ir[:20000]) and np.array_equal(hq, q_ir[:20000])
    assert np.array_equal(hr, rot[ridx[:20000]]) and np.array_equal(hinv, rot[invridx[:20000]])
    g.close()
