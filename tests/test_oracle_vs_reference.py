"""CPU: the oracle restatement pinned against the reference itself (oracle/_ref) on freshly built grids."""
import numpy as np
import pytest

from oracle.oracle import Oracle
from brille_b200 import workloads as W


def _compare(host, probe, bridge, wl, n, seed, sort=False):
    g = wl.grid
    if sort:
        g.sort()
    orc = Oracle(bridge.flatten(g), bridge.flatten_data(g))
    Q = wl.make_q(n, seed)
    rv, rw = g.ir_interpolate_at(Q, True, 4)
    rc, ov, ow, pr = orc.interpolate_at(Q)
    assert rc == 0
    q, x, tau, r, ri = probe.ir_moveinto_idx(wl.bz, Q, 1)
    assert np.array_equal(tau, pr.tau) and np.array_equal(r, pr.ridx) and np.array_equal(ri, pr.invridx)
    assert np.array_equal(q, pr.q_ir) and np.array_equal(x, pr.x_ir)
    cnt, idx, wgt = probe.indices_weights(g, x)
    assert np.array_equal(cnt, pr.n_vert) and np.array_equal(idx, pr.vertex) and np.array_equal(wgt, pr.weight)
    assert np.array_equal(rv.reshape(ov.shape), ov)
    assert np.array_equal(rw.reshape(ow.shape), ow)


def test_c1_fd3m_scalar(host, probe, bridge):
    _compare(host, probe, bridge, W.c1_fd3m_scalar(host, density=500), 20000, 1)


def test_c2_nacl(host, probe, bridge):
    _compare(host, probe, bridge, W.c2_nacl(host, density=300), 20000, 2)


def test_c3_p63mmc(host, probe, bridge):
    _compare(host, probe, bridge, W.c3_p63mmc(host, density=500), 20000, 3)


def test_c3_p63mmc_sorted(host, probe, bridge):
    _compare(host, probe, bridge, W.c3_p63mmc(host, density=100, seed=5), 5000, 4, sort=True)


def test_powder_q_large_tau(host, probe, bridge):
    wl = W.c3_p63mmc(host, density=200)
    B = np.asarray(bridge.flatten_bz(wl.bz)["to_xyz"])
    wl.make_q = lambda n, seed: W.powder_q(B, n, seed)
    _compare(host, probe, bridge, wl, 10000, 6)


@pytest.mark.parametrize("cls,args", [("BZNestQdc", (5,)), ("BZMeshQdc", (3,)), ("BZMeshQdc", (1,))])
def test_nest_and_mesh(host, probe, bridge, cls, args):
    lat = W.p63mmc_lattice(host)
    bz = host.BrillouinZone(lat)
    g = getattr(host, cls)(bz, bz.ir_polyhedron.volume / 300, *args)
    W._gamma_fill(g, 12, 4, 9)
    wl = W.Workload(cls, g, bz, 12, 4, lambda n, seed: np.random.default_rng(seed).uniform(-3, 3, (n, 3)))
    _compare(host, probe, bridge, wl, 10000, 8)


def test_c4_nest_low_symmetry(host, probe, bridge):
    _compare(host, probe, bridge, W.c4_p21c_nest(host, density=200), 3000, 9)


SPECIAL = np.array([[0, 0, 0], [0.5, 0, 0], [0.5, 0.5, 0], [0.5, 0.5, 0.5], [1, 0, 0], [0.25, 0.25, 0], [1 / 3, 1 / 3, 0], [0, 0, 0.5],
                    [-0.5, 0, 0], [2, 1, 0], [0.1, 0.1, 0.1], [-1.5, 2.5, 0.5], [0.75, 0.25, 0.5]], dtype=float)


@pytest.mark.parametrize("cls", ["BZTrellisQcc", "BZTrellisQdd", "BZTrellisQdc"])
@pytest.mark.parametrize("name", sorted(W.ZOO))
def test_lattice_zoo(host, probe, bridge, name, cls):
    """centred / rhombohedral / triclinic lattices, added time reversal; pseudovector, reciprocal-vector and matrix data"""
    wl = W.zoo_grid(host, name, cls)
    base = wl.make_q
    wl.make_q = lambda n, seed: np.vstack([SPECIAL, base(n, seed)])
    _compare(host, probe, bridge, wl, 3000, 5)


@pytest.mark.parametrize("which", ["C2", "C3", "F23", "P3_timereversal", "P-1"])
def test_bz_methods(host, bridge, which):
    """BrillouinZone.isinside / moveinto / ir_moveinto_wedge (wrap/_bz.cpp:378-520) restated by the oracle, bit for bit."""
    from brille_b200 import tables as T

    wl = W.BUILDERS[which](host, density=100) if which in W.BUILDERS else W.zoo_grid(host, which)
    bz = wl.bz
    orc = Oracle(bridge.flatten(wl.grid), bridge.flatten_data(wl.grid))
    rng = np.random.default_rng(8)
    Q = np.vstack([rng.uniform(-1.5, 1.5, (6000, 3)), rng.integers(-4, 5, (1000, 3)) / 4.0, np.zeros((1, 3))])
    rots = np.asarray(bridge.flatten_bz(bz)["rotations"]).reshape(-1, 3, 3)
    rc, pr = orc.moveinto(Q, 3)
    assert rc == 0 and np.array_equal((pr.status & T.ST_OUTSIDE_BZ) == 0, np.asarray(bz.isinside(Q), dtype=bool))
    rc, pr = orc.moveinto(Q, 0)
    rq, rtau = bz.moveinto(Q)
    assert rc == 0 and np.array_equal(pr.tau, rtau) and np.array_equal(pr.q_ir, rq)
    rc, pr = orc.moveinto(Q, 2)
    rqw, rRw = bz.ir_moveinto_wedge(Q)
    assert rc == 0 and np.array_equal(pr.q_ir, rqw) and np.array_equal(rots[pr.ridx], rRw) and not pr.tau.any()


def far_points(seed=11, n=40000):
    rng = np.random.default_rng(seed)
    base = rng.uniform(-1, 1, (n, 3))
    scale = 10.0 ** rng.uniform(-3, 3, (n, 1))
    special = rng.integers(-4, 5, (n // 10, 3)).astype(float)  # on the wedge planes, scaled exactly by powers of two
    return np.vstack([base * scale, special * 2.0 ** rng.integers(-3, 9, (n // 10, 1)), np.zeros((1, 3))])


@pytest.mark.parametrize("which", ["C3", "C2", "F23", "P3_timereversal", "Im-3m"])
def test_wedge_rotation_of_far_points(host, bridge, which):
    """ir_moveinto_wedge does not translate: |Q| up to 1e3 rlu.  Far points ON a wedge plane can fail every operation's tolerance
    test; the reference then returns the zero-initialised row (q = 0, operation 0) without an error (bz_move.cpp:348-354)."""
    from brille_b200 import tables as T

    wl = W.BUILDERS[which](host, density=100) if which in W.BUILDERS else W.zoo_grid(host, which)
    bz = wl.bz
    orc = Oracle(bridge.flatten(wl.grid), bridge.flatten_data(wl.grid))
    Q = far_points()
    rots = np.asarray(bridge.flatten_bz(bz)["rotations"]).reshape(-1, 3, 3)
    rc, pr = orc.moveinto(Q, 2)
    rqw, rRw = bz.ir_moveinto_wedge(Q)
    assert rc == 0 and np.array_equal(pr.q_ir, rqw) and np.array_equal(rots[pr.ridx], rRw)
    rc, pr = orc.moveinto(Q, 3)
    assert np.array_equal((pr.status & T.ST_OUTSIDE_BZ) == 0, np.asarray(bz.isinside(Q), dtype=bool))


@pytest.mark.parametrize("which,cartesian", [("prim", False), ("conv", False), ("conv", True)])
def test_gamma_vectors_and_matrices(host, probe, bridge, which, cartesian):
    """RotatesLike::Gamma data with 3x3 matrices (interpolator_gamma.tpp:116-134): the reference rotates the first Nmat x Nmat
    matrices and copies the rest of the mode back from the work array it shares with the vector pass."""
    wl = W.gamma_matrix_grid(host, which, cartesian=cartesian)
    _compare(host, probe, bridge, wl, 4000, 23)
