"""GPU: the CUDA path, called through the C ABI, against the oracle, the golden fixtures and the reference."""
import numpy as np
import pytest

import brille_b200
from brille_b200 import tables as T
from brille_b200 import workloads as W
from oracle.oracle import Oracle
from helpers import RTOL, assert_decisions_equal, assert_values_close, load_golden, ref_decisions, remove_mode_phase

pytestmark = pytest.mark.gpu

FIXTURES = ["nacl_prim_trellis.npz", "nacl_prim_trellis_sorted.npz", "fd3m_scalar_trellis.npz", "p63mmc_trellis.npz", "p1_trellis_dd.npz",
            "p63mmc_nest.npz", "p63mmc_nest_sorted.npz", "p63mmc_mesh.npz"]


# kernel selections every parity case runs under: the general per-(Q,mode) kernel, the cell-batched kernel that stages the
# vertex rows on the fly, and the persistent pipelined cell kernel fed from the per-cell record table (the default)
PATHS = [(1, 0), (2, 1), (2, 2)]
PATH_IDS = ["general", "cell-onthefly", "cell-pipelined"]


def apply_path(g, path):
    g.set_option("interp_path", path[0])
    g.set_option("cell_kernel", path[1])


def rel_close(a, b):
    return float(np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(b).max(), 1e-300))


def probe_dict(pr):
    return {"tau": pr.tau, "q_ir": pr.q_ir, "ridx": pr.ridx, "invridx": pr.invridx, "n_vert": pr.n_vert, "vertex": pr.vertex, "weight": pr.weight}


@pytest.mark.parametrize("path", PATHS, ids=PATH_IDS)
@pytest.mark.parametrize("name", FIXTURES)
def test_golden_fixtures(name, path):
    s, d, _, rest = load_golden(name)
    g = brille_b200.B200Grid(None, structure=s, data=d)
    apply_path(g, path)
    vals, vecs, pr = g.ir_interpolate_at(rest["Q"], probe=True)
    assert_decisions_equal(pr, ref_decisions(rest), "cuda", adaptive_ulps=4 if str(s["kind"]) in ("nest", "mesh") else 0)
    assert_values_close(vals, rest["ref_values"])
    assert_values_close(vecs, rest["ref_vectors"])
    assert g.launch_count >= 2
    g.close()


def test_interpolate_at_without_rotation():
    s, d, _, rest = load_golden("p1_trellis_dd.npz")
    g = brille_b200.B200Grid(None, structure=s, data=d)
    vals, vecs, pr = g.interpolate_at(rest["Q"], probe=True)
    assert np.array_equal(pr.tau, rest["ref0_tau"])
    assert_values_close(vals, rest["ref0_values"])
    assert_values_close(vecs, rest["ref0_vectors"])
    q, tau = g.moveinto(rest["Q"])
    assert np.array_equal(tau, rest["ref0_tau"]) and np.array_equal(q, rest["ref0_q_ir"])


def test_nacl_gamma_reference_golden_vectors():
    s, d, d2, rest = load_golden("nacl_gamma.npz")
    g = brille_b200.B200Grid(None, structure=s, data=d)
    vals, vecs = g.ir_interpolate_at(rest["Q"])
    assert_values_close(vals, rest["ref_values"])
    assert_values_close(vecs, rest["ref_vectors"])
    assert np.allclose(vals.reshape(48, 24), rest["golden_euphonic_values"])
    br_vec = np.einsum("ba,ijkb->ijka", rest["golden_basis_vectors"], vecs.reshape(48, 24, 8, 3))
    eu = rest["golden_euphonic_vectors"]
    assert np.allclose(remove_mode_phase(br_vec, eu), eu)
    # Cartesian eigenvectors (LengthUnit::angstrom)
    g._set_data(d2)
    g.set_option("interp_path", 2)
    vals2, vecs2 = g.ir_interpolate_at(rest["Q"])
    assert_values_close(vecs2, rest["ref2_vectors"])
    assert np.allclose(remove_mode_phase(vecs2.reshape(48, 24, 8, 3), eu), eu)


def test_unsupported_combinations_raise():
    s, d, d2, rest = load_golden("nacl_gamma.npz")
    for rl, lu in [(0, 1), (1, 0), (2, 2), (2, 4)]:
        dd = dict(d)
        dd["vectors_rotlike"], dd["vectors_lenunit"] = rl, lu
        g = brille_b200.B200Grid(None, structure=s, data=dd)
        with pytest.raises(RuntimeError):
            g.ir_interpolate_at(rest["Q"])
        g.close()


def test_errors_and_edge_cases():
    s, d, _, rest = load_golden("nacl_prim_trellis.npz")
    g = brille_b200.B200Grid(None, structure=s)
    with pytest.raises(RuntimeError, match="must be filled"):
        g.ir_interpolate_at(rest["Q"])
    g._set_data(d)
    vals, vecs = g.ir_interpolate_at(np.zeros((0, 3)))
    assert vals.shape[0] == 0 and vecs.shape[0] == 0
    with pytest.raises(RuntimeError):
        g.ir_interpolate_at(np.zeros((4, 2)))
    # points far outside the gridded irreducible zone, not moved: all-or-nothing failure like the reference
    with pytest.raises(RuntimeError, match="failed to find|null"):
        g.ir_interpolate_at(np.full((3, 3), 7.3), do_not_move_points=True)
    one = g.ir_interpolate_at(rest["Q"][:1])
    assert one[0].shape[0] == 1


@pytest.mark.parametrize("path", PATHS, ids=PATH_IDS)
@pytest.mark.parametrize("builder,n", [("C1", 200000), ("C2", 100000), ("C3", 100000)])
def test_against_oracle_and_reference(host, bridge, builder, n, path):
    wl = W.BUILDERS[builder](host)
    g = brille_b200.accelerate(wl.grid)
    apply_path(g, path)
    Q = wl.make_q(n, 11)
    vals, vecs, pr = g.ir_interpolate_at(Q, probe=True)
    orc = Oracle(bridge.flatten(wl.grid), bridge.flatten_data(wl.grid))
    rc, ov, ow, opr = orc.interpolate_at(Q)
    assert rc == 0
    assert_decisions_equal(pr, probe_dict(opr), "cuda vs oracle")
    assert np.array_equal(pr.cell, opr.cell) and np.array_equal(pr.tet, opr.tet) and np.array_equal(pr.status, opr.status)
    assert_values_close(vals, ov)
    assert_values_close(vecs, ow)
    # and the reference itself on a slice
    m = min(n, 20000)
    rv, rw = wl.grid.ir_interpolate_at(Q[:m], True, 8)
    assert_values_close(vals[:m], rv)
    assert_values_close(vecs[:m], rw)


@pytest.mark.parametrize("path", PATHS, ids=PATH_IDS)
@pytest.mark.parametrize("cls,args", [("BZNestQdc", (5,)), ("BZMeshQdc", (3,))])
def test_nest_and_mesh_against_oracle_and_reference(host, bridge, cls, args, path):
    """BZNestQdc (nest.hpp) and BZMeshQdc (mesh.hpp, triangulation_layers.hpp) on the C3 lattice."""
    lat = W.p63mmc_lattice(host)
    bz = host.BrillouinZone(lat)
    hg = getattr(host, cls)(bz, bz.ir_polyhedron.volume / 1000, *args)
    W._gamma_fill(hg, 12, 4, 17)
    g = brille_b200.accelerate(hg)
    apply_path(g, path)
    Q = np.random.default_rng(5).uniform(-3, 3, (100000, 3))
    vals, vecs, pr = g.ir_interpolate_at(Q, probe=True)
    orc = Oracle(bridge.flatten(hg), bridge.flatten_data(hg))
    rc, ov, ow, opr = orc.interpolate_at(Q)
    assert rc == 0
    assert_decisions_equal(pr, probe_dict(opr), "cuda vs oracle")
    assert np.array_equal(pr.tet, opr.tet)
    assert_values_close(vals, ov)
    assert_values_close(vecs, ow)
    rv, rw = hg.ir_interpolate_at(Q[:20000], True, 8)
    assert_values_close(vals[:20000], rv)
    assert_values_close(vecs[:20000], rw)


SPECIAL = np.array([[0, 0, 0], [0.5, 0, 0], [0.5, 0.5, 0], [0.5, 0.5, 0.5], [1, 0, 0], [0.25, 0.25, 0], [1 / 3, 1 / 3, 0], [0, 0, 0.5],
                    [-0.5, 0, 0], [2, 1, 0], [0.1, 0.1, 0.1], [-1.5, 2.5, 0.5], [0.75, 0.25, 0.5]], dtype=float)


@pytest.mark.parametrize("cls", ["BZTrellisQcc", "BZTrellisQdd", "BZTrellisQdc"])
@pytest.mark.parametrize("name", sorted(W.ZOO))
def test_lattice_zoo(host, bridge, name, cls):
    """centred / rhombohedral / triclinic lattices, added time reversal; complex values, pseudovector, reciprocal-vector
    and matrix data (rip_axial, rip_recip, rip_real with matrices) through the general kernel"""
    wl = W.zoo_grid(host, name, cls)
    g = brille_b200.accelerate(wl.grid)
    Q = np.vstack([SPECIAL, wl.make_q(20000, 5)])
    vals, vecs, pr = g.ir_interpolate_at(Q, probe=True)
    orc = Oracle(bridge.flatten(wl.grid), bridge.flatten_data(wl.grid))
    rc, ov, ow, opr = orc.interpolate_at(Q)
    assert rc == 0
    assert_decisions_equal(pr, probe_dict(opr), "cuda vs oracle")
    assert_values_close(vals, ov)
    assert_values_close(vecs, ow)
    rv, rw = wl.grid.ir_interpolate_at(Q[:2000], False, 1)
    assert_values_close(vals[:2000], rv)
    assert_values_close(vecs[:2000], rw)


def test_host_pipeline_chunking_is_invisible(host):
    """many small chunks through the two-stream host pipeline give the same bits as one chunk"""
    wl = W.c2_nacl(host, density=300)
    g = brille_b200.accelerate(wl.grid)
    Q = wl.make_q(300001, 4)
    v1, w1 = g.ir_interpolate_at(Q)
    g.set_option("host_chunk", 7001)
    v2, w2 = g.ir_interpolate_at(Q, pinned=True)
    assert np.array_equal(v1, v2) and np.array_equal(w1, w2)
    # a failing point anywhere fails the whole call, like the reference
    Qbad = Q.copy()
    Qbad[123456] = 7.3  # far outside the gridded irreducible zone when points are not moved
    with pytest.raises(RuntimeError):
        g.ir_interpolate_at(Qbad, do_not_move_points=True)


def test_c4_p21c_nest_72_modes(host, bridge):
    """BASELINE config 4: P2_1/c, 24 atoms, 72 modes, BZNestQdc (mode-tiled staging in the cell kernel)."""
    wl = W.c4_p21c_nest(host, density=300)
    g = brille_b200.accelerate(wl.grid)
    Q = wl.make_q(20000, 21)
    vals, vecs = g.ir_interpolate_at(Q)
    g.set_option("interp_path", 1)
    v1, w1 = g.ir_interpolate_at(Q[:4000])
    assert rel_close(v1, vals[:4000]) <= 1e-12 and rel_close(w1, vecs[:4000]) <= 1e-12
    rv, rw = wl.grid.ir_interpolate_at(Q[:4000], True, 8)
    assert_values_close(vals[:4000], rv)
    assert_values_close(vecs[:4000], rw)


@pytest.mark.parametrize("which", ["C2", "C3", "C3-sorted", "C4"])
def test_cell_kernels_agree_bitwise(host, which):
    """The pipelined cell kernel reads per-cell records built once per fill; the on-the-fly kernel derives the same
    numbers per work item with the same operations: identical bits, also across mode passes (C4) and permutations."""
    if which == "C4":
        wl = W.c4_p21c_nest(host, density=300)
        n = 30000
    elif which == "C2":
        wl = W.c2_nacl(host, density=300)
        n = 200000
    else:
        wl = W.c3_p63mmc(host, density=300, seed=5)
        n = 200000
        if which == "C3-sorted":
            wl.grid.sort()
    g = brille_b200.accelerate(wl.grid)
    Q = wl.make_q(n, 31)
    out = {}
    for ck in (1, 2):
        g.set_option("interp_path", 2)
        g.set_option("cell_kernel", ck)
        out[ck] = g.ir_interpolate_at(Q)
    assert np.array_equal(out[1][0], out[2][0])
    assert np.array_equal(out[1][1], out[2][1])
    # a second fill replaces the cell records
    if which == "C2":
        vals, vecs = out[2]
        g2 = brille_b200.accelerate(wl.grid)
        g2.set_option("cell_kernel", 2)
        a = g2.ir_interpolate_at(Q[:50000])
        g2._sync_data()
        b = g2.ir_interpolate_at(Q[:50000])
        assert np.array_equal(a[1], b[1]) and np.array_equal(a[1], vecs[:50000])
    g.close()


@pytest.mark.parametrize("path", PATHS, ids=PATH_IDS)
def test_sorted_permutations_against_reference(host, bridge, path):
    wl = W.c3_p63mmc(host, density=150, seed=9)
    wl.grid.sort()
    g = brille_b200.accelerate(wl.grid)
    apply_path(g, path)
    Q = wl.make_q(20000, 12)
    vals, vecs = g.ir_interpolate_at(Q)
    rv, rw = wl.grid.ir_interpolate_at(Q, True, 8)
    assert_values_close(vals, rv)
    assert_values_close(vecs, rw)


def test_full_size_properties(host):
    """C3 at 1e6 Q (too many for the scalar oracle): size-independent properties.
    (1) every point is placed, Q == R^T-equivalent of q_ir + tau is implied by (3);
    (2) the device-resident and the host-buffer entry points agree bit for bit, as do differently chunked calls;
    (3) Q and Q+G (G a reciprocal lattice vector) give the same eigenvalues;
    (4) eigenvalues are invariant under the point group: Q and R^T Q interpolate to the same values."""
    import torch

    wl = W.c3_p63mmc(host)
    g = brille_b200.accelerate(wl.grid)
    n = 1_000_000
    Q = wl.make_q(n, 3)
    vals, vecs, pr = g.ir_interpolate_at(Q, probe=True)
    assert (pr.status & 7).max() == 0
    dQ = torch.from_numpy(Q).cuda()
    dv, dw = g.ir_interpolate_at_device(dQ)
    assert np.array_equal(dv.cpu().numpy(), vals) and np.array_equal(dw.cpu().numpy(), vecs)
    v3, w3 = g.ir_interpolate_at(Q[: n // 3])
    assert np.array_equal(v3, vals[: n // 3]) and np.array_equal(w3, vecs[: n // 3])
    g.set_option("interp_path", 1)  # the general kernel and the cell-batched kernel agree to rounding
    v1, w1 = g.ir_interpolate_at(Q[:200000])
    assert rel_close(v1, vals[:200000]) <= 1e-12 and rel_close(w1, vecs[:200000]) <= 1e-12
    g.set_option("interp_path", 0)
    shift = np.random.default_rng(1).integers(-3, 4, (n, 3)).astype(float)
    v2, _ = g.ir_interpolate_at(Q + shift)
    assert rel_close(v2, vals) <= 1e-9
    ops = np.asarray(wl.bz.lattice.pointgroup.W)
    sub = Q[:50000]
    for j in (1, 5, 11, 17):
        vr, _ = g.ir_interpolate_at(sub @ ops[j].astype(float))  # rows are R^T q
        assert rel_close(vr, vals[:50000]) <= 1e-9


# ---------------------------------------------------------------------------------------------------------------------
# sort() on the device: interpolatordual.hpp:398-434, interpolator_cost.tpp, lapjv.hpp
# ---------------------------------------------------------------------------------------------------------------------
def _sort_workload(host, which):
    if which == "C2":
        return W.c2_nacl(host, density=300)
    if which == "C3":
        return W.c3_p63mmc(host, density=300, seed=5)
    if which == "C4":
        return W.c4_p21c_nest(host, density=40)
    lat = W.p63mmc_lattice(host)
    bz = host.BrillouinZone(lat)
    if which == "C3nest":
        g = host.BZNestQdc(bz, bz.ir_polyhedron.volume / 200, 5)
    else:
        g = host.BZMeshQdc(bz, bz.ir_polyhedron.volume / 200, 3)
    return W.Workload(which, g, bz, 12, 4, W._uniform_q(-3, 3), W._gamma_fill(g, 12, 4, 3))


@pytest.mark.parametrize("which", ["C2", "C3", "C3nest", "C3mesh", "C4"])
def test_device_sort_matches_reference_and_oracle(host, bridge, which):
    """sort() on the device against the oracle and the reference's own sort().

    The path has a floating-point part (cost matrices: atan2, cos, sin, acos of CUDA's libm instead of glibc's) and an
    integer part (the Jonker-Volgenant solver).  Bars: (1) cost matrices within 4 ulp of the oracle's; (2) the solver is
    bit-identical to the restated reference solver on the same (device) cost matrices; (3) the permutations equal the
    reference's own.  The reference's solver breaks ties with a cost-proportional epsilon (lapjv.hpp:312-314), which makes
    a pair in ~1000 depend on the last bit of a cost: for those the two assignments must be equally good (total cost within
    1e-3) -- the analogue of a Q within tolerance of a face taking either side."""
    from oracle import oracle as orc

    wl = _sort_workload(host, which)
    plan = bridge.sort_plan(wl.grid)
    data = bridge.flatten_data(wl.grid)
    g = brille_b200.accelerate(wl.grid)
    row, col, cost = g.sort_pairs(plan["pairs"], plan, want_cost=True)
    rc, orow, ocol, ocost = orc.sort_pairs(data, plan, want_cost=True)
    assert rc == 0
    assert np.abs(cost - ocost).max() <= 4 * 2.3e-16 * np.abs(ocost).max(), "(1) cost matrices"
    lrow, lcol = orc.lapjv_batch(cost)
    assert np.array_equal(row, lrow) and np.array_equal(col, lcol), "(2) solver on the device's cost matrices"
    B = row.shape[1]
    assert np.array_equal(np.take_along_axis(col, row, axis=1), np.broadcast_to(np.arange(B), row.shape))
    # (3) the reference's own sort()
    g.sort()  # device: installs the permutations in the device tables only
    Q = wl.make_q(20000, 77)
    vals, vecs = g.ir_interpolate_at(Q)
    wl.grid.sort()
    ref = bridge.pair_permutations(wl.grid, plan["pairs"])
    assert np.array_equal(orow.astype(np.uint32), ref[:, 0, :]) and np.array_equal(ocol.astype(np.uint32), ref[:, 1, :]), "oracle vs reference"
    differ = np.where((row.astype(np.uint32) != ref[:, 0, :]).any(axis=1))[0]
    assert len(differ) <= max(1, len(row) // 200), f"{len(differ)} of {len(row)} pairs differ"
    ar = np.arange(B)
    for k in differ:
        mine, theirs = ocost[k][ar, row[k]].sum(), ocost[k][ar, orow[k]].sum()
        assert abs(mine - theirs) <= 1e-3 * abs(theirs), (k, mine, theirs)
    # interpolation through the device-sorted tables: against the oracle with the same tables, and against the reference
    # (after its own sort) wherever no differing pair is involved
    from brille_b200.grid import install_permutations

    o = Oracle(bridge.flatten(wl.grid), install_permutations(data, plan["pairs"], row, col, int(plan["n_vertices"])))
    rc, ov, ow, _ = o.interpolate_at(Q)
    assert rc == 0
    assert_values_close(vals, ov)
    assert_values_close(vecs, ow)
    if len(differ) == 0:
        rv, rw = wl.grid.ir_interpolate_at(Q, True, 8)
        assert_values_close(vals, rv)
        assert_values_close(vecs, rw)
        g2 = brille_b200.accelerate(wl.grid)  # tables flattened from the host-sorted object: same bits
        v2, w2 = g2.ir_interpolate_at(Q)
        assert np.array_equal(v2, vals) and np.array_equal(w2, vecs)


def test_device_sort_cost_functions_and_errors(host, bridge):
    wl = W.c2_nacl(host, density=100)
    wl.grid.set_flags_weights(np.array([0, 0, 0, 1], dtype=np.int32), np.array([1.0, 1.0, 1.0]),
                              np.array([2, 3, 0, 4], dtype=np.int32), np.array([0.5, 2.0, 1.0]))
    plan = bridge.sort_plan(wl.grid)
    g = brille_b200.accelerate(wl.grid)
    row, col = g.sort_pairs(plan["pairs"], plan)
    wl.grid.sort()
    ref = bridge.pair_permutations(wl.grid, plan["pairs"])
    same = (row.astype(np.uint32) == ref[:, 0, :]).all(axis=1) & (col.astype(np.uint32) == ref[:, 1, :]).all(axis=1)
    assert same.mean() >= 0.995
    with pytest.raises(RuntimeError, match="out of range"):
        g.sort_pairs(np.array([[0, 10**6]], dtype=np.uint32), plan)
    # real eigenvectors: not offloaded
    s, d, _, rest = load_golden("p1_trellis_dd.npz")
    gd = brille_b200.B200Grid(None, structure=s, data=d)
    with pytest.raises(RuntimeError, match="real-valued"):
        gd.sort_pairs(np.array([[0, 1]], dtype=np.uint32), plan)


@pytest.mark.parametrize("which", ["C2", "C3", "C3nest", "C3mesh", "C4"])
def test_two_kernel_location_is_bit_invisible(host, which):
    """The location in two kernels (points regrouped in between: by trellis node, or by a spatial bin for the Nest / Mesh descents)
    and in one kernel: same decisions, same weights, same results, bit for bit -- with and without the probe (the probe switches
    the lean outputs off)."""
    if which in ("C3nest", "C3mesh"):
        lat = W.p63mmc_lattice(host)
        bz = host.BrillouinZone(lat)
        hg = host.BZNestQdc(bz, bz.ir_polyhedron.volume / 500, 5) if which == "C3nest" else host.BZMeshQdc(bz, bz.ir_polyhedron.volume / 500, 3)
        W._gamma_fill(hg, 12, 4, 17)
        wl = W.Workload(which, hg, bz, 12, 4, W._uniform_q(-3, 3))
    elif which == "C4":
        wl = W.c4_p21c_nest(host, density=300)
    else:
        wl = W.c2_nacl(host, density=300) if which == "C2" else W.c3_p63mmc(host, density=300, seed=5)
    g = brille_b200.accelerate(wl.grid)
    Q = wl.make_q(60000 if which == "C4" else 300000, 41)
    Q[:10] = 0.0  # the zone centre sits on node corners: neighbour search, compact emission, general kernel
    out = {}
    for split in (0, 1):
        g.set_option("split_locate", split)
        out[split] = g.ir_interpolate_at(Q, probe=True) + g.ir_interpolate_at(Q)
    a, b = out[0], out[1]
    for k in (0, 1, 3, 4):
        assert np.array_equal(a[k], b[k])
    for name in ("tau", "q_ir", "ridx", "invridx", "cell", "tet", "n_vert", "vertex", "weight", "status"):
        assert np.array_equal(getattr(a[2], name), getattr(b[2], name)), name
    g.close()


def test_degenerate_point_sets(host, bridge):
    """Point sets that stress the bucketing: all points identical (one node, one (cell, operation) bucket holds everything),
    points on a line through the zone centre (node faces / edges: neighbour search, compact emission, general kernel), and a
    call whose size sits just below / above the thresholds that switch the two-kernel location and the cell kernels on."""
    wl = W.c3_p63mmc(host, density=300, seed=5)
    g = brille_b200.accelerate(wl.grid)
    orc = Oracle(bridge.flatten(wl.grid), bridge.flatten_data(wl.grid))
    rng = np.random.default_rng(5)
    one = np.tile(np.array([[0.137, 0.211, 0.303]]), (200_000, 1))
    line = np.outer(np.linspace(-2.0, 2.0, 200_001), np.array([1.0, 0.0, 0.0]))
    grid_pts = rng.integers(-8, 9, (100_000, 3)) / 8.0  # rational points: many sit exactly on cell faces
    for name, Q in (("identical", one), ("line", line), ("rational", grid_pts)):
        vals, vecs, pr = g.ir_interpolate_at(Q, probe=True)
        v2, w2 = g.ir_interpolate_at(Q)  # without the probe: lean outputs of the location
        assert np.array_equal(vals, v2) and np.array_equal(vecs, w2), name
        sel = np.unique(np.concatenate([np.arange(0, len(Q), max(1, len(Q) // 3000)), [len(Q) - 1]]))
        rc, ov, ow, opr = orc.interpolate_at(Q[sel])
        assert rc == 0, name
        assert np.array_equal(pr.tau[sel], opr.tau) and np.array_equal(pr.ridx[sel], opr.ridx), name
        assert np.array_equal(pr.vertex[sel], opr.vertex) and np.array_equal(pr.weight[sel], opr.weight), name
        assert_values_close(vals[sel], ov)
        assert_values_close(vecs[sel], ow)
        if name == "identical":
            assert (vals == vals[0]).all() and (vecs == vecs[0]).all()
    # sizes around the switches (number of nodes / cells times a small factor)
    base = wl.make_q(60_000, 9)
    ref_v, ref_w = g.ir_interpolate_at(base)
    for n in (1, 31, 257, 5_000, 20_001, 59_999):
        v, w = g.ir_interpolate_at(base[:n])
        # (the general and the cell-batched kernel agree to rounding, not bitwise)
        assert rel_close(v, ref_v[:n]) <= 1e-12 and rel_close(w, ref_w[:n]) <= 1e-12, n
    g.close()


def test_c5_powder_q_sharded(host, bridge):
    """BASELINE config 5 at reduced size: powder-average Q (|Q| up to 10 1/angstrom: large tau) on the C3 trellis, Q sharded in
    contiguous row blocks (here: two shards on the one device through ShardedGrid) -- decisions bit-identical to the oracle,
    the sharded call bit-identical to the single call, the reference itself on a slice."""
    from brille_b200.sharding import ShardedGrid

    wl = W.c3_p63mmc(host, density=500)
    B = np.asarray(bridge.flatten_bz(wl.bz)["to_xyz"])
    Q = W.powder_q(B, 200001, 77)
    g = brille_b200.accelerate(wl.grid)
    vals, vecs, pr = g.ir_interpolate_at(Q, probe=True)
    assert np.abs(pr.tau).max() >= 3
    orc = Oracle(bridge.flatten(wl.grid), bridge.flatten_data(wl.grid))
    rc, ov, ow, opr = orc.interpolate_at(Q)
    assert rc == 0
    assert_decisions_equal(pr, probe_dict(opr), "cuda vs oracle")
    assert_values_close(vals, ov)
    assert_values_close(vecs, ow)
    rv, rw = wl.grid.ir_interpolate_at(Q[:20000], True, 8)
    assert_values_close(vals[:20000], rv)
    assert_values_close(vecs[:20000], rw)
    sg = ShardedGrid(wl.grid, [0, 0])
    sv, sw = sg.ir_interpolate_at(Q)
    assert np.array_equal(sv, vals) and np.array_equal(sw, vecs)
    sg.close()
    g.close()


@pytest.mark.parametrize("which", ["C3", "C2", "F23", "P3_timereversal"])
def test_bz_methods_against_reference(host, bridge, which):
    """BrillouinZone.isinside / moveinto / ir_moveinto / ir_moveinto_wedge (wrap/_bz.cpp:378-520) routed to the locate kernel
    (SURVEY 8f rank 3), against the reference's own methods: booleans, tau and q bit for bit, the same rotation matrices."""
    wl = W.BUILDERS[which](host) if which in W.BUILDERS else W.zoo_grid(host, which)
    g = brille_b200.accelerate(wl.grid)
    bz = wl.bz
    rng = np.random.default_rng(8)
    Q = np.vstack([rng.uniform(-1.5, 1.5, (50000, 3)), rng.integers(-4, 5, (2000, 3)) / 4.0, np.zeros((1, 3))])
    # isinside
    inside = g.isinside(Q)
    assert inside.dtype == bool and np.array_equal(inside, np.asarray(bz.isinside(Q), dtype=bool))
    assert 0 < inside.sum() < len(Q)
    # moveinto
    q, tau = g.moveinto(Q)
    rq, rtau = bz.moveinto(Q)
    assert np.array_equal(tau, rtau) and np.array_equal(q, rq)
    assert g.isinside(q).all()
    # ir_moveinto: q, tau, and the rotation matrices of the returned indices
    rots = np.asarray(bridge.flatten_bz(bz)["rotations"]).reshape(-1, 3, 3)
    q, tau, ridx, invridx = g.ir_moveinto(Q)
    rq, rtau, rR, rinvR = bz.ir_moveinto(Q)
    assert np.array_equal(tau, rtau) and np.array_equal(q, rq)
    assert np.array_equal(rots[ridx], rR) and np.array_equal(rots[invridx], rinvR)
    # ir_moveinto_wedge: no translation
    qw, rw = g.ir_moveinto_wedge(Q)
    rqw, rRw = bz.ir_moveinto_wedge(Q)
    assert np.array_equal(qw, rqw) and np.array_equal(rots[rw], rRw)
    # ... and the oracle's restatement of the same methods (indices instead of matrices)
    orc = Oracle(bridge.flatten(wl.grid), bridge.flatten_data(wl.grid))
    rc, opr = orc.moveinto(Q, 3)
    assert np.array_equal(inside, (opr.status & T.ST_OUTSIDE_BZ) == 0)
    rc, opr = orc.moveinto(Q, 2)
    assert np.array_equal(qw, opr.q_ir) and np.array_equal(rw, opr.ridx)
    rc, opr = orc.moveinto(Q, 1)
    assert np.array_equal(q, opr.q_ir) and np.array_equal(tau, opr.tau) and np.array_equal(ridx, opr.ridx) and np.array_equal(invridx, opr.invridx)
    g.close()


def test_sharded_grid_on_two_devices(host):
    """One process driving two GPUs (ShardedGrid: a handle and a host thread per device, Q cut into contiguous row blocks, no
    collective): bit-identical to the single-device call.  Needs two visible devices."""
    import torch

    from brille_b200.sharding import ShardedGrid

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    wl = W.c3_p63mmc(host, density=500)
    Q = wl.make_q(400_001, 5)
    g = brille_b200.accelerate(wl.grid, device=0)
    vals, vecs = g.ir_interpolate_at(Q)
    g.close()
    sg = ShardedGrid(wl.grid, [0, 1])
    sv, sw = sg.ir_interpolate_at(Q)
    assert np.array_equal(sv, vals) and np.array_equal(sw, vecs)
    sg.close()
    g1 = brille_b200.accelerate(wl.grid, device=1)  # the second device on its own: structure factor, fused and not
    rng = np.random.default_rng(1)
    g1.set_structure_factor(rng.normal(size=4) + 1j * rng.normal(size=4), positions=rng.uniform(0, 1, (4, 3)))
    v1, sf1 = g1.ir_structure_factor(Q)
    g1.set_option("sf_fused", 0)
    v0, sf0 = g1.ir_structure_factor(Q)
    assert np.array_equal(v1, vals) and rel_close(sf1, sf0) <= 1e-12
    g1.close()
