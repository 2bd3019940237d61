"""CPU: the plain-C restatement of DualInterpolator::sort() (oracle/sort_oracle.c) against the reference's own sort().

The reference is run here through oracle/_ref (the unmodified sources); the bridge exposes the vertex pairs sort() walks
and the permutations it stored.  Permutations are integers: the bar is equality for every pair."""
import numpy as np
import pytest

from brille_b200 import workloads as W
from oracle import oracle as orc


def reference_sort(host, bridge, wl):
    plan = bridge.sort_plan(wl.grid)
    data = bridge.flatten_data(wl.grid)  # the data sort() sees (unsorted table)
    wl.grid.sort()
    ref = bridge.pair_permutations(wl.grid, plan["pairs"])
    return plan, data, ref


@pytest.mark.parametrize("which", ["C2", "C3", "C3nest", "C4"])
def test_sort_oracle_reproduces_reference_permutations(host, bridge, which):
    if which == "C2":
        wl = W.c2_nacl(host, density=300)
    elif which == "C3":
        wl = W.c3_p63mmc(host, density=300, seed=5)
    elif which == "C4":
        wl = W.c4_p21c_nest(host, density=40)
    else:
        lat = W.p63mmc_lattice(host)
        bz = host.BrillouinZone(lat)
        g = host.BZNestQdc(bz, bz.ir_polyhedron.volume / 200, 5)
        wl = W.Workload(which, g, bz, 12, 4, W._uniform_q(-3, 3), W._gamma_fill(g, 12, 4, 3))
    plan, data, ref = reference_sort(host, bridge, wl)
    assert plan["pairs"].shape[0] > 0
    rc, row, col = orc.sort_pairs(data, plan)
    assert rc == 0
    assert np.array_equal(row.astype(np.uint32), ref[:, 0, :]), "permutation stored for (i, j)"
    assert np.array_equal(col.astype(np.uint32), ref[:, 1, :]), "permutation stored for (j, i)"
    # every row is a permutation, and col is the inverse of row
    B = row.shape[1]
    assert np.array_equal(np.sort(row, axis=1), np.broadcast_to(np.arange(B), row.shape))
    assert np.array_equal(np.take_along_axis(col, row, axis=1), np.broadcast_to(np.arange(B), row.shape))


def test_sort_oracle_vector_cost_functions(host, bridge):
    """the other vector cost functions of set_cost_info (interpolator.hpp:246-299) and non-unit weights"""
    for vcf in (1, 2, 3, 4):
        wl = W.c2_nacl(host, density=100)
        # flags: RotatesLike, LengthUnit, scalar cost function, vector cost function (wrap/_interpolator.cpp:42-85)
        wl.grid.set_flags_weights(np.array([0, 0, 0, vcf], dtype=np.int32), np.array([1.0, 1.0, 1.0]),
                                  np.array([2, 3, 0, vcf], dtype=np.int32), np.array([0.5, 2.0, 1.0]))
        plan, data, ref = reference_sort(host, bridge, wl)
        assert int(plan["vectors_vector_cost"]) == vcf and list(plan["vectors_costmult"]) == [0.5, 2.0, 1.0]
        rc, row, col = orc.sort_pairs(data, plan)
        assert rc == 0
        assert np.array_equal(row.astype(np.uint32), ref[:, 0, :]) and np.array_equal(col.astype(np.uint32), ref[:, 1, :]), vcf


def tie_heavy_matrices(seed=0):
    """cost matrices full of ties (small integers, rounded decimals, constant rows / columns) next to generic ones"""
    rng = np.random.default_rng(seed)
    out = []
    for B, n, kind in [(1, 3, "int"), (2, 100, "int"), (3, 200, "int"), (6, 200, "int"), (12, 200, "int"), (12, 200, "float"), (33, 40, "int"),
                       (72, 12, "float"), (72, 12, "int"), (40, 40, "bin"), (12, 200, "dec"), (65, 10, "bin"), (12, 50, "neg")]:
        if kind == "int":
            c = rng.integers(0, 4, (n, B, B)).astype(float)
        elif kind == "bin":
            c = rng.integers(0, 2, (n, B, B)).astype(float)
        elif kind == "dec":
            c = np.round(rng.uniform(0, 2, (n, B, B)), 1)
        elif kind == "neg":
            c = rng.integers(-3, 3, (n, B, B)).astype(float)
        else:
            c = rng.uniform(0, 1, (n, B, B))
        out.append(c)
    return out


def test_lane_parallel_formulation_equals_sequential_solver():
    """The device solver (brille_b200/csrc/sortpairs.cu: match_pair) replaces the sequential scans of lapjv.hpp by lane-parallel
    formulations; oracle/lap_model.py states them in numpy.  They must take the decisions of the sequential restatement
    (oracle/sort_oracle.c, pinned on the reference above) on every matrix, ties included."""
    from oracle import lap_model

    for c in tie_heavy_matrices():
        r, cc = orc.lapjv_batch(c)
        r2, c2 = lap_model.solve_batch(c)
        assert np.array_equal(r, r2) and np.array_equal(cc, c2), c.shape


@pytest.mark.gpu
def test_device_solver_on_tie_heavy_matrices():
    from brille_b200 import capi

    for c in tie_heavy_matrices():
        r, cc = orc.lapjv_batch(c)
        r2, c2 = capi.solve_assignments(c)
        assert np.array_equal(r, r2) and np.array_equal(cc, c2), c.shape
    big = np.random.default_rng(3).integers(0, 5, (3, 200, 200)).astype(float)  # more modes than a grid ever has
    r, cc = orc.lapjv_batch(big)
    r2, c2 = capi.solve_assignments(big)
    assert np.array_equal(r, r2) and np.array_equal(cc, c2)
