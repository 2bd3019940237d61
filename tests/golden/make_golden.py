#!/usr/bin/env python3
"""Generate the committed golden fixtures (run in the build container, where oracle/_ref exists).

    python tests/golden/make_golden.py

Each fixture holds the flat tables of a small grid (brille_b200.tables.save_tables layout: 's.*' structure,
'd.*' data), a set of Q points and what the REFERENCE ITSELF (oracle/_ref/_brille, built from
/root/reference by oracle/build_ref.sh) returned for them, including the intermediate decisions exposed by
oracle/probe.cpp.  nacl_gamma.npz additionally carries the reference's own golden vectors
(wrap/tests/test_5_gamma.npz: euphonic_values/vectors and brille_values/vectors).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402
from brille_b200 import _bridge as br  # noqa: E402
from brille_b200 import tables as T  # noqa: E402
from brille_b200 import workloads as W  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
b = ref.host()
pr = ref.probe()


def flat(prefix, d, out):
    for k, v in d.items():
        if isinstance(v, dict):
            flat(f"{prefix}{k}.", v, out)
        elif isinstance(v, str):
            out[f"{prefix}{k}"] = np.array(v)
        else:
            out[f"{prefix}{k}"] = np.asarray(v)


def reference_run(grid, bz, Q, ir=True):
    out = {}
    if ir:
        v, w = grid.ir_interpolate_at(Q, False, 1)
        q, x, tau, r, invr = pr.ir_moveinto_idx(bz, Q, 1)
        out.update(ref_q_ir=q, ref_x_ir=x, ref_tau=tau, ref_ridx=r, ref_invridx=invr)
    else:
        v, w = grid.interpolate_at(Q, False, 1)
        q, tau = pr.moveinto(bz, Q, 1)
        x = q @ np.asarray(br.flatten_bz(bz)["to_xyz"]).reshape(3, 3).T
        out.update(ref_q_ir=q, ref_x_ir=x, ref_tau=tau)
    cnt, idx, wgt = pr.indices_weights(grid, out["ref_x_ir"])
    out.update(ref_values=v, ref_vectors=w, ref_n_vert=cnt, ref_vertex=idx, ref_weight=wgt)
    return out


def nacl_gamma():
    src = os.path.join(os.environ.get("BRILLE_REFERENCE", "/root/reference"), "wrap", "tests", "test_5_gamma.npz")
    nacl = np.load(src)
    bas = b.Basis(nacl["atom_positions"], [int(i) for i in nacl["atom_index"]])
    sym = b.Symmetry(nacl["spacegroup_mat"], nacl["spacegroup_vec"])
    lat = b.Lattice(nacl["basis_vectors"], sym, bas)
    bz = b.BrillouinZone(lat)
    grid = b.BZTrellisQdc(bz, float(nacl["grid_max_volume"]), bool(nacl["grid_always_triangulate"]))
    perm = np.hstack([np.argwhere(np.all(np.isclose(nacl["grid_rlu"], x), axis=1)) for x in grid.rlu]).flatten()
    out = {}
    # (a) real_lattice Gamma, exactly as wrap/tests/test_5_gamma.py:95-128
    vec_els = np.array([0, 24, 0, 2, 3, 0, 0], dtype=np.int32)
    grid.fill(nacl["grid_values"][perm], nacl["grid_values_elements"].astype(np.int32), nacl["grid_values_weights"],
              nacl["grid_vectors"][perm], vec_els, nacl["grid_vectors_weights"], bool(nacl["grid_sort"]))
    flat("s.", br.flatten(grid), out)
    flat("d.", br.flatten_data(grid), out)
    Q = np.ascontiguousarray(nacl["q_nu"])
    out["Q"] = Q
    out.update(reference_run(grid, bz, Q))
    # (b) Cartesian eigenvectors, LengthUnit::angstrom (test_5_gamma.py:180-221)
    cart = np.einsum("ba,ijkb->ijka", nacl["basis_vectors"], nacl["grid_vectors"])
    vec_els2 = np.array([0, 24, 0, 2, 1, 0, 0], dtype=np.int32)
    grid.fill(nacl["grid_values"][perm], nacl["grid_values_elements"].astype(np.int32), nacl["grid_values_weights"],
              cart[perm], vec_els2, nacl["grid_vectors_weights"])
    flat("d2.", br.flatten_data(grid), out)
    r2 = reference_run(grid, bz, Q)
    out["ref2_values"], out["ref2_vectors"] = r2["ref_values"], r2["ref_vectors"]
    # the reference's own golden vectors
    for k in ("euphonic_values", "euphonic_vectors", "brille_values", "brille_vectors", "basis_vectors"):
        out[f"golden_{k}"] = nacl[k]
    np.savez_compressed(os.path.join(HERE, "nacl_gamma.npz"), **out)
    print("nacl_gamma.npz", os.path.getsize(os.path.join(HERE, "nacl_gamma.npz")))


def small_trellis(name, wl, nq, seed, sort=False, extra_q=None):
    g, bz = wl.grid, wl.bz
    if sort:
        g.sort()
    out = {}
    flat("s.", br.flatten(g), out)
    flat("d.", br.flatten_data(g), out)
    Q = wl.make_q(nq, seed)
    if extra_q is not None:
        Q = np.vstack([extra_q, Q])
    out["Q"] = Q
    out.update(reference_run(g, bz, Q))
    np.savez_compressed(os.path.join(HERE, name), **out)
    print(name, os.path.getsize(os.path.join(HERE, name)))


def special_points():
    """High-symmetry and on-face points: the ties where first-match order decides."""
    pts = [[0, 0, 0], [0.5, 0, 0], [0.5, 0.5, 0], [0.5, 0.5, 0.5], [1, 0, 0], [0.25, 0.25, 0], [1 / 3, 1 / 3, 0], [1 / 3, 1 / 3, 0.5],
           [0, 0, 0.5], [0.5, 0, 0.5], [-0.5, 0, 0], [0, -0.5, 0.5], [2, 1, 0], [0.1, 0.1, 0.1], [-1.5, 2.5, 0.5], [0.75, 0.25, 0.5]]
    return np.array(pts, dtype=float)


def p1_trellis():
    """P1 lattice: no point symmetry, the irreducible zone IS the first zone, so `interpolate_at` (moveinto without
    wedge rotation, bz_trellis.hpp:105-121) is meaningful; vector-like values + matrix-like vectors exercise rip_real."""
    lat = b.Lattice((3.1, 4.2, 5.3), (90.0, 97.0, 90.0), "P 1")  # fully triclinic P 1 cells crash the reference trellis constructor
    bz = b.BrillouinZone(lat)
    g = b.BZTrellisQdd(bz, bz.ir_polyhedron.volume / 120)
    nv = g.rlu.shape[0]
    rng = np.random.default_rng(31)
    vals = rng.normal(size=(nv, 3, 1 + 3))       # per mode: 1 scalar + one 3-vector
    vecs = rng.normal(size=(nv, 3, 2 + 9))       # per mode: 2 scalars + one 3x3 matrix
    g.fill(vals, (1, 3, 0, 0, 3), vecs, (2, 0, 9, 0, 3))
    out = {}
    flat("s.", br.flatten(g), out)
    flat("d.", br.flatten_data(g), out)
    Q = np.vstack([special_points(), rng.uniform(-2, 2, (300, 3))])
    out["Q"] = Q
    out.update(reference_run(g, bz, Q))
    r0 = reference_run(g, bz, Q, ir=False)
    for k, v in r0.items():
        out[k.replace("ref_", "ref0_")] = v
    np.savez_compressed(os.path.join(HERE, "p1_trellis_dd.npz"), **out)
    print("p1_trellis_dd.npz", os.path.getsize(os.path.join(HERE, "p1_trellis_dd.npz")))


def tet_grid(name, cls, args, nq, seed, sort=False):
    """Nest / Mesh grids on the P6_3/mmc lattice (BZNestQdc: nest.hpp, BZMeshQdc: mesh.hpp)."""
    lat = W.p63mmc_lattice(b)
    bz = b.BrillouinZone(lat)
    g = getattr(b, cls)(bz, *[a(bz) if callable(a) else a for a in args])
    W._gamma_fill(g, 12, 4, seed)
    if sort:
        g.sort()
    out = {}
    flat("s.", br.flatten(g), out)
    flat("d.", br.flatten_data(g), out)
    Q = np.vstack([special_points(), np.random.default_rng(seed).uniform(-3, 3, (nq, 3))])
    out["Q"] = Q
    out.update(reference_run(g, bz, Q))
    np.savez_compressed(os.path.join(HERE, name), **out)
    print(name, os.path.getsize(os.path.join(HERE, name)))


def main():
    tet_grid("p63mmc_nest.npz", "BZNestQdc", (lambda bz: bz.ir_polyhedron.volume / 40, 5), 300, 41)
    tet_grid("p63mmc_nest_sorted.npz", "BZNestQdc", (lambda bz: bz.ir_polyhedron.volume / 30, 5), 200, 42, sort=True)
    tet_grid("p63mmc_mesh.npz", "BZMeshQdc", (lambda bz: bz.ir_polyhedron.volume / 60, 3), 300, 43)
    if "--tet-only" in sys.argv:
        return
    nacl_gamma()
    p1_trellis()
    small_trellis("nacl_prim_trellis.npz", W.c2_nacl(b, density=150, seed=5), 400, 21, extra_q=special_points())
    small_trellis("nacl_prim_trellis_sorted.npz", W.c2_nacl(b, density=150, seed=6), 300, 22, sort=True)
    small_trellis("fd3m_scalar_trellis.npz", W.c1_fd3m_scalar(b, density=300), 500, 23, extra_q=special_points())
    w3 = W.c3_p63mmc(b, density=60, seed=7)
    small_trellis("p63mmc_trellis.npz", w3, 300, 24, extra_q=special_points())


if __name__ == "__main__":
    main()
