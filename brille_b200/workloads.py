"""Synthetic workloads of BASELINE.json / SURVEY.md section 8(d), built with brille's host module.

Every builder takes the host module ``b`` (brille's own ``_brille``: construction stays brille's C++)
and returns a :class:`Workload` with a constructed + filled host grid and a Q generator.  Eigen-data are
random: the arithmetic of the path is data independent.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable

import numpy as np


@dataclass
class Workload:
    name: str
    grid: object  # host BZTrellisQ*/BZNestQ*/BZMeshQ* object, already filled
    bz: object
    modes: int
    n_atoms: int
    make_q: Callable[[int, int], np.ndarray]  # (n, seed) -> (n,3) rlu
    fill_args: tuple = field(default_factory=tuple)
    note: str = ""

    @property
    def bytes_per_q(self) -> int:
        """ALGORITHMIC HBM bytes per Q: Q in + values out + vectors out."""
        g = self.fill_args
        vals, vecs = np.asarray(g[0]), np.asarray(g[2])
        return 24 + vals[0].nbytes + vecs[0].nbytes


def _uniform_q(lo, hi):
    def make(n, seed):
        return np.random.default_rng(seed).uniform(lo, hi, (int(n), 3))

    return make


def _gamma_fill(grid, modes, n_atoms, seed):
    nv = grid.rlu.shape[0]
    rng = np.random.default_rng(seed)
    vals = rng.uniform(1.0, 50.0, (nv, modes, 1))
    vecs = rng.normal(size=(nv, modes, n_atoms, 3)) + 1j * rng.normal(size=(nv, modes, n_atoms, 3))
    # elements: (scalars, vector elements, matrix elements, RotatesLike, LengthUnit)
    args = (vals, (1, 0, 0, 0, 3), vecs, (0, 3 * n_atoms, 0, 2, 3))
    grid.fill(*args)
    return args


def c1_fd3m_scalar(b, density=2000):
    """C1: Fd-3m cubic a=4.96, BZTrellisQdc, one scalar eigenvalue (+ complex scalar 'vector')."""
    lat = b.Lattice((4.96, 4.96, 4.96), (90, 90, 90), "Fd-3m")
    bz = b.BrillouinZone(lat)
    g = b.BZTrellisQdc(bz, bz.ir_polyhedron.volume / density)
    vals = np.cos(np.pi * g.rlu).sum(axis=1)
    args = (vals, (1,), vals.astype(complex), (1,))
    g.fill(*args)
    return Workload("C1 Fd-3m scalar trellis", g, bz, 1, 0, _uniform_q(-5, 5), args)


def nacl_primitive_lattice(b, a=5.64):
    conv = b.Lattice((a, a, a), (90, 90, 90), "Fm-3m")
    W = np.array(conv.spacegroup.W)
    w = np.array(conv.spacegroup.w)
    P = np.array([[0, 0.5, 0.5], [0.5, 0, 0.5], [0.5, 0.5, 0]]).T  # columns = primitive vectors (conv. basis)
    Pi = np.linalg.inv(P)
    ops = {}
    for Wi, wi in zip(W, w):
        Wp = np.rint(Pi @ Wi @ P).astype(int)
        wp = np.mod(Pi @ wi + 1e-9, 1.0) - 1e-9
        wp[np.abs(wp) < 1e-8] = 0.0
        key = (tuple(Wp.ravel()), tuple(np.round(wp, 6)))
        ops.setdefault(key, (Wp, wp))
    Wp = np.array([v[0] for v in ops.values()], dtype=np.int32)
    wp = np.array([v[1] for v in ops.values()], dtype=np.float64)
    vectors = a * np.array([[0, 0.5, 0.5], [0.5, 0, 0.5], [0.5, 0.5, 0]])
    return b.Lattice(vectors, b.Symmetry(Wp, wp), b.Basis([[0, 0, 0], [0.5, 0.5, 0.5]], [0, 1]))


def c2_nacl(b, density=1000, seed=12):
    """C2: NaCl-like primitive 2-atom cell, 6 modes, Gamma eigenvectors."""
    lat = nacl_primitive_lattice(b)
    bz = b.BrillouinZone(lat)
    g = b.BZTrellisQdc(bz, bz.ir_polyhedron.volume / density)
    args = _gamma_fill(g, 6, 2, seed)
    return Workload("C2 NaCl primitive 2 atoms / 6 modes trellis", g, bz, 6, 2, _uniform_q(-3, 3), args)


def p63mmc_lattice(b):
    return b.Lattice(
        (3.6, 3.6, 5.0), (90, 90, 120), "P6_3/mmc",
        b.Basis([[0, 0, 0], [0, 0, 0.5], [1 / 3, 2 / 3, 0.25], [2 / 3, 1 / 3, 0.75]], [0, 0, 1, 1]),
    )


def c3_p63mmc(b, density=2000, seed=13, cls="BZTrellisQdc"):
    """C3/C5: P6_3/mmc hexagonal 4-atom cell, 12 modes, hybrid cube/tetrahedron trellis."""
    lat = p63mmc_lattice(b)
    bz = b.BrillouinZone(lat)
    g = getattr(b, cls)(bz, bz.ir_polyhedron.volume / density)
    args = _gamma_fill(g, 12, 4, seed)
    return Workload(f"C3 P6_3/mmc 4 atoms / 12 modes {cls} V_ir/{density}", g, bz, 12, 4, _uniform_q(-3, 3), args)


def p21c_lattice(b):
    x = np.array([[0.11, 0.13, 0.17], [0.31, 0.07, 0.41], [0.23, 0.37, 0.09], [0.43, 0.29, 0.33], [0.07, 0.43, 0.27], [0.37, 0.19, 0.47]])
    pos, typ = [], []
    for t, p in enumerate(x):
        orbit = [p, np.array([-p[0], p[1] + 0.5, -p[2] + 0.5]), -p, np.array([p[0], -p[1] + 0.5, p[2] + 0.5])]
        for o in orbit:
            pos.append(np.mod(o, 1.0))
            typ.append(t)
    return b.Lattice((7.1, 9.3, 11.2), (90, 104, 90), "-P 2ybc", b.Basis(np.array(pos), typ))


def c4_p21c_nest(b, density=2000, seed=14):
    """C4: P2_1/c monoclinic 24-atom cell, 72 modes, BZNestQdc tetrahedral nest."""
    lat = p21c_lattice(b)
    bz = b.BrillouinZone(lat)
    g = b.BZNestQdc(bz, bz.ir_polyhedron.volume / density, 5)
    args = _gamma_fill(g, 72, 24, seed)
    return Workload("C4 P2_1/c 24 atoms / 72 modes nest", g, bz, 72, 24, _uniform_q(-3, 3), args)


def nacl_conventional_lattice(b, a=5.64):
    """NaCl conventional cell (Fm-3m, 8 atoms): the point group leaves the Na at the origin in place."""
    pos = [[0, 0, 0], [0, .5, .5], [.5, 0, .5], [.5, .5, 0], [.5, .5, .5], [.5, 0, 0], [0, .5, 0], [0, 0, .5]]
    return b.Lattice((a, a, a), (90, 90, 90), "Fm-3m", b.Basis(pos, [0, 0, 0, 0, 1, 1, 1, 1]))


def gamma_matrix_grid(b, which="prim", density=60, seed=21, modes=3, cartesian=False):
    """Gamma-rotated 3-vectors AND 3x3 matrices in one mode (interpolator_gamma.tpp:100-134): complex values with `no1` vectors
    and 9 matrices (Nmat = 1: the reference rotates matrix 0 and copies the rest of the mode back from its work array).
    ``prim``: NaCl primitive cell, 2 vectors; ``conv``: conventional cell, 8 vectors -- 24 vector elements reach into the
    matrix part of the work array."""
    lat = nacl_primitive_lattice(b) if which == "prim" else nacl_conventional_lattice(b)
    n_at = 2 if which == "prim" else 8
    bz = b.BrillouinZone(lat)
    g = b.BZTrellisQcc(bz, bz.ir_polyhedron.volume / density)
    nv = g.rlu.shape[0]
    rng = np.random.default_rng(seed)

    def rnd(shape):
        return rng.normal(size=shape) + 1j * rng.normal(size=shape)

    lu = 1 if cartesian else 3
    vals = rnd((nv, modes, 1 + 3 * n_at))               # scalar + Gamma vectors
    vecs = rnd((nv, modes, 2 + 3 * n_at + 81))          # 2 scalars + Gamma vectors + 9 matrices
    args = (vals, (1, 3 * n_at, 0, 2, lu), vecs, (2, 3 * n_at, 81, 2, lu))
    g.fill(*args)
    return Workload(f"Gamma vectors+matrices NaCl {which}", g, bz, modes, n_at, _uniform_q(-2, 2), args)


def powder_q(lat_to_xyz, n, seed):
    """C5 Q generator: |Q| ~ U(0.1, 10) 1/angstrom, isotropic directions, converted to rlu with B^-1."""
    rng = np.random.default_rng(seed)
    mod = rng.uniform(0.1, 10.0, int(n))
    v = rng.normal(size=(int(n), 3))
    v /= np.linalg.norm(v, axis=1)[:, None]
    xyz = v * mod[:, None]
    Binv = np.linalg.inv(np.asarray(lat_to_xyz, dtype=float).reshape(3, 3))
    return xyz @ Binv.T


BUILDERS = {"C1": c1_fd3m_scalar, "C2": c2_nacl, "C3": c3_p63mmc, "C4": c4_p21c_nest}


# ---------------------------------------------------------------------------------------------------------------------
# "lattice zoo": small grids over lattices / settings that exercise the remaining branches of the path (centred lattices
# with a primitive transform, time-reversal symmetry added to a non-centrosymmetric group, rhombohedral and triclinic
# cells) with data that transform like real/reciprocal vectors, pseudovectors and matrices (dd and cc classes)
# ---------------------------------------------------------------------------------------------------------------------
ZOO = {
    "P3_timereversal": dict(lengths=(4.0, 4.0, 5.0), angles=(90, 90, 120), sg="P 3", tr=True),
    "I4_timereversal": dict(lengths=(4.0, 4.0, 5.0), angles=(90, 90, 90), sg="I 4", tr=True),
    "F23": dict(lengths=(5.0, 5.0, 5.0), angles=(90, 90, 90), sg="F 2 2 3", tr=False),
    "R-3m_hex": dict(lengths=(4.0, 4.0, 12.0), angles=(90, 90, 120), sg='-R 3 2"', tr=False),
    "P-1": dict(lengths=(4.0, 5.0, 6.0), angles=(80, 85, 95), sg="-P 1", tr=False),
    "Im-3m": dict(lengths=(2.87, 2.87, 2.87), angles=(90, 90, 90), sg="Im-3m", tr=False),
}


def zoo_grid(b, name, cls="BZTrellisQcc", density=60, seed=1):
    """values: complex, per mode 1 scalar + one pseudovector (real_lattice); vectors: complex, per mode 2 scalars + one
    reciprocal-lattice vector ... or, for the dd class, real vectors + one 3x3 matrix."""
    z = ZOO[name]
    lat = b.Lattice(z["lengths"], z["angles"], z["sg"])
    bz = b.BrillouinZone(lat, True, 1, z["tr"])
    g = getattr(b, cls)(bz, bz.ir_polyhedron.volume / density)
    nv = g.rlu.shape[0]
    rng = np.random.default_rng(seed)
    modes = 3

    def rnd(shape, cplx):
        a = rng.normal(size=shape)
        return a + 1j * rng.normal(size=shape) if cplx else a

    if cls.endswith("cc"):
        vals = rnd((nv, modes, 1 + 3), True)       # scalar + pseudovector
        vecs = rnd((nv, modes, 2 + 3), True)       # 2 scalars + reciprocal-lattice vector
        args = (vals, (1, 3, 0, 1, 3), vecs, (2, 3, 0, 0, 4))
    elif cls.endswith("dd"):
        vals = rnd((nv, modes, 3), False)          # real-lattice vector
        vecs = rnd((nv, modes, 1 + 9), False)      # scalar + matrix
        args = (vals, (0, 3, 0, 0, 3), vecs, (1, 0, 9, 0, 3))
    else:  # dc: real eigenvalue-like scalars, complex pseudovector + matrix
        vals = rnd((nv, modes, 2), False)
        vecs = rnd((nv, modes, 3 + 9), True)
        args = (vals, (2, 0, 0, 0, 3), vecs, (0, 3, 9, 1, 3))
    g.fill(*args)
    return Workload(f"zoo {name} {cls}", g, bz, modes, 0, _uniform_q(-2, 2), args)
