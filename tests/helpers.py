"""Shared comparison helpers of the parity tests."""
import os

import numpy as np

from brille_b200 import tables as T

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

#: relative tolerance BASELINE.json's north_star states for eigenvalues / rotated eigenvectors
RTOL = 1e-10


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name), allow_pickle=False)
    out = {"s": {}, "d": {}, "d2": {}}
    rest = {}
    for key in z.files:
        parts = key.split(".")
        if parts[0] in out and len(parts) > 1:
            d = out[parts[0]]
            for p in parts[1:-1]:
                d = d.setdefault(p, {})
            v = z[key]
            if v.dtype.kind in "US" and v.ndim == 0:
                v = str(v)
            elif v.ndim == 0:
                v = v.item()
            d[parts[-1]] = v
        else:
            rest[key] = z[key]
    return out["s"], out["d"], out["d2"], rest


def rel_err(a, b):
    """Relative error PER (Q, mode) row: ||got - want||_2 / ||want||_2 over the elements of one mode of one point (for
    eigenvalues, one element per mode, that is the element-wise relative error; for eigenvectors it is the error of the
    eigenvector), maximum over all rows.  north_star: "1e-10 relative".  Rows whose reference norm is below 1e-3 of the median
    row norm (random test data that cancel in the weighted sum) are measured against that floor instead of their own norm.

    The arrays are (nQ, modes, ...) as the grid returns them; if only one of them has that shape the other is reshaped to it; two
    flat (nQ, k) arrays are compared per point."""
    a = np.asarray(a)
    b = np.asarray(b)
    if a.size == 0:
        return 0.0
    shape = a.shape if a.ndim >= 3 else (b.shape if b.ndim >= 3 else None)
    if shape is None:
        shape = (a.shape[0], 1, -1) if a.ndim >= 1 and a.shape[0] > 0 else (1, 1, -1)
    a3 = a.reshape(shape[0], shape[1], -1)
    b3 = b.reshape(a3.shape)
    num = np.linalg.norm(a3 - b3, axis=2)
    den = np.linalg.norm(b3, axis=2)
    floor = max(1e-3 * float(np.median(den)), 1e-300)
    return float((num / np.maximum(den, floor)).max())


def assert_values_close(got, want, rtol=RTOL):
    e = rel_err(got, want)
    assert e <= rtol, f"relative error per (Q, mode) {e:.3e} > {rtol}"


def assert_decisions_equal(pr, ref, what="oracle", adaptive_ulps=0):
    """tau, rotation indices, vertex lists and weights must be identical (bit-exact).

    ``adaptive_ulps`` > 0 (Nest / Mesh grids only): points ON a face/edge/vertex of their tetrahedron (fewer than 4
    emitted vertices) have determinants inside the error bound of TetGen's orient3d, where the reference refines the
    value with Shewchuk's exact expansions and this project with double-double arithmetic ("orient3d"); the
    vanishing weight is folded into the largest one (nest.hpp:203-217), which may then differ in its last bits."""
    n = len(ref["tau"])
    assert np.array_equal(pr.tau[:n], ref["tau"]), f"{what}: tau differs"
    assert np.array_equal(pr.ridx[:n], ref["ridx"]), f"{what}: Ridx differs"
    assert np.array_equal(pr.invridx[:n], ref["invridx"]), f"{what}: invRidx differs"
    assert np.array_equal(pr.q_ir[:n], ref["q_ir"]), f"{what}: q_ir not bit-identical"
    if "n_vert" in ref:
        assert np.array_equal(pr.n_vert[:n], ref["n_vert"]), f"{what}: vertex counts differ"
        assert np.array_equal(pr.vertex[:n], ref["vertex"]), f"{what}: vertex lists differ"
        if adaptive_ulps:
            full = np.asarray(ref["n_vert"]) == 4
            assert np.array_equal(pr.weight[:n][full], ref["weight"][full]), f"{what}: weights of interior points not bit-identical"
            assert np.abs(pr.weight[:n] - ref["weight"]).max() <= adaptive_ulps * 2.3e-16, f"{what}: weights of on-face points differ"
        else:
            assert np.array_equal(pr.weight[:n], ref["weight"]), f"{what}: weights not bit-identical"


def ref_decisions(rest, prefix="ref_"):
    out = {"tau": rest[prefix + "tau"], "q_ir": rest[prefix + "q_ir"]}
    if prefix + "ridx" in rest:
        out["ridx"] = rest[prefix + "ridx"]
        out["invridx"] = rest[prefix + "invridx"]
    out["n_vert"] = rest[prefix + "n_vert"]
    out["vertex"] = rest[prefix + "vertex"]
    out["weight"] = rest[prefix + "weight"]
    return out


def remove_mode_phase(vec, like):
    """Multiply every (Q, mode) eigenvector by the phase that makes <like|vec> real (the freedom brille leaves)."""
    vec = np.asarray(vec)
    like = np.asarray(like).reshape(vec.shape)
    ax = tuple(range(2, vec.ndim))
    ph = np.exp(-1j * np.angle(np.sum(np.conj(like) * vec, axis=ax)))
    return vec * ph.reshape(ph.shape + (1,) * (vec.ndim - 2))
