"""ctypes front end of the plain-C oracle (oracle/brille_oracle.c).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from brille_b200 import tables as T

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "brille_oracle.c")
SRC_SORT = os.path.join(HERE, "sort_oracle.c")
LIB = os.path.join(HERE, "liboracle.so")


def build(force=False):
    """gcc -O2 without FMA contraction (the reference is x86-64 baseline code: no fused multiply-add)."""
    hdr = os.path.join(ROOT, "include", "brille_b200.h")
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= max(os.path.getmtime(SRC), os.path.getmtime(SRC_SORT), os.path.getmtime(hdr)):
        return LIB
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-fopenmp", "-I", os.path.join(ROOT, "include"), SRC, SRC_SORT, "-o", LIB, "-lm"]
    subprocess.run(cmd, check=True)
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.oracle_interpolate_at.restype = C.c_int
        _lib.oracle_interpolate_at.argtypes = [
            C.c_int, C.POINTER(T.BZTables), C.c_void_p, C.POINTER(T.DataTables), C.c_void_p, C.c_size_t,
            C.c_uint32, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(T.Probe),
        ]
        _lib.oracle_sort_pairs.restype = C.c_int
        _lib.oracle_sort_pairs.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                           C.c_int, C.c_uint32, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.oracle_lapjv_batch.restype = C.c_int
        _lib.oracle_lapjv_batch.argtypes = [C.c_void_p, C.c_size_t, C.c_uint32, C.c_void_p, C.c_void_p]
        _lib.oracle_moveinto.restype = C.c_int
        _lib.oracle_moveinto.argtypes = [C.POINTER(T.BZTables), C.c_void_p, C.c_size_t, C.c_int, C.POINTER(T.Probe)]
    return _lib


class Oracle:
    """The oracle over one set of flat tables (bridge dictionaries)."""

    def __init__(self, structure, data=None):
        self.bz = T.pack_bz(structure["bz"])
        self.kind, self.structure = T.pack_structure(structure)
        self.data = T.pack_data(data) if data is not None else None

    def set_data(self, data):
        self.data = T.pack_data(data)

    def row_shapes(self):
        d = self.data
        vs = d.values.branches * sum(d.values.elements)
        ws = d.vectors.branches * sum(d.vectors.elements)
        return vs, ws

    def interpolate_at(self, Q, ir=True, no_move=False, probe=True):
        Q = np.ascontiguousarray(Q, dtype=np.float64).reshape(-1, 3)
        n = Q.shape[0]
        vs, ws = self.row_shapes()
        vals = np.zeros((n, vs), dtype=np.complex128 if self.data.values.is_complex else np.float64)
        vecs = np.zeros((n, ws), dtype=np.complex128 if self.data.vectors.is_complex else np.float64)
        pr = T.ProbeArrays(n) if probe else None
        rc = lib().oracle_interpolate_at(
            self.kind, C.byref(self.bz), C.cast(C.pointer(self.structure), C.c_void_p), C.byref(self.data),
            Q.ctypes.data, n, T.FLAG_NO_MOVE if no_move else 0, 1 if ir else 0, vals.ctypes.data, vecs.ctypes.data,
            pr.byref() if pr is not None else None,
        )
        return rc, vals, vecs, pr

    def moveinto(self, Q, ir=True):
        """``ir``: False / 0 moveinto, True / 1 ir_moveinto, 2 ir_moveinto_wedge, 3 isinside (status bit ST_OUTSIDE_BZ)."""
        Q = np.ascontiguousarray(Q, dtype=np.float64).reshape(-1, 3)
        pr = T.ProbeArrays(Q.shape[0], fields=("q_ir", "x_ir", "tau", "ridx", "invridx", "status"))
        rc = lib().oracle_moveinto(C.byref(self.bz), Q.ctypes.data, Q.shape[0], int(ir), pr.byref())
        return rc, pr


def sort_pairs(data, plan, pairs=None, want_cost=False):
    """DualInterpolator::sort() restated (sort_oracle.c) for the vertex pairs of a bridge ``sort_plan``.

    ``data`` is the bridge ``flatten_data`` dictionary.  Returns (rc, row, col[, cost]): row[k] is the permutation the
    reference stores for the ordered pair (i, j) = pairs[k], col[k] the one for (j, i)."""
    pairs = np.ascontiguousarray(plan["pairs"] if pairs is None else pairs, dtype=np.uint32).reshape(-1, 2)
    n = pairs.shape[0]

    def interp(prefix):
        a = np.asarray(data[f"{prefix}_data"])
        cplx = np.iscomplexobj(a)
        a = np.ascontiguousarray(a, dtype=np.complex128 if cplx else np.float64)
        el = np.ascontiguousarray(np.asarray(data[f"{prefix}_elements"]).ravel(), dtype=np.uint32)
        cm = np.ascontiguousarray(plan[f"{prefix}_costmult"], dtype=np.float64)
        return a, int(cplx), el, cm, int(plan[f"{prefix}_vector_cost"])

    va, vc, vel, vcm, vfun = interp("values")
    wa, wc, wel, wcm, wfun = interp("vectors")
    B = int(data["values_branches"])
    row = np.zeros((n, B), dtype=np.int32)
    col = np.zeros((n, B), dtype=np.int32)
    cost = np.zeros((n, B, B), dtype=np.float64) if want_cost else None
    rc = lib().oracle_sort_pairs(va.ctypes.data, vc, vel.ctypes.data, vcm.ctypes.data, vfun, wa.ctypes.data, wc, wel.ctypes.data,
                                 wcm.ctypes.data, wfun, B, pairs.ctypes.data, n, row.ctypes.data, col.ctypes.data,
                                 cost.ctypes.data if want_cost else None)
    return (rc, row, col, cost) if want_cost else (rc, row, col)


def lapjv_batch(cost):
    """Jonker-Volgenant assignment (lapjv.hpp:281-538 restated) of every (branches, branches) cost matrix in ``cost``."""
    cost = np.ascontiguousarray(cost, dtype=np.float64)
    n, B, _ = cost.shape
    row = np.zeros((n, B), dtype=np.int32)
    col = np.zeros((n, B), dtype=np.int32)
    lib().oracle_lapjv_batch(cost.ctypes.data, n, B, row.ctypes.data, col.ctypes.data)
    return row, col
