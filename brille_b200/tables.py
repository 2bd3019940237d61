"""ctypes mirror of ``include/brille_b200.h`` and packing of bridge dictionaries into it.

The bridge (``brille_b200/bridge/flatten.cpp``) walks brille's host objects and returns plain
dictionaries of numpy arrays; this module turns them into the ``b200_*_tables_t`` structures the C ABI
takes.  Every packed structure keeps references to the (C-contiguous, correctly typed) numpy arrays it
points at in ``_keep`` so they outlive the call.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)
c_uint32_p = C.POINTER(C.c_uint32)
c_uint8_p = C.POINTER(C.c_uint8)

# numeric values of brille's enums (src/enums.hpp:37,43; src/rotates.hpp:27)
GRID_TRELLIS, GRID_NEST, GRID_MESH = 0, 1, 2
NODE_ASSUMED_NULL, NODE_FOUND_NULL, NODE_NULL, NODE_CUBE, NODE_POLY = 0, 1, 2, 3, 4
ROT_VECTOR, ROT_PSEUDOVECTOR, ROT_GAMMA = 0, 1, 2
LEN_NONE, LEN_ANGSTROM, LEN_INVERSE_ANGSTROM, LEN_REAL_LATTICE, LEN_RECIPROCAL_LATTICE = 0, 1, 2, 3, 4

FLAG_NO_MOVE = 1
ST_OUTSIDE_BZ, ST_OUTSIDE_WEDGE, ST_NOT_FOUND, ST_FALLBACK_TET, ST_NEIGHBOUR = 1, 2, 4, 8, 16

E_INVALID, E_CUDA, E_OUTSIDE_BZ, E_OUTSIDE_WEDGE, E_NOT_FOUND, E_UNSUPPORTED, E_NODATA = -1, -2, -3, -4, -5, -6, -7


class BZTables(C.Structure):
    _fields_ = [
        ("transform_needed", C.c_int32),
        ("P6t", C.c_int32 * 9),
        ("invPt", C.c_int32 * 9),
        ("w_recip_metric", C.c_double * 9),
        ("w_real_metric", C.c_double * 9),
        ("w_recip_volume", C.c_double),
        ("o_recip_metric", C.c_double * 9),
        ("o_real_metric", C.c_double * 9),
        ("o_recip_volume", C.c_double),
        ("to_xyz", C.c_double * 9),
        ("n_faces", C.c_int32),
        ("pa", c_double_p),
        ("pb", c_double_p),
        ("pc", c_double_p),
        ("normals", c_double_p),
        ("taus", c_int32_p),
        ("tau_lens", c_double_p),
        ("ca", c_double_p),
        ("cb", c_double_p),
        ("cc", c_double_p),
        ("n_wedge", C.c_int32),
        ("wedge_normals", c_double_p),
        ("no_ir_mirroring", C.c_int32),
        ("float_tolerance", C.c_double),
        ("approx_tolerance", C.c_int32),
        ("n_ops", C.c_int32),
        ("rotations", c_int32_p),
        ("inverse_index", c_int32_p),
        ("identity_index", C.c_int32),
    ]


class TrellisTables(C.Structure):
    _fields_ = [
        ("n_knots", C.c_int32 * 3),
        ("knots", c_double_p * 3),
        ("n_nodes", C.c_uint32),
        ("node_type", c_uint8_p),
        ("node_index", c_uint32_p),
        ("n_cubes", C.c_uint32),
        ("cube_vertices", c_uint32_p),
        ("n_polys", C.c_uint32),
        ("poly_offsets", c_uint32_p),
        ("n_tets", C.c_uint32),
        ("tet_vertices", c_uint32_p),
        ("tet_circum", c_double_p),
        ("tet_volume", c_double_p),
        ("n_vertices", C.c_uint32),
        ("vertices", c_double_p),
    ]


class NestTables(C.Structure):
    _fields_ = [
        ("n_nodes", C.c_uint32),
        ("node_vertices", c_uint32_p),
        ("node_circum", c_double_p),
        ("node_volume", c_double_p),
        ("node_is_leaf", c_uint8_p),
        ("child_begin", c_uint32_p),
        ("child_end", c_uint32_p),
        ("n_vertices", C.c_uint32),
        ("vertices", c_double_p),
        ("tolerance", C.c_double),
        ("digit", C.c_int32),
    ]


class MeshTables(C.Structure):
    _fields_ = [
        ("n_layers", C.c_uint32),
        ("tet_offset", c_uint32_p),
        ("vert_offset", c_uint32_p),
        ("tets", c_uint32_p),
        ("centres", c_double_p),
        ("radii", c_double_p),
        ("vol6", c_double_p),
        ("vertices", c_double_p),
        ("conn_offset", c_uint32_p),
        ("conn_index", c_uint32_p),
    ]


class InterpDesc(C.Structure):
    _fields_ = [
        ("data", C.c_void_p),
        ("is_complex", C.c_int32),
        ("branches", C.c_uint32),
        ("elements", C.c_uint32 * 3),
        ("rotates_like", C.c_int32),
        ("length_unit", C.c_int32),
    ]


class DataTables(C.Structure):
    _fields_ = [
        ("n_vertices", C.c_uint32),
        ("values", InterpDesc),
        ("vectors", InterpDesc),
        ("n_perm_rows", C.c_uint32),
        ("perm_rows", c_uint32_p),
        ("cube_perm", c_uint32_p),
        ("tet_perm", c_uint32_p),
        ("n_atoms", C.c_uint32),
        ("gamma_F0", c_uint32_p),
        ("gamma_vidx", c_uint32_p),
        ("n_gamma_vectors", C.c_uint32),
        ("gamma_vectors", c_double_p),
        ("rot_cart", c_double_p),
    ]


class Probe(C.Structure):
    _fields_ = [
        ("q_ir", c_double_p),
        ("x_ir", c_double_p),
        ("tau", c_int32_p),
        ("ridx", c_int32_p),
        ("invridx", c_int32_p),
        ("cell", c_uint32_p),
        ("tet", c_int32_p),
        ("n_vert", c_int32_p),
        ("vertex", c_uint32_p),
        ("weight", c_double_p),
        ("status", c_uint32_p),
    ]


def _arr(a, dtype, shape=None):
    out = np.ascontiguousarray(np.asarray(a), dtype=dtype)
    if shape is not None:
        out = out.reshape(shape)
    return out


def _ptr(a, typ):
    return a.ctypes.data_as(typ) if a.size else C.cast(None, typ)


def pack_bz(d) -> BZTables:
    """bridge ``flatten_bz`` dictionary -> ``b200_bz_tables_t``"""
    t = BZTables()
    keep = []
    t.transform_needed = int(d["transform_needed"])
    for name in ("P6t", "invPt"):
        getattr(t, name)[:] = [int(x) for x in np.asarray(d[name]).ravel()]
    for name in ("w_recip_metric", "w_real_metric", "o_recip_metric", "o_real_metric", "to_xyz"):
        getattr(t, name)[:] = [float(x) for x in np.asarray(d[name]).ravel()]
    t.w_recip_volume = float(d["w_recip_volume"])
    t.o_recip_volume = float(d["o_recip_volume"])
    F = int(np.asarray(d["pa"]).shape[0])
    t.n_faces = F
    for name in ("pa", "pb", "pc", "normals", "ca", "cb", "cc"):
        a = _arr(d[name], np.float64, (F, 3))
        keep.append(a)
        setattr(t, name, _ptr(a, c_double_p))
    a = _arr(d["taus"], np.int32, (F, 3))
    keep.append(a)
    t.taus = _ptr(a, c_int32_p)
    a = _arr(d["tau_lens"], np.float64, (F,))
    keep.append(a)
    t.tau_lens = _ptr(a, c_double_p)
    wn = _arr(d["wedge_normals"], np.float64).reshape(-1, 3)
    keep.append(wn)
    t.n_wedge = int(wn.shape[0])
    t.wedge_normals = _ptr(wn, c_double_p)
    t.no_ir_mirroring = int(d["no_ir_mirroring"])
    t.float_tolerance = float(d["float_tolerance"])
    t.approx_tolerance = int(d["approx_tolerance"])
    rot = _arr(d["rotations"], np.int32).reshape(-1, 9)
    inv = _arr(d["inverse_index"], np.int32).reshape(-1)
    keep += [rot, inv]
    t.n_ops = int(rot.shape[0])
    t.rotations = _ptr(rot, c_int32_p)
    t.inverse_index = _ptr(inv, c_int32_p)
    t.identity_index = int(d["identity_index"])
    t._keep = keep
    return t


def pack_trellis(d) -> TrellisTables:
    """bridge ``flatten`` dictionary (kind == 'trellis') -> ``b200_trellis_tables_t``"""
    t = TrellisTables()
    keep = []
    for i in range(3):
        k = _arr(d[f"knots{i}"], np.float64)
        keep.append(k)
        t.n_knots[i] = int(k.size)
        t.knots[i] = _ptr(k, c_double_p)
    nt = _arr(d["node_type"], np.uint8)
    ni = _arr(d["node_index"], np.uint32)
    cv = _arr(d["cube_vertices"], np.uint32).reshape(-1, 8)
    po = _arr(d["poly_offsets"], np.uint32)
    tv = _arr(d["tet_vertices"], np.uint32).reshape(-1, 4)
    tc = _arr(d["tet_circum"], np.float64).reshape(-1, 4)
    tvol = _arr(d["tet_volume"], np.float64)
    vx = _arr(d["vertices"], np.float64).reshape(-1, 3)
    keep += [nt, ni, cv, po, tv, tc, tvol, vx]
    t.n_nodes = int(nt.size)
    t.node_type = _ptr(nt, c_uint8_p)
    t.node_index = _ptr(ni, c_uint32_p)
    t.n_cubes = int(cv.shape[0])
    t.cube_vertices = _ptr(cv, c_uint32_p)
    t.n_polys = int(po.size - 1)
    t.poly_offsets = _ptr(po, c_uint32_p)
    t.n_tets = int(tv.shape[0])
    t.tet_vertices = _ptr(tv, c_uint32_p)
    t.tet_circum = _ptr(tc, c_double_p)
    t.tet_volume = _ptr(tvol, c_double_p)
    t.n_vertices = int(vx.shape[0])
    t.vertices = _ptr(vx, c_double_p)
    t._keep = keep
    return t


def pack_nest(d) -> NestTables:
    """bridge ``flatten`` dictionary (kind == 'nest') -> ``b200_nest_tables_t``"""
    t = NestTables()
    nv = _arr(d["node_vertices"], np.uint32).reshape(-1, 4)
    nc = _arr(d["node_circum"], np.float64).reshape(-1, 4)
    vol = _arr(d["node_volume"], np.float64)
    leaf = _arr(d["node_is_leaf"], np.uint8)
    cb = _arr(d["child_begin"], np.uint32)
    ce = _arr(d["child_end"], np.uint32)
    vx = _arr(d["vertices"], np.float64).reshape(-1, 3)
    t.n_nodes = int(vol.size)
    t.node_vertices = _ptr(nv, c_uint32_p)
    t.node_circum = _ptr(nc, c_double_p)
    t.node_volume = _ptr(vol, c_double_p)
    t.node_is_leaf = _ptr(leaf, c_uint8_p)
    t.child_begin = _ptr(cb, c_uint32_p)
    t.child_end = _ptr(ce, c_uint32_p)
    t.n_vertices = int(vx.shape[0])
    t.vertices = _ptr(vx, c_double_p)
    t.tolerance = float(d["approx_reciprocal"])
    t.digit = int(d["approx_digit"])
    t._keep = [nv, nc, vol, leaf, cb, ce, vx]
    return t


def pack_mesh(d) -> MeshTables:
    """bridge ``flatten`` dictionary (kind == 'mesh') -> ``b200_mesh_tables_t``"""
    t = MeshTables()
    to = _arr(d["tet_offset"], np.uint32)
    vo = _arr(d["vert_offset"], np.uint32)
    tets = _arr(d["tets"], np.uint32).reshape(-1, 4)
    cen = _arr(d["centres"], np.float64).reshape(-1, 3)
    rad = _arr(d["radii"], np.float64)
    vol6 = _arr(d["vol6"], np.float64)
    vx = _arr(d["vertices"], np.float64).reshape(-1, 3)
    co = _arr(d["conn_offset"], np.uint32)
    ci = _arr(d["conn_index"], np.uint32)
    t.n_layers = int(d["n_layers"])
    t.tet_offset = _ptr(to, c_uint32_p)
    t.vert_offset = _ptr(vo, c_uint32_p)
    t.tets = _ptr(tets, c_uint32_p)
    t.centres = _ptr(cen, c_double_p)
    t.radii = _ptr(rad, c_double_p)
    t.vol6 = _ptr(vol6, c_double_p)
    t.vertices = _ptr(vx, c_double_p)
    t.conn_offset = _ptr(co, c_uint32_p)
    t.conn_index = _ptr(ci, c_uint32_p)
    t._keep = [to, vo, tets, cen, rad, vol6, vx, co, ci]
    return t


def pack_structure(d):
    """dispatch on the bridge dictionary's ``kind`` -> (grid kind, packed structure tables)"""
    kind = str(d["kind"])
    if kind == "trellis":
        return GRID_TRELLIS, pack_trellis(d)
    if kind == "nest":
        return GRID_NEST, pack_nest(d)
    if kind == "mesh":
        return GRID_MESH, pack_mesh(d)
    raise ValueError(f"unknown grid kind {kind!r}")


def _pack_interp(desc: InterpDesc, d, prefix, keep):
    data = np.asarray(d[f"{prefix}_data"])
    is_complex = np.iscomplexobj(data)
    data = np.ascontiguousarray(data, dtype=np.complex128 if is_complex else np.float64)
    keep.append(data)
    desc.data = data.ctypes.data if data.size else None
    desc.is_complex = 1 if is_complex else 0
    desc.branches = int(d[f"{prefix}_branches"])
    el = [int(x) for x in np.asarray(d[f"{prefix}_elements"]).ravel()]
    desc.elements[:] = el
    desc.rotates_like = int(d[f"{prefix}_rotlike"])
    desc.length_unit = int(d[f"{prefix}_lenunit"])
    span = sum(el)
    if data.ndim != 2 or (data.size and data.shape[1] != desc.branches * span):
        raise ValueError(f"{prefix}: data shape {data.shape} inconsistent with branches={desc.branches} span={span}")
    return data


def pack_data(d) -> DataTables:
    """bridge ``flatten_data`` dictionary -> ``b200_data_tables_t``"""
    t = DataTables()
    keep = []
    vals = _pack_interp(t.values, d, "values", keep)
    _pack_interp(t.vectors, d, "vectors", keep)
    t.n_vertices = int(vals.shape[0])
    rows = _arr(d.get("perm_rows", np.zeros((0, 0))), np.uint32)
    keep.append(rows)
    t.n_perm_rows = int(rows.shape[0]) if rows.ndim == 2 else 0
    t.perm_rows = _ptr(rows, c_uint32_p)
    if int(d.get("perm_nonidentity", 0)):
        for name in ("cube_perm", "tet_perm"):
            if name not in d:
                continue
            a = _arr(d[name], np.uint32)
            keep.append(a)
            setattr(t, name, _ptr(a, c_uint32_p))
    else:
        t.n_perm_rows = min(t.n_perm_rows, 1)
    nat = int(d.get("gamma_natoms", 0))
    t.n_atoms = nat
    if nat:
        f0 = _arr(d["gamma_F0"], np.uint32)
        vi = _arr(d["gamma_vidx"], np.uint32)
        gv = _arr(d["gamma_vectors"], np.float64).reshape(-1, 3)
        keep += [f0, vi, gv]
        t.gamma_F0 = _ptr(f0, c_uint32_p)
        t.gamma_vidx = _ptr(vi, c_uint32_p)
        t.n_gamma_vectors = int(gv.shape[0])
        t.gamma_vectors = _ptr(gv, c_double_p)
    if "rot_cart" in d:
        rc = _arr(d["rot_cart"], np.float64).reshape(-1, 9)
        keep.append(rc)
        t.rot_cart = _ptr(rc, c_double_p)
    t._keep = keep
    return t


class ProbeArrays:
    """Host arrays for the optional per-Q intermediate results + the ctypes view of them."""

    FIELDS = {
        "q_ir": (np.float64, 3),
        "x_ir": (np.float64, 3),
        "tau": (np.int32, 3),
        "ridx": (np.int32, 0),
        "invridx": (np.int32, 0),
        "cell": (np.uint32, 0),
        "tet": (np.int32, 0),
        "n_vert": (np.int32, 0),
        "vertex": (np.uint32, 8),
        "weight": (np.float64, 8),
        "status": (np.uint32, 0),
    }

    def __init__(self, n, fields=None):
        self.n = int(n)
        self.struct = Probe()
        for name, (dt, w) in self.FIELDS.items():
            if fields is not None and name not in fields:
                continue
            a = np.zeros((self.n, w) if w else (self.n,), dtype=dt)
            setattr(self, name, a)
            ptr_t = dict(Probe._fields_)[name]
            setattr(self.struct, name, a.ctypes.data_as(ptr_t))

    def byref(self):
        return C.byref(self.struct)


def save_tables(path, structure, data=None):
    """Serialise bridge dictionaries into one ``.npz`` (used for the committed golden fixtures)."""
    flat = {}

    def put(prefix, d):
        for k, v in d.items():
            if isinstance(v, dict):
                put(f"{prefix}{k}.", v)
            elif isinstance(v, str):
                flat[f"{prefix}{k}"] = np.array(v)
            else:
                flat[f"{prefix}{k}"] = np.asarray(v)

    put("s.", structure)
    if data is not None:
        put("d.", data)
    np.savez_compressed(path, **flat)


def load_tables(path):
    """Inverse of :func:`save_tables` -> (structure dict, data dict or None)"""
    z = np.load(path, allow_pickle=False)
    out = {"s": {}, "d": {}}
    for key in z.files:
        parts = key.split(".")
        d = out[parts[0]]
        for p in parts[1:-1]:
            d = d.setdefault(p, {})
        v = z[key]
        if v.dtype.kind in "US" and v.ndim == 0:
            v = str(v)
        elif v.ndim == 0:
            v = v.item()
        d[parts[-1]] = v
    return out["s"], (out["d"] or None)
