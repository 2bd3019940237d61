"""CPU: the C-ABI library loads and exports every symbol include/brille_b200.h declares; no compute without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from brille_b200 import capi
from brille_b200 import tables as T
from helpers import load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "brille_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_[a-z_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    L = capi.lib()
    names = declared_symbols()
    assert len(names) >= 14
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/brille_b200.h but not exported"
    assert set(names) == set(capi.EXPORTS)
    assert L.b200_abi_version() == 1


def test_struct_layout_matches_header_sizes():
    # spot check of the ctypes mirror: pointer-sized alignment, array members
    assert C.sizeof(T.Probe) == 11 * C.sizeof(C.c_void_p)
    assert T.BZTables.P6t.size == 36 and T.BZTables.to_xyz.size == 72
    assert T.InterpDesc.elements.size == 12
    # b200_sf_config_t: uint32 + 3 pointers + 9 doubles + int32, natural alignment
    assert C.sizeof(capi.SFConfig) == 112 and capi.SFConfig.q_transform.offset == 32 and capi.SFConfig.conjugate.offset == 104
    assert C.sizeof(capi.SortConfig) == 56


def test_no_cpu_fallback_without_device():
    if capi.device_count() > 0:
        pytest.skip("a CUDA device is present")
    s, d, _, rest = load_golden("fd3m_scalar_trellis.npz")
    import brille_b200

    with pytest.raises(brille_b200.B200Error) as e:
        brille_b200.B200Grid(None, structure=s, data=d)
    assert e.value.code == T.E_CUDA and "no CPU fallback" in str(e.value)


def test_table_packing_roundtrip(tmp_path):
    s, d, _, rest = load_golden("nacl_prim_trellis.npz")
    T.save_tables(tmp_path / "t.npz", s, d)
    s2, d2 = T.load_tables(tmp_path / "t.npz")
    bz, bz2 = T.pack_bz(s["bz"]), T.pack_bz(s2["bz"])
    assert bytes(bz.P6t) == bytes(bz2.P6t) and bz.n_faces == bz2.n_faces and bz.n_ops == bz2.n_ops
    tr, tr2 = T.pack_trellis(s), T.pack_trellis(s2)
    assert tr.n_tets == tr2.n_tets and tr.n_cubes == tr2.n_cubes and list(tr.n_knots) == list(tr2.n_knots)
    dt = T.pack_data(d2)
    assert dt.vectors.is_complex == 1 and dt.values.is_complex == 0 and dt.n_atoms == 2
    assert np.array_equal(np.asarray(d["vectors_data"]), np.asarray(d2["vectors_data"]))
