"""ctypes binding of the C ABI (``include/brille_b200.h``) exported by ``libbrille_b200.so``.

There is no CPU fallback: if the CUDA library is missing, or no CUDA device is present when a grid is
created, the calls below raise.
"""
from __future__ import annotations

import ctypes as C
import os

from . import tables as T

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BRILLE_B200_LIB") or os.path.join(HERE, "libbrille_b200.so")  # (the override is for A/B timing of builds)

#: every symbol include/brille_b200.h declares
EXPORTS = (
    "b200_grid_create",
    "b200_grid_set_data",
    "b200_grid_destroy",
    "b200_ir_interpolate_at",
    "b200_ir_interpolate_at_device",
    "b200_interpolate_at",
    "b200_moveinto",
    "b200_alloc_pinned",
    "b200_free_pinned",
    "b200_last_error",
    "b200_abi_version",
    "b200_device_count",
    "b200_grid_launch_count",
    "b200_grid_last_path",
    "b200_grid_enable_timing",
    "b200_grid_kernel_ms",
    "b200_grid_row_bytes",
    "b200_grid_set_option",
    "b200_grid_sort_pairs",
    "b200_solve_assignments",
    "b200_grid_set_structure_factor",
    "b200_ir_structure_factor",
    "b200_ir_structure_factor_device",
    "b200_ir_powder_bin",
    "b200_ir_powder_sweep",
    "b200_powder_points",
)


class SortConfig(C.Structure):
    """``b200_sort_config_t``"""

    _fields_ = [
        ("values_costmult", C.c_double * 3),
        ("vectors_costmult", C.c_double * 3),
        ("values_vector_cost", C.c_int32),
        ("vectors_vector_cost", C.c_int32),
    ]


class SFConfig(C.Structure):
    """``b200_sf_config_t``"""

    _fields_ = [
        ("n_atoms", C.c_uint32),
        ("coef", C.c_void_p),
        ("positions", C.c_void_p),
        ("debye_waller", C.c_void_p),
        ("q_transform", C.c_double * 9),
        ("conjugate", C.c_int32),
    ]


class PowderConfig(C.Structure):
    """``b200_powder_config_t``"""

    _fields_ = [
        ("n_qbins", C.c_uint32),
        ("n_wbins", C.c_uint32),
        ("q_lo", C.c_double),
        ("q_hi", C.c_double),
        ("w_lo", C.c_double),
        ("w_hi", C.c_double),
        ("weight", C.c_int32),
    ]


class B200Error(RuntimeError):
    """Raised for every non-zero return code; ``code`` is the B200_E_* value (mirrors brille's RuntimeError)."""

    def __init__(self, code, message):
        super().__init__(message)
        self.code = code


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m brille_b200.build` (nvcc, sm_100a). "
            "brille_b200 has no CPU fallback."
        )
    L = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    L.b200_grid_create.restype = C.c_int
    L.b200_grid_create.argtypes = [C.c_int, C.POINTER(T.BZTables), vp, C.c_int, C.POINTER(vp)]
    L.b200_grid_set_data.restype = C.c_int
    L.b200_grid_set_data.argtypes = [vp, C.POINTER(T.DataTables)]
    L.b200_grid_destroy.restype = None
    L.b200_grid_destroy.argtypes = [vp]
    for name in ("b200_ir_interpolate_at", "b200_interpolate_at"):
        f = getattr(L, name)
        f.restype = C.c_int
        f.argtypes = [vp, vp, C.c_size_t, C.c_uint32, vp, vp, C.POINTER(T.Probe)]
    L.b200_ir_interpolate_at_device.restype = C.c_int
    L.b200_ir_interpolate_at_device.argtypes = [vp, vp, C.c_size_t, C.c_uint32, vp, vp, C.POINTER(T.Probe), vp, C.POINTER(C.c_uint64)]
    L.b200_moveinto.restype = C.c_int
    L.b200_moveinto.argtypes = [vp, vp, C.c_size_t, C.c_int, C.POINTER(T.Probe)]
    L.b200_alloc_pinned.restype = vp
    L.b200_alloc_pinned.argtypes = [C.c_size_t]
    L.b200_free_pinned.restype = None
    L.b200_free_pinned.argtypes = [vp]
    L.b200_last_error.restype = C.c_char_p
    L.b200_abi_version.restype = C.c_int
    L.b200_device_count.restype = C.c_int
    L.b200_grid_launch_count.restype = C.c_uint64
    L.b200_grid_launch_count.argtypes = [vp]
    L.b200_grid_last_path.restype = C.c_uint32
    L.b200_grid_last_path.argtypes = [vp]
    L.b200_grid_enable_timing.restype = C.c_int
    L.b200_grid_enable_timing.argtypes = [vp, C.c_int]
    L.b200_grid_kernel_ms.restype = C.c_double
    L.b200_grid_kernel_ms.argtypes = [vp, C.c_char_p]
    L.b200_grid_set_option.restype = C.c_int
    L.b200_grid_set_option.argtypes = [vp, C.c_char_p, C.c_double]
    L.b200_grid_sort_pairs.restype = C.c_int
    L.b200_grid_sort_pairs.argtypes = [vp, vp, C.c_size_t, C.POINTER(SortConfig), vp, vp, vp]
    L.b200_solve_assignments.restype = C.c_int
    L.b200_solve_assignments.argtypes = [vp, C.c_size_t, C.c_uint32, vp, vp, C.c_int]
    L.b200_grid_set_structure_factor.restype = C.c_int
    L.b200_grid_set_structure_factor.argtypes = [vp, C.POINTER(SFConfig)]
    L.b200_ir_structure_factor.restype = C.c_int
    L.b200_ir_structure_factor.argtypes = [vp, vp, C.c_size_t, C.c_uint32, vp, vp]
    L.b200_ir_structure_factor_device.restype = C.c_int
    L.b200_ir_structure_factor_device.argtypes = [vp, vp, C.c_size_t, C.c_uint32, vp, vp, vp, vp, C.POINTER(C.c_uint64)]
    L.b200_ir_powder_bin.restype = C.c_int
    L.b200_ir_powder_bin.argtypes = [vp, vp, C.c_size_t, C.c_uint32, C.POINTER(PowderConfig), vp, vp]
    L.b200_ir_powder_sweep.restype = C.c_int
    L.b200_ir_powder_sweep.argtypes = [vp, C.POINTER(PowderConfig), C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, vp, vp]
    L.b200_powder_points.restype = C.c_int
    L.b200_powder_points.argtypes = [vp, C.POINTER(PowderConfig), C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, vp]
    L.b200_grid_row_bytes.restype = C.c_int
    L.b200_grid_row_bytes.argtypes = [vp, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
    _lib = L
    return L


def check(rc):
    if rc != 0:
        msg = lib().b200_last_error()
        raise B200Error(rc, msg.decode() if msg else f"brille_b200 error {rc}")


def solve_assignments(cost, device=0):
    """(row, col) solutions of the assignment problems ``cost`` (n, modes, modes) on the device (``b200_solve_assignments``)."""
    import numpy as np

    cost = np.ascontiguousarray(cost, dtype=np.float64)
    n, B, B2 = cost.shape
    if B != B2:
        raise ValueError("cost matrices must be square")
    row = np.zeros((n, B), dtype=np.int32)
    col = np.zeros((n, B), dtype=np.int32)
    check(lib().b200_solve_assignments(cost.ctypes.data, n, B, row.ctypes.data, col.ctypes.data, int(device)))
    return row, col


def device_count() -> int:
    return int(lib().b200_device_count())
