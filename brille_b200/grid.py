"""Host-side mirror of brille's grid interface for the accelerated path.

``B200Grid`` wraps a brille host grid object (``BZTrellisQdd/dc/cc`` ...; construction, ``fill`` and
``sort`` stay brille's C++) and routes

* ``ir_interpolate_at(Q, useparallel=False, threads=-1, do_not_move_points=False)``  wrap/_common_grid.hpp:272-341
* ``interpolate_at(Q, useparallel=False, threads=-1, do_not_move_points=False)``     wrap/_common_grid.hpp:408-437
* ``ir_moveinto(Q)`` / ``moveinto(Q)`` of the grid's Brillouin zone                    wrap/_bz.cpp:378-463

through the C ABI of ``include/brille_b200.h`` to the CUDA kernels.  Names, argument meaning, output shapes and
error behaviour (``RuntimeError``, all-or-nothing) are the reference's.  Every other attribute is forwarded to
the wrapped host object, so user code written against brille keeps working.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi
from . import tables as T


def _bridge():
    try:
        from . import _bridge as br
    except ImportError as e:  # pragma: no cover - build problem
        raise RuntimeError(
            "brille_b200._bridge is not built: run brille_b200/bridge/build_bridge.sh against the brille sources"
        ) from e
    return br


class PinnedArray:
    """numpy view of page-locked host memory obtained from the library (freed with the object)."""

    def __init__(self, shape, dtype):
        self.shape = tuple(int(s) for s in shape)
        self.dtype = np.dtype(dtype)
        nbytes = int(np.prod(self.shape, dtype=np.int64)) * self.dtype.itemsize
        self._ptr = capi.lib().b200_alloc_pinned(max(nbytes, 1))
        if not self._ptr:
            raise MemoryError("b200_alloc_pinned failed")
        buf = (C.c_char * max(nbytes, 1)).from_address(self._ptr)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=int(np.prod(self.shape, dtype=np.int64))).reshape(self.shape)

    def __del__(self):
        try:
            if getattr(self, "_ptr", None):
                self.array = None
                capi.lib().b200_free_pinned(self._ptr)
                self._ptr = None
        except Exception:
            pass


class B200Grid:
    """A brille grid whose interpolation path runs on one B200."""

    _KINDS = {"trellis": T.GRID_TRELLIS, "nest": T.GRID_NEST, "mesh": T.GRID_MESH}

    def __init__(self, host_grid, device=0, structure=None, data=None):
        """``host_grid``: a brille BZ{Trellis,Nest,Mesh}Q{dd,dc,cc} object (may be None when flat tables are given).

        ``structure`` / ``data``: bridge dictionaries (e.g. loaded with ``tables.load_tables``) to use instead of
        flattening ``host_grid``.
        """
        self._host = host_grid
        self._handle = None
        self.device = int(device)
        if structure is None:
            structure = _bridge().flatten(host_grid)
        self._structure = structure
        self._bz_tables = T.pack_bz(structure["bz"])
        self._kind, self._struct_tables = T.pack_structure(structure)
        h = C.c_void_p()
        capi.check(
            capi.lib().b200_grid_create(
                self._kind, C.byref(self._bz_tables), C.cast(C.pointer(self._struct_tables), C.c_void_p), self.device, C.byref(h)
            )
        )
        self._handle = h
        self._vals_shape = self._vecs_shape = None
        self._vals_dtype = self._vecs_dtype = None
        if data is not None:
            self._set_data(data)
        elif host_grid is not None:
            self._sync_data()

    # ------------------------------------------------------------------ life cycle
    def close(self):
        if self._handle is not None:
            capi.lib().b200_grid_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __getattr__(self, name):  # everything else is brille's own object
        host = self.__dict__.get("_host")
        if host is None:
            raise AttributeError(name)
        return getattr(host, name)

    # ------------------------------------------------------------------ data
    def _set_data(self, d):
        self._data_tables = T.pack_data(d)
        capi.check(capi.lib().b200_grid_set_data(self._handle, C.byref(self._data_tables)))
        self._vals_shape = tuple(int(x) for x in d["values_shape"])[1:]
        self._vecs_shape = tuple(int(x) for x in d["vectors_shape"])[1:]
        self._vals_dtype = np.complex128 if self._data_tables.values.is_complex else np.float64
        self._vecs_dtype = np.complex128 if self._data_tables.vectors.is_complex else np.float64

    def _sync_data(self):
        d = _bridge().flatten_data(self._host)
        if int(np.asarray(d["values_data"]).size) == 0 and int(np.asarray(d["vectors_data"]).size) == 0:
            return  # nothing filled yet
        self._set_data(d)

    def fill(self, *args, **kwargs):
        """brille ``fill`` (wrap/_common_grid.hpp:26-47,127-147) followed by the table upload."""
        out = self._host.fill(*args, **kwargs)
        self._sync_data()
        return out

    def sort(self, *args, **kwargs):
        out = self._host.sort(*args, **kwargs)
        self._sync_data()
        return out

    # ------------------------------------------------------------------ the path
    def _check_q(self, Q):
        Q = np.asarray(Q, dtype=np.float64)
        if Q.ndim != 2 or Q.shape[1] != 3:
            raise RuntimeError("Interpolation requires one or more 3-vectors")
        return np.ascontiguousarray(Q)

    def _outputs(self, n, pinned):
        if self._vals_shape is None:
            raise RuntimeError("The interpolation data must be filled before interpolating.")
        vs, ws = (n,) + self._vals_shape, (n,) + self._vecs_shape
        if pinned:
            pv, pw = PinnedArray(vs, self._vals_dtype), PinnedArray(ws, self._vecs_dtype)
            return pv.array, pw.array, (pv, pw)
        return np.empty(vs, self._vals_dtype), np.empty(ws, self._vecs_dtype), None

    def _run(self, fn, Q, no_move, probe, pinned, out=None):
        Q = self._check_q(Q)
        n = Q.shape[0]
        if out is not None:  # caller-provided (e.g. pinned, reused) output buffers
            vals, vecs = out
            keep = None
            vs, ws = (n,) + self._vals_shape, (n,) + self._vecs_shape
            if vals.shape != vs or vecs.shape != ws or vals.dtype != self._vals_dtype or vecs.dtype != self._vecs_dtype:
                raise RuntimeError(f"out buffers must have shapes {vs} / {ws} and dtypes {self._vals_dtype} / {self._vecs_dtype}")
            if not (vals.flags.c_contiguous and vecs.flags.c_contiguous):
                raise RuntimeError("out buffers must be C-contiguous")
        else:
            vals, vecs, keep = self._outputs(n, pinned)
        pr = T.ProbeArrays(n) if probe else None
        rc = fn(
            self._handle, Q.ctypes.data, n, T.FLAG_NO_MOVE if no_move else 0, vals.ctypes.data, vecs.ctypes.data,
            pr.byref() if pr is not None else None,
        )
        capi.check(rc)
        if keep is not None:  # keep the pinned allocation alive as long as the arrays
            vals = _Owned(vals, keep[0])
            vecs = _Owned(vecs, keep[1])
        return (vals, vecs, pr) if probe else (vals, vecs)

    def ir_interpolate_at(self, Q, useparallel=False, threads=-1, do_not_move_points=False, *, probe=False, pinned=False, out=None):
        """Same signature as brille; ``useparallel`` / ``threads`` are accepted and ignored (the GPU is the parallelism).
        Extras (keyword only): ``probe`` also returns the per-Q decisions, ``pinned`` allocates page-locked outputs,
        ``out=(vals, vecs)`` writes into caller-owned buffers."""
        return self._run(capi.lib().b200_ir_interpolate_at, Q, do_not_move_points, probe, pinned, out)

    def interpolate_at(self, Q, useparallel=False, threads=-1, do_not_move_points=False, *, probe=False, pinned=False, out=None):
        return self._run(capi.lib().b200_interpolate_at, Q, do_not_move_points, probe, pinned, out)

    def ir_moveinto(self, Q, threads=0):
        """(q_ir, tau, R, invR) like ``BrillouinZone.ir_moveinto`` (wrap/_bz.cpp:434-463) but with the operation INDICES."""
        Q = self._check_q(Q)
        pr = T.ProbeArrays(Q.shape[0], fields=("q_ir", "x_ir", "tau", "ridx", "invridx", "status"))
        capi.check(capi.lib().b200_moveinto(self._handle, Q.ctypes.data, Q.shape[0], 1, pr.byref()))
        return pr.q_ir, pr.tau, pr.ridx, pr.invridx

    def moveinto(self, Q, threads=0):
        """(q, tau) like ``BrillouinZone.moveinto`` (wrap/_bz.cpp:378-405)."""
        Q = self._check_q(Q)
        pr = T.ProbeArrays(Q.shape[0], fields=("q_ir", "tau", "ridx", "invridx", "status"))
        capi.check(capi.lib().b200_moveinto(self._handle, Q.ctypes.data, Q.shape[0], 0, pr.byref()))
        return pr.q_ir, pr.tau

    # ------------------------------------------------------------------ device-resident variant
    def ir_interpolate_at_device(self, dQ, vals_out=None, vecs_out=None, do_not_move_points=False, check=True, stream=None):
        """torch CUDA tensors in and out (no host copies).  ``dQ``: float64 (n,3) on this grid's device."""
        import torch

        if not dQ.is_cuda or dQ.dtype != torch.float64 or dQ.dim() != 2 or dQ.shape[1] != 3:
            raise RuntimeError("dQ must be a CUDA float64 tensor of shape (n,3)")
        dQ = dQ.contiguous()
        n = int(dQ.shape[0])
        if self._vals_shape is None:
            raise RuntimeError("The interpolation data must be filled before interpolating.")
        tv = torch.complex128 if self._vals_dtype == np.complex128 else torch.float64
        tw = torch.complex128 if self._vecs_dtype == np.complex128 else torch.float64
        if vals_out is None:
            vals_out = torch.empty((n,) + self._vals_shape, dtype=tv, device=dQ.device)
        if vecs_out is None:
            vecs_out = torch.empty((n,) + self._vecs_shape, dtype=tw, device=dQ.device)
        s = stream if stream is not None else torch.cuda.current_stream(dQ.device)
        nf = C.c_uint64(0)
        rc = capi.lib().b200_ir_interpolate_at_device(
            self._handle, dQ.data_ptr(), n, T.FLAG_NO_MOVE if do_not_move_points else 0, vals_out.data_ptr(),
            vecs_out.data_ptr(), None, s.cuda_stream, C.byref(nf) if check else None,
        )
        capi.check(rc)
        return vals_out, vecs_out

    # ------------------------------------------------------------------ introspection
    @property
    def launch_count(self):
        return int(capi.lib().b200_grid_launch_count(self._handle))

    def enable_timing(self, on=True):
        capi.check(capi.lib().b200_grid_enable_timing(self._handle, 1 if on else 0))

    def set_option(self, name, value):
        """``interp_path``: 0 auto, 1 general kernel, 2 cell-batched kernel; ``chunk``: points per CTA item."""
        capi.check(capi.lib().b200_grid_set_option(self._handle, name.encode(), float(value)))

    def kernel_ms(self, name):
        return float(capi.lib().b200_grid_kernel_ms(self._handle, name.encode()))

    @property
    def row_bytes(self):
        a, b = C.c_size_t(0), C.c_size_t(0)
        capi.check(capi.lib().b200_grid_row_bytes(self._handle, C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    @property
    def bytes_per_q(self):
        """ALGORITHMIC HBM bytes per Q: Q in + values row out + vectors row out."""
        a, b = self.row_bytes
        return 24 + a + b


class _Owned(np.ndarray):
    """ndarray subclass that keeps a pinned allocation alive."""

    def __new__(cls, arr, owner):
        obj = np.asarray(arr).view(cls)
        obj._owner = owner
        return obj

    def __array_finalize__(self, obj):
        self._owner = getattr(obj, "_owner", None)


def accelerate(host_grid, device=0):
    """Wrap an existing brille grid object."""
    return B200Grid(host_grid, device=device)
