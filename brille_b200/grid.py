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

    def sort(self, *args, device=True, **kwargs):
        """brille ``sort`` (wrap/_common_grid.hpp:486, interpolatordual.hpp:398-434): the equivalent-mode permutation of every
        connected vertex pair.

        ``device=True`` (default) solves the pairs on the GPU (``b200_grid_sort_pairs``) and installs the permutations in
        the device tables; brille's host object is left untouched (its own ``ir_interpolate_at`` keeps interpolating
        unsorted data).  ``device=False`` runs brille's host ``sort`` and re-uploads the tables, as do grids whose
        eigenvectors are real (not offloaded, see the header)."""
        if self._host is None:
            raise RuntimeError("sort() needs the brille host grid (this grid was built from flat tables only)")
        data = _bridge().flatten_data(self._host)
        if not device or not np.iscomplexobj(np.asarray(data["vectors_data"])):
            out = self._host.sort(*args, **kwargs)
            self._sync_data()
            return out
        self._set_data(data)  # the costs are evaluated on the data as filled
        plan = _bridge().sort_plan(self._host)
        row, col = self.sort_pairs(plan["pairs"], plan)
        self._set_data(install_permutations(data, plan["pairs"], row, col, int(plan["n_vertices"])))
        return None

    def sort_pairs(self, pairs, plan, want_cost=False):
        """Permutations (row for (i, j), col for (j, i)) of the given vertex pairs, computed on the device."""
        pairs = np.ascontiguousarray(pairs, dtype=np.uint32).reshape(-1, 2)
        cfg = capi.SortConfig()
        cfg.values_costmult[:] = [float(x) for x in plan["values_costmult"]]
        cfg.vectors_costmult[:] = [float(x) for x in plan["vectors_costmult"]]
        cfg.values_vector_cost = int(plan["values_vector_cost"])
        cfg.vectors_vector_cost = int(plan["vectors_vector_cost"])
        B = int(self._data_tables.values.branches)
        row = np.zeros((pairs.shape[0], B), dtype=np.int32)
        col = np.zeros((pairs.shape[0], B), dtype=np.int32)
        cost = np.zeros((pairs.shape[0], B, B), dtype=np.float64) if want_cost else None
        capi.check(capi.lib().b200_grid_sort_pairs(self._handle, pairs.ctypes.data, pairs.shape[0], C.byref(cfg), row.ctypes.data, col.ctypes.data,
                                                   cost.ctypes.data if want_cost else None))
        return (row, col, cost) if want_cost else (row, col)

    # ------------------------------------------------------------------ the path
    def _check_q(self, Q):
        Q = np.asarray(Q, dtype=np.float64)
        if Q.ndim != 2 or Q.shape[1] != 3:
            raise RuntimeError("Interpolation requires one or more 3-vectors")
        return np.ascontiguousarray(Q)

    def _check_filled(self):
        if self._vals_shape is None:
            raise RuntimeError("The interpolation data must be filled before interpolating.")

    def _check_device_tensor(self, t, shape, dtype, name, like):
        """caller-provided device outputs: a wrong tensor would mean out-of-bounds or strided device writes"""
        if not t.is_cuda or t.device != like.device:
            raise RuntimeError(f"{name} must be a CUDA tensor on {like.device}")
        if t.dtype != dtype or tuple(t.shape) != tuple(shape) or not t.is_contiguous():
            raise RuntimeError(f"{name} must be a contiguous {dtype} tensor of shape {tuple(shape)}")

    def _outputs(self, n, pinned):
        self._check_filled()
        vs, ws = (n,) + self._vals_shape, (n,) + self._vecs_shape
        if pinned:
            pv, pw = PinnedArray(vs, self._vals_dtype), PinnedArray(ws, self._vecs_dtype)
            return pv.array, pw.array, (pv, pw)
        return np.empty(vs, self._vals_dtype), np.empty(ws, self._vecs_dtype), None

    def _run(self, fn, Q, no_move, probe, pinned, out=None):
        Q = self._check_q(Q)
        n = Q.shape[0]
        self._check_filled()
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

    def ir_moveinto_wedge(self, Q, threads=0):
        """(q_ir, Ridx) like ``BrillouinZone.ir_moveinto_wedge`` (wrap/_bz.cpp:498-520): the rotation into the irreducible wedge of
        Q itself (no translation), with the operation INDEX instead of the matrix (Q = R[Ridx] q_ir)."""
        Q = self._check_q(Q)
        pr = T.ProbeArrays(Q.shape[0], fields=("q_ir", "tau", "ridx", "invridx", "status"))
        capi.check(capi.lib().b200_moveinto(self._handle, Q.ctypes.data, Q.shape[0], 2, pr.byref()))
        return pr.q_ir, pr.ridx

    def isinside(self, Q):
        """bool per point like ``BrillouinZone.isinside`` (wrap/_bz.cpp:378-384): inside (or on the surface of) the first zone."""
        Q = self._check_q(Q)
        pr = T.ProbeArrays(Q.shape[0], fields=("q_ir", "tau", "ridx", "invridx", "status"))
        capi.check(capi.lib().b200_moveinto(self._handle, Q.ctypes.data, Q.shape[0], 3, pr.byref()))
        return (pr.status & T.ST_OUTSIDE_BZ) == 0

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
        if dQ.device.index != self.device:
            raise RuntimeError(f"dQ is on {dQ.device}, the grid on cuda:{self.device}")
        if vals_out is None:
            vals_out = torch.empty((n,) + self._vals_shape, dtype=tv, device=dQ.device)
        else:
            self._check_device_tensor(vals_out, (n,) + self._vals_shape, tv, "vals_out", dQ)
        if vecs_out is None:
            vecs_out = torch.empty((n,) + self._vecs_shape, dtype=tw, device=dQ.device)
        else:
            self._check_device_tensor(vecs_out, (n,) + self._vecs_shape, tw, "vecs_out", dQ)
        s = stream if stream is not None else torch.cuda.current_stream(dQ.device)
        nf = C.c_uint64(0)
        rc = capi.lib().b200_ir_interpolate_at_device(
            self._handle, dQ.data_ptr(), n, T.FLAG_NO_MOVE if do_not_move_points else 0, vals_out.data_ptr(),
            vecs_out.data_ptr(), None, s.cuda_stream, C.byref(nf) if check else None,
        )
        capi.check(rc)
        return vals_out, vecs_out

    # ------------------------------------------------------------------ device-resident consumer
    def set_structure_factor(self, coef, positions=None, q_transform=None, debye_waller=None, conjugate=True):
        """Configure the one-phonon structure factor reduction ``|sum_k coef_k e^{-qv.W_k.qv} e^{2 pi i Q.r_k} (qv . eps_k^*)|^2``
        (``b200_grid_set_structure_factor``): ``coef`` complex (n_atoms,), ``positions`` fractional (n_atoms,3) or None,
        ``q_transform`` 3x3 with ``qv = q_transform @ Q`` (default identity), ``debye_waller`` (n_atoms,3,3) or None."""
        coef = np.ascontiguousarray(coef, dtype=np.complex128).reshape(-1)
        cfg = capi.SFConfig()
        cfg.n_atoms = coef.size
        keep = [coef]
        cfg.coef = coef.ctypes.data
        if positions is not None:
            pos = np.ascontiguousarray(positions, dtype=np.float64).reshape(coef.size, 3)
            keep.append(pos)
            cfg.positions = pos.ctypes.data
        if debye_waller is not None:
            dw = np.ascontiguousarray(debye_waller, dtype=np.float64).reshape(coef.size, 9)
            keep.append(dw)
            cfg.debye_waller = dw.ctypes.data
        T_ = np.eye(3) if q_transform is None else np.asarray(q_transform, dtype=np.float64).reshape(3, 3)
        cfg.q_transform[:] = [float(x) for x in T_.reshape(-1)]
        cfg.conjugate = 1 if conjugate else 0
        capi.check(capi.lib().b200_grid_set_structure_factor(self._handle, C.byref(cfg)))

    def ir_structure_factor(self, Q, do_not_move_points=False, *, pinned=False, out=None):
        """``(vals, sf)``: the eigenvalues of ``ir_interpolate_at`` and ``sf (nQ, modes)``, the structure factor of the interpolated,
        rotated eigenvectors, which stay on the device (``b200_ir_structure_factor``).  Host arrays in and out."""
        Q = self._check_q(Q)
        n = Q.shape[0]
        if self._vals_shape is None:
            raise RuntimeError("The interpolation data must be filled before interpolating.")
        M = int(self._data_tables.vectors.branches)
        if out is not None:
            vals, sf = out
            if vals.shape != (n,) + self._vals_shape or sf.shape != (n, M) or sf.dtype != np.float64 or vals.dtype != self._vals_dtype:
                raise RuntimeError("out buffers have the wrong shape or dtype")
            if not (vals.flags.c_contiguous and sf.flags.c_contiguous):
                raise RuntimeError("out buffers must be C-contiguous")
            keep = None
        elif pinned:
            pv, ps = PinnedArray((n,) + self._vals_shape, self._vals_dtype), PinnedArray((n, M), np.float64)
            vals, sf, keep = pv.array, ps.array, (pv, ps)
        else:
            vals, sf, keep = np.empty((n,) + self._vals_shape, self._vals_dtype), np.empty((n, M), np.float64), None
        capi.check(capi.lib().b200_ir_structure_factor(self._handle, Q.ctypes.data, n, T.FLAG_NO_MOVE if do_not_move_points else 0,
                                                       vals.ctypes.data, sf.ctypes.data))
        if keep is not None:
            vals, sf = _Owned(vals, keep[0]), _Owned(sf, keep[1])
        return vals, sf

    def ir_structure_factor_device(self, dQ, vals_out=None, sf_out=None, scratch=None, do_not_move_points=False, check=True, stream=None):
        """torch CUDA tensors in and out; ``scratch``: optional complex128 tensor with room for the eigenvectors of all points."""
        import torch

        if not dQ.is_cuda or dQ.dtype != torch.float64 or dQ.dim() != 2 or dQ.shape[1] != 3:
            raise RuntimeError("dQ must be a CUDA float64 tensor of shape (n,3)")
        dQ = dQ.contiguous()
        n = int(dQ.shape[0])
        if self._vals_shape is None:
            raise RuntimeError("The interpolation data must be filled before interpolating.")
        M = int(self._data_tables.vectors.branches)
        tv = torch.complex128 if self._vals_dtype == np.complex128 else torch.float64
        if dQ.device.index != self.device:
            raise RuntimeError(f"dQ is on {dQ.device}, the grid on cuda:{self.device}")
        if vals_out is None:
            vals_out = torch.empty((n,) + self._vals_shape, dtype=tv, device=dQ.device)
        else:
            self._check_device_tensor(vals_out, (n,) + self._vals_shape, tv, "vals_out", dQ)
        if sf_out is None:
            sf_out = torch.empty((n, M), dtype=torch.float64, device=dQ.device)
        else:
            self._check_device_tensor(sf_out, (n, M), torch.float64, "sf_out", dQ)
        if scratch is not None and (not scratch.is_cuda or scratch.device != dQ.device or not scratch.is_contiguous()
                                    or scratch.numel() * scratch.element_size() < n * self.row_bytes[1]):
            raise RuntimeError("scratch is too small for the eigenvectors of all points (or not a contiguous CUDA tensor on the grid's device)")
        s = stream if stream is not None else torch.cuda.current_stream(dQ.device)
        nf = C.c_uint64(0)
        capi.check(capi.lib().b200_ir_structure_factor_device(
            self._handle, dQ.data_ptr(), n, T.FLAG_NO_MOVE if do_not_move_points else 0, vals_out.data_ptr(), sf_out.data_ptr(),
            scratch.data_ptr() if scratch is not None else None, s.cuda_stream, C.byref(nf) if check else None))
        return vals_out, sf_out

    # ------------------------------------------------------------------ device-resident consumer: powder average
    @staticmethod
    def _powder_cfg(q_range, n_qbins, w_range, n_wbins, weight):
        cfg = capi.PowderConfig()
        cfg.n_qbins, cfg.n_wbins = int(n_qbins), int(n_wbins)
        cfg.q_lo, cfg.q_hi = float(q_range[0]), float(q_range[1])
        cfg.w_lo, cfg.w_hi = float(w_range[0]), float(w_range[1])
        cfg.weight = int(weight)
        return cfg

    def ir_powder_bin(self, Q, q_range, n_qbins, w_range, n_wbins, weight=0, do_not_move_points=False, out=None):
        """``(hist, counts)``: the structure factor |F(Q, mode)|^2 of every point binned on (|B Q|, eigenvalue) on the device
        (``b200_ir_powder_bin``; ``set_structure_factor`` first).  ``hist`` (n_qbins, n_wbins), ``counts`` (n_qbins,) points per
        |Q| bin; ``out=(hist, counts)`` accumulates into existing arrays.  Only the histogram leaves the GPU."""
        Q = self._check_q(Q)
        self._check_filled()
        cfg = self._powder_cfg(q_range, n_qbins, w_range, n_wbins, weight)
        hist, counts = out if out is not None else (np.zeros((cfg.n_qbins, cfg.n_wbins)), np.zeros(cfg.n_qbins))
        if hist.shape != (cfg.n_qbins, cfg.n_wbins) or counts.shape != (cfg.n_qbins,) or hist.dtype != np.float64 or counts.dtype != np.float64 \
                or not (hist.flags.c_contiguous and counts.flags.c_contiguous):
            raise RuntimeError("out must be C-contiguous float64 arrays of shapes (n_qbins, n_wbins) and (n_qbins,)")
        capi.check(capi.lib().b200_ir_powder_bin(self._handle, Q.ctypes.data, Q.shape[0], T.FLAG_NO_MOVE if do_not_move_points else 0, C.byref(cfg),
                                                 hist.ctypes.data, counts.ctypes.data))
        return hist, counts

    def ir_powder_sweep(self, q_range, n_qbins, w_range, n_wbins, n_dir, seed=0, weight=0, dir_range=None, out=None):
        """The powder sweep with the points generated on the device: ``n_dir`` isotropic directions at the centre of every |Q| bin
        (``dir_range=(lo, hi)``: only that slice of the direction sequence -- ranks / GPUs take disjoint slices and add their
        histograms).  Returns ``(hist, counts)``; nothing per Q crosses PCIe."""
        self._check_filled()
        cfg = self._powder_cfg(q_range, n_qbins, w_range, n_wbins, weight)
        lo, hi = (0, int(n_dir)) if dir_range is None else (int(dir_range[0]), int(dir_range[1]))
        hist, counts = out if out is not None else (np.zeros((cfg.n_qbins, cfg.n_wbins)), np.zeros(cfg.n_qbins))
        capi.check(capi.lib().b200_ir_powder_sweep(self._handle, C.byref(cfg), int(n_dir), int(seed), lo, hi, hist.ctypes.data, counts.ctypes.data))
        return hist, counts

    def powder_points(self, q_range, n_qbins, n_dir, seed=0, dir_range=None):
        """The points (rlu) of the sweep, |Q|-bin major: (n_qbins * (hi - lo), 3)."""
        self._check_filled()
        cfg = self._powder_cfg(q_range, n_qbins, (0.0, 1.0), 1, 0)
        lo, hi = (0, int(n_dir)) if dir_range is None else (int(dir_range[0]), int(dir_range[1]))
        Q = np.zeros((cfg.n_qbins * (hi - lo), 3))
        capi.check(capi.lib().b200_powder_points(self._handle, C.byref(cfg), int(n_dir), int(seed), lo, hi, Q.ctypes.data))
        return Q

    # ------------------------------------------------------------------ introspection
    @property
    def launch_count(self):
        return int(capi.lib().b200_grid_launch_count(self._handle))

    @property
    def last_path(self):
        """B200_PATH_* bits of the kernels the last interpolation call took (1 two-kernel location, 2 on-the-fly cell kernel,
        4 pipelined cell kernel, 8 general kernel for all points, 16 fused structure factor, 32 cooperative second location kernel)."""
        return int(capi.lib().b200_grid_last_path(self._handle))

    @property
    def hot_path_taken(self):
        """True when the last call ran the kernels the bench runs: two-kernel location + pipelined cell kernel."""
        return (self.last_path & 5) == 5

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


def install_permutations(data, pairs, row, col, n_vertices):
    """A copy of the bridge data dictionary whose permutation tables hold ``row``/``col`` of the vertex ``pairs``.

    Table row 0 is the identity (pairs that are not connected, i == j), row 1 + 2k the permutation of (i, j) = pairs[k] and
    row 2 + 2k the one of (j, i) -- the lookup of PermutationTable::safe_get (permutation_table.hpp:193-198,214) without
    its de-duplication of equal permutations."""
    pairs = np.ascontiguousarray(pairs, dtype=np.uint64).reshape(-1, 2)
    B = row.shape[1]
    rows = np.empty((1 + 2 * pairs.shape[0], B), dtype=np.uint32)
    rows[0] = np.arange(B, dtype=np.uint32)
    rows[1::2] = row
    rows[2::2] = col
    keys = pairs[:, 0] * np.uint64(n_vertices) + pairs[:, 1]
    order = np.argsort(keys, kind="stable")
    skeys = keys[order]

    def table(cells):
        cells = np.ascontiguousarray(cells, dtype=np.uint64)
        if cells.size == 0:
            return np.zeros((0, cells.shape[1] ** 2 if cells.ndim == 2 else 0), dtype=np.uint32)
        a = cells[:, :, None]
        b = cells[:, None, :]
        lo, hi = np.minimum(a, b), np.maximum(a, b)
        key = lo * np.uint64(n_vertices) + hi
        pos = np.searchsorted(skeys, key)
        pos = np.minimum(pos, max(len(skeys) - 1, 0))
        found = (len(skeys) > 0) & (skeys[pos] == key) & (a != b) if len(skeys) else np.zeros(key.shape, dtype=bool)
        k = order[pos] if len(skeys) else np.zeros(key.shape, dtype=np.int64)
        idx = np.where(found, 1 + 2 * k + (a > b), 0)
        return idx.reshape(cells.shape[0], -1).astype(np.uint32)

    d = dict(data)
    d["perm_rows"] = rows
    d["perm_nonidentity"] = 1
    d["cube_perm"] = table(np.asarray(data["perm_cube_vertices"]).reshape(-1, 8)) if np.asarray(data["perm_cube_vertices"]).size else np.zeros((0, 64), dtype=np.uint32)
    d["tet_perm"] = table(np.asarray(data["perm_tet_vertices"]).reshape(-1, 4))
    return d


def accelerate(host_grid, device=0):
    """Wrap an existing brille grid object."""
    return B200Grid(host_grid, device=device)
