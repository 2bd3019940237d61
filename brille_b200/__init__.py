"""brille_b200 -- B200-native batched Q-point interpolation for brille grids.

One hot path of brille (``BZ{Trellis,Nest,Mesh}Q*.ir_interpolate_at``) re-built as hand-written sm_100a CUDA
kernels behind a C ABI (``include/brille_b200.h``).  Construction stays brille's host C++.

>>> import brille_b200
>>> grid = brille_b200.BZTrellisQdc(bz, max_volume)          # brille constructs, the tables go to the GPU
>>> grid.fill(vals, vals_elements, vecs, vecs_elements)
>>> vals, vecs = grid.ir_interpolate_at(Q)                   # same signature and results as brille
"""
from __future__ import annotations

from . import tables  # noqa: F401
from .capi import B200Error  # noqa: F401
from .grid import B200Grid, PinnedArray, accelerate  # noqa: F401
from .sharding import ShardedGrid  # noqa: F401

__version__ = "0.1.0"


def _factory(name):
    def make(bz, *args, device=0, **kwargs):
        from . import host

        return B200Grid(getattr(host.get(), name)(bz, *args, **kwargs), device=device)

    make.__name__ = name
    make.__doc__ = f"Construct brille's ``{name}`` on the host and move its interpolation path to the GPU."
    return make


for _n in ("BZTrellisQdd", "BZTrellisQdc", "BZTrellisQcc", "BZNestQdd", "BZNestQdc", "BZNestQcc", "BZMeshQdd", "BZMeshQdc", "BZMeshQcc"):
    globals()[_n] = _factory(_n)
del _n


def install(module=None):
    """Bind the nine grid class names of an imported ``brille`` package (default: ``import brille``) to the GPU subclasses of
    ``brille_b200._accel`` and route ``BrillouinZone.isinside / moveinto / ir_moveinto / ir_moveinto_wedge`` to the device --
    what the drop-in package ``brille_b200/dropin`` does at import time, for a brille that is already installed.  User code keeps
    calling ``brille.BZTrellisQdc(bz, ...)``, ``grid.fill(...)``, ``grid.ir_interpolate_at(Q)``; returns the module."""
    import importlib

    from . import host

    host.get()
    from . import _accel

    module = module or importlib.import_module("brille")
    targets = [module] + [getattr(module, n) for n in ("bound", "_brille") if hasattr(module, n)]
    for t in targets:
        for name in _accel.GRID_CLASSES:
            if hasattr(t, name):
                setattr(t, name, getattr(_accel, name))
        if hasattr(t, "__grid_types__"):
            t.__grid_types__ = tuple(getattr(_accel, n) for n in _accel.GRID_CLASSES)
    _accel.patch_brillouinzone()
    return module


def dropin_path():
    """Directory to put on ``sys.path`` (before any installed brille) for ``import brille`` to be the accelerated package
    assembled by ``brille_b200/accel/build_package.sh``."""
    import os

    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "dropin", "site")
