"""``brille._brille`` of the drop-in package: brille's own compiled module with the grid classes on the GPU.

brille's Python package (``brille/__init__.py``, ``brille/bound.py``) imports every public name from ``brille._brille``.
This file takes the place of that compiled module inside the package directory ``brille_b200/accel/build_package.sh``
assembles (brille's unmodified ``*.py`` next to a copy of this file): it re-exports brille's host module unchanged, except
that the nine grid classes ``BZ{Trellis,Nest,Mesh}Q{dd,dc,cc}`` are the subclasses registered by ``brille_b200._accel``
(same names, same constructors, everything inherited; ``ir_interpolate_at`` / ``interpolate_at`` run on the GPU) and that
``BrillouinZone.isinside / moveinto / ir_moveinto / ir_moveinto_wedge`` run the device kernel.  ``import brille`` then
behaves as before -- brille's own tests run against it unmodified (tests/test_dropin.py).
"""
from brille_b200 import host as _host

_h = _host.get()
from brille_b200 import _accel  # noqa: E402  (brille's module must be loaded first: its classes are the bases)

for _n in dir(_h):
    if not _n.startswith("__") or _n in ("__version__",):
        globals()[_n] = getattr(_h, _n)
for _n in _accel.GRID_CLASSES:
    globals()[_n] = getattr(_accel, _n)
_accel.patch_brillouinzone()
del _n
