"""The pybind11 drop-in: brille's grid classes, registered as subclasses of brille's own by brille_b200._accel, behind the
unchanged names of an unchanged `brille` package.

CPU: the module loads next to brille's, the classes are subclasses with brille's constructors / properties / host methods, and
without a CUDA device the interpolation fails loudly (no CPU fallback).  GPU: brille's OWN test files (wrap/tests, copied
unmodified into brille_b200/dropin/site/reference_tests by brille_b200/accel/build_package.sh -- the GPU box has no
/root/reference) run against the drop-in package in a fresh interpreter, and the results equal the host classes'."""
import os
import subprocess
import sys

import numpy as np
import pytest

import brille_b200
from brille_b200 import workloads as W

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SITE = brille_b200.dropin_path()


@pytest.fixture(scope="module")
def accel(host):
    try:
        from brille_b200 import _accel
    except ImportError:
        pytest.skip("brille_b200._accel not built")
    return _accel


def test_classes_are_subclasses_of_brilles_own(host, accel):
    for name in accel.GRID_CLASSES:
        cls, base = getattr(accel, name), getattr(host, name)
        assert issubclass(cls, base) and cls.__name__ == name
        assert cls.ir_interpolate_at is not base.ir_interpolate_at
        for inherited in ("rlu", "invA", "tetrahedra", "BrillouinZone", "values", "vectors", "bytes_per_point"):
            assert hasattr(cls, inherited), (name, inherited)
    lat = W.p63mmc_lattice(host)
    bz = host.BrillouinZone(lat)
    g = accel.BZTrellisQdc(bz, bz.ir_polyhedron.volume / 100)
    h = host.BZTrellisQdc(bz, bz.ir_polyhedron.volume / 100)
    assert isinstance(g, host.BZTrellisQdc) and np.array_equal(g.rlu, h.rlu)
    args = W._gamma_fill(g, 12, 4, 1)
    assert g.values.shape == (g.rlu.shape[0], 12, 1) and g.vectors.shape == (g.rlu.shape[0], 12, 4, 3)
    g.sort()  # brille's host sort through the subclass
    plain = g.host()
    assert type(plain) is host.BZTrellisQdc and np.shares_memory(plain.values, g.values)
    assert type(accel.BZTrellisQdc(plain)) is accel.BZTrellisQdc
    assert accel.BZNestQdc(bz, bz.ir_polyhedron.volume / 100, 5).rlu.shape[1] == 3
    assert accel.BZMeshQdd(bz, bz.ir_polyhedron.volume / 100, 3).rlu.shape[1] == 3
    assert "Q" in g.ir_interpolate_at.__doc__ and g.gpu_launches == 0
    del args


def test_no_cpu_fallback(host, accel):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    lat = W.p63mmc_lattice(host)
    bz = host.BrillouinZone(lat)
    g = accel.BZTrellisQdc(bz, bz.ir_polyhedron.volume / 50)
    W._gamma_fill(g, 12, 4, 1)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        g.ir_interpolate_at(np.zeros((3, 3)))


def test_dropin_package_is_assembled():
    if not os.path.isdir(os.path.join(SITE, "brille")):
        pytest.skip("drop-in package not assembled (brille_b200/accel/build_package.sh)")
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r); import brille; from brille_b200 import _accel; "
            "assert brille.BZTrellisQdc is _accel.BZTrellisQdc and brille.BZNestQcc is _accel.BZNestQcc; "
            "assert hasattr(brille.BrillouinZone, 'host_ir_moveinto'); print('ok', brille.__version__)") % (ROOT, SITE)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr[-2000:]


@pytest.mark.gpu
def test_brilles_own_tests_pass_against_the_dropin():
    """wrap/tests/test_4_interpolation.py, test_5_gamma.py (the golden-file test of the Gamma rotation) and test_2_brillouinzone.py
    (isinside / moveinto on the device), unmodified, in a fresh interpreter whose `brille` is the drop-in package."""
    tests = os.path.join(SITE, "reference_tests")
    if not os.path.isdir(tests):
        pytest.skip("drop-in package not assembled")
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([SITE, ROOT, os.environ.get("PYTHONPATH", "")]))
    files = [os.path.join(tests, f) for f in ("test_4_interpolation.py", "test_5_gamma.py", "test_2_brillouinzone.py", "test_1_lattice.py", "test_6_utils.py")]
    out = subprocess.run([sys.executable, "-m", "pytest", "-q", "-p", "no:cacheprovider", *files], capture_output=True, text=True, env=env, timeout=1800,
                         cwd=tests)
    tail = out.stdout[-3000:] + out.stderr[-2000:]
    assert out.returncode == 0, tail
    assert " passed" in out.stdout and "failed" not in out.stdout, tail


@pytest.mark.gpu
def test_dropin_launches_kernels_and_equals_the_host_classes(host, accel):
    lat = W.p63mmc_lattice(host)
    bz = host.BrillouinZone(lat)
    rng = np.random.default_rng(3)
    Q = rng.uniform(-3, 3, (200_000, 3))
    for name, args in (("BZTrellisQdc", ()), ("BZNestQdc", (5,)), ("BZMeshQdc", (3,))):
        g = getattr(accel, name)(bz, bz.ir_polyhedron.volume / 300, *args)
        W._gamma_fill(g, 12, 4, 2)
        vals, vecs = g.ir_interpolate_at(Q)
        assert g.gpu_launches > 0
        assert vals.shape == (len(Q), 12, 1) and vecs.shape == (len(Q), 12, 4, 3) and vecs.dtype == np.complex128
        rv, rw = g.host().ir_interpolate_at(Q[:20000], True, 8)
        from helpers import assert_values_close

        assert_values_close(vals[:20000], rv)
        assert_values_close(vecs[:20000], rw)
        # a refill through the subclass re-uploads the data
        W._gamma_fill(g, 6, 4, 5)
        v2, w2 = g.ir_interpolate_at(Q[:1000])
        rv, rw = g.host().ir_interpolate_at(Q[:1000], False, 1)
        assert v2.shape == (1000, 6, 1)
        assert_values_close(v2, rv)
        assert_values_close(w2, rw)
        with pytest.raises(RuntimeError, match="3-vectors"):
            g.ir_interpolate_at(np.zeros((4, 2)))
    # BrillouinZone methods on the device return what brille returns: matrices, booleans
    accel.patch_brillouinzone()
    try:
        _bz_methods(bz, Q)
    finally:
        accel.unpatch_brillouinzone()  # (the other test modules use brille's own methods as the reference)


def _bz_methods(bz, Q):
    q, tau, R, invR = bz.ir_moveinto(Q[:50000])
    hq, htau, hR, hinvR = bz.host_ir_moveinto(Q[:50000])
    assert np.array_equal(q, hq) and np.array_equal(tau, htau) and np.array_equal(R, hR) and np.array_equal(invR, hinvR)
    assert R.shape == (50000, 3, 3)
    q, tau = bz.moveinto(Q[:50000])
    hq, htau = bz.host_moveinto(Q[:50000])
    assert np.array_equal(q, hq) and np.array_equal(tau, htau)
    assert np.array_equal(bz.isinside(Q[:50000] / 6), np.asarray(bz.host_isinside(Q[:50000] / 6), dtype=bool))
    qw, Rw = bz.ir_moveinto_wedge(Q[:50000])
    hqw, hRw = bz.host_ir_moveinto_wedge(Q[:50000])
    assert np.array_equal(qw, hqw) and np.array_equal(Rw, hRw)
