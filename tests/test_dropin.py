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
    import torch

    if not torch.cuda.is_available():  # sort() of complex eigenvectors is a device call like the interpolation
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            g.sort()
    d = accel.BZTrellisQdd(bz, bz.ir_polyhedron.volume / 100)
    d.fill(np.ones((d.rlu.shape[0], 2, 1)), (1,), np.ones((d.rlu.shape[0], 2, 3)), (0, 3, 0, 0, 3), True)
    d.sort()  # real eigenvectors: brille's host sort through the subclass
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


@pytest.mark.gpu
@pytest.mark.parametrize("name,args", [("BZTrellisQdc", ()), ("BZNestQdc", (5,)), ("BZMeshQdc", (3,)), ("BZTrellisQcc", ())])
def test_dropin_sort_runs_on_the_device(host, accel, bridge, name, args):
    """grid.sort(), fill(..., sort=True) and set_flags_weights(..., sort=True) of the drop-in classes: the cost matrices and
    assignments come from the device (b200_grid_sort_pairs), brille's own PermutationTable holds the result.  Against brille's
    host sort() on a copy of the same object: the same permutation for every vertex pair but the few whose assignment hangs
    on the last bit of a cost (tests/test_gpu_parity.py::test_device_sort_matches_reference_and_oracle), and the table the
    device sort leaves behind drives brille's HOST interpolation to the device's results."""
    from helpers import assert_values_close

    lat = W.p63mmc_lattice(host)
    bz = host.BrillouinZone(lat)
    cls = getattr(accel, name)
    g = cls(bz, bz.ir_polyhedron.volume / 300, *args)
    nv = g.rlu.shape[0]
    rng = np.random.default_rng(8)
    vals = rng.uniform(1.0, 50.0, (nv, 12, 1))
    if name.endswith("cc"):
        vals = vals + 1j * rng.uniform(0.0, 1.0, vals.shape)
    vecs = rng.normal(size=(nv, 12, 4, 3)) + 1j * rng.normal(size=(nv, 12, 4, 3))
    fill = (vals, (1, 0, 0, 0, 3), vecs, (0, 12, 0, 2, 3))
    g.fill(*fill)
    ref = g.host()          # a plain brille object with its own copy of the (still unsorted) permutation table
    before = g.gpu_launches
    g.sort()
    assert g.gpu_launches > before
    ref.sort()              # brille's OpenMP sort
    plan = bridge.sort_plan(ref)
    mine, theirs = bridge.pair_permutations(g, plan["pairs"]), bridge.pair_permutations(ref, plan["pairs"])
    assert (mine != np.arange(12)).any(), "the data must need sorting"
    differ = np.flatnonzero((mine != theirs).any(axis=(1, 2)))
    assert len(differ) <= max(1, len(mine) // 200), f"{len(differ)} of {len(mine)} pairs differ"
    Q = rng.uniform(-3, 3, (50_000, 3))
    v, w = g.ir_interpolate_at(Q)
    hv, hw = g.host().ir_interpolate_at(Q[:5000], False, 1)   # brille's CPU path reading the table the device sort wrote
    assert_values_close(v[:5000], hv)
    assert_values_close(w[:5000], hw)
    if len(differ) == 0:
        rv, rw = ref.ir_interpolate_at(Q[:5000], False, 1)
        assert_values_close(v[:5000], rv)
        assert_values_close(w[:5000], rw)
    # the `sort` argument of fill (positional and keyword) and of set_flags_weights
    for how in ("positional", "keyword", "flags"):
        g2 = cls(bz, bz.ir_polyhedron.volume / 300, *args)
        if how == "positional":
            g2.fill(*fill, True)
        elif how == "keyword":
            g2.fill(*fill, sort=True)
        else:
            g2.fill(*fill)
            g2.set_flags_weights([0, 3, 0, 0], [1.0, 1.0, 1.0], [2, 3, 0, 0], [1.0, 1.0, 1.0], sort=True)
        assert g2.gpu_launches > 0
        assert np.array_equal(bridge.pair_permutations(g2, plan["pairs"]), mine), how
    # a second sort() of sorted data changes nothing (the costs are those of the data as filled)
    g.sort()
    assert np.array_equal(bridge.pair_permutations(g, plan["pairs"]), mine)


@pytest.mark.gpu
def test_dropin_consumers_equal_the_ctypes_mirror(host, accel, bridge):
    """set_structure_factor / ir_structure_factor / ir_powder_bin / ir_powder_sweep on the drop-in classes are the same C-ABI
    calls as on the ctypes mirror (whose results tests/test_consumer.py checks against numpy on the reference's eigenvectors):
    same bits."""
    lat = W.p63mmc_lattice(host)
    bz = host.BrillouinZone(lat)
    g = accel.BZTrellisQdc(bz, bz.ir_polyhedron.volume / 300)
    W._gamma_fill(g, 12, 4, 2)
    m = brille_b200.accelerate(g.host())
    rng = np.random.default_rng(4)
    cfg = dict(coef=rng.normal(size=4) + 1j * rng.normal(size=4), positions=rng.uniform(0, 1, (4, 3)),
               q_transform=np.asarray(bridge.flatten_bz(bz)["to_xyz"]).reshape(3, 3), debye_waller=None, conjugate=True)
    with pytest.raises(RuntimeError, match="set_structure_factor"):
        g.ir_structure_factor(np.zeros((2, 3)))
    g.set_structure_factor(**cfg)
    m.set_structure_factor(**cfg)
    Q = rng.uniform(-3, 3, (100_000, 3))
    v, sf = g.ir_structure_factor(Q)
    mv, msf = m.ir_structure_factor(Q)
    assert sf.shape == (len(Q), 12) and np.array_equal(v, mv) and np.array_equal(sf, msf)
    qr, nqb, wr, nwb = (0.2, 6.2), 24, (0.0, 52.0), 40
    h, c = g.ir_powder_sweep(qr, nqb, wr, nwb, 2000, seed=3, weight=1)
    mh, mc = m.ir_powder_sweep(qr, nqb, wr, nwb, 2000, seed=3, weight=1)
    assert h.shape == (nqb, nwb) and c.sum() == nqb * 2000 and np.array_equal(c, mc)
    assert np.abs(h - mh).max() <= 1e-12 * mh.max()       # FP64 atomics: the order of the additions is not fixed
    ha, ca = g.ir_powder_sweep(qr, nqb, wr, nwb, 2000, seed=3, weight=1, dir_range=(0, 700))
    hb, cb = g.ir_powder_sweep(qr, nqb, wr, nwb, 2000, seed=3, weight=1, dir_range=(700, 2000))
    assert np.array_equal(ca + cb, c) and np.abs(ha + hb - h).max() <= 1e-12 * h.max()
    Qp = m.powder_points(qr, nqb, 2000, seed=3)
    hp, cp = g.ir_powder_bin(Qp, qr, nqb, wr, nwb, weight=1)
    assert np.array_equal(cp, c) and np.abs(hp - h).max() <= 1e-12 * h.max()
    # a refill keeps the configuration (it belongs to the grid, not to the data)
    W._gamma_fill(g, 12, 4, 9)
    v2, sf2 = g.ir_structure_factor(Q[:1000])
    assert not np.array_equal(sf2, sf[:1000])
    m.close()


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
