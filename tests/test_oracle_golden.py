"""CPU: the plain-C oracle against the committed golden fixtures (outputs of the reference itself, and the
reference's own golden file wrap/tests/test_5_gamma.npz)."""
import numpy as np
import pytest

from oracle.oracle import Oracle
from helpers import RTOL, assert_decisions_equal, assert_values_close, load_golden, ref_decisions, remove_mode_phase

FIXTURES = ["nacl_prim_trellis.npz", "nacl_prim_trellis_sorted.npz", "fd3m_scalar_trellis.npz", "p63mmc_trellis.npz", "p1_trellis_dd.npz",
            "p63mmc_nest.npz", "p63mmc_nest_sorted.npz", "p63mmc_mesh.npz"]


@pytest.mark.parametrize("name", FIXTURES)
def test_oracle_matches_reference_outputs(name):
    s, d, _, rest = load_golden(name)
    orc = Oracle(s, d)
    rc, vals, vecs, pr = orc.interpolate_at(rest["Q"])
    assert rc == 0
    tet_grid = str(s["kind"]) in ("nest", "mesh")
    assert_decisions_equal(pr, ref_decisions(rest), adaptive_ulps=4 if tet_grid else 0)
    # the oracle keeps the reference's operation order: results are bit-identical (for points off the cell faces)
    ok = np.ones(len(rest["Q"]), bool) if not tet_grid else np.asarray(rest["ref_n_vert"]) == 4
    assert np.array_equal(vals[ok], rest["ref_values"].reshape(vals.shape)[ok])
    assert np.array_equal(vecs[ok], rest["ref_vectors"].reshape(vecs.shape)[ok])
    assert_values_close(vals, rest["ref_values"])
    assert_values_close(vecs, rest["ref_vectors"])


def test_oracle_interpolate_at_no_rotation():
    s, d, _, rest = load_golden("p1_trellis_dd.npz")
    orc = Oracle(s, d)
    rc, vals, vecs, pr = orc.interpolate_at(rest["Q"], ir=False)
    assert rc == 0
    assert np.array_equal(pr.tau, rest["ref0_tau"])
    assert np.array_equal(vals, rest["ref0_values"].reshape(vals.shape))
    assert np.array_equal(vecs, rest["ref0_vectors"].reshape(vecs.shape))


def test_sorted_fixture_has_nonidentity_permutations():
    s, d, _, rest = load_golden("nacl_prim_trellis_sorted.npz")
    assert int(d["perm_nonidentity"]) == 1 and np.asarray(d["perm_rows"]).shape[0] > 1


def test_nacl_gamma_reference_golden_vectors():
    """wrap/tests/test_5_gamma.py:95-175 restated on the oracle: eigenvalues allclose to Euphonic's, eigenvectors equal up
    to a per-(Q, mode) phase; additionally identical to what the reference build returns today."""
    s, d, d2, rest = load_golden("nacl_gamma.npz")
    orc = Oracle(s, d)
    rc, vals, vecs, pr = orc.interpolate_at(rest["Q"])
    assert rc == 0
    assert np.array_equal(vals, rest["ref_values"].reshape(vals.shape))
    assert np.array_equal(vecs, rest["ref_vectors"].reshape(vecs.shape))
    assert np.allclose(vals.reshape(48, 24), rest["golden_euphonic_values"])
    br_vec = np.einsum("ba,ijkb->ijka", rest["golden_basis_vectors"], vecs.reshape(48, 24, 8, 3))
    eu = rest["golden_euphonic_vectors"]
    assert np.allclose(remove_mode_phase(br_vec, eu), eu)
    # stored brille v0.5 output, also only defined up to the per-mode phase
    old = rest["golden_brille_vectors"]
    assert np.allclose(remove_mode_phase(vecs.reshape(48, 24, 8, 3), old), old)
    assert_values_close(vals.reshape(48, 24), rest["golden_brille_values"], 1e-9)


def test_nacl_gamma_cartesian_eigenvectors():
    """test_5_gamma.py:180-221: LengthUnit::angstrom Gamma rotation with Cartesian rotation matrices."""
    s, d, d2, rest = load_golden("nacl_gamma.npz")
    orc = Oracle(s, d2)
    rc, vals, vecs, pr = orc.interpolate_at(rest["Q"])
    assert rc == 0
    assert np.array_equal(vecs, rest["ref2_vectors"].reshape(vecs.shape))
    eu = rest["golden_euphonic_vectors"]
    assert np.allclose(remove_mode_phase(vecs.reshape(48, 24, 8, 3), eu), eu)


def test_unsupported_rotateslike_lengthunit_combinations_error():
    """test_5_gamma.py:223-240"""
    s, d, d2, rest = load_golden("nacl_gamma.npz")
    for rl, lu in [(0, 1), (1, 0), (2, 2), (2, 4)]:
        dd = dict(d)
        dd["vectors_rotlike"], dd["vectors_lenunit"] = rl, lu
        rc, *_ = Oracle(s, dd).interpolate_at(rest["Q"])
        assert rc == -6
