"""ORACLE (test infrastructure, never imported by brille_b200): numpy restatement of the structure-factor consumer.

The consumer is not part of the reference tree: it is what brille's callers do with the output of ``ir_interpolate_at``
(Euphonic 1.x ``QpointPhononModes.calculate_structure_factor`` -- third party, not vendored under /root/reference, restated
here from its published algorithm:  ``F(Q,nu) = sum_k b_k/sqrt(m_k) e^{-W_k(Q)} e^{i Q.r_k} (Q . eps*_{nu,k})``,
``S = |F|^2``; brille's own ``validation/profiling.md:30-67`` times this loop through brilleu's ``s_qw``).  Parity is anchored on
the reference's own ``ir_interpolate_at`` (or the plain-C oracle of it): the reduction below is applied to ITS eigenvectors.
"""
import numpy as np


def structure_factor(Q, vecs, coef, positions=None, q_transform=None, debye_waller=None, conjugate=True):
    """``|sum_k coef_k exp(-qv^T W_k qv) exp(2 pi i Q.r_k) (qv . eps_k^[*])|^2`` -> (nQ, modes)."""
    Q = np.asarray(Q, dtype=np.float64)
    coef = np.asarray(coef, dtype=np.complex128).reshape(-1)
    n_at = coef.size
    e = np.asarray(vecs).reshape(Q.shape[0], -1, n_at, 3)
    T = np.eye(3) if q_transform is None else np.asarray(q_transform, dtype=np.float64).reshape(3, 3)
    qv = Q @ T.T
    f = np.broadcast_to(coef, (Q.shape[0], n_at)).astype(np.complex128)
    if positions is not None:
        f = f * np.exp(2j * np.pi * (Q @ np.asarray(positions, dtype=np.float64).reshape(n_at, 3).T))
    if debye_waller is not None:
        W = np.asarray(debye_waller, dtype=np.float64).reshape(n_at, 3, 3)
        f = f * np.exp(-np.einsum("qi,kij,qj->qk", qv, W, qv))
    if conjugate:
        e = np.conj(e)
    F = np.einsum("qmkc,qc,qk->qm", e, qv, f)
    return np.abs(F) ** 2


# ---------------------------------------------------------------------------------------------------------------------
# powder average (brille_b200/csrc/consumer.cu: k_powder_q, k_powder_bin), restated with numpy
# ---------------------------------------------------------------------------------------------------------------------
_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(z):
    z = np.asarray(z, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = z + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def powder_points(B, q_range, n_qbins, n_dir, seed=0, dir_range=None):
    """The sweep's points in rlu, |Q|-bin major: |Q| at the centre of bin i, direction j of the counter-based sequence
    (splitmix64 of seed, bin and j), Q = B^-1 (|Q| d).  ``B``: row-major 3x3 with x = B q."""
    lo, hi = (0, int(n_dir)) if dir_range is None else (int(dir_range[0]), int(dir_range[1]))
    i = np.repeat(np.arange(n_qbins, dtype=np.uint64), hi - lo)
    j = np.tile(np.arange(lo, hi, dtype=np.uint64), n_qbins)
    with np.errstate(over="ignore"):
        z1 = _splitmix64(np.uint64(seed) ^ _splitmix64(i * np.uint64(n_dir) + j))
    z2 = _splitmix64(z1)
    u1 = (z1 >> np.uint64(11)).astype(np.float64) * 2.0 ** -53
    u2 = (z2 >> np.uint64(11)).astype(np.float64) * 2.0 ** -53
    ct = 1.0 - 2.0 * u1
    st = np.sqrt(np.maximum(0.0, 1.0 - ct * ct))
    dq = (q_range[1] - q_range[0]) / n_qbins
    qn = q_range[0] + (i.astype(np.float64) + 0.5) * dq
    xyz = np.stack([qn * st * np.cos(2 * np.pi * u2), qn * st * np.sin(2 * np.pi * u2), qn * ct], axis=1)
    return xyz @ np.linalg.inv(np.asarray(B, dtype=np.float64).reshape(3, 3)).T


def powder_histogram(Q, B, vals, sf, q_range, n_qbins, w_range, n_wbins, weight=0):
    """(hist, counts) of the points ``Q`` (rlu) with energies ``vals`` (nQ, modes[, span]: first element) and intensities ``sf``."""
    Q = np.asarray(Q, dtype=np.float64)
    qn = np.linalg.norm(Q @ np.asarray(B, dtype=np.float64).reshape(3, 3).T, axis=1)
    w = np.asarray(vals, dtype=np.float64).reshape(Q.shape[0], np.asarray(sf).shape[1], -1)[:, :, 0]
    sf = np.asarray(sf, dtype=np.float64)
    fq = (qn - q_range[0]) * (n_qbins / (q_range[1] - q_range[0]))
    okq = (fq >= 0) & (fq < n_qbins)
    iq = np.where(okq, fq, 0).astype(np.int64)
    counts = np.bincount(iq[okq], minlength=n_qbins).astype(np.float64)
    fw = (w - w_range[0]) * (n_wbins / (w_range[1] - w_range[0]))
    ok = okq[:, None] & (fw >= 0) & (fw < n_wbins)
    v = sf.copy()
    if weight == 1:
        ok &= w > 0
        v = np.where(w > 0, sf / np.where(w > 0, w, 1.0), 0.0)
    iw = np.where(ok, fw, 0).astype(np.int64)
    hist = np.zeros((n_qbins, n_wbins))
    np.add.at(hist, (np.broadcast_to(iq[:, None], iw.shape)[ok], iw[ok]), v[ok])
    return hist, counts
