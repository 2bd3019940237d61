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
