"""Device-resident consumer: the structure-factor reduction behind ir_interpolate_at.

CPU part: the numpy oracle of the reduction against a literal loop.  GPU part: b200_ir_structure_factor[_device] against the
oracle applied to the REFERENCE's eigenvectors (golden fixtures / the reference run on the spot)."""
import numpy as np
import pytest

from helpers import RTOL, assert_values_close, load_golden, rel_err
from oracle.consumer import structure_factor


def sf_config(n_atoms, seed, cartesian=False, dw=True, pos=True):
    rng = np.random.default_rng(seed)
    cfg = {"coef": rng.normal(size=n_atoms) + 1j * rng.normal(size=n_atoms)}
    cfg["positions"] = rng.uniform(0, 1, (n_atoms, 3)) if pos else None
    cfg["q_transform"] = rng.normal(size=(3, 3)) if cartesian else None
    if dw:
        a = rng.normal(size=(n_atoms, 3, 3)) * 0.05
        cfg["debye_waller"] = a @ a.transpose(0, 2, 1)
    else:
        cfg["debye_waller"] = None
    return cfg


def test_oracle_consumer_against_literal_loop():
    rng = np.random.default_rng(0)
    nq, M, K = 7, 5, 3
    Q = rng.uniform(-2, 2, (nq, 3))
    vecs = rng.normal(size=(nq, M, K, 3)) + 1j * rng.normal(size=(nq, M, K, 3))
    for conj in (True, False):
        cfg = sf_config(K, 4, cartesian=True)
        got = structure_factor(Q, vecs, conjugate=conj, **cfg)
        want = np.zeros((nq, M))
        for i in range(nq):
            qv = cfg["q_transform"] @ Q[i]
            for m in range(M):
                F = 0j
                for k in range(K):
                    e = np.conj(vecs[i, m, k]) if conj else vecs[i, m, k]
                    F += cfg["coef"][k] * np.exp(-qv @ cfg["debye_waller"][k] @ qv) * np.exp(2j * np.pi * Q[i] @ cfg["positions"][k]) * (qv @ e)
                want[i, m] = abs(F) ** 2
        assert np.allclose(got, want, rtol=1e-13, atol=0)
    # without the optional factors: |sum_k c_k Q.eps_k|^2
    got = structure_factor(Q, vecs, np.ones(K), conjugate=False)
    assert np.allclose(got, np.abs(np.einsum("qmkc,qc->qm", vecs, Q)) ** 2, rtol=1e-13)


@pytest.mark.gpu
@pytest.mark.parametrize("name,n_atoms", [("p63mmc_trellis.npz", 4), ("nacl_prim_trellis.npz", 2), ("nacl_prim_trellis_sorted.npz", 2),
                                          ("p63mmc_nest.npz", 4), ("p63mmc_mesh.npz", 4), ("nacl_gamma.npz", 8)])
@pytest.mark.parametrize("variant", ["full", "plain"])
def test_structure_factor_on_reference_eigenvectors(name, n_atoms, variant):
    import brille_b200

    s, d, _, rest = load_golden(name)
    g = brille_b200.B200Grid(None, structure=s, data=d)
    cfg = sf_config(n_atoms, 7, cartesian=variant == "full", dw=variant == "full", pos=variant == "full")
    conj = variant == "full"
    g.set_structure_factor(conjugate=conj, **cfg)
    Q = rest["Q"]
    vals, sf = g.ir_structure_factor(Q)
    assert_values_close(vals, rest["ref_values"])
    want = structure_factor(Q, rest["ref_vectors"], conjugate=conj, **cfg)
    assert sf.shape == want.shape
    assert_values_close(sf, want)
    g.close()


@pytest.mark.gpu
def test_structure_factor_device_chunks_and_errors(host):
    import torch

    import brille_b200
    from brille_b200 import workloads as W

    wl = W.BUILDERS["C3"](host)
    g = brille_b200.accelerate(wl.grid)
    Q = wl.make_q(300000, 21)
    with pytest.raises(RuntimeError, match="set_structure_factor"):
        g.ir_structure_factor(Q[:10])
    cfg = sf_config(wl.n_atoms, 3, cartesian=True)
    g.set_structure_factor(**cfg)
    vals, vecs = g.ir_interpolate_at(Q)
    want = structure_factor(Q, vecs, **cfg)
    m = 20000
    rv, rw = wl.grid.ir_interpolate_at(Q[:m], True, 8)  # the reference itself
    want_ref = structure_factor(Q[:m], rw, **cfg)
    # fused into the pipelined cell kernel (the default) ...
    v1, sf1 = g.ir_structure_factor(Q)
    assert np.array_equal(v1, vals)
    assert_values_close(sf1, want)
    assert_values_close(sf1[:m], want_ref)
    # ... and reduced from the eigenvector scratch by the separate kernel: the same numbers to rounding
    g.set_option("sf_fused", 0)
    v0, sf0 = g.ir_structure_factor(Q)
    assert np.array_equal(v0, vals)
    assert_values_close(sf0, want)
    assert_values_close(sf0, sf1, rtol=1e-12)
    for fused, ref_sf in ((1, sf1), (0, sf0)):
        g.set_option("sf_fused", fused)
        # the chunking of the host pipeline is invisible
        g.set_option("host_chunk", 70001)
        v2, sf2 = g.ir_structure_factor(Q, pinned=True)
        assert np.array_equal(sf2, ref_sf) and np.array_equal(v2, v1)
        g.set_option("host_chunk", 0)
        # device buffers with the library's own scratch
        dQ = torch.from_numpy(Q).cuda()
        dv, dsf = g.ir_structure_factor_device(dQ)
        assert np.array_equal(dsf.cpu().numpy(), ref_sf) and np.array_equal(dv.cpu().numpy(), vals)
    # caller scratch: never fused, the scratch holds the eigenvectors of the call
    g.set_option("sf_fused", 1)
    scratch = torch.empty((Q.shape[0], wl.modes, wl.n_atoms, 3), dtype=torch.complex128, device="cuda")
    dv, dsf = g.ir_structure_factor_device(dQ, scratch=scratch)
    assert np.array_equal(dsf.cpu().numpy(), sf0)
    assert np.array_equal(scratch.cpu().numpy().reshape(vecs.shape), vecs)
    with pytest.raises(RuntimeError, match="too small"):
        g.ir_structure_factor_device(dQ, scratch=scratch[:10])
    # a call too small for the cell-batched kernels takes the general kernel + the separate reduction
    vs, sfs = g.ir_structure_factor(Q[:500])
    assert_values_close(sfs, want[:500])
    # a Q outside the gridded zone fails the whole call like ir_interpolate_at
    with pytest.raises(RuntimeError):
        g.ir_structure_factor(np.full((3, 3), 7.3), do_not_move_points=True)
    Qbad = Q.copy()
    Qbad[123456] = 7.3
    with pytest.raises(RuntimeError):
        g.ir_structure_factor(Qbad, do_not_move_points=True)
    # wrong atom count / data that is not made of per-atom 3-vectors
    g.set_structure_factor(np.ones(3))
    with pytest.raises(RuntimeError, match="3-vectors"):
        g.ir_structure_factor(Q[:10])
    g.close()
    s, d, _, rest = load_golden("fd3m_scalar_trellis.npz")
    g = brille_b200.B200Grid(None, structure=s, data=d)
    g.set_structure_factor(np.ones(1))
    with pytest.raises(RuntimeError, match="3-vectors"):
        g.ir_structure_factor(rest["Q"])
    g.close()


@pytest.mark.gpu
def test_structure_factor_fused_on_degenerate_point_sets(host):
    """Points on cell faces / edges / vertices do not take the cell kernel: their eigenvectors go to the compact scratch and are
    reduced by the list mode of the consumer kernel; when there are more of them than compact rows the call falls back to the
    unfused reduction.  Powder Q (large tau), sorted data (permutations) and a Nest grid run fused as well."""
    import brille_b200
    from brille_b200 import workloads as W

    wl = W.c3_p63mmc(host, density=300, seed=5)
    wl.grid.sort()
    g = brille_b200.accelerate(wl.grid)
    cfg = sf_config(wl.n_atoms, 11, cartesian=True)
    g.set_structure_factor(**cfg)
    rng = np.random.default_rng(5)
    line = np.outer(np.linspace(-2.0, 2.0, 200_001), np.array([1.0, 0.0, 0.0]))  # node faces and edges all along
    rational = rng.integers(-8, 9, (100_000, 3)) / 8.0                           # many exactly on cell faces
    mixed = np.vstack([wl.make_q(150_000, 4), rational[:5000]])                  # a few general-kernel points among many
    one = np.tile(np.array([[0.137, 0.211, 0.303]]), (100_000, 1))
    for name, Q in (("line", line), ("rational", rational), ("mixed", mixed), ("identical", one)):
        vals, vecs = g.ir_interpolate_at(Q)
        want = structure_factor(Q, vecs, **cfg)
        v1, sf1 = g.ir_structure_factor(Q)
        assert np.array_equal(v1, vals), name
        assert_values_close(sf1, want)
        import torch

        dv, dsf = g.ir_structure_factor_device(torch.from_numpy(Q).cuda())
        # (bit-identical unless one of the two calls had to fall back to the unfused reduction: the host pipeline sizes its
        # compact scratch per chunk, and the general-kernel points of "mixed" all sit in the last chunk)
        assert rel_err(dsf.cpu().numpy(), sf1) <= 1e-12, name
        if name != "mixed":
            assert np.array_equal(dsf.cpu().numpy(), sf1), name
    g.close()
    lat = W.p63mmc_lattice(host)
    bz = host.BrillouinZone(lat)
    hg = host.BZNestQdc(bz, bz.ir_polyhedron.volume / 500, 5)
    W._gamma_fill(hg, 12, 4, 17)
    g = brille_b200.accelerate(hg)
    g.set_structure_factor(**cfg)
    Q = rng.uniform(-3, 3, (100_000, 3))
    vals, vecs = g.ir_interpolate_at(Q)
    v1, sf1 = g.ir_structure_factor(Q)
    assert_values_close(sf1, structure_factor(Q, vecs, **cfg))
    g.set_option("sf_fused", 0)
    v0, sf0 = g.ir_structure_factor(Q)
    assert_values_close(sf1, sf0, rtol=1e-12)
    g.close()


@pytest.mark.gpu
def test_structure_factor_c4_72_modes(host):
    """BASELINE config 4: 24 atoms (not a power of two: the lane groups of the fused finish are padded to 32), 72 modes staged in
    several passes by the cell kernel, Nest grid; fused and unfused against the reference's eigenvectors."""
    import brille_b200
    from brille_b200 import workloads as W

    wl = W.c4_p21c_nest(host, density=300)
    g = brille_b200.accelerate(wl.grid)
    cfg = sf_config(wl.n_atoms, 13, cartesian=True)
    g.set_structure_factor(**cfg)
    Q = wl.make_q(40000, 21)
    vals, vecs = g.ir_interpolate_at(Q)
    want = structure_factor(Q, vecs, **cfg)
    v1, sf1 = g.ir_structure_factor(Q)
    assert np.array_equal(v1, vals)
    assert_values_close(sf1, want)
    rv, rw = wl.grid.ir_interpolate_at(Q[:3000], True, 8)
    assert_values_close(sf1[:3000], structure_factor(Q[:3000], rw, **cfg))
    g.set_option("sf_fused", 0)
    v0, sf0 = g.ir_structure_factor(Q)
    assert_values_close(sf0, want)
    assert_values_close(sf1, sf0, rtol=1e-12)
    assert not np.array_equal(sf1, sf0)  # (two different kernels did run)
    g.close()


@pytest.mark.gpu
def test_structure_factor_sharded(host):
    """Q cut into contiguous row blocks over several handles (here: twice the one device), each shard reduced where it was computed."""
    import brille_b200
    from brille_b200 import workloads as W
    from brille_b200.sharding import ShardedGrid

    wl = W.c3_p63mmc(host, density=500)
    cfg = sf_config(wl.n_atoms, 5, cartesian=True)
    Q = wl.make_q(250_001, 2)
    g = brille_b200.accelerate(wl.grid)
    g.set_structure_factor(**cfg)
    vals, sf = g.ir_structure_factor(Q)
    g.close()
    sg = ShardedGrid(wl.grid, [0, 0])
    sg.set_structure_factor(**cfg)
    sv, ssf = sg.ir_structure_factor(Q)
    assert np.array_equal(sv, vals) and np.array_equal(ssf, sf)
    with pytest.raises(RuntimeError):
        sg.ir_structure_factor(np.full((4, 3), 7.3), do_not_move_points=True)
    sg.close()


# ---------------------------------------------------------------------------------------------------------------------
# powder average (b200_ir_powder_bin / b200_ir_powder_sweep)
# ---------------------------------------------------------------------------------------------------------------------
def test_oracle_powder_histogram_against_literal_loop():
    from oracle.consumer import powder_histogram, powder_points

    rng = np.random.default_rng(2)
    B = rng.normal(size=(3, 3)) + 2 * np.eye(3)
    Q = powder_points(B, (0.5, 4.5), 8, 50, seed=7)
    assert Q.shape == (400, 3)
    qn = np.linalg.norm(Q @ B.T, axis=1)
    assert np.allclose(qn.reshape(8, 50), (0.5 + (np.arange(8) + 0.5) * 0.5)[:, None], rtol=1e-12)  # bin centres, |Q|-bin major
    d = (Q @ B.T) / qn[:, None]
    assert abs(d.mean(axis=0)).max() < 0.15  # isotropic directions
    assert np.array_equal(powder_points(B, (0.5, 4.5), 8, 50, seed=7, dir_range=(10, 30)), Q.reshape(8, 50, 3)[:, 10:30].reshape(-1, 3))
    vals = rng.uniform(-1, 12, (400, 5, 1))
    sf = rng.uniform(0, 3, (400, 5))
    for weight in (0, 1):
        hist, counts = powder_histogram(Q, B, vals, sf, (0.5, 4.5), 8, (0.0, 10.0), 20, weight)
        want = np.zeros((8, 20))
        for i in range(400):
            iq = int((qn[i] - 0.5) / 0.5)
            for m in range(5):
                w = vals[i, m, 0]
                if 0 <= w < 10 and (weight == 0 or w > 0):
                    want[iq, int(w / 0.5)] += sf[i, m] / (w if weight else 1.0)
        assert np.allclose(hist, want, rtol=1e-13) and np.array_equal(counts, np.full(8, 50.0))


@pytest.mark.gpu
@pytest.mark.parametrize("which", ["C3", "C3nest", "C2"])
def test_powder_average_on_reference_eigenvectors(host, bridge, which):
    """The histogram of the device (structure factor fused into the cell kernel, binned with atomics) against the numpy histogram
    of the numpy structure factor of the REFERENCE's ir_interpolate_at output on the same points; the sweep with points generated
    on the device against the same points binned from the host; direction slices add up to the whole."""
    import brille_b200
    from brille_b200 import workloads as W
    from brille_b200.sharding import ShardedGrid
    from oracle.consumer import powder_histogram, powder_points

    if which == "C2":
        wl = W.c2_nacl(host, density=300)
    else:
        wl = W.c3_p63mmc(host, density=300, seed=5, cls="BZNestQdc" if which == "C3nest" else "BZTrellisQdc")
    g = brille_b200.accelerate(wl.grid)
    B = np.asarray(bridge.flatten_bz(wl.bz)["to_xyz"]).reshape(3, 3)
    cfg = sf_config(wl.n_atoms, 11, cartesian=False, dw=False)
    cfg["q_transform"] = B  # Cartesian Q, as Euphonic uses it
    g.set_structure_factor(**cfg)
    qr, nqb, wr, nwb, n_dir = (0.2, 6.2), 24, (0.0, 52.0), 40, 4000
    Q = g.powder_points(qr, nqb, n_dir, seed=3)
    Qo = powder_points(B, qr, nqb, n_dir, seed=3)
    assert Q.shape == (nqb * n_dir, 3) and np.abs(Q - Qo).max() <= 1e-13 * np.abs(Qo).max()
    for weight in (0, 1):
        hist, counts = g.ir_powder_bin(Q, qr, nqb, wr, nwb, weight)
        assert g.last_path & 16, "the structure factor was not fused into the cell kernel"
        assert np.array_equal(counts, np.full(nqb, float(n_dir)))
        # reference eigenvectors -> numpy structure factor -> numpy histogram
        m = 30000
        sel = np.random.default_rng(1).choice(len(Q), m, replace=False)
        rv, rw = wl.grid.ir_interpolate_at(Q[sel], True, 8)
        want_sf = structure_factor(Q[sel], rw, **cfg)
        h_ref, c_ref = powder_histogram(Q[sel], B, rv, want_sf, qr, nqb, wr, nwb, weight)
        h_dev, c_dev = g.ir_powder_bin(Q[sel], qr, nqb, wr, nwb, weight)
        assert np.array_equal(c_dev, c_ref)
        # (an eigenvalue within rounding of a bin edge may fall on either side: compare bin-wise with a tolerance scaled to the
        # largest bin, and the totals tightly)
        assert np.abs(h_dev - h_ref).max() <= 1e-9 * h_ref.max()
        assert abs(h_dev.sum() - h_ref.sum()) <= 1e-10 * h_ref.sum()
        # against the device's own (vals, sf) through numpy: only the order of the additions differs
        dv, dsf = g.ir_structure_factor(Q)
        h_np, _ = powder_histogram(Q, B, dv, dsf, qr, nqb, wr, nwb, weight)
        assert np.abs(hist - h_np).max() <= 1e-12 * h_np.max()
        # the sweep (points generated on the device) == the same points binned from the host
        hs, cs = g.ir_powder_sweep(qr, nqb, wr, nwb, n_dir, seed=3, weight=weight)
        assert np.array_equal(cs, counts) and np.abs(hs - hist).max() <= 1e-12 * hist.max()
        # direction slices add up
        ha, ca = g.ir_powder_sweep(qr, nqb, wr, nwb, n_dir, seed=3, weight=weight, dir_range=(0, 1500))
        hb, cb = g.ir_powder_sweep(qr, nqb, wr, nwb, n_dir, seed=3, weight=weight, dir_range=(1500, n_dir))
        assert np.array_equal(ca + cb, counts) and np.abs(ha + hb - hist).max() <= 1e-12 * hist.max()
    # unfused reduction: the same histogram to rounding
    g.set_option("sf_fused", 0)
    hu, cu = g.ir_powder_sweep(qr, nqb, wr, nwb, n_dir, seed=3, weight=1)
    assert not (g.last_path & 16) and np.abs(hu - hist).max() <= 1e-11 * hist.max()
    g.set_option("sf_fused", 1)
    sg = ShardedGrid(wl.grid, [0, 0])
    sg.set_structure_factor(**cfg)
    hsg, csg = sg.ir_powder_sweep(qr, nqb, wr, nwb, n_dir, seed=3, weight=1)
    assert np.array_equal(csg, counts) and np.abs(hsg - hist).max() <= 1e-12 * hist.max()
    sg.close()
    with pytest.raises(RuntimeError, match="q_lo"):
        g.ir_powder_sweep((2.0, 1.0), 4, wr, nwb, 10)
    g.close()
