"""GPU parity of whole shuffles: fit -> TFCE -> scaled max against the oracle pipeline that restates
vertex_tfce_multiple_regression_randomise.py:90-117 + pyfunc.py:107-126 and the mediation twin."""
import numpy as np
import pytest

import oracle
from tests import helpers
from tfce_mediation_b200 import synth

pytestmark = pytest.mark.gpu


def _two_hemi_setup(level, n, k, seed, weight=False):
    from tfce_mediation_b200.engine import Surface
    from tfce_mediation_b200.tfce import CreateAdjSet
    _, _, csr = helpers.ico(level)
    V = csr[0].shape[0] - 1
    rs = np.random.RandomState(seed)
    y = np.concatenate([synth.subject_data(n, csr, seed, 3), synth.subject_data(n, csr, seed + 1, 3)], axis=1)
    X = np.column_stack([np.ones(n), rs.standard_normal((n, k - 1))])
    w = synth.vertex_density(synth.kring_csr(csr, 2)) if weight else None
    surfs = [Surface(CreateAdjSet(2, 0.67, csr), 0, w), Surface(CreateAdjSet(2, 0.67, csr), V, w)]
    return csr, V, y, X, surfs, w


@pytest.mark.parametrize("k,weight", [(2, False), (4, True)])
def test_regression_block_matches_oracle_pipeline(k, weight):
    from tfce_mediation_b200.engine import PermutationEngine
    n, P = 60, 12
    csr, V, y, X, surfs, w = _two_hemi_setup(4, n, k, 10, weight)
    eng = PermutationEngine(y, surfs, two_sided=True)
    idx = np.stack([oracle.permutation_indices(2000 + p, n) for p in range(P)])
    got = eng.regression_block(X, perm_idx=idx)                       # [P, C, S, 2]
    assert got.shape == (P, k - 1, 2, 2)
    run = helpers.oracle_run(2, 0.67, csr)
    mask = np.ones(V, dtype=bool)
    dens = 1 if w is None else w
    for p in range(P):
        nx = X[idx[p]]
        t = oracle.tval_int(nx, np.linalg.inv(nx.T @ nx), y, n, k, 2 * V)
        for c in range(k - 1):
            for sg, sign in enumerate((1.0, -1.0)):
                want = oracle.perm_max_vertex(t[c + 1] * sign, V, mask, mask, run, run, dens, dens)
                have = max(got[p, c, 0, sg], got[p, c, 1, sg])
                assert np.float32(want) == np.float32(have), (p, c, sg, want, have)
                assert "%.4f" % want == "%.4f" % have


@pytest.mark.parametrize("medtype", ["M", "I", "Y"])
def test_mediation_block_matches_oracle_pipeline(medtype):
    from tfce_mediation_b200.engine import PermutationEngine
    n, P = 50, 6
    csr, V, y, X, surfs, _ = _two_hemi_setup(4, n, 2, 20)
    rs = np.random.RandomState(5)
    pred_x = rs.standard_normal(n)
    dep = 0.5 * pred_x + rs.standard_normal(n)
    y = (y + 0.3 * pred_x[:, None] + 0.2 * dep[:, None]).astype(np.float32)
    eng = PermutationEngine(y, surfs, two_sided=False)
    idx = np.stack([oracle.permutation_indices(4000 + p, n) for p in range(P)])
    got, z32, _ = eng.mediation_block(medtype, pred_x, dep, idx, want_maps=True)
    got, z32 = got.cpu().numpy(), z32.cpu().numpy()[:, :2 * V]
    run = helpers.oracle_run(2, 0.67, csr)
    mask = np.ones(V, dtype=bool)
    for p in range(P):
        xp = pred_x[idx[p]]
        dp = dep[idx[p]] if medtype == "Y" else dep
        z = oracle.sobelz(medtype, xp, dp, y, n, 2 * V)
        np.testing.assert_allclose(z32[p], z.astype(np.float32), rtol=1e-5, atol=1e-7)
        want = oracle.perm_max_vertex(z, V, mask, mask, run, run)
        have = max(got[p, 0], got[p, 1])
        np.testing.assert_allclose(have, want, rtol=1e-5)


def test_voxel_style_single_surface_nan_to_zero():
    from tfce_mediation_b200.engine import PermutationEngine, Surface
    from tfce_mediation_b200.tfce import CreateAdjSet
    adj = helpers.grid_csr(30, 20)
    csr = oracle.adjacency_to_csr(adj)
    V, n, k, P = 600, 40, 3, 5
    rs = np.random.RandomState(2)
    y = rs.standard_normal((n, V)).astype(np.float32)
    y[:, 17] = 0.0                                           # zero-variance voxel -> NaN t in the reference
    X = np.column_stack([np.ones(n), rs.standard_normal((n, k - 1))])
    eng = PermutationEngine(y, [Surface(CreateAdjSet(2, 0.5, adj), 0)], two_sided=True, nan_to_zero=True)
    idx = np.stack([oracle.permutation_indices(3000 + p, n) for p in range(P)])
    got = eng.regression_block(X, perm_idx=idx)
    run = helpers.oracle_run(2, 0.5, csr)
    for p in range(P):
        nx = X[idx[p]]
        with np.errstate(divide="ignore", invalid="ignore"):
            t = oracle.tval_int(nx, np.linalg.inv(nx.T @ nx), y, n, k, V)
        t[np.isnan(t)] = 0                                   # voxel_tfce_multiple_regression_randomise.py:109
        for c in range(k - 1):
            assert np.float32(oracle.perm_max_voxel(t[c + 1], run)) == got[p, c, 0, 0]
            assert np.float32(oracle.perm_max_voxel(t[c + 1] * -1, run)) == got[p, c, 0, 1]


def test_observed_statistics_match_step1_writer_math():
    """Identity permutation with full maps: pyfunc.py:80-91 write_vertStat_img's array math."""
    from tfce_mediation_b200.engine import PermutationEngine
    n, k = 40, 3
    csr, V, y, X, surfs, w = _two_hemi_setup(4, n, k, 30, True)
    eng = PermutationEngine(y, surfs, two_sided=True)
    obs = eng.observed_statistics(X)
    t = oracle.tval_int(X, np.linalg.inv(X.T @ X), y, n, k, 2 * V)[1:].astype(np.float32)
    assert np.array_equal(obs["t"], t)
    for c in range(k - 1):
        for h in range(2):
            seg = np.ascontiguousarray(t[c, h * V:(h + 1) * V])
            want_pos = oracle.tfce_run(2, 0.67, csr, seg) * (seg.max() / 100) * w
            want_neg = oracle.tfce_run(2, 0.67, csr, -seg) * ((-seg).max() / 100) * w
            assert np.array_equal(obs["tfce_pos"][c, h * V:(h + 1) * V], want_pos)
            assert np.array_equal(obs["tfce_neg"][c, h * V:(h + 1) * V], want_neg)
            assert obs["max_pos"][c, h] == want_pos.max() and obs["max_neg"][c, h] == want_neg.max()


def test_mediation_blocks_pipelined_equals_block_by_block():
    """engine.mediation_blocks (software-pipelined: host design algebra of block i+1 behind the sweep of block i) returns
    the rows of mediation_block, and those are the oracle pipeline's (pyfunc.py:130-162 -> :107-119)."""
    import oracle
    from tests import helpers
    from tfce_mediation_b200 import synth
    from tfce_mediation_b200.engine import PermutationEngine, Surface
    from tfce_mediation_b200.tfce import CreateAdjSet
    _, _, csr = helpers.ico(4)
    V = csr[0].shape[0] - 1
    n, N = 36, 11
    rs = np.random.RandomState(12)
    px = rs.standard_normal(n)
    dep = 0.5 * px + rs.standard_normal(n)
    y = (synth.subject_data(n, csr, 3, 2) + np.float32(0.3) * px[:, None].astype(np.float32)
         + np.float32(0.3) * dep[:, None].astype(np.float32)).astype(np.float32)
    eng = PermutationEngine(y, [Surface(CreateAdjSet(2, 0.67, csr), 0)], two_sided=False)
    idx = np.stack([oracle.permutation_indices(500 + p, n) for p in range(N)])
    for med in ("M", "I", "Y"):
        whole = eng.mediation_blocks(med, px, dep, idx, block=4)
        parts = np.concatenate([eng.mediation_block(med, px, dep, idx[a:a + 4]) for a in range(0, N, 4)])
        assert whole.shape == (N, 1) and np.array_equal(whole, parts)
    run = helpers.oracle_run(2, 0.67, csr)
    mask = np.ones(V, dtype=bool)
    whole = eng.mediation_blocks("M", px, dep, idx, block=4)
    for p in (0, 5, 10):
        z = oracle.sobelz("M", px[idx[p]], dep, y, n, V).astype(np.float32)
        tf = np.zeros(V, dtype=np.float32)
        run(np.ascontiguousarray(z), tf)
        want = np.float32((tf * (z.max() / 100)).max())
        assert "%.4f" % whole[p, 0] == "%.4f" % want


@pytest.mark.parametrize("medtype", ["M", "I"])
@pytest.mark.parametrize("alg", ["aroian", "sobel", "goodman"])
def test_sobelz_from_cross_products_fast_epilogue_is_bit_identical_to_exact(monkeypatch, medtype, alg):
    """Medtype 'M' / 'I' run from one contraction row per shuffle (tmb_sobelz_cross).  Its float32 epilogue takes float32
    seeds + Newton steps and falls back to the exact sequence near rounding boundaries: same bits as TMB_GLM_EPILOGUE=exact
    on 2.6 M values; and the float64 z agrees with the two-design formulation (tmb_sobelz, TMB_SOBEL=designs) and with the
    oracle to 1e-10."""
    from tfce_mediation_b200.engine import PermutationEngine
    n, P = 60, 128
    csr, V, y, X, surfs, _ = _two_hemi_setup(5, n, 2, 31)
    rs = np.random.RandomState(17)
    pred_x = rs.standard_normal(n)
    dep = 0.4 * pred_x + rs.standard_normal(n)
    y = (y + 0.25 * pred_x[:, None] + 0.2 * dep[:, None]).astype(np.float32)
    y[:, 7] = 0.0                                   # a constant column: NaN in the reference, NaN here
    eng = PermutationEngine(y, surfs, two_sided=False)
    idx = np.stack([oracle.permutation_indices(7000 + p, n) for p in range(P)])
    assert eng.sobelz_cross_ok(medtype)
    fast = eng.sobelz(medtype, pred_x, dep, idx, alg).cpu().numpy()
    monkeypatch.setenv("TMB_GLM_EPILOGUE", "exact")
    exact = eng.sobelz(medtype, pred_x, dep, idx, alg).cpu().numpy()
    monkeypatch.delenv("TMB_GLM_EPILOGUE")
    assert fast.shape[0] == P and np.array_equal(fast, exact, equal_nan=True)
    z32, z64 = eng.sobelz(medtype, pred_x, dep, idx[:6], alg, want_f64=True)
    z32, z64 = z32.cpu().numpy()[:, :2 * V], z64.cpu().numpy()[:, :2 * V]
    assert np.array_equal(z32, fast[:6, :2 * V], equal_nan=True)
    monkeypatch.setenv("TMB_SOBEL", "designs")
    assert not eng.sobelz_cross_ok(medtype)
    _, d64 = eng.sobelz(medtype, pred_x, dep, idx[:6], alg, want_f64=True)
    d64 = d64.cpu().numpy()[:, :2 * V]
    ok = np.isfinite(d64)
    assert np.array_equal(ok, np.isfinite(z64))
    assert np.all(np.abs(z64[ok] - d64[ok]) <= 1e-10 * np.maximum(1.0, np.abs(d64[ok])))
    for p in range(3):
        with np.errstate(divide="ignore", invalid="ignore"):
            want = oracle.sobelz(medtype, pred_x[idx[p]], dep, y, n, 2 * V, alg=alg)
        okp = np.isfinite(want)
        assert np.all(np.abs(z64[p][okp] - want[okp]) <= 1e-10 * np.maximum(1.0, np.abs(want[okp])))
