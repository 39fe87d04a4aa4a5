"""GPU parity: fit / t-statistic kernels against the numpy oracle (restating cynumstats.pyx).

Tolerances: float64 results |dt| <= 1e-10 * max(1, |t|) (north_star: <= 1e-10 fp64);
float32 t-maps fed to TFCE: <= 1e-5 relative hard gate and, as measured against the compiled
reference, bit-identical (the number of differing float32 values is asserted to be tiny)."""
import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu


def _data(n, V, k, seed, mean=0.0, dtype=np.float32):
    rs = np.random.RandomState(seed)
    y = (rs.standard_normal((n, V)) + mean).astype(dtype)
    X = np.column_stack([np.ones(n), rs.standard_normal((n, k - 1))])
    return X, y


def _close64(a, b, tol=1e-10):
    return np.all(np.abs(a - b) <= tol * np.maximum(1, np.abs(b)))


@pytest.mark.parametrize("n,V,k,mean,dtype", [(40, 777, 2, 0.0, np.float32), (100, 5000, 4, 2.5, np.float32),
                                               (57, 1031, 3, -1.0, np.float64), (30, 129, 9, 0.0, np.float32)])
def test_tval_int_matches_oracle(n, V, k, mean, dtype):
    from tfce_mediation_b200 import cynumstats as cs
    X, y = _data(n, V, k, 1, mean, dtype)
    invXX = np.linalg.inv(X.T @ X)
    got = cs.tval_int(X, invXX, y, n, k, V)
    want = oracle.tval_int(X, invXX, y, n, k, V)
    assert got.shape == want.shape and got.dtype == np.float64
    assert _close64(got, want)
    assert np.mean(got.astype(np.float32) != want.astype(np.float32)) < 1e-4


def test_lstsq_beta_residuals_calcF_se():
    from tfce_mediation_b200 import cynumstats as cs
    n, V, k = 64, 2000, 4
    X, y = _data(n, V, k, 2, 1.0)
    assert _close64(cs.cy_lin_lstsqr_mat(X, y), oracle.lstsq_beta(X, y))
    assert _close64(cs.cy_lin_lstsqr_mat(X, y[:, 3]), oracle.lstsq_beta(X, y[:, 3]))
    b, sse = cs.cy_lin_lstsqr_mat_residual(X, y)
    wb, wsse = oracle.lstsq_residual(X, y)
    assert _close64(b, wb) and _close64(sse, wsse)
    assert _close64(cs.calcF(X, y, n, k), oracle.calcF(X, y, n, k), 1e-9)
    data = np.ascontiguousarray(y.T)                      # V x n like the reference's step-1 input
    assert _close64(cs.resid_covars(X, data), oracle.resid_covars(X, data))
    beta, se = cs.calc_beta_se(X[:, 1], y, n, V)
    wbeta, wse = oracle.calc_beta_se(X[:, 1], y, n, V)
    assert _close64(beta, wbeta) and se.dtype == np.float32
    assert np.mean(se != wse) < 1e-4 and np.allclose(se, wse, rtol=1e-6)
    sigma2 = np.abs(np.random.RandomState(0).standard_normal(V))
    invXX = np.linalg.inv(X.T @ X)
    assert np.array_equal(cs.se_of_slope(V, invXX, sigma2, k), oracle.se_of_slope(V, invXX, sigma2, k))


@pytest.mark.parametrize("n,V,k,P", [(100, 10242, 2, 70), (48, 3001, 4, 9), (33, 515, 6, 130)])
def test_engine_tstat_batch_matches_oracle(n, V, k, P):
    import torch
    from tfce_mediation_b200 import engine as eng
    from tfce_mediation_b200._device import DeviceMatrix
    X, y = _data(n, V, k, 3, 2.0)
    rs = np.random.RandomState(7)
    idx = np.stack([rs.permutation(n) for _ in range(P)])
    Y = DeviceMatrix(y)
    e = eng.PermutationEngine.__new__(eng.PermutationEngine)      # fit only: no TFCE plan needed
    e.device, e.Y, e.nan_to_zero, e.h2d_bytes, e.d2h_bytes, e._pinned, e.colperm = Y.t.device, Y, False, 0, 0, {}, None
    t32, t64 = e.tstat(eng.row_permuted_stack(X, idx), want_f64=True)
    t32, t64 = t32.cpu().numpy()[:, :, :V], t64.cpu().numpy()[:, :, :V]
    # and through explicit per-shuffle designs (the -v / mediation route)
    t32b = e.tstat(eng.design_stack(np.stack([X[i] for i in idx]))).cpu().numpy()[:, :, :V]
    bad = 0
    for p in range(0, P, max(1, P // 6)):
        nx = X[idx[p]]
        want = oracle.tval_int(nx, np.linalg.inv(nx.T @ nx), y, n, k, V)[1:]
        assert _close64(t64[p], want)
        np.testing.assert_allclose(t32[p], want.astype(np.float32), rtol=1e-5, atol=1e-7)
        np.testing.assert_allclose(t32b[p], want.astype(np.float32), rtol=1e-5, atol=1e-7)
        bad += int(np.sum(t32[p] != want.astype(np.float32)))
    assert bad <= 2, "float32 t-maps are expected to be bit-identical to the reference (%d differ)" % bad
    assert torch.cuda.is_available()


def test_dmma_fast_epilogue_bit_identical_to_exact(monkeypatch):
    """The headline fit's fp32 t-maps: the cheap epilogue (fp32 MUFU seeds + one fp64 Newton step, exact path whenever the
    value is within 2^-41 of an fp32 rounding boundary) must reproduce the exact fp64 sqrt/division path bit for bit --
    including constant columns (sse = 0 -> inf / NaN), tiny and huge scales, and zero slopes."""
    import torch
    from tfce_mediation_b200.engine import PermutationEngine
    n, V, P = 96, 40000, 192
    rs = np.random.RandomState(5)
    y = rs.standard_normal((n, V)).astype(np.float32)
    y[:, :64] = 1.25                                         # constant columns: beta = 0, sse = 0
    y[:, 64:128] *= np.float32(1e-18)                        # se far below the fp32-normal window of the fast path
    y[:, 128:192] *= np.float32(1e17)
    x = rs.standard_normal(n)
    y[:, 192:256] = (np.outer(x, np.ones(64)) * 3).astype(np.float32)   # exact fit of the unpermuted design: sse ~ 0
    X = np.column_stack([np.ones(n), x])
    perm = np.stack([np.arange(n)] + [rs.permutation(n) for _ in range(P - 1)])
    eng = PermutationEngine(y, None)
    monkeypatch.setenv("TMB_GLM_EPILOGUE", "exact")
    want = eng.tstat_rowperm(X, perm).clone()
    monkeypatch.delenv("TMB_GLM_EPILOGUE")
    got = eng.tstat_rowperm(X, perm).clone()
    torch.cuda.synchronize()
    a, b = want.view(torch.int32), got.view(torch.int32)
    assert int((a != b).sum()) == 0
    # and the exact path is the oracle's (spot check on the regular columns)
    nx = X[perm[3]]
    ref = oracle.tval_int(nx, np.linalg.inv(nx.T @ nx), y[:, 256:2256], n, 2, 2000)[1].astype(np.float32)
    assert np.mean(got[3, 0, 256:2256].cpu().numpy() != ref) < 1e-3


def _fit_engine(y):
    from tfce_mediation_b200 import engine as eng
    from tfce_mediation_b200._device import DeviceMatrix
    Y = DeviceMatrix(y)
    e = eng.PermutationEngine.__new__(eng.PermutationEngine)      # fit only: no TFCE plan needed
    e.device, e.Y, e.nan_to_zero, e.h2d_bytes, e.d2h_bytes, e._pinned, e.colperm = Y.t.device, Y, False, 0, 0, {}, None
    e._rings = {}
    return e


@pytest.mark.parametrize("k", [2, 3, 4, 5, 6, 7, 8, 9])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_every_regressor_count_on_both_fit_kernels(k, dtype):
    """r = k - 1 = 1..8 regressors: float32 data runs on the tensor-core kernels (tile8 column order, one CTA shape per
    r, no padded rows), float64 data on the fp64 vector kernel (r padded to 1/2/4/8).  Ragged P (not a multiple of 8)
    and a ragged last subject chunk (n % 32 != 0).  cynumstats.pyx:59-64 via the oracle."""
    from tfce_mediation_b200 import engine as eng
    n, V, P = 45, 1300, 21
    X, y = _data(n, V, k, 11 + k, 0.5, dtype)
    rs = np.random.RandomState(k)
    idx = np.stack([rs.permutation(n) for _ in range(P)])
    e = _fit_engine(y)
    t32, t64 = e.tstat(eng.row_permuted_stack(X, idx), want_f64=True)
    t32r = e.tstat_rowperm(X, idx)                                  # device-side gather of the pseudo-inverse columns
    t32, t64, t32r = t32.cpu().numpy()[:, :, :V], t64.cpu().numpy()[:, :, :V], t32r.cpu().numpy()[:, :, :V]
    assert np.array_equal(t32, t32r)
    bad = 0
    for p in range(P):
        nx = X[idx[p]]
        want = oracle.tval_int(nx, np.linalg.inv(nx.T @ nx), y, n, k, V)[1:]
        assert _close64(t64[p], want)
        bad += int(np.sum(t32[p] != want.astype(np.float32)))
    assert bad <= 2, bad
    # a sub-range of rows, as the drivers request it for `-v first last`
    if k >= 4:
        sub = e.tstat(eng.row_permuted_stack(X, idx), rows=(1, 2)).cpu().numpy()[:, :, :V]
        assert np.array_equal(sub, t32[:, 1:3])


@pytest.mark.parametrize("k", [10, 21, 40])
def test_many_regressors_stored_beta_path(k):
    """More than 8 non-intercept regressors (site dummies + covariates): the reference has no limit
    (cynumstats.pyx:28-29,59-64); here the betas make one round trip through HBM (tmb_glm_beta + tmb_glm_tstat_beta)."""
    from tfce_mediation_b200 import engine as eng
    n, V, P = 90, 700, 5
    X, y = _data(n, V, k, 50 + k, 1.0)
    rs = np.random.RandomState(k)
    idx = np.stack([rs.permutation(n) for _ in range(P)])
    e = _fit_engine(y)
    t32, t64 = e.tstat(eng.row_permuted_stack(X, idx), want_f64=True)
    t32r = e.tstat_rowperm(X, idx, rows=(0, 3)).cpu().numpy()[:, :, :V]
    t32, t64 = t32.cpu().numpy()[:, :, :V], t64.cpu().numpy()[:, :, :V]
    assert np.array_equal(t32[:, :3], t32r)
    for p in range(P):
        nx = X[idx[p]]
        want = oracle.tval_int(nx, np.linalg.inv(nx.T @ nx), y, n, k, V)[1:]
        assert _close64(t64[p], want, 1e-9)
        np.testing.assert_allclose(t32[p], want.astype(np.float32), rtol=1e-5, atol=1e-7)


def test_regression_block_with_twenty_regressors_end_to_end():
    """A 20-column design through the whole engine (fit -> TFCE -> max): rows equal the oracle pipeline's."""
    from tests import helpers
    from tfce_mediation_b200.engine import PermutationEngine, Surface
    from tfce_mediation_b200.tfce import CreateAdjSet
    from tfce_mediation_b200 import synth
    _, _, csr = helpers.ico(4)
    V = csr[0].shape[0] - 1
    n, k, P = 80, 20, 3
    y = synth.subject_data(n, csr, 5, 2)
    rs = np.random.RandomState(9)
    site = np.eye(6)[rs.randint(0, 6, n)][:, 1:]                      # dummy-coded site
    X = np.column_stack([np.ones(n), rs.standard_normal((n, k - 1 - site.shape[1])), site])
    eng = PermutationEngine(y, [Surface(CreateAdjSet(2, 0.67, csr), 0)])
    idx = np.stack([oracle.permutation_indices(77 + p, n) for p in range(P)])
    got = eng.regression_block(X, perm_idx=idx)
    assert got.shape == (P, k - 1, 1, 2)
    for p in range(P):
        nx = X[idx[p]]
        t = oracle.tval_int(nx, np.linalg.inv(nx.T @ nx), y, n, k, V)
        for c in (0, 7, k - 2):
            want = helpers.oracle_signed_max(2, 0.67, csr, t[c + 1].astype(np.float32))
            assert "%.4f" % got[p, c, 0, 0] == "%.4f" % want[0] and "%.4f" % got[p, c, 0, 1] == "%.4f" % want[1]


@pytest.mark.parametrize("k,first,last", [(5, 1, 2), (4, 3, 3), (3, 1, 1)])
def test_partial_column_permutation_from_cross_products(monkeypatch, k, first, last):
    """The drivers' `-v first last` mode (vertex_tfce_multiple_regression_randomise.py:84-97): only regressors first..last
    are permuted, cumulatively; the other columns stay.  The engine contracts only the changing columns with the data
    (tmb_glm_tstat_cross_rows): float64 t within 1e-10 of the whole-design fit and of the oracle's tval_int, float32 maps
    equal to the float64 ones rounded, and the block's maxima equal those of the whole-design path."""
    from tests import helpers
    from tfce_mediation_b200.engine import PermutationEngine, Surface, design_stack
    from tfce_mediation_b200.tfce import CreateAdjSet
    n, P = 60, 12
    _, _, csr = helpers.ico(4)
    V = csr[0].shape[0] - 1
    rs = np.random.RandomState(k * 10 + first)
    X = np.column_stack([np.ones(n), rs.standard_normal((n, k - 1))])
    y = (rs.standard_normal((n, V)) + 0.3 * X[:, first][:, None]).astype(np.float32)
    designs = []
    for p in range(P):                                     # the reference permutes the chosen columns in place
        X[:, first:last + 1] = X[oracle.permutation_indices(5000 + p, n), first:last + 1]
        designs.append(X.copy())
    designs = np.stack(designs)
    eng = PermutationEngine(y, [Surface(CreateAdjSet(2, 0.67, csr), 0)], two_sided=True)
    changing = eng._partial_columns(designs)
    assert changing is not None and list(changing) == list(range(first - 1, last))
    t32, t64 = eng.tstat_partial(designs, changing, want_f64=True)
    t32, t64 = eng.to_caller_order(t32).cpu().numpy()[:, :, :V], eng.to_caller_order(t64).cpu().numpy()[:, :, :V]
    _, w64 = eng.tstat(design_stack(designs, center=True), want_f64=True)
    w64 = w64.cpu().numpy()[:, :, :V]
    tol = lambda a, b: np.all(np.abs(a - b) <= 1e-10 * np.maximum(1.0, np.abs(b)))  # noqa: E731
    assert tol(t64, w64) and np.array_equal(t32, t64.astype(np.float32))
    for p in (0, P - 1):
        nx = designs[p]
        assert tol(t64[p], oracle.tval_int(nx, np.linalg.inv(nx.T @ nx), y, n, k, V)[1:])
    got = eng.regression_block(None, designs=designs)
    monkeypatch.setenv("TMB_GLM_PARTIAL", "0")
    assert eng._partial_columns(designs) is None
    want = eng.regression_block(None, designs=designs)
    assert got.shape == (P, k - 1, 1, 2) and np.allclose(got, want, rtol=1e-6, atol=0)
