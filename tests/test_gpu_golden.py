"""GPU parity against the committed golden fixtures (outputs of the REAL reference, see
tests/golden/make_golden.py): TFCE maps bit-exact, FWER rows textually identical."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(G, name), allow_pickle=False)


def test_tfce_maps_vs_reference_outputs():
    from tfce_mediation_b200.tfce import CreateAdjSet
    g = load("tfce_maps.npz")
    csr = (g["indptr"], g["indices"])
    for a, (H, E) in enumerate(g["he"]):
        c = CreateAdjSet(H, E, csr)
        for b, m in enumerate(g["maps"]):
            out = np.zeros_like(m)
            c.run(np.ascontiguousarray(m), out)
            if H == 2:
                assert np.array_equal(out, g["tfce"][a, b]), (H, E, b)
            else:  # general H: device pow, tolerance 1e-5 (north_star)
                np.testing.assert_allclose(out, g["tfce"][a, b], rtol=1e-5)


def test_cynumstats_vs_reference_outputs():
    from tfce_mediation_b200 import cynumstats as cs
    g = load("cynumstats.npz")
    X, y = g["X"], g["y"]
    n, V = y.shape
    k = X.shape[1]
    invXX = np.linalg.inv(X.T @ X)
    tol = dict(rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(cs.tval_int(X, invXX, y, n, k, V), g["tval"], **tol)
    np.testing.assert_allclose(cs.cy_lin_lstsqr_mat(X, y), g["beta"], **tol)
    np.testing.assert_allclose(cs.resid_covars(X, np.ascontiguousarray(y.T)), g["resid"], **tol)
    np.testing.assert_allclose(cs.calcF(X, y, n, k), g["calcF"], rtol=1e-9)
    b, se = cs.calc_beta_se(X[:, 1], y, n, V)
    np.testing.assert_allclose(b, g["cbs_beta"], **tol)
    assert np.mean(se != g["cbs_se"]) < 1e-3 and np.allclose(se, g["cbs_se"], rtol=1e-6)
    assert np.array_equal(cs.se_of_slope(V, invXX, g["sigma2"], k), g["se"])


def _vertex_engine(g, two_sided=True):
    from tfce_mediation_b200._graph import induced_subgraph
    from tfce_mediation_b200.engine import PermutationEngine, Surface
    from tfce_mediation_b200.tfce import CreateAdjSet
    surfs, off = [], 0
    for keep in (g["keep_lh"], g["keep_rh"]):
        ip, ix = induced_subgraph(g["indptr"], g["indices"], keep)
        c = CreateAdjSet(2.0, 0.67, (ip, ix))
        surfs.append(Surface(c, off, g["density"][keep]))
        off += c.num_vertices
    return PermutationEngine(g["merge_y"], surfs, two_sided=two_sided)


def test_vertex_randomise_rows_identical_to_reference_csv():
    import oracle
    g = load("vertex_randomise.npz")
    eng = _vertex_engine(g)
    n = g["merge_y"].shape[0]
    X = np.column_stack([np.ones(n), g["pred_x"]])
    idx = np.stack([oracle.permutation_indices(s, n) for s in g["seeds"]])
    mx = eng.regression_block(X, perm_idx=idx)             # [P, C, S, 2]
    for c, want in ((0, g["rows_con1"]), (1, g["rows_con2"])):
        rows = []
        for p in range(len(idx)):
            for sg in (0, 1):
                rows.append("%.4f" % max(mx[p, c, 0, sg], mx[p, c, 1, sg]))
        assert rows == list(want)


def test_vertex_rows_through_dropin_functions(tmp_path, monkeypatch):
    """Same rows through the reference-shaped single-map API (CreateAdjSet.run + write_perm_maxTFCE_vertex)."""
    import oracle
    from tfce_mediation_b200 import cynumstats as cs, pyfunc
    from tfce_mediation_b200.tfce import CreateAdjSet
    g = load("vertex_randomise.npz")
    monkeypatch.chdir(tmp_path)
    indptr, indices = g["indptr"], g["indices"]
    adj = [indices[indptr[i]:indptr[i + 1]].tolist() for i in range(len(indptr) - 1)]
    c_lh, c_rh = CreateAdjSet(2.0, 0.67, adj), CreateAdjSet(2.0, 0.67, adj)
    y = g["merge_y"]
    n = y.shape[0]
    X = np.column_stack([np.ones(n), g["pred_x"]])
    k = X.shape[1]
    for seed in g["seeds"][:3]:
        nx = X[oracle.permutation_indices(seed, n)]
        t = cs.tval_int(nx, np.linalg.inv(nx.T @ nx), y, n, k, y.shape[1])
        for j in (1, 2):
            for sign in (1, -1):
                pyfunc.write_perm_maxTFCE_vertex("tstat_con%d" % j, t[j] * sign, int(g["keep_lh"].sum()), g["keep_lh"],
                                                 g["keep_rh"], c_lh, c_rh, g["density"], g["density"])
    for j, want in ((1, g["rows_con1"]), (2, g["rows_con2"])):
        got = [l.strip() for l in open("perm_tstat_con%d_TFCE_maxVertex.csv" % j)]
        assert got == list(want[:6])


def test_voxel_adjacency_and_rows_vs_reference():
    import oracle
    from tfce_mediation_b200 import pyfunc
    from tfce_mediation_b200.engine import PermutationEngine, Surface
    from tfce_mediation_b200.tfce import CreateAdjSet
    g = load("voxel.npz")
    mask = g["mask"]
    for conn, ip, ix in ((26, g["indptr26"], g["indices26"]), (6, g["indptr6"], g["indices6"])):
        adj = pyfunc.create_adjac_voxel(mask, mask.astype(np.float32), int(mask.sum()), dirtype=conn)
        p, i = oracle.adjacency_to_csr(list(adj))
        assert np.array_equal(p, ip) and np.array_equal(i, ix)
    adj = pyfunc.create_adjac_voxel(mask, mask.astype(np.float32), int(mask.sum()), dirtype=26)
    eng = PermutationEngine(g["y"], [Surface(CreateAdjSet(2.0, 0.5, adj), 0)], two_sided=True, nan_to_zero=True)
    n = g["y"].shape[0]
    idx = np.stack([oracle.permutation_indices(s, n) for s in g["seeds"]])
    mx = eng.regression_block(g["X"], perm_idx=idx)
    rows = ["%1.4f" % mx[p, 0, 0, sg] for p in range(len(idx)) for sg in (0, 1)]
    assert rows == list(g["rows"])


def test_sobelz_vs_reference_outputs():
    from tfce_mediation_b200 import pyfunc
    g = load("sobel.npz")
    n, V = g["merge_y"].shape
    for med in ("I", "M", "Y"):
        for alg in ("aroian", "sobel", "goodman"):
            got = pyfunc.calc_sobelz(med, g["pred_x"], g["depend_y"], g["merge_y"], n, V, alg=alg)
            want = g["%s_%s" % (med, alg)]
            ok = np.isfinite(want)
            np.testing.assert_allclose(got[ok], want[ok], rtol=1e-9)


def test_mmr_lowram_rows_identical_to_reference_csv():
    import oracle
    from tfce_mediation_b200._graph import induced_subgraph
    from tfce_mediation_b200.engine import PermutationEngine, Surface
    from tfce_mediation_b200.tfce import CreateAdjSet
    g = load("mmr_lowram.npz")
    keep = g["mask"] == 1
    ip, ix = induced_subgraph(g["indptr"], g["indices"], keep)
    eng = PermutationEngine(g["data"], [Surface(CreateAdjSet(2.0, 0.67, (ip, ix)), 0, g["vdensity"])], two_sided=True)
    n = g["data"].shape[0]
    X = np.column_stack([np.ones(n), g["pred_x"]])
    idx = np.stack([oracle.permutation_indices(int(pn) + int(g["perm_seed"]), n) for pn in g["perm_numbers"]])
    mx = eng.regression_block(X, perm_idx=idx)
    for c, want in ((0, g["rows_tcon1"]), (1, g["rows_tcon2"])):
        rows = ["%f" % mx[p, c, 0, sg] for p in range(len(idx)) for sg in (0, 1)]
        assert rows == list(want)


def test_fwe_lookup_vs_reference_outputs():
    from tfce_mediation_b200.tmanalysis.calculate_fweP import fwe_corrected_p, fwe_image
    import oracle
    g = load("fwe.npz")
    assert np.array_equal(fwe_corrected_p(g["perm_max"], g["values"]), g["corrp"])
    rs = np.random.RandomState(0)
    big = (np.abs(rs.standard_normal(50000)) * 100).astype(np.float32)
    big[::7] = 0
    pm = np.round(np.abs(rs.standard_normal(2000)) * 90, 4)
    got = fwe_image(pm, big)
    want = np.zeros(big.shape)
    want[big > 0] = oracle.fwe_p(pm, big[big > 0])
    assert np.array_equal(got, want)
