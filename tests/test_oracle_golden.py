"""CPU: pin the oracle against the golden fixtures generated from the REAL reference
(tests/golden/make_golden.py) and, when oracle/_ref is present, against the compiled reference."""
import os

import numpy as np
import pytest

import oracle
from oracle import build_ref

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(G, name), allow_pickle=False)


def test_tfce_maps_bitexact():
    g = load("tfce_maps.npz")
    csr = (g["indptr"], g["indices"])
    for a, (H, E) in enumerate(g["he"]):
        for b, m in enumerate(g["maps"]):
            assert np.array_equal(oracle.tfce_run(H, E, csr, m), g["tfce"][a, b]), (H, E, b)


def test_pure_python_restatement_agrees_on_small_case():
    g = load("tfce_maps.npz")
    indptr, indices = g["indptr"], g["indices"]
    adj = [indices[indptr[i]:indptr[i + 1]].tolist() for i in range(len(indptr) - 1)]
    assert np.array_equal(oracle.tfce_run_pure(2, 0.67, adj, g["maps"][1]), g["tfce"][0, 1])


def test_cynumstats_restatement():
    g = load("cynumstats.npz")
    X, y = g["X"], g["y"]
    n, V = y.shape
    k = X.shape[1]
    invXX = np.linalg.inv(X.T @ X)
    np.testing.assert_allclose(oracle.tval_int(X, invXX, y, n, k, V), g["tval"], rtol=1e-12, atol=0)
    np.testing.assert_allclose(oracle.lstsq_beta(X, y), g["beta"], rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(oracle.resid_covars(X, np.ascontiguousarray(y.T)), g["resid"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(oracle.calcF(X, y, n, k), g["calcF"], rtol=1e-9)
    b, se = oracle.calc_beta_se(X[:, 1], y, n, V)
    np.testing.assert_allclose(b, g["cbs_beta"], rtol=1e-12, atol=1e-14)
    assert np.array_equal(se, g["cbs_se"])
    assert np.array_equal(oracle.se_of_slope(V, invXX, g["sigma2"], k), g["se"])


def test_vertex_randomise_rows_identical():
    g = load("vertex_randomise.npz")
    csr = (g["indptr"], g["indices"])
    y, pred_x = g["merge_y"], g["pred_x"]
    n = y.shape[0]
    X = np.column_stack([np.ones(n), pred_x])
    k = X.shape[1]
    run = lambda img, out: oracle.tfce_run(2.0, 0.67, csr, img, out)  # noqa: E731
    keep_lh, keep_rh, dens = g["keep_lh"], g["keep_rh"], g["density"]
    rows = {1: [], 2: []}
    for seed in g["seeds"]:
        nx = X[oracle.permutation_indices(seed, n)]
        t = oracle.tval_int(nx, np.linalg.inv(nx.T @ nx), y, n, k, y.shape[1])
        for j in (1, 2):
            for sign in (1, -1):
                rows[j].append("%.4f" % oracle.perm_max_vertex(t[j] * sign, int(keep_lh.sum()), keep_lh, keep_rh, run,
                                                               run, dens, dens))
    assert rows[1] == list(g["rows_con1"])
    assert rows[2] == list(g["rows_con2"])


def test_voxel_adjacency_and_rows():
    g = load("voxel.npz")
    mask = g["mask"]
    for conn, ip, ix in ((26, g["indptr26"], g["indices26"]), (6, g["indptr6"], g["indices6"])):
        adj = oracle.voxel_adjacency(mask, conn)
        p, i = oracle.adjacency_to_csr(adj)
        assert np.array_equal(p, ip) and np.array_equal(i, ix)
    csr = (g["indptr26"], g["indices26"])
    y, X = g["y"], g["X"]
    n, V = y.shape
    run = lambda img, out: oracle.tfce_run(2.0, 0.5, csr, img, out)  # noqa: E731
    rows = []
    for seed in g["seeds"]:
        nx = X[oracle.permutation_indices(seed, n)]
        t = oracle.tval_int(nx, np.linalg.inv(nx.T @ nx), y, n, 2, V)
        t[np.isnan(t)] = 0
        rows.append("%1.4f" % oracle.perm_max_voxel(t[1], run))
        rows.append("%1.4f" % oracle.perm_max_voxel(t[1] * -1, run))
    assert rows == list(g["rows"])


def test_sobelz_restatement():
    g = load("sobel.npz")
    n, V = g["merge_y"].shape
    for med in ("I", "M", "Y"):
        for alg in ("aroian", "sobel", "goodman"):
            got = oracle.sobelz(med, g["pred_x"], g["depend_y"], g["merge_y"], n, V, alg)
            np.testing.assert_allclose(got, g["%s_%s" % (med, alg)], rtol=1e-11, equal_nan=True)


def test_mmr_lowram_rows_identical():
    g = load("mmr_lowram.npz")
    csr = (g["indptr"], g["indices"])
    run = lambda img, out: oracle.tfce_run(2.0, 0.67, csr, img, out)  # noqa: E731
    r1, r2 = [], []
    for pn in g["perm_numbers"]:
        rows = oracle.low_ram_max(g["data"], g["mask"], g["pred_x"], run, g["vdensity"], int(pn), int(g["perm_seed"]))
        r1 += ["%f" % rows[0][0], "%f" % rows[0][1]]
        r2 += ["%f" % rows[1][0], "%f" % rows[1][1]]
    assert r1 == list(g["rows_tcon1"])
    assert r2 == list(g["rows_tcon2"])


def test_fwe_lookup():
    g = load("fwe.npz")
    assert np.array_equal(oracle.fwe_p(g["perm_max"], g["values"]), g["corrp"])


@pytest.mark.skipif(build_ref.load() is None, reason="oracle/_ref (compiled reference) not present")
def test_c_oracle_vs_compiled_reference_random():
    ref_tfce, ref_stats = build_ref.load()
    from tests import helpers
    _, _, csr = helpers.ico(4)
    adj = [csr[1][csr[0][i]:csr[0][i + 1]].tolist() for i in range(len(csr[0]) - 1)]
    for seed, (H, E) in enumerate([(2, 0.67), (2, 1), (2, 0.5), (2, 2), (1.5, 0.8), (3, 0.67)]):
        c = ref_tfce.CreateAdjSet(H, E, adj)
        for rounds in (0, 2, 5):
            img = helpers.smooth_map(csr, 50 + seed * 7 + rounds, rounds, scale=1 + seed)
            want = np.zeros_like(img)
            c.run(img, want)
            assert np.array_equal(oracle.tfce_run(H, E, csr, img), want)
    # directed / asymmetric adjacency (tools builder quirk)
    g = helpers.grid_csr(9, 7)
    g = [[a for a in lst if a != 0] for lst in g]
    c = ref_tfce.CreateAdjSet(2, 0.67, g)
    rs = np.random.RandomState(0)
    for _ in range(5):
        img = rs.standard_normal(63).astype(np.float32)
        want = np.zeros_like(img)
        c.run(img, want)
        assert np.array_equal(oracle.tfce_run(2, 0.67, oracle.adjacency_to_csr(g), img), want)


def test_threshold_sequence_properties():
    for mx in (4.4231, 0.001, 873.25, 1e-30, 3.0e38):
        T = oracle.tfce_thresholds(mx)
        assert T[0] == np.float32(mx) and len(T) in (100, 101, 102)
        assert np.all(np.diff(T) < 0) and T[-1] >= 0


def test_apply_mfwer_restatement_vs_reference(tmp_path, monkeypatch):
    """tm_func.apply_mfwer (host post-processing) against the reference's output."""
    from tfce_mediation_b200 import tm_func
    g = load("mfwer.npz")
    monkeypatch.chdir(tmp_path)
    os.mkdir("output_t")
    for sf in range(3):
        for c in (1, 2):
            np.savetxt("output_t/perm_maxTFCE_surf%d_tcon%d.csv" % (sf, c), g["csv_s%d_c%d" % (sf, c)], fmt="%f")
    pa = g["position_array"].tolist()
    for wname, w in (("none", None), ("logmasksize", "logmasksize")):
        pos, neg = tm_func.apply_mfwer([g["image"].copy()], 2, [0, 1, 2], int(g["num_perm"]), 3, "t", pa,
                                       pos_range=[0, 1], neg_range=[2, 3], weight=w)
        assert np.array_equal(pos, g["pos_" + wname])
        assert np.array_equal(neg, g["neg_" + wname])
    assert tm_func.lowest_length(2, [0, 1, 2], "t") == int(g["num_perm"])


def test_host_libm_powf_matches_the_fixtures_libm():
    """Bit-identity of TFCE values rests on libm's powf (fast_tfce.hpp:70 compiles to it, and it is NOT correctly rounded:
    for H = 2 a few thresholds in 10^4 differ from T*T by one ulp).  The golden fixtures were generated with the build
    container's glibc; this pins that the library on THIS machine returns the same bits on a 4,007-point grid for the
    exponents the reference uses, so a fixture mismatch elsewhere cannot be a silent libm difference."""
    import ctypes
    import ctypes.util
    g = np.load(os.path.join(G, "libm_powf.npz"))
    libm = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
    libm.powf.restype = ctypes.c_float
    libm.powf.argtypes = [ctypes.c_float, ctypes.c_float]
    got = np.array([[libm.powf(float(t), float(h)) for t in g["T"]] for h in g["H"]], dtype=np.float32)
    assert np.array_equal(got.view(np.int32), g["powf"].view(np.int32)), \
        "this machine's libm powf differs from the one the golden fixtures were generated with"
    assert int((g["powf"][0] != (g["T"] * g["T"]).astype(np.float32)).sum()) > 0      # the non-trivial case is covered
