"""tm-models GLM statistics: pyfunc.py:2282-2401 glm_typeI and the GLM branch of
tmanalysis/tm_models_randomise.py:197-272.  CPU: the oracle restatement against the golden fixture produced by the real
reference (tests/golden/make_golden_glm.py) and the host algebra of the one-fit F formulation.  GPU: the fused F/t
kernels, the batched block and the driver against the oracle pipeline."""
import argparse
import os

import numpy as np
import pytest

import oracle
from tests import helpers
from tfce_mediation_b200 import synth

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
F64_TOL = 1e-10        # |delta| <= tol * max(1, |value|) for float64 statistics (BASELINE.json north_star)


def _golden():
    g = np.load(os.path.join(G, "glm_typeI.npz"))
    return g, [g["exog0"], g["exog1"], g["exog2"]]


def _close64(got, want):
    return np.all(np.abs(got - want) <= F64_TOL * np.maximum(1.0, np.abs(want)))


# ------------------------------------------------------------------------------------------- CPU
def test_oracle_glm_typeI_matches_reference_golden():
    g, exog = _golden()
    F, Fvar, T = oracle.glm_typeI(g["data"], exog, g["cov"], output_tvalues=True)
    assert np.array_equal(F, g["F"]) and np.array_equal(Fvar, g["Fvar"]) and np.array_equal(T, g["T"])
    F0, Fvar0, T0 = oracle.glm_typeI(g["data"], exog, None, output_tvalues=True)
    assert np.array_equal(F0, g["F_nocov"]) and np.array_equal(Fvar0, g["Fvar_nocov"]) and np.array_equal(T0, g["T_nocov"])
    for p, r in enumerate(g["perms"]):
        assert np.array_equal(oracle.glm_typeI(g["data"], exog, g["cov"], rand_array=r)[1], g["perm_Fvar"][p])
        assert np.array_equal(oracle.glm_typeI(g["data"], exog, g["cov"], output_fvalues=False, output_tvalues=True,
                                               rand_array=r), g["perm_T"][p])


def test_extra_sum_of_squares_identity_host():
    """RSS_without_S - RSS == b_S' inv(C_SS) b_S with C = inv(X'X) on centred designs: the algebra tmb_glm_fstat uses."""
    from tfce_mediation_b200.engine import design_stack, fstat_blocks
    from tfce_mediation_b200.pyfunc import typeI_design
    g, exog = _golden()
    y = g["data"].astype(np.float64)
    X, kvars = typeI_design(exog, g["cov"], y.shape[0])
    assert kvars == [1, 2, 1] and X.shape[1] == 7
    st = design_stack(X[None], center=True)
    b = st["pinv"][0] @ y                                                  # [r, V] slopes
    var_lo = [0, 1, 3]
    M = fstat_blocks(st["G"], var_lo, kvars)[0]
    off = 0
    for i, (lo, k) in enumerate(zip(var_lo, kvars)):
        Mi = M[off:off + k * k].reshape(k, k); off += k * k
        num = np.einsum("av,ab,bv->v", b[lo:lo + k], Mi, b[lo:lo + k])
        rss_full = oracle.lstsq_residual(X, y)[1]
        rss_red = oracle.lstsq_residual(np.delete(X, np.s_[1 + lo:1 + lo + k], 1), y)[1]
        assert np.allclose(num, rss_red - rss_full, rtol=1e-9, atol=1e-9)
        F = num / (rss_full / (y.shape[0] - X.shape[1]) * k)
        assert np.allclose(F, g["Fvar"][i], rtol=1e-9, atol=1e-9)


def test_rand_blocks_follows_reference_rng_calls():
    from tfce_mediation_b200.pyfunc import check_blocks, rand_blocks
    blocks = np.array(["a", "b", "a", "c", "b", "c", "a", "b", "c"])
    assert check_blocks(blocks) is True
    np.random.seed(5)
    got = rand_blocks(blocks, True)
    np.random.seed(5)
    idx = np.arange(9)
    want = np.concatenate([np.random.permutation(idx[blocks == b]) for b in np.random.permutation(list(np.unique(blocks)))])
    assert np.array_equal(got, want) and sorted(got.tolist()) == list(range(9))
    uneq = np.array(["a", "a", "b"])
    assert check_blocks(uneq) is False
    r = rand_blocks(uneq, False)
    assert sorted(r[:2].tolist()) == [0, 1] and r[2] == 2


# ------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
def test_glm_typeI_dropin_matches_reference_golden():
    from tfce_mediation_b200 import pyfunc
    g, exog = _golden()
    F, Fvar, T = pyfunc.glm_typeI(g["data"], exog, dmy_covariates=g["cov"], output_tvalues=True, verbose=False)
    assert _close64(Fvar, g["Fvar"]) and _close64(T, g["T"])
    # The MODEL F of float32 data: pyfunc.py:2331 accumulates SS_Total with numpy float32 arithmetic (mean, squares and sum
    # of the float32 data, row after row) and the residual SS in float64.  tmb_rm_totals restates that accumulation
    # (Device matrix sstotal_reference), so the model F agrees to float64 accuracy as well.
    assert _close64(F, g["F"])
    F64, Fvar64 = pyfunc.glm_typeI(g["data"].astype(np.float64), exog, dmy_covariates=g["cov"], verbose=False)
    assert _close64(F64, g["F_f64"]) and _close64(Fvar64, g["Fvar_f64"])
    F0, Fvar0 = pyfunc.glm_typeI(g["data"], exog, verbose=False)
    assert _close64(F0, g["F_nocov"]) and _close64(Fvar0, g["Fvar_nocov"])
    for p, r in enumerate(g["perms"][:3]):
        Fp = pyfunc.glm_typeI(g["data"], exog, dmy_covariates=g["cov"], verbose=False, rand_array=r)[1]
        assert _close64(Fp, g["perm_Fvar"][p])
        Tp = pyfunc.glm_typeI(g["data"], exog, dmy_covariates=g["cov"], output_fvalues=False, output_tvalues=True,
                              verbose=False, rand_array=r)
        assert _close64(Tp, g["perm_T"][p])


def _glm_state(n=40, seed=9):
    v, f, csr = helpers.ico(3)
    keep_lh, keep_rh = synth.cap_mask(v, 600), synth.cap_mask(-v, 590)
    dens = synth.vertex_density(synth.kring_csr(csr, 2))
    y = np.hstack([synth.subject_data(n, csr, seed, 2)[:, keep_lh], synth.subject_data(n, csr, seed + 1, 2)[:, keep_rh]])
    rs = np.random.RandomState(seed)
    grp = rs.randint(0, 3, n)
    exog = [rs.standard_normal((n, 1)), np.column_stack([(grp == 1) * 1.0, (grp == 2) * 1.0])]
    cov = rs.standard_normal((n, 2))
    y = y.astype(np.float32)
    y[:, :100] += np.float32(0.7) * exog[0].astype(np.float32)
    return dict(v=v, csr=csr, keep_lh=keep_lh, keep_rh=keep_rh, dens=dens, y=y, exog=exog, cov=cov, n=n)


def _oracle_rows(st, perms, stat):
    run = helpers.oracle_run(2, 0.67, st["csr"])
    nlh = int(st["keep_lh"].sum())
    rows_f, rows_t = [], []
    for r in perms:
        if stat in ("f", "both"):
            Fvar = oracle.glm_typeI(st["y"], st["exog"], st["cov"], rand_array=r)[1]
            rows_f.append([oracle.perm_max_vertex(Fvar[j], nlh, st["keep_lh"], st["keep_rh"], run, run, st["dens"], st["dens"])
                           for j in range(Fvar.shape[0])])
        if stat in ("t", "both"):
            T = oracle.glm_typeI(st["y"], st["exog"], st["cov"], output_fvalues=False, output_tvalues=True, rand_array=r)
            rows_t.append([[oracle.perm_max_vertex(T[j] * s, nlh, st["keep_lh"], st["keep_rh"], run, run, st["dens"], st["dens"])
                            for s in (1, -1)] for j in range(1, 4)])
    return np.array(rows_f), np.array(rows_t)


@pytest.mark.gpu
def test_glm_typeI_block_maxima_match_oracle_pipeline():
    from tfce_mediation_b200.engine import PermutationEngine
    from tfce_mediation_b200.pyfunc import typeI_design
    from tfce_mediation_b200.tmanalysis import _common as C
    st = _glm_state()
    adj = synth.csr_to_lists(st["csr"])
    surfs = [C.masked_surface(adj, 2, 0.67, st["keep_lh"], st["dens"], 0),
             C.masked_surface(adj, 2, 0.67, st["keep_rh"], st["dens"], int(st["keep_lh"].sum()))]
    eng = PermutationEngine(st["y"], surfs, two_sided=True)
    X, kvars = typeI_design(st["exog"], st["cov"], st["n"])
    perms = np.stack([oracle.permutation_indices(300 + p, st["n"]) for p in range(5)])
    f, t = eng.glm_typeI_block(X, kvars, perms, stat="both")
    want_f, want_t = _oracle_rows(st, perms, "both")
    assert f.shape == (5, 2, 2) and t.shape == (5, 3, 2, 2)
    assert np.allclose(f.max(axis=2), want_f, rtol=1e-5, atol=0)
    assert np.allclose(t.max(axis=2), want_t, rtol=1e-5, atol=0)


def _obj(lists):
    a = np.empty(len(lists), dtype=object)
    for i, l in enumerate(lists):
        a[i] = list(l)
    return a


@pytest.mark.gpu
@pytest.mark.parametrize("gstat", ["f", "t", "all"])
def test_tm_models_randomise_glm_driver_rows(tmp_path, monkeypatch, gstat):
    from tfce_mediation_b200.tmanalysis import tm_models_randomise as drv
    st = _glm_state()
    d = os.path.join(str(tmp_path), "tmtemp_GLM_area")
    os.makedirs(d)
    adj = synth.csr_to_lists(st["csr"])
    np.save(d + "/exog_flat.npy", np.column_stack(st["exog"])); np.save(d + "/exog_shape.npy", np.array([1, 2]))
    np.save(d + "/varnames.npy", np.array(["age", "group"])); np.save(d + "/gstat.npy", np.array(gstat))
    np.save(d + "/data.npy", st["y"]); np.save(d + "/optstfce.npy", np.array([2, 0.67]))
    np.save(d + "/dmy_covariates.npy", st["cov"]); np.save(d + "/num_vertex_lh.npy", int(st["keep_lh"].sum()))
    np.save(d + "/mask_lh.npy", st["keep_lh"]); np.save(d + "/mask_rh.npy", st["keep_rh"])
    np.save(d + "/adjac_lh.npy", _obj(adj), allow_pickle=True); np.save(d + "/adjac_rh.npy", _obj(adj), allow_pickle=True)
    np.save(d + "/vdensity_lh.npy", st["dens"]); np.save(d + "/vdensity_rh.npy", st["dens"])
    monkeypatch.chdir(tmp_path)
    opts = drv.getArgumentParser(argparse.ArgumentParser()).parse_args(["-r", "1", "4", "-s", "area", "-glm", "--seed", "3"])
    drv.run(opts)
    perms = [oracle.permutation_indices(p * 1000 + 3, st["n"]) for p in range(1, 5)]
    want_f, want_t = _oracle_rows(st, perms, {"f": "f", "t": "t", "all": "both"}[gstat])
    out = "output_GLM_area/perm_GLM"
    if gstat != "t":
        for j, name in enumerate(["age", "group"]):
            got = np.array([float(l) for l in open("%s/perm_Fstat_%s_TFCE_maxVertex.csv" % (out, name))])
            assert np.allclose(got, want_f[:, j], rtol=1e-5, atol=6e-5)
    else:
        assert not os.path.exists("%s/perm_Fstat_age_TFCE_maxVertex.csv" % out)
    if gstat != "f":
        for j in range(3):
            got = np.array([float(l) for l in open("%s/perm_Tstat_con%d_TFCE_maxVertex.csv" % (out, j + 1))])
            assert np.allclose(got, want_t[:, j, :].reshape(-1), rtol=1e-5, atol=6e-5)


@pytest.mark.gpu
def test_tm_models_randomise_glm_driver_volume_rows(tmp_path, monkeypatch):
    """-v (volume) input of the GLM branch: one voxel graph, no density weights, '%.4f' rows of maxVoxel files."""
    from tfce_mediation_b200.tmanalysis import tm_models_randomise as drv
    mask = synth.skeleton_mask(shape=(24, 28, 24), frac=0.35, seed=5, sigma=2.0, margin=3)
    csr = oracle.adjacency_to_csr(oracle.voxel_adjacency(mask, 26))
    V = int(mask.sum())
    n = 32
    rs = np.random.RandomState(12)
    y = rs.standard_normal((n, V)).astype(np.float32)
    exog = [rs.standard_normal((n, 1)), rs.standard_normal((n, 2))]
    d = os.path.join(str(tmp_path), "tmtemp_GLM_volume")
    os.makedirs(d)
    np.save(d + "/exog_flat.npy", np.column_stack(exog)); np.save(d + "/exog_shape.npy", np.array([1, 2]))
    np.save(d + "/varnames.npy", np.array(["x", "pair"])); np.save(d + "/gstat.npy", np.array("f"))
    np.save(d + "/data.npy", y); np.save(d + "/optstfce.npy", np.array([2, 0.5]))
    np.save(d + "/dmy_covariates.npy", np.array(None, dtype=object), allow_pickle=True)
    np.save(d + "/adjac.npy", _obj(synth.csr_to_lists(csr)), allow_pickle=True)
    monkeypatch.chdir(tmp_path)
    opts = drv.getArgumentParser(argparse.ArgumentParser()).parse_args(["-r", "1", "3", "-v", "-glm", "--seed", "8"])
    drv.run(opts)
    run = helpers.oracle_run(2, 0.5, csr)
    for j, name in enumerate(["x", "pair"]):
        got = np.array([float(l) for l in open("output_GLM_volume/perm_GLM/perm_Fstat_%s_TFCE_maxVoxel.csv" % name)])
        want = []
        for p in range(1, 4):
            Fvar = oracle.glm_typeI(y, exog, None, rand_array=oracle.permutation_indices(p * 1000 + 8, n))[1]
            want.append(oracle.perm_max_voxel(Fvar[j], run))
        assert np.allclose(got, np.array(want), rtol=1e-5, atol=6e-5)


# ------------------------------------------------------------------------------------------- tm-models mediation
def test_oracle_tm_models_sobelz_matches_reference_golden():
    g, _ = _golden()
    for m in "IMY":
        for i, r in enumerate(g["perms"][:3]):
            lv = g["med_left"][r]
            rv = g["med_right"][r] if m == "Y" else g["med_right"]
            assert np.array_equal(oracle.tm_models_sobelz(m, g["data"], lv, rv, g["cov"]), g["med_%s" % m][i])


@pytest.mark.gpu
@pytest.mark.parametrize("medtype", ["I", "M", "Y"])
def test_sobelz_designs_matches_reference_golden(medtype):
    """glm_typeI t-values + calc_indirect of the reference (golden) against the fused two-fit kernel on general designs."""
    from tfce_mediation_b200.engine import PermutationEngine
    g, _ = _golden()
    data, cov, left, right = g["data"], g["cov"], g["med_left"], g["med_right"]
    n = data.shape[0]
    eng = PermutationEngine(data, None)
    perms = g["perms"][:3]
    ones = np.ones((3, n, 1))
    lv = left[perms]
    rv = right[perms] if medtype == "Y" else np.broadcast_to(right, (3,) + right.shape)
    cv = np.broadcast_to(cov, (3,) + cov.shape)
    XA = np.concatenate([ones, lv, cv], axis=2)
    XB = np.concatenate([ones, lv, rv, cv], axis=2) if medtype == "I" else np.concatenate([ones, rv, lv, cv], axis=2)
    ta = None
    if medtype == "Y":
        ta = np.array([oracle.glm_typeI(rv[p], [lv[p]], cov, output_fvalues=False, output_tvalues=True)[1][0] for p in range(3)])
        XA = None
    _, z64 = eng.sobelz_designs(XA, XB, ta, "aroian", want_f64=True)
    z = z64[:, :data.shape[1]].cpu().numpy()
    assert np.all(np.abs(z - g["med_%s" % medtype]) <= 1e-9 * np.maximum(1.0, np.abs(g["med_%s" % medtype])))


@pytest.mark.gpu
@pytest.mark.parametrize("medtype", ["I", "M", "Y"])
def test_tm_models_sobelz_from_cross_products_matches_reference_golden(medtype):
    """The tm-models mediation block contracts only the PERMUTED columns with the data (tmb_sobelz_cross_rows; the fixed
    columns' cross-products are fitted once): float64 z against the reference's glm_typeI t-values + calc_indirect."""
    from tfce_mediation_b200.engine import PermutationEngine
    g, _ = _golden()
    data, cov, left, right = g["data"], g["cov"], g["med_left"], g["med_right"]
    n = data.shape[0]
    eng = PermutationEngine(data, None)
    perms = g["perms"][:3]
    ones = np.ones((3, n, 1))
    lv = left[perms]
    rv = right[perms] if medtype == "Y" else np.broadcast_to(right, (3,) + right.shape)
    cv = np.broadcast_to(cov, (3,) + cov.shape)
    XB = np.concatenate([ones, lv, rv, cv], axis=2) if medtype == "I" else np.concatenate([ones, rv, lv, cv], axis=2)
    ta = None
    if medtype == "Y":
        ta = np.array([oracle.glm_typeI(rv[p], [lv[p]], cov, output_fvalues=False, output_tvalues=True)[1][0] for p in range(3)])
    z32, z64 = eng._tm_models_sobelz_cross(medtype, left, right, cov, perms, XB, ta, "aroian", want_f64=True)
    z = z64[:, :data.shape[1]].cpu().numpy()
    assert np.all(np.abs(z - g["med_%s" % medtype]) <= 1e-9 * np.maximum(1.0, np.abs(g["med_%s" % medtype])))
    assert np.array_equal(z32[:, :data.shape[1]].cpu().numpy(), z.astype(np.float32))


@pytest.mark.gpu
@pytest.mark.parametrize("medtype", ["M", "Y"])
def test_tm_models_randomise_mediation_driver_rows(tmp_path, monkeypatch, medtype):
    from tfce_mediation_b200.tmanalysis import tm_models_randomise as drv
    st = _glm_state()
    n = st["n"]
    rs = np.random.RandomState(21)
    left = rs.standard_normal((n, 1))
    right = 0.5 * left + rs.standard_normal((n, 1))
    d = os.path.join(str(tmp_path), "tmtemp_mediation_area")
    os.makedirs(d)
    adj = synth.csr_to_lists(st["csr"])
    np.save(d + "/dmy_leftvar.npy", left); np.save(d + "/dmy_rightvar.npy", right); np.save(d + "/medtype.npy", np.array(medtype))
    np.save(d + "/data.npy", st["y"]); np.save(d + "/optstfce.npy", np.array([2, 0.67]))
    np.save(d + "/dmy_covariates.npy", st["cov"]); np.save(d + "/num_vertex_lh.npy", int(st["keep_lh"].sum()))
    np.save(d + "/mask_lh.npy", st["keep_lh"]); np.save(d + "/mask_rh.npy", st["keep_rh"])
    np.save(d + "/adjac_lh.npy", _obj(adj), allow_pickle=True); np.save(d + "/adjac_rh.npy", _obj(adj), allow_pickle=True)
    np.save(d + "/vdensity_lh.npy", st["dens"]); np.save(d + "/vdensity_rh.npy", st["dens"])
    monkeypatch.chdir(tmp_path)
    opts = drv.getArgumentParser(argparse.ArgumentParser()).parse_args(["-r", "1", "4", "-s", "area", "-med", "--seed", "6"])
    drv.run(opts)
    run = helpers.oracle_run(2, 0.67, st["csr"])
    nlh = int(st["keep_lh"].sum())
    want, lv, rv = [], left, right
    for p in range(1, 5):                                   # the reference permutes its variables in place (:436,:459,:480-481)
        r = oracle.permutation_indices(p * 1000 + 6, n)
        lv = lv[r]
        if medtype == "Y":
            rv = rv[r]
        z = oracle.tm_models_sobelz(medtype, st["y"], lv, rv, st["cov"])
        want.append(oracle.perm_max_vertex(z, nlh, st["keep_lh"], st["keep_rh"], run, run, st["dens"], st["dens"]))
    got = np.array([float(l) for l in open("output_mediation_area/perm_mediation/perm_Zstat_%s_TFCE_maxVertex.csv" % medtype)])
    assert np.allclose(got, np.array(want), rtol=1e-5, atol=6e-5)


@pytest.mark.gpu
def test_glm_typeI_with_many_covariates_matches_oracle():
    """tm-models GLM with dummy-coded covariates beyond 8 regressors (pyfunc.py:2282-2401 has no limit): the partial F
    and t maps of the stored-beta path agree with the oracle's restatement to 1e-9."""
    from tfce_mediation_b200 import pyfunc
    n, V = 90, 400
    rs = np.random.RandomState(4)
    data = rs.standard_normal((n, V)).astype(np.float64)
    site = np.eye(8)[rs.randint(0, 8, n)][:, 1:]
    exog = [rs.standard_normal((n, 1)), rs.standard_normal((n, 2))]
    cov = np.column_stack([site, rs.standard_normal((n, 4))])           # 3 + 11 = 14 regressors
    perm = rs.permutation(n)
    for r in (None, perm):
        F, Fvar, T = pyfunc.glm_typeI(data, exog, dmy_covariates=cov, output_tvalues=True, verbose=False, rand_array=r)
        wF, wFvar = oracle.glm_typeI(data, exog, cov, rand_array=r)
        wT = oracle.glm_typeI(data, exog, cov, output_fvalues=False, output_tvalues=True, rand_array=r)
        tol = lambda a, b: np.all(np.abs(a - b) <= 1e-9 * np.maximum(1, np.abs(b)))   # noqa: E731
        assert tol(F, wF) and tol(Fvar, wFvar) and tol(T, wT)
