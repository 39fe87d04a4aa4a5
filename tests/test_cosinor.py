"""tm-models cosinor and cosinor mediation: pyfunc.py:2406-2563 glm_cosinor as the
permutation driver calls it (tmanalysis/tm_models_randomise.py:274-426).  CPU: the oracle restatement and the host
path-A algebra against the golden fixture produced by the real reference (tests/golden/make_golden_cosinor.py).
GPU: the statistics kernel against the golden fixture, the batched blocks and the drivers against the oracle pipeline."""
import argparse
import os

import numpy as np
import pytest

import oracle
from tests import helpers
from tfce_mediation_b200 import synth

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
F64_TOL = 1e-10        # |delta| <= tol * max(1, |value|) for float64 statistics (BASELINE.json north_star)
CASES = ("full", "noexog", "exog1")


def _golden():
    return np.load(os.path.join(G, "cosinor.npz"))


def _case(g, tag):
    """(exog list or None, covariates or None, periods) of a golden case."""
    if tag == "full":
        return [g["exog0"], g["exog1"]], g["cov"], list(g["period"])
    if tag == "noexog":
        return None, None, [24.0]
    return [g["exog0"]], None, [24.0]


def _close64(got, want, tol=F64_TOL):
    return np.all(np.abs(got - want) <= tol * np.maximum(1.0, np.abs(want)))


# ------------------------------------------------------------------------------------------- CPU
@pytest.mark.parametrize("tag", CASES)
def test_oracle_glm_cosinor_matches_reference_golden(tag):
    g = _golden()
    exog, cov, period = _case(g, tag)
    for p, r in enumerate(g["perms"]):
        F, ta, tc, te = oracle.glm_cosinor(g["data"], g["time_var"], exog, cov, rand_array=r, period=period)
        assert np.array_equal(F, g["F_" + tag][p])
        assert np.array_equal(ta, g["tamp_" + tag][p]) and np.array_equal(tc, g["tacr_" + tag][p])
        if exog is not None:
            assert np.array_equal(te, g["texog_" + tag][p])


def test_oracle_cosinor_mediation_matches_reference_golden():
    g = _golden()
    ta = oracle.glm_cosinor(g["mediator"], g["time_var"], period=[24.0])[1]
    assert np.array_equal(ta, g["med_ta"])
    for p, r in enumerate(g["perms"]):
        tb = oracle.glm_cosinor(g["data"], g["time_var"], [g["mediator"]], None, rand_array=r, period=[24.0])[3]
        assert np.array_equal(oracle.calc_indirect(ta[0], tb[0]), g["med_z"][p])


def test_host_path_a_amplitude_t_matches_reference_golden():
    from tfce_mediation_b200.engine import cosinor_amplitude_t, cosinor_design
    g = _golden()
    got = cosinor_amplitude_t(g["mediator"], g["time_var"], [24.0])
    assert got.shape == (1,) and _close64(got, g["med_ta"].reshape(-1))
    X, nper, nexog = cosinor_design(g["time_var"], list(g["period"]), [g["exog0"], g["exog1"]], g["cov"])
    assert X.shape == (g["time_var"].shape[0], 1 + 4 + 3 + 2) and (nper, nexog) == (2, 3)
    assert np.array_equal(X[:, 1], np.cos(2.0 * np.pi * g["time_var"] / 24.0))
    assert np.array_equal(X[:, 4], np.sin(2.0 * np.pi * g["time_var"] / 12.0))


# ------------------------------------------------------------------------------------------- GPU
def _line_engine(data, two_sided=True):
    """Engine over a chain graph (the statistics tests do not look at the TFCE result)."""
    from tfce_mediation_b200.engine import PermutationEngine
    from tfce_mediation_b200.tmanalysis import _common as C
    V = data.shape[1]
    adj = [[j for j in (i - 1, i + 1) if 0 <= j < V] for i in range(V)]
    return PermutationEngine(data, [C.masked_surface(adj, 2.0, 0.67)], two_sided=two_sided)


@pytest.mark.gpu
@pytest.mark.parametrize("tag", CASES)
def test_cosinor_stats_kernel_matches_reference_golden(tag):
    from tfce_mediation_b200.engine import cosinor_design
    g = _golden()
    exog, cov, period = _case(g, tag)
    eng = _line_engine(g["data"])
    X, nper, nexog = cosinor_design(g["time_var"], period, exog, cov)
    V = g["data"].shape[1]
    s32, s64 = eng.cosinor_stats(X, nper, nexog, g["perms"], want_f64=True)
    s32, s64 = s32.cpu().numpy()[:, :, :V], s64.cpu().numpy()[:, :, :V]
    assert s64.shape == (6, 1 + 2 * nper + nexog, V)
    # Model F: the reference accumulates SS_Total in float32 for float32 data (pyfunc.py:2492); tmb_rm_totals restates
    # that accumulation, so the row agrees to float64 accuracy too
    assert _close64(s64[:, 0], g["F_" + tag], 1e-9)
    for i in range(nper):
        assert _close64(s64[:, 1 + 2 * i], g["tamp_" + tag][:, i])
        assert _close64(s64[:, 2 + 2 * i], g["tacr_" + tag][:, i])
    if nexog:
        assert _close64(s64[:, 1 + 2 * nper:], g["texog_" + tag][:, :nexog])       # the golden also holds the covariates' t
    assert np.array_equal(s32, s64.astype(np.float32))


@pytest.mark.gpu
def test_cosinor_mediation_stat_matches_reference_golden():
    from tfce_mediation_b200.engine import cosinor_amplitude_t, cosinor_design
    g = _golden()
    eng = _line_engine(g["data"], two_sided=False)
    X, nper, nexog = cosinor_design(g["time_var"], [24.0], [g["mediator"]])
    ta = cosinor_amplitude_t(g["mediator"], g["time_var"], [24.0])[0]
    V = g["data"].shape[1]
    _, z64 = eng.cosinor_stats(X, nper, nexog, g["perms"], mediation_ta=ta, want_f64=True)
    assert _close64(z64.cpu().numpy()[:, 0, :V], g["med_z"], 1e-9)


def _state(n=40, seed=21):
    v, f, csr = helpers.ico(3)
    keep_lh, keep_rh = synth.cap_mask(v, 600), synth.cap_mask(-v, 590)
    dens = synth.vertex_density(synth.kring_csr(csr, 2))
    y = np.hstack([synth.subject_data(n, csr, seed, 2)[:, keep_lh], synth.subject_data(n, csr, seed + 1, 2)[:, keep_rh]])
    rs = np.random.RandomState(seed)
    time_var = rs.uniform(0, 24, n)
    exog = [rs.standard_normal((n, 1)), rs.standard_normal((n, 1))]
    cov = rs.standard_normal((n, 2))
    y = y.astype(np.float32)
    y[:, :150] += (np.float32(0.9) * np.cos(2 * np.pi * (time_var - 4.0) / 24.0))[:, None].astype(np.float32)
    mediator = 0.7 * np.cos(2 * np.pi * (time_var - 2.0) / 24.0) + 0.6 * rs.standard_normal(n)
    return dict(v=v, csr=csr, keep_lh=keep_lh, keep_rh=keep_rh, dens=dens, y=y, exog=exog, cov=cov, n=n,
                time_var=time_var, mediator=mediator - mediator.mean(), period=[24.0, 8.0])


def _surfaces(st):
    from tfce_mediation_b200.tmanalysis import _common as C
    adj = synth.csr_to_lists(st["csr"])
    return [C.masked_surface(adj, 2, 0.67, st["keep_lh"], st["dens"], 0),
            C.masked_surface(adj, 2, 0.67, st["keep_rh"], st["dens"], int(st["keep_lh"].sum()))]


def _oracle_rows(st, perms, exog, cov, period):
    run = helpers.oracle_run(2, 0.67, st["csr"])
    nlh = int(st["keep_lh"].sum())
    mx = lambda stat: oracle.perm_max_vertex(stat, nlh, st["keep_lh"], st["keep_rh"], run, run, st["dens"], st["dens"])  # noqa: E731
    pos, tex = [], []
    for r in perms:
        F, ta, tc, te = oracle.glm_cosinor(st["y"], st["time_var"], exog, cov, rand_array=r, period=period)
        row = [mx(F)]
        for i in range(len(period)):
            row += [mx(ta[i]), mx(tc[i])]
        pos.append(row)
        if exog is not None:
            ncon = sum(np.asarray(e).reshape(st["n"], -1).shape[1] for e in exog)
            tex.append([[mx(te[j] * s) for s in (1, -1)] for j in range(ncon)])
    return np.array(pos), np.array(tex)


def _oracle_mediation_rows(st, perms, period):
    run = helpers.oracle_run(2, 0.67, st["csr"])
    nlh = int(st["keep_lh"].sum())
    ta = oracle.glm_cosinor(st["mediator"], st["time_var"], period=period)[1]
    rows = []
    for r in perms:
        tb = oracle.glm_cosinor(st["y"], st["time_var"], [st["mediator"]], None, rand_array=r, period=period)[3]
        rows.append(oracle.perm_max_vertex(oracle.calc_indirect(ta[0], tb[0]), nlh, st["keep_lh"], st["keep_rh"], run, run,
                                           st["dens"], st["dens"]))
    return np.array(rows)


@pytest.mark.gpu
def test_cosinor_block_maxima_match_oracle_pipeline():
    from tfce_mediation_b200.engine import PermutationEngine
    st = _state()
    eng = PermutationEngine(st["y"], _surfaces(st), two_sided=True)
    perms = np.stack([oracle.permutation_indices(700 + p, st["n"]) for p in range(5)])
    pos, tex = eng.cosinor_block(st["time_var"], st["period"], st["exog"], st["cov"], perms)
    want_pos, want_t = _oracle_rows(st, perms, st["exog"], st["cov"], st["period"])
    assert pos.shape == (5, 5, 2) and tex.shape == (5, 2, 2, 2)
    assert np.allclose(pos.max(axis=2), want_pos, rtol=1e-5, atol=0)
    assert np.allclose(tex.max(axis=2), want_t, rtol=1e-5, atol=0)
    pos0, tex0 = eng.cosinor_block(st["time_var"], [24.0], None, None, perms[:2])
    assert tex0 is None and np.allclose(pos0.max(axis=2), _oracle_rows(st, perms[:2], None, None, [24.0])[0], rtol=1e-5, atol=0)
    z = eng.cosinor_mediation_block(st["time_var"], [24.0], st["mediator"], perms)
    assert z.shape == (5, 2) and np.allclose(z.max(axis=1), _oracle_mediation_rows(st, perms, [24.0]), rtol=1e-5, atol=0)


def _obj(lists):
    a = np.empty(len(lists), dtype=object)
    for i, l in enumerate(lists):
        a[i] = list(l)
    return a


def _write_common(d, st):
    adj = synth.csr_to_lists(st["csr"])
    np.save(d + "/data.npy", st["y"]); np.save(d + "/optstfce.npy", np.array([2, 0.67]))
    np.save(d + "/num_vertex_lh.npy", int(st["keep_lh"].sum()))
    np.save(d + "/mask_lh.npy", st["keep_lh"]); np.save(d + "/mask_rh.npy", st["keep_rh"])
    np.save(d + "/adjac_lh.npy", _obj(adj), allow_pickle=True); np.save(d + "/adjac_rh.npy", _obj(adj), allow_pickle=True)
    np.save(d + "/vdensity_lh.npy", st["dens"]); np.save(d + "/vdensity_rh.npy", st["dens"])
    np.save(d + "/time_var.npy", st["time_var"])


def _rows(path):
    return np.array([float(l) for l in open(path)])


@pytest.mark.gpu
@pytest.mark.parametrize("named", [True, False])
def test_tm_models_randomise_cosinor_driver_rows(tmp_path, monkeypatch, named):
    from tfce_mediation_b200.tmanalysis import tm_models_randomise as drv
    st = _state()
    d = os.path.join(str(tmp_path), "tmtemp_cosinor_area")
    os.makedirs(d)
    _write_common(d, st)
    np.save(d + "/period.npy", np.array(st["period"])); np.save(d + "/dmy_covariates.npy", st["cov"])
    np.save(d + "/exog_flat.npy", np.column_stack(st["exog"])); np.save(d + "/exog_shape.npy", np.array([1, 1]))
    np.save(d + "/varnames.npy", np.array(["age", "bmi"] if named else ["pair"]))
    monkeypatch.chdir(tmp_path)
    opts = drv.getArgumentParser(argparse.ArgumentParser()).parse_args(["-r", "2", "5", "-s", "area", "-cos", "--seed", "11"])
    drv.run(opts)
    perms = [oracle.permutation_indices(p * 1000 + 11, st["n"]) for p in range(2, 6)]
    want_pos, want_t = _oracle_rows(st, perms, st["exog"], st["cov"], st["period"])
    out = "output_cosinor_area/perm_cosinor"
    tol = dict(rtol=1e-5, atol=6e-5)
    assert np.allclose(_rows(out + "/perm_Fstat_model_TFCE_maxVertex.csv"), want_pos[:, 0], **tol)
    for i, per in enumerate(st["period"]):
        assert np.allclose(_rows(out + "/perm_Tstat_amplitude_%2.2f_TFCE_maxVertex.csv" % per), want_pos[:, 1 + 2 * i], **tol)
        assert np.allclose(_rows(out + "/perm_Tstat_acrophase_%2.2f_TFCE_maxVertex.csv" % per), want_pos[:, 2 + 2 * i], **tol)
    for j, name in enumerate(["age", "bmi"] if named else ["con1", "con2"]):
        assert np.allclose(_rows(out + "/perm_Tstat_%s_TFCE_maxVertex.csv" % name), want_t[:, j, :].reshape(-1), **tol)


@pytest.mark.gpu
def test_tm_models_randomise_cosinor_mediation_driver_rows(tmp_path, monkeypatch):
    from tfce_mediation_b200.tmanalysis import tm_models_randomise as drv
    st = _state()
    d = os.path.join(str(tmp_path), "tmtemp_medcosinor_area")
    os.makedirs(d)
    _write_common(d, st)
    np.save(d + "/period.npy", np.array([24.0])); np.save(d + "/dmy_covariates.npy", np.array(None, dtype=object), allow_pickle=True)
    np.save(d + "/dmy_mediator.npy", st["mediator"]); np.save(d + "/medtype.npy", np.array("M"))
    monkeypatch.chdir(tmp_path)
    opts = drv.getArgumentParser(argparse.ArgumentParser()).parse_args(["-r", "1", "4", "-s", "area", "-mcos", "--seed", "5"])
    drv.run(opts)
    perms = [oracle.permutation_indices(p * 1000 + 5, st["n"]) for p in range(1, 5)]
    got = _rows("output_medcosinor_area/perm_cosinor/perm_Zstat_M_TFCE_maxVertex.csv")
    assert np.allclose(got, _oracle_mediation_rows(st, perms, [24.0]), rtol=1e-5, atol=6e-5)


@pytest.mark.gpu
def test_glm_cosinor_dropin_matches_reference_golden():
    """pyfunc.glm_cosinor with the reference's signature: all twelve outputs of the un-permuted call, the fit-only form, a
    permuted call, and the 1-D path-A call of the cosinor mediation, against the real reference (golden)."""
    from tfce_mediation_b200 import pyfunc
    g = _golden()
    exog, cov, period = _case(g, "full")
    names = ("R2", "MESOR", "SE_MESOR", "AMPLITUDE", "SE_AMPLITUDE", "ACROPHASE", "SE_ACROPHASE", "Fmodel", "tMESOR",
             "tAMPLITUDE", "tACROPHASE", "tEXOG")
    got = pyfunc.glm_cosinor(endog=g["data"], time_var=g["time_var"], exog=exog, dmy_covariates=cov, period=period)
    assert len(got) == 12
    for nm, val in zip(names, got):
        want = g["obs_" + nm]
        val = np.asarray(val, dtype=np.float64)
        assert val.shape == want.shape, nm
        # SE_MESOR is float32 in the reference (se_of_slope); R2 and Fmodel carry its float32 SS_Total, restated on the host
        assert _close64(val, want, 1e-9), nm
    fit = pyfunc.glm_cosinor(endog=g["data"], time_var=g["time_var"], exog=exog, dmy_covariates=cov, period=period,
                             output_fit_only=True)
    for nm, val in zip(("MESOR", "AMPLITUDE", "ACROPHASE"), fit):
        assert _close64(np.asarray(val), g["fit_" + nm], 1e-9), nm
    r = g["perms"][2]
    perm = pyfunc.glm_cosinor(endog=g["data"], time_var=g["time_var"], exog=exog, dmy_covariates=cov, rand_array=r,
                              period=period, calc_MESOR=False)
    assert perm[0] is None and _close64(perm[7], g["F_full"][2], 1e-9) and _close64(perm[9], g["tamp_full"][2], 1e-9)
    assert _close64(perm[10], g["tacr_full"][2], 1e-9) and _close64(perm[11], g["texog_full"][2], 1e-9)
    ta = pyfunc.glm_cosinor(endog=g["mediator"], time_var=g["time_var"], period=[24.0])[9]
    assert _close64(np.asarray(ta), g["med_ta"], 1e-9)
