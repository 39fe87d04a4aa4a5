"""GPU: the randomise drivers end to end (reference on-disk contract in, CSV rows out) against the rows the
oracle pipeline / the real reference produce for the same permutation stream."""
import argparse
import os

import numpy as np
import pytest

import oracle
from tests import helpers
from tfce_mediation_b200 import synth

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _obj(lists):
    a = np.empty(len(lists), dtype=object)
    for i, l in enumerate(lists):
        a[i] = list(l)
    return a


def _vertex_state(tmp, surface, n=36, k=3, seed=3, med=False):
    v, f, csr = helpers.ico(3)
    V = v.shape[0]
    adj = synth.csr_to_lists(csr)
    keep_lh, keep_rh = synth.cap_mask(v, 600), synth.cap_mask(-v, 590)
    dens = synth.vertex_density(synth.kring_csr(csr, 2))
    y = np.hstack([synth.subject_data(n, csr, seed, 2)[:, keep_lh], synth.subject_data(n, csr, seed + 1, 2)[:, keep_rh]])
    rs = np.random.RandomState(seed)
    pred_x = rs.standard_normal((n, k - 1)) if not med else rs.standard_normal(n)
    d = os.path.join(tmp, ("python_temp_med_%s" if med else "python_temp_%s") % surface)
    os.makedirs(d)
    os.makedirs(os.path.join(tmp, ("output_med_%s" if med else "output_%s") % surface))
    np.save(d + "/merge_y.npy", y.astype(np.float32))
    np.save(d + "/num_vertex.npy", y.shape[1]); np.save(d + "/num_vertex_lh.npy", int(keep_lh.sum()))
    np.save(d + "/all_vertex.npy", V); np.save(d + "/num_subjects.npy", n)
    np.save(d + "/bin_mask_lh.npy", keep_lh); np.save(d + "/bin_mask_rh.npy", keep_rh)
    np.save(d + "/pred_x.npy", pred_x)
    np.save(d + "/adjac_lh.npy", _obj(adj), allow_pickle=True); np.save(d + "/adjac_rh.npy", _obj(adj), allow_pickle=True)
    np.save(d + "/optstfce.npy", np.array([2, 0.67])); np.save(d + "/vdensity_lh.npy", dens); np.save(d + "/vdensity_rh.npy", dens)
    dep = None
    if med:
        dep = 0.5 * pred_x + rs.standard_normal(n)
        np.save(d + "/depend_y.npy", dep)
    return dict(csr=csr, keep_lh=keep_lh, keep_rh=keep_rh, dens=dens, y=y.astype(np.float32), pred_x=pred_x, n=n, dep=dep)


def test_vertex_regression_driver_rows(tmp_path, monkeypatch):
    from tfce_mediation_b200.tmanalysis import vertex_tfce_multiple_regression_randomise as drv
    st = _vertex_state(str(tmp_path), "area")
    monkeypatch.chdir(tmp_path)
    opts = drv.getArgumentParser(argparse.ArgumentParser()).parse_args(["-r", "1", "7", "-s", "area", "--seed", "11"])
    drv.run(opts)
    n, y = st["n"], st["y"]
    X = np.column_stack([np.ones(n), st["pred_x"]])
    run = helpers.oracle_run(2, 0.67, st["csr"])
    want = {1: [], 2: []}
    for p in range(1, 8):
        nx = X[oracle.permutation_indices(p * 1000 + 11, n)]
        t = oracle.tval_int(nx, np.linalg.inv(nx.T @ nx), y, n, 3, y.shape[1])
        for j in (1, 2):
            for sign in (1, -1):
                want[j].append("%.4f" % oracle.perm_max_vertex(t[j] * sign, int(st["keep_lh"].sum()), st["keep_lh"],
                                                               st["keep_rh"], run, run, st["dens"], st["dens"]))
    for j in (1, 2):
        got = [l.strip() for l in open("output_area/perm_Tstat_area/perm_tstat_con%d_TFCE_maxVertex.csv" % j)]
        assert got == want[j]


def test_vertex_regression_driver_specifyvars_and_blocks(tmp_path, monkeypatch):
    from tfce_mediation_b200.tmanalysis import vertex_tfce_multiple_regression_randomise as drv
    st = _vertex_state(str(tmp_path), "thickness", n=36)
    monkeypatch.chdir(tmp_path)
    # -v: only regressor 2 permuted, cumulatively, contrasts written 1..(stop-start+1) like the reference
    opts = drv.getArgumentParser(argparse.ArgumentParser()).parse_args(
        ["-r", "1", "3", "-s", "thickness", "-v", "2", "2", "--seed", "5"])
    drv.run(opts)
    n, y = st["n"], st["y"]
    X = np.column_stack([np.ones(n), st["pred_x"]])
    run = helpers.oracle_run(2, 0.67, st["csr"])
    want = []
    for p in range(1, 4):
        np.random.seed(p * 1000 + 5)
        X[:, 2:3] = X[:, 2:3][np.random.permutation(list(range(n)))]
        t = oracle.tval_int(X, np.linalg.inv(X.T @ X), y, n, 3, y.shape[1])
        for sign in (1, -1):
            want.append("%.4f" % oracle.perm_max_vertex(t[1] * sign, int(st["keep_lh"].sum()), st["keep_lh"],
                                                        st["keep_rh"], run, run, st["dens"], st["dens"]))
    got = [l.strip() for l in open("output_thickness/perm_Tstat_thickness/perm_tstat_con1_TFCE_maxVertex.csv")]
    assert got == want


@pytest.mark.parametrize("medtype", ["M", "Y"])
def test_vertex_mediation_driver_rows(tmp_path, monkeypatch, medtype):
    from tfce_mediation_b200.tmanalysis import vertex_tfce_mediation_randomise as drv
    st = _vertex_state(str(tmp_path), "area", k=2, med=True)
    monkeypatch.chdir(tmp_path)
    opts = drv.getArgumentParser(argparse.ArgumentParser()).parse_args(
        ["-r", "1", "5", "-s", "area", "-m", medtype, "--seed", "2"])
    drv.run(opts)
    n, y = st["n"], st["y"]
    run = helpers.oracle_run(2, 0.67, st["csr"])
    got = np.array([float(l) for l in open("output_med_area/perm_SobelZ_%s/perm_Zstat_%s_TFCE_maxVertex.csv" % (medtype, medtype))])
    for i, p in enumerate(range(1, 6)):
        idx = oracle.permutation_indices(p * 1000 + 2, n)
        xp = st["pred_x"][idx]
        dp = st["dep"][idx] if medtype == "Y" else st["dep"]
        z = oracle.sobelz(medtype, xp, dp, y, n, y.shape[1])
        want = oracle.perm_max_vertex(z, int(st["keep_lh"].sum()), st["keep_lh"], st["keep_rh"], run, run, st["dens"], st["dens"])
        assert abs(got[i] - want) <= 1e-5 * max(1.0, abs(want)) + 1e-4      # printed with 4 decimals


def test_voxel_regression_driver_rows_identical_to_reference_csv(tmp_path, monkeypatch):
    from tfce_mediation_b200 import pyfunc
    from tfce_mediation_b200.tmanalysis import voxel_tfce_multiple_regression_randomise as drv
    g = np.load(os.path.join(G, "voxel.npz"))
    monkeypatch.chdir(tmp_path)
    os.makedirs("python_temp"); os.makedirs("output")
    mask = g["mask"]
    adj = pyfunc.create_adjac_voxel(mask, mask.astype(np.float32), int(mask.sum()), dirtype=26)
    np.save("python_temp/num_voxel.npy", int(mask.sum())); np.save("python_temp/num_subjects.npy", g["y"].shape[0])
    np.save("python_temp/raw_nonzero_corr.npy", np.ascontiguousarray(g["y"].T))
    np.save("python_temp/pred_x.npy", g["X"][:, 1]); np.save("python_temp/adjac.npy", adj, allow_pickle=True)
    np.save("python_temp/ancova.npy", 0); np.save("python_temp/optstfce.npy", np.array([2, 0.5]))
    # the golden rows were generated with seeds 3001.. = iter_perm*1000 + seed  ->  iter_perm 3, seed 1..4
    rows = []
    for s in range(1, 5):
        opts = drv.getArgumentParser(argparse.ArgumentParser()).parse_args(["-r", "3", "3", "--seed", str(s)])
        drv.run(opts)
    rows = [l.strip() for l in open("output/perm_Tstat/perm_tstat_con1_TFCE_maxVoxel.csv")]
    assert rows == list(g["rows"])


def test_mmr_lr_driver_rows_identical_to_reference_csv(tmp_path, monkeypatch):
    from tfce_mediation_b200.tm_multisurface import mmr_lr_randomise as drv
    g = np.load(os.path.join(G, "mmr_lowram.npz"))
    monkeypatch.chdir(tmp_path)
    os.makedirs("tmi_temp")
    indptr, indices = g["indptr"], g["indices"]
    adj = [indices[indptr[i]:indptr[i + 1]].tolist() for i in range(len(indptr) - 1)]
    for s in (0, 1):                                   # two copies of the surface: a 2-surface mmr-lr job
        np.save("tmi_temp/%d_data_temp.npy" % s, g["data"]); np.save("tmi_temp/%d_mask_temp.npy" % s, g["mask"])
        np.save("tmi_temp/%d_adjacency_temp.npy" % s, _obj(adj), allow_pickle=True)
        np.save("tmi_temp/%d_vdensity_temp.npy" % s, g["vdensity"])
    np.savetxt("pred.csv", g["pred_x"], delimiter=",")
    pn = g["perm_numbers"]
    opts = drv.getArgumentParser(argparse.ArgumentParser()).parse_args(
        ["--path", "out", "-pr", str(pn[0]), str(pn[-1]), "--seed", str(int(g["perm_seed"])), "-i", "pred.csv",
         "--tfce", "2", "0.67"])
    drv.run(opts)
    for s in (0, 1):
        assert [l.strip() for l in open("out/perm_maxTFCE_surf%d_tcon1.csv" % s)] == list(g["rows_tcon1"])
        assert [l.strip() for l in open("out/perm_maxTFCE_surf%d_tcon2.csv" % s)] == list(g["rows_tcon2"])


@pytest.mark.parametrize("medtype", ["M", "Y"])
def test_mmr_lr_mediation_driver_matches_per_surface_dropin(tmp_path, monkeypatch, medtype):
    """mmr-lr mediation: the batched driver (all surfaces of a block of shuffles at once) writes the rows the
    per-surface drop-in low_ram_calculate_mediation_tfce (tm_func.py:269-305) writes shuffle by shuffle."""
    from tfce_mediation_b200 import tm_func
    from tfce_mediation_b200.tfce import CreateAdjSet
    from tfce_mediation_b200.tm_multisurface import mmr_lr_randomise as drv
    g = np.load(os.path.join(G, "mmr_lowram.npz"))
    monkeypatch.chdir(tmp_path)
    os.makedirs("tmi_temp"); os.makedirs("want")
    indptr, indices = g["indptr"], g["indices"]
    adj = [indices[indptr[i]:indptr[i + 1]].tolist() for i in range(len(indptr) - 1)]
    data, mask, vdens = g["data"], g["mask"], g["vdensity"]
    n = data.shape[0]
    rs = np.random.RandomState(11)
    pred = rs.standard_normal(n); dep = 0.4 * pred + rs.standard_normal(n)
    for s in (0, 1):
        np.save("tmi_temp/%d_data_temp.npy" % s, data * (1 + s)); np.save("tmi_temp/%d_mask_temp.npy" % s, mask)
        np.save("tmi_temp/%d_adjacency_temp.npy" % s, _obj(adj), allow_pickle=True)
        np.save("tmi_temp/%d_vdensity_temp.npy" % s, vdens)
    np.savetxt("pred.csv", pred, delimiter=","); np.savetxt("dep.csv", dep, delimiter=",")
    opts = drv.getArgumentParser(argparse.ArgumentParser()).parse_args(
        ["--path", "out", "-pr", "3", "8", "--seed", "42", "-im", medtype, "pred.csv", "dep.csv", "--tfce", "2", "0.67"])
    drv.run(opts)
    calc = CreateAdjSet(2, 0.67, adj)
    for s in (0, 1):
        for p in range(3, 9):
            tm_func.low_ram_calculate_mediation_tfce(medtype, data * (1 + s), mask, pred, dep, calc, vdens, set_surf_count=s,
                                                     perm_number=p, randomise=True, output_dir="want", perm_seed=42)
        got = [float(l) for l in open("out/perm_maxTFCE_surf%d_%s_zstat.csv" % (s, medtype))]
        want = [float(l) for l in open("want/perm_maxTFCE_surf%d_%s_zstat.csv" % (s, medtype))]
        assert len(got) == len(want) == 6
        np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-5)
