#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by running the REAL reference.

Runs only in the build container (needs /root/reference).  What executes here is the
reference's own code: its compiled extension modules (oracle/_ref, built by
oracle/build_ref.py from the unmodified tfce.pyx / cynumstats.pyx / lib/fast_tfce.hpp) and its
own pure-Python glue (pyfunc.py, tm_func.py, calculate_fweP_vertex.py) imported from
/root/reference with import shims for what this container lacks:
nibabel / matplotlib stubs, np.int / np.str aliases, and a ragged-array shim for the one
np.array(...) call at pyfunc.py:74 that numpy 2 rejects.  Nothing from this repository's product
or oracle is involved in producing the numbers.

Usage:  python tests/golden/make_golden.py        (rewrites tests/golden/*.npz)
"""
import os
import sys
import tempfile
import types
from unittest import mock

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference/tfce_mediation"

from oracle import build_ref  # noqa: E402  (only its build recipe/loader is used)
from tfce_mediation_b200 import synth  # noqa: E402  (input generators only)


def load_reference():
    assert build_ref.build(), "reference extensions could not be built"
    ref_tfce, ref_stats = build_ref.load()
    for name in ("nibabel", "nibabel.freesurfer", "nibabel.freesurfer.mghformat", "matplotlib", "matplotlib.pyplot",
                 "matplotlib.colors", "matplotlib.patches"):
        sys.modules.setdefault(name, mock.MagicMock())
    if not hasattr(np, "int"):
        np.int = int
    if not hasattr(np, "str"):
        np.str = str
    pkg = types.ModuleType("tfce_mediation")
    pkg.__path__ = [REF]
    sys.modules["tfce_mediation"] = pkg
    sys.modules["tfce_mediation.tfce"] = ref_tfce
    sys.modules["tfce_mediation.cynumstats"] = ref_stats
    import importlib
    pyfunc = importlib.import_module("tfce_mediation.pyfunc")
    tm_func = importlib.import_module("tfce_mediation.tm_func")
    sub = types.ModuleType("tfce_mediation.tmanalysis")
    sub.__path__ = [os.path.join(REF, "tmanalysis")]
    sys.modules["tfce_mediation.tmanalysis"] = sub
    fwe = importlib.import_module("tfce_mediation.tmanalysis.calculate_fweP_vertex")
    return ref_tfce, ref_stats, pyfunc, tm_func, fwe


class _RaggedNumpy(object):
    """numpy proxy whose array() falls back to dtype=object for ragged input (pyfunc.py:74 under numpy 2)."""

    def __getattr__(self, name):
        return getattr(np, name)

    @staticmethod
    def array(obj, *a, **k):
        try:
            return np.array(obj, *a, **k)
        except ValueError:
            out = np.empty(len(obj), dtype=object)
            for i, o in enumerate(obj):
                out[i] = o
            return out


def read_rows(path):
    with open(path) as f:
        return np.array([line.strip() for line in f if line.strip()])


def main():
    ref_tfce, ref_stats, pyfunc, tm_func, fwe = load_reference()
    rs = np.random.RandomState(20261017)

    # ---- meshes ---------------------------------------------------------------------------------
    v3, f3 = synth.icosphere(3)                       # 642 vertices
    csr3 = synth.faces_to_csr(v3.shape[0], f3)
    adj3 = synth.csr_to_lists(csr3)
    V3 = v3.shape[0]

    # ---- 1. raw TFCE maps (tfce.pyx:44-45 -> fast_tfce.hpp) ------------------------------------------
    maps = []
    for i in range(5):
        m = rs.standard_normal((1, V3)).astype(np.float32)
        m = synth.smooth_columns(m, csr3, i)[0]
        maps.append((m / m.std() * (1 + 3 * i)).astype(np.float32))
    maps = np.stack(maps)
    he = np.array([[2, 0.67], [2, 1.0], [2, 0.5], [1.5, 0.8]], dtype=np.float64)
    out = np.zeros((len(he), len(maps), V3), dtype=np.float32)
    for a, (H, E) in enumerate(he):
        c = ref_tfce.CreateAdjSet(float(H), float(E), adj3)
        for b, m in enumerate(maps):
            c.run(np.ascontiguousarray(m), out[a, b])
    np.savez_compressed(os.path.join(HERE, "tfce_maps.npz"), indptr=csr3[0], indices=csr3[1], maps=maps, he=he,
                        tfce=out)

    # ---- 2. cynumstats on one design -------------------------------------------------------------
    n, V, k = 30, 257, 4
    y = (rs.standard_normal((n, V)) * 0.7 + 1.5).astype(np.float32)
    X = np.column_stack([np.ones(n), rs.standard_normal((n, k - 1))])
    invXX = np.linalg.inv(X.T @ X)
    beta_se = ref_stats.calc_beta_se(X[:, 1], y, n, V)
    np.savez_compressed(
        os.path.join(HERE, "cynumstats.npz"), X=X, y=y, tval=ref_stats.tval_int(X, invXX, y, n, k, V),
        beta=ref_stats.cy_lin_lstsqr_mat(X, y), resid=ref_stats.resid_covars(X, np.ascontiguousarray(y.T)),
        calcF=np.asarray(ref_stats.calcF(X, y, n, k)), cbs_beta=beta_se[0], cbs_se=beta_se[1],
        se=ref_stats.se_of_slope(V, invXX, np.abs(y[0].astype(np.float64)), k),
        sigma2=np.abs(y[0].astype(np.float64)))

    # ---- 3. vertex regression shuffles through the reference's own write_perm_maxTFCE_vertex -------
    n, k = 32, 3
    keep_lh = synth.cap_mask(v3, 600)
    keep_rh = synth.cap_mask(-v3, 590)
    dens = synth.vertex_density(synth.kring_csr(csr3, 2))
    ylh = synth.subject_data(n, csr3, 11, 2)[:, keep_lh]
    yrh = synth.subject_data(n, csr3, 12, 2)[:, keep_rh]
    merge_y = np.ascontiguousarray(np.hstack([ylh, yrh]), dtype=np.float32)
    pred_x = rs.standard_normal((n, k - 1))
    X = np.column_stack([np.ones(n), pred_x])
    seeds = np.arange(2001, 2007)
    cwd = os.getcwd()
    rows = {}
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)
        try:
            c_lh = ref_tfce.CreateAdjSet(2.0, 0.67, adj3)
            c_rh = ref_tfce.CreateAdjSet(2.0, 0.67, adj3)
            num_vertex = merge_y.shape[1]
            for seed in seeds:
                np.random.seed(int(seed))
                nx = X[np.random.permutation(list(range(n)))]
                invXX = np.linalg.inv(np.dot(nx.T, nx))
                tvals = ref_stats.tval_int(nx, invXX, merge_y, n, k, num_vertex)
                for j in range(k - 1):           # vertex_tfce_multiple_regression_randomise.py:113-117
                    pyfunc.write_perm_maxTFCE_vertex("tstat_con%d" % (j + 1), tvals[j + 1], int(keep_lh.sum()),
                                                     keep_lh, keep_rh, c_lh, c_rh, dens, dens)
                    pyfunc.write_perm_maxTFCE_vertex("tstat_con%d" % (j + 1), tvals[j + 1] * -1, int(keep_lh.sum()),
                                                     keep_lh, keep_rh, c_lh, c_rh, dens, dens)
            for j in range(k - 1):
                rows["con%d" % (j + 1)] = read_rows("perm_tstat_con%d_TFCE_maxVertex.csv" % (j + 1))
        finally:
            os.chdir(cwd)
    np.savez_compressed(os.path.join(HERE, "vertex_randomise.npz"), indptr=csr3[0], indices=csr3[1], keep_lh=keep_lh,
                        keep_rh=keep_rh, density=dens, merge_y=merge_y, pred_x=pred_x, seeds=seeds,
                        rows_con1=rows["con1"], rows_con2=rows["con2"])

    # ---- 4. voxel adjacency (pyfunc.py:48-76) + voxel shuffles (pyfunc.py:121-126) ----------------------
    mask = rs.rand(9, 10, 8) < 0.45
    nvox = int(mask.sum())
    pyfunc.np = _RaggedNumpy()
    try:
        adj26 = pyfunc.create_adjac_voxel(mask, mask.astype(np.float32), nvox, dirtype=26)
        adj6 = pyfunc.create_adjac_voxel(mask, mask.astype(np.float32), nvox, dirtype=6)
    finally:
        pyfunc.np = np
    csr26 = synth_csr([list(a) for a in adj26])
    csr6 = synth_csr([list(a) for a in adj6])
    n, k = 28, 2
    yv = rs.standard_normal((n, nvox)).astype(np.float32)
    Xv = np.column_stack([np.ones(n), rs.standard_normal(n)])
    vseeds = np.arange(3001, 3005)
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)
        try:
            c = ref_tfce.CreateAdjSet(2.0, 0.5, [list(a) for a in adj26])
            for seed in vseeds:
                np.random.seed(int(seed))
                nx = Xv[np.random.permutation(list(range(n)))]
                t = ref_stats.tval_int(nx, np.linalg.inv(np.dot(nx.T, nx)), yv, n, k, nvox)
                t[np.isnan(t)] = 0
                pyfunc.write_perm_maxTFCE_voxel("tstat_con1", t[1], c)
                pyfunc.write_perm_maxTFCE_voxel("tstat_con1", t[1] * -1, c)
            vrows = read_rows("perm_tstat_con1_TFCE_maxVoxel.csv")
        finally:
            os.chdir(cwd)
    np.savez_compressed(os.path.join(HERE, "voxel.npz"), mask=mask, indptr26=csr26[0], indices26=csr26[1],
                        indptr6=csr6[0], indices6=csr6[1], y=yv, X=Xv, seeds=vseeds, rows=vrows)

    # ---- 5. Sobel z (pyfunc.py:130-162) -------------------------------------------------------------------
    n, V = 40, 301
    px = rs.standard_normal(n)
    dy = 0.6 * px + rs.standard_normal(n)
    ym = (rs.standard_normal((n, V)) + 0.4 * px[:, None] + 0.3 * dy[:, None]).astype(np.float32)
    sob = {}
    for med in ("I", "M", "Y"):
        for alg in ("aroian", "sobel", "goodman"):
            with np.errstate(all="ignore"):
                sob["%s_%s" % (med, alg)] = pyfunc.calc_sobelz(med, px, dy, ym, n, V, alg=alg)
    np.savez_compressed(os.path.join(HERE, "sobel.npz"), pred_x=px, depend_y=dy, merge_y=ym, **sob)

    # ---- 6. mmr-lr deterministic shuffles (tm_func.py:144-185 with perm_seed) ----------------------------------
    n = 30
    mask_lr = keep_lh.astype(np.float32)                   # mask array as stored in tmi_temp (==1 test)
    data_lr = synth.subject_data(n, csr3, 21, 2)[:, keep_lh]
    pred_lr = rs.standard_normal((n, 2))
    dens_lr = dens[keep_lh]
    perm_numbers = np.arange(1, 5)
    perm_seed = 5000
    with tempfile.TemporaryDirectory() as tmp:
        c = ref_tfce.CreateAdjSet(2.0, 0.67, adj3)
        for pn in perm_numbers:
            tm_func.low_ram_calculate_tfce(data_lr, mask_lr, pred_lr, c, dens_lr, set_surf_count=0,
                                           perm_number=int(pn), randomise=True, output_dir=tmp, perm_seed=perm_seed)
        lr1 = read_rows(os.path.join(tmp, "perm_maxTFCE_surf0_tcon1.csv"))
        lr2 = read_rows(os.path.join(tmp, "perm_maxTFCE_surf0_tcon2.csv"))
    np.savez_compressed(os.path.join(HERE, "mmr_lowram.npz"), indptr=csr3[0], indices=csr3[1], mask=mask_lr,
                        data=data_lr, pred_x=pred_lr, vdensity=dens_lr, perm_numbers=perm_numbers,
                        perm_seed=perm_seed, rows_tcon1=lr1, rows_tcon2=lr2)

    # ---- 7. FWER lookup (calculate_fweP_vertex.py:37-42,61-69) -----------------------------------------------
    perm_max = np.round(np.abs(rs.standard_normal(200)) * 100, 4)
    vals = np.abs(rs.standard_normal(64)) * 120
    srt = np.sort(perm_max)
    p_array = np.zeros_like(srt)
    for j in range(len(srt)):
        p_array[j] = np.true_divide(j, len(srt))
    corr = np.array([fwe.find_nearest(srt, v, p_array) for v in vals])
    np.savez_compressed(os.path.join(HERE, "fwe.npz"), perm_max=perm_max, values=vals, corrp=corr)
    # ---- 8. mmr study-wide FWER (tm_func.py:403-491 apply_mfwer) ---------------------------------------------
    sizes = [120, 80, 100]
    position_array = np.concatenate([[0], np.cumsum(sizes)]).tolist()
    nperm, ncon = 60, 2
    img = np.abs(rs.standard_normal((sum(sizes), 4))) * 300
    img[rs.rand(*img.shape) < 0.2] = 0.5          # log <= 0 -> excluded by the reference's mask
    perm_csv = {}
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)
        try:
            os.mkdir("output_t")
            for sf in range(3):
                for c in range(ncon):
                    vals = np.abs(rs.standard_normal(nperm)) * 200 * (1 + sf) + 1
                    perm_csv["s%d_c%d" % (sf, c + 1)] = vals
                    np.savetxt("output_t/perm_maxTFCE_surf%d_tcon%d.csv" % (sf, c + 1), vals, fmt="%f")
            out = {}
            for wname, w in (("none", None), ("logmasksize", "logmasksize")):
                with np.errstate(all="ignore"):
                    pos, neg = tm_func.apply_mfwer([img.copy()], ncon, list(range(3)), nperm, 3, "t", position_array,
                                                   pos_range=[0, 1], neg_range=[2, 3], weight=w)
                out["pos_" + wname] = pos
                out["neg_" + wname] = neg
        finally:
            os.chdir(cwd)
    np.savez_compressed(os.path.join(HERE, "mfwer.npz"), image=img, position_array=np.array(position_array),
                        num_perm=nperm, **{"csv_" + k: v for k, v in perm_csv.items()}, **out)
    print("golden fixtures written to", HERE)


def synth_csr(lists):
    counts = np.array([len(a) for a in lists], dtype=np.int64)
    indptr = np.zeros(len(lists) + 1, dtype=np.int64)
    np.cumsum(counts, out=indptr[1:])
    flat = [x for a in lists for x in a]
    return indptr, np.array(flat, dtype=np.int32)


if __name__ == "__main__":
    main()
