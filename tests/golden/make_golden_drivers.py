#!/usr/bin/env python
"""Golden rows produced by the REFERENCE'S OWN DRIVERS and mmr glue, run unmodified in the build container.

What executes: /root/reference/tfce_mediation/tmanalysis/{voxel_tfce_multiple_regression_randomise,
voxel_tfce_mediation_randomise,vertex_tfce_multiple_regression_randomise,vertex_tfce_mediation_randomise}.py `run(opts)`
and tm_func.{calculate_tfce,calculate_mediation_tfce,calc_mixed_tfce,create_full_mask,merge_adjacency_array,
create_position_array}, pyfunc.create_adjac_vertex -- on top of the reference's compiled tfce / cynumstats
(oracle/_ref).  Shims: nibabel / matplotlib stubs, np.int / np.str, np.load(allow_pickle=True),
the ragged-array fallback of pyfunc.py:74, and `time()` frozen to SEED in the modules that seed numpy's RNG from the
clock (so that `int(iter_perm*1000 + time())` == iter_perm*1000 + SEED, the stream our drivers reproduce with --seed).

Usage:  python tests/golden/make_golden_drivers.py      (rewrites tests/golden/drivers.npz and mmr_full.npz)
"""
import argparse
import importlib
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from make_golden import _RaggedNumpy, load_reference, read_rows  # noqa: E402
from tfce_mediation_b200 import synth  # noqa: E402  (input generators only)

SEED = 4242
MMR_TIME = 1700000000.25        # str(...)[-6:] == '000.25' -> time seed offset 25 (tm_func.py:57)


class _LoadNumpy(_RaggedNumpy):
    @staticmethod
    def load(path, *a, **k):
        k.setdefault("allow_pickle", True)
        return np.load(path, *a, **k)


def _driver(name):
    mod = importlib.import_module("tfce_mediation.tmanalysis." + name)
    mod.np = _LoadNumpy()
    mod.time = lambda: float(SEED)
    return mod


def _obj(lists):
    out = np.empty(len(lists), dtype=object)
    for i, a in enumerate(lists):
        out[i] = a
    return out


def main():
    ref_tfce, ref_stats, pyfunc, tm_func, fwe = load_reference()
    rs = np.random.RandomState(20261018)
    cwd = os.getcwd()
    out = {}

    # ================= voxel state (python_temp/) =====================================================
    mask = rs.rand(10, 9, 8) < 0.5
    nvox = int(mask.sum())
    pyfunc.np = _RaggedNumpy()
    try:
        adj26 = pyfunc.create_adjac_voxel(mask, mask.astype(np.float32), nvox, dirtype=26)
    finally:
        pyfunc.np = np
    n = 26
    raw = rs.standard_normal((nvox, n)).astype(np.float32)          # V x n on disk (voxel_..._randomise.py:63)
    pred2 = rs.standard_normal((n, 2))
    px = rs.standard_normal(n)
    dy = 0.5 * px + rs.standard_normal(n)
    raw_med = (raw + np.float32(0.5) * px[None, :].astype(np.float32) + np.float32(0.4) * dy[None, :].astype(np.float32))
    out.update(vox_mask=mask, vox_raw=raw, vox_pred=pred2, vox_px=px, vox_dy=dy, vox_raw_med=raw_med,
               vox_adj_indptr=np.concatenate([[0], np.cumsum([len(a) for a in adj26])]).astype(np.int64),
               vox_adj_indices=np.array([x for a in adj26 for x in a], dtype=np.int32), seed=SEED)

    def voxel_state(tmp, data, pred, ancova, depend=None):
        os.makedirs(os.path.join(tmp, "python_temp"))
        os.makedirs(os.path.join(tmp, "output"))
        sv = lambda name, a: np.save(os.path.join(tmp, "python_temp", name), a, allow_pickle=True)   # noqa: E731
        sv("num_voxel.npy", nvox); sv("num_subjects.npy", n); sv("raw_nonzero_corr.npy", data); sv("pred_x.npy", pred)
        sv("adjac.npy", _obj([list(a) for a in adj26])); sv("ancova.npy", ancova); sv("optstfce.npy", np.array([2.0, 0.5]))
        if depend is not None:
            sv("depend_y.npy", depend)

    drv = _driver("voxel_tfce_multiple_regression_randomise")
    for tag, ancova in (("vox_reg", 0), ("vox_ancova", 1)):
        with tempfile.TemporaryDirectory() as tmp:
            voxel_state(tmp, raw, pred2, ancova)
            os.chdir(tmp)
            try:
                drv.run(argparse.Namespace(range=[1, 3], specifyvars=None, exchangeblock=None))
                os.chdir(tmp)
                if ancova:
                    out[tag + "_rows"] = read_rows("output/perm_Tstat/perm_fstat_TFCE_maxVoxel.csv")
                else:
                    out[tag + "_rows_con1"] = read_rows("output/perm_Tstat/perm_tstat_con1_TFCE_maxVoxel.csv")
                    out[tag + "_rows_con2"] = read_rows("output/perm_Tstat/perm_tstat_con2_TFCE_maxVoxel.csv")
            finally:
                os.chdir(cwd)
    # -v 2 2: only regressor column 2 is permuted, cumulatively (the reference mutates X in place, :93-97)
    with tempfile.TemporaryDirectory() as tmp:
        voxel_state(tmp, raw, pred2, 0)
        os.chdir(tmp)
        try:
            drv.run(argparse.Namespace(range=[1, 3], specifyvars=[2, 2], exchangeblock=None))
            os.chdir(tmp)
            out["vox_reg_v22_rows_con1"] = read_rows("output/perm_Tstat/perm_tstat_con1_TFCE_maxVoxel.csv")
        finally:
            os.chdir(cwd)

    drv = _driver("voxel_tfce_mediation_randomise")
    for med in ("M", "I", "Y"):
        with tempfile.TemporaryDirectory() as tmp:
            voxel_state(tmp, raw_med, px, 0, dy)
            os.makedirs(os.path.join(tmp, "output_med_%s" % med))
            os.chdir(tmp)
            try:
                with np.errstate(all="ignore"):
                    drv.run(argparse.Namespace(range=[1, 4], medtype=[med]))
                os.chdir(tmp)
                out["vox_med_%s_rows" % med] = read_rows("output_med_%s/perm_SobelZ/perm_Zstat_%s_TFCE_maxVoxel.csv" % (med, med))
            finally:
                os.chdir(cwd)

    # ================= vertex state (python_temp_area/, python_temp_med_area/) ===========================
    v3, f3 = synth.icosphere(3)
    csr3 = synth.faces_to_csr(v3.shape[0], f3)
    adj_sets = pyfunc.create_adjac_vertex(v3, f3)                       # pyfunc.py:37-46
    out.update(vert_v=v3, vert_f=f3, vert_adj_indptr=np.concatenate([[0], np.cumsum([len(a) for a in adj_sets])]).astype(np.int64),
               vert_adj_indices=np.array([x for a in adj_sets for x in sorted(a)], dtype=np.int32))
    keep_lh, keep_rh = synth.cap_mask(v3, 600), synth.cap_mask(-v3, 590)
    dens = synth.vertex_density(synth.kring_csr(csr3, 2))
    nv = 30
    merge_y = np.ascontiguousarray(np.hstack([synth.subject_data(nv, csr3, 31, 2)[:, keep_lh],
                                              synth.subject_data(nv, csr3, 32, 2)[:, keep_rh]]), dtype=np.float32)
    vpred = rs.standard_normal((nv, 2))
    vpx = rs.standard_normal(nv)
    vdy = 0.6 * vpx + rs.standard_normal(nv)
    merge_y_med = (merge_y + np.float32(0.4) * vpx[:, None].astype(np.float32) + np.float32(0.3) * vdy[:, None].astype(np.float32))
    out.update(vert_keep_lh=keep_lh, vert_keep_rh=keep_rh, vert_dens=dens, vert_merge_y=merge_y, vert_pred=vpred,
               vert_px=vpx, vert_dy=vdy, vert_merge_y_med=merge_y_med)

    def vertex_state(tmp, tdir, y, pred, depend=None):
        os.makedirs(os.path.join(tmp, tdir))
        sv = lambda name, a: np.save(os.path.join(tmp, tdir, name), a, allow_pickle=True)   # noqa: E731
        sv("merge_y.npy", y); sv("num_vertex.npy", y.shape[1]); sv("num_vertex_lh.npy", int(keep_lh.sum()))
        sv("bin_mask_lh.npy", keep_lh); sv("bin_mask_rh.npy", keep_rh); sv("num_subjects.npy", nv); sv("pred_x.npy", pred)
        sv("adjac_lh.npy", _obj([sorted(a) for a in adj_sets])); sv("adjac_rh.npy", _obj([sorted(a) for a in adj_sets]))
        sv("all_vertex.npy", v3.shape[0]); sv("optstfce.npy", np.array([2.0, 0.67]))
        sv("vdensity_lh.npy", dens); sv("vdensity_rh.npy", dens)
        if depend is not None:
            sv("depend_y.npy", depend)

    drv = _driver("vertex_tfce_multiple_regression_randomise")
    with tempfile.TemporaryDirectory() as tmp:
        vertex_state(tmp, "python_temp_area", merge_y, vpred)
        os.makedirs(os.path.join(tmp, "output_area"))
        os.chdir(tmp)
        try:
            drv.run(argparse.Namespace(range=[1, 6], surface=["area"], specifyvars=None, exchangeblock=None))
            os.chdir(tmp)
            out["vert_reg_rows_con1"] = read_rows("output_area/perm_Tstat_area/perm_tstat_con1_TFCE_maxVertex.csv")
            out["vert_reg_rows_con2"] = read_rows("output_area/perm_Tstat_area/perm_tstat_con2_TFCE_maxVertex.csv")
        finally:
            os.chdir(cwd)
    drv = _driver("vertex_tfce_mediation_randomise")
    for med in ("M", "Y"):
        with tempfile.TemporaryDirectory() as tmp:
            vertex_state(tmp, "python_temp_med_area", merge_y_med, vpx, vdy)
            os.makedirs(os.path.join(tmp, "output_med_area"))
            os.chdir(tmp)
            try:
                with np.errstate(all="ignore"):
                    drv.run(argparse.Namespace(range=[1, 4], surface=["area"], medtype=[med]))
                os.chdir(tmp)
                out["vert_med_%s_rows" % med] = read_rows("output_med_area/perm_SobelZ_%s/perm_Zstat_%s_TFCE_maxVertex.csv" % (med, med))
            finally:
                os.chdir(cwd)
    np.savez_compressed(os.path.join(HERE, "drivers.npz"), **out)

    # ================= non-low-RAM mmr glue (tm_func.py:54-123,207-249,327-378,502-562) ====================
    tm_func.time = lambda: MMR_TIME
    v2, f2 = synth.icosphere(2)
    csr2 = synth.faces_to_csr(v2.shape[0], f2)
    adjA = _obj(synth.csr_to_lists(csr3))
    adjB = _obj(synth.csr_to_lists(csr2))
    maskA = keep_lh.reshape(-1, 1, 1)
    maskB = np.ones((v2.shape[0], 1, 1), dtype=bool)
    maskC = keep_rh.reshape(-1, 1, 1)
    masking_array = [maskA, maskB, maskC]
    adjacency_array = [adjA, adjB, adjA]
    full_mask = tm_func.create_full_mask(masking_array)
    position_array = tm_func.create_position_array(masking_array)
    merged = tm_func.merge_adjacency_array([0, 1, 2], adjacency_array)
    nm = 28
    yA = synth.subject_data(nm, csr3, 41, 2)[:, keep_lh]
    yB = synth.subject_data(nm, csr2, 42, 1) * np.float32(1.7)          # another scale: the global threshold step matters
    yC = synth.subject_data(nm, csr3, 43, 2)[:, keep_rh]
    merge = np.ascontiguousarray(np.hstack([yA, yB, yC]), dtype=np.float32)
    mpred = rs.standard_normal((nm, 2))
    vdens = np.hstack([dens[keep_lh], synth.vertex_density(csr2), dens[keep_rh]]).astype(np.float64)   # mmr: float64 (App. A.2)
    mm = dict(merge_y=merge, pred_x=mpred, vdensity=vdens, full_mask=np.asarray(full_mask), position_array=np.array(position_array),
              merged_indptr=np.concatenate([[0], np.cumsum([len(a) for a in merged])]).astype(np.int64),
              merged_indices=np.array([x for a in merged for x in a], dtype=np.int32),
              maskA=maskA, maskB=maskB, maskC=maskC, csr3_indptr=csr3[0], csr3_indices=csr3[1], csr2_indptr=csr2[0],
              csr2_indices=csr2[1], time_offset=25)
    calc = ref_tfce.CreateAdjSet(2.0, 0.67, merged)
    t, p, q = tm_func.calculate_tfce(merge, masking_array, mpred, calc, vdens, position_array, full_mask)
    mm.update(obs_t=t, obs_pos=p, obs_neg=q)
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)
        try:
            for pn in (1, 2, 3):
                tm_func.calculate_tfce(merge, masking_array, mpred, calc, vdens, position_array, full_mask, perm_number=pn,
                                       randomise=True)
            for sf in range(3):
                for c in (1, 2):
                    mm["rows_surf%d_tcon%d" % (sf, c)] = read_rows("perm_maxTFCE_surf%d_tcon%d.csv" % (sf, c))
        finally:
            os.chdir(cwd)
    # mediation
    mpx = rs.standard_normal(nm)
    mdy = 0.5 * mpx + rs.standard_normal(nm)
    merge_med = (merge + np.float32(0.4) * mpx[:, None].astype(np.float32) + np.float32(0.3) * mdy[:, None].astype(np.float32))
    mm.update(med_merge_y=merge_med, med_px=mpx, med_dy=mdy)
    with np.errstate(all="ignore"):
        z, tz = tm_func.calculate_mediation_tfce("M", merge_med, masking_array, mpx, mdy, calc, vdens, position_array, full_mask)
    mm.update(med_obs_z=z, med_obs_tfce=tz)
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)
        try:
            for pn in (1, 2, 3):
                with np.errstate(all="ignore"):
                    tm_func.calculate_mediation_tfce("M", merge_med, masking_array, mpx, mdy, calc, vdens, position_array,
                                                     full_mask, perm_number=pn, randomise=True)
            for sf in range(3):
                mm["med_rows_surf%d" % sf] = read_rows("perm_maxTFCE_surf%d_M_zstat.csv" % sf)
        finally:
            os.chdir(cwd)
    # mixed TFCE settings: surfaces 0, 2 -> setting 0 (2, 0.67), surface 1 -> setting 1 (2, 1.0); ndarray so that the
    # reference's `assigntfcesettings == i` selects something
    assign = np.array([0, 1, 0])
    calc_list = [ref_tfce.CreateAdjSet(2.0, 0.67, tm_func.merge_adjacency_array([0, 2], [adjA, adjB, adjA])),
                 ref_tfce.CreateAdjSet(2.0, 1.0, adjB)]
    # merge_adjacency_array always starts from adjacency_array[0] (App. B.7): for the group {0, 2} that is A, then A + offset
    tm_func.np = _RaggedNumpy()      # np.array(masking_array) is ragged (tm_func.py:350): object array under numpy 2
    mt, mp, mq = tm_func.calc_mixed_tfce(assign, merge, masking_array, position_array, vdens, mpred, calc_list)
    # NB the reference returns ONE array three times (`tvals = tfce_tvals = neg_tfce_tvals = np.zeros(...)`, :368): all
    # three hold the negative-direction TFCE values, the last ones written
    mm.update(mixed_assign=assign, mixed_t=mt, mixed_pos=mp, mixed_neg=mq)
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)
        try:
            for pn in (1, 2):
                tm_func.calc_mixed_tfce(assign, merge, masking_array, position_array, vdens, mpred, calc_list, perm_number=pn,
                                        randomise=True)
            for sf in range(3):
                for c in (1, 2):
                    mm["mixed_rows_surf%d_tcon%d" % (sf, c)] = read_rows("perm_maxTFCE_surf%d_tcon%d.csv" % (sf, c))
        finally:
            os.chdir(cwd)
    tm_func.np = np
    np.savez_compressed(os.path.join(HERE, "mmr_full.npz"), **mm)

    # ================= the `mmr -p a b` flow on a TMI container (tm_multimodality_multisurface_regression.py:409-572) ====
    # The driver itself needs nibabel/matplotlib objects at import; its randomise branch is a straight sequence of
    # reference functions, called here in its order: read_tm_filetype -> create_position_array -> merge_adjacency_array
    # -> CreateAdjSet -> create_full_mask -> density weights (:449-459) -> calculate_tfce / calculate_mediation_tfce.
    tm_io = importlib.import_module("tfce_mediation.tm_io")
    tmi = os.path.join(HERE, "sample.tmi")
    _, image_array, masking_array, _, _, _, _, _, adjacency_array, _, _ = tm_io.read_tm_filetype(tmi, verbose=False)
    position_array = tm_func.create_position_array(masking_array)
    adjacent_range = list(range(len(adjacency_array)))
    calc = ref_tfce.CreateAdjSet(2.0, 0.67, tm_func.merge_adjacency_array(adjacent_range, adjacency_array))
    fullmask = tm_func.create_full_mask(masking_array)
    vdensity = []
    for i in range(len(masking_array)):
        temp_vdensity = np.zeros((adjacency_array[adjacent_range[i]].shape[0]))
        for j in range(adjacency_array[adjacent_range[i]].shape[0]):
            temp_vdensity[j] = len(adjacency_array[adjacent_range[i]][j])
        if masking_array[i].shape[2] == 1:
            temp_vdensity = temp_vdensity[masking_array[i][:, 0, 0] == True]   # noqa: E712
        vdensity = np.hstack((vdensity, np.array((1 - (temp_vdensity / temp_vdensity.max()) + (temp_vdensity.mean() / temp_vdensity.max())), dtype=np.float32)))
    nsub = image_array[0].shape[1]
    dpred = rs.standard_normal(nsub)
    ddep = 0.5 * dpred + rs.standard_normal(nsub)
    mapped_y = image_array[0].T.astype(np.float32, order="C")
    md = dict(pred=dpred, dep=ddep, time_offset=25)
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)
        try:
            for pn in range(1, 5):
                tm_func.calculate_tfce(mapped_y, masking_array, dpred, calc, vdensity, position_array, fullmask,
                                       perm_number=pn, randomise=True)
            for sf in range(len(masking_array)):
                md["rows_surf%d_tcon1" % sf] = read_rows("perm_maxTFCE_surf%d_tcon1.csv" % sf)
        finally:
            os.chdir(cwd)
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)
        try:
            for pn in range(1, 5):
                with np.errstate(all="ignore"):
                    tm_func.calculate_mediation_tfce("M", mapped_y, masking_array, dpred, ddep, calc, vdensity, position_array,
                                                     fullmask, perm_number=pn, randomise=True)
            for sf in range(len(masking_array)):
                md["med_rows_surf%d" % sf] = read_rows("perm_maxTFCE_surf%d_M_zstat.csv" % sf)
        finally:
            os.chdir(cwd)
    np.savez_compressed(os.path.join(HERE, "mmr_driver.npz"), **md)
    print("driver / mmr goldens written to", HERE)


if __name__ == "__main__":
    main()
