#!/usr/bin/env python
"""mmr-lr randomisation -- replaces the reference's fan-out of one `tm_multimodal mmr-lr-run` job per
(100-permutation block, surface) (tm_mmr_rand_low_ram.py:237-260, tm_mmr_rand_low_ram_parallel.py:166-207).

Reads the same tmi_temp/ state the reference's `mmr-lr` front end writes
({i}_data_temp.npy float32 n x V_i, {i}_mask_temp.npy, {i}_adjacency_temp.npy, {i}_vdensity_temp.npy,
opts.npy) and appends the same rows to <path>/perm_maxTFCE_surf{i}_tcon{c}.csv ('%f', positive then
negative per shuffle).  All surfaces of a shuffle are processed together: one fused fit over the
concatenated data and one TFCE launch over shuffles x contrasts x surfaces with per-surface (H, E).
The permutation of shuffle p is the reference's deterministic rule np.random.seed(p + seed);
np.random.permutation(range(n)) (tm_func.py:146-151), so rows are reproducible and row i of every
surface file belongs to the same shuffle (the reference's apply_mfwer assumes exactly that).
Under torchrun the range -pr a b is sharded over ranks; rank 0 writes.
"""
import argparse as ap
import os
from time import time

import numpy as np

from .. import parallel
from ..tmanalysis import _common as C

DESCRIPTION = "Batched GPU companion of mmr-lr (all surfaces of a permutation block at once)"


def getArgumentParser(ap=ap.ArgumentParser(description=DESCRIPTION)):
    ap.add_argument("-sn", "--surfacenumber", nargs="+", type=int, metavar=('INT'), required=False,
                    help="Restrict to these surfaces (default: all, as one batch)")
    ap.add_argument("-pr", "--permutationrange", nargs=2, type=int, metavar=('INT', 'INT'), required=True)
    ap.add_argument("--path", nargs=1, type=str, metavar=('STR'), required=True)
    ap.add_argument("--seed", nargs=1, type=int, metavar=('INT'), required=True)
    ap.add_argument("--tmitemp", default="tmi_temp", help="Directory written by mmr-lr (default tmi_temp)")
    ap.add_argument("-i", "--input", nargs="+", metavar=('*.csv'),
                    help="Predictor file(s); default: those recorded in tmi_temp/opts.npy")
    ap.add_argument("-im", "--inputmediation", nargs=3, metavar=('{I|M|Y}', 'pred.csv', 'dep.csv'),
                    help="Mediation instead of regression (default: tmi_temp/opts.npy): one row per shuffle and surface "
                         "in perm_maxTFCE_surf{i}_{medtype}_zstat.csv (tm_func.py:269-305)")
    ap.add_argument("--tmifile", nargs=1, metavar=('*.tmi'),
                    help="Build the tmi_temp/ state from this TMI container first (what the reference's `mmr-lr` front end "
                         "does, tm_mmr_rand_low_ram.py:138-199) when it does not exist yet")
    ap.add_argument("-c", "--covariates", nargs=1, metavar=('*.csv'), help="With --tmifile: covariates of no interest")
    ap.add_argument("--subset", nargs=1, metavar=('*.csv'), help="With --tmifile: keep the subjects whose entry is finite")
    ap.add_argument("--noweight", action="store_true", help="With --tmifile: no vertex-density weighting")
    ap.add_argument("-sa", "--setadjacencyobjs", nargs="+", type=int, metavar=('INT'),
                    help="With --tmifile: adjacency object of every mask")
    ap.add_argument("--tfce", nargs="+", type=float, help="H E [H E ...]; default: tmi_temp/opts.npy")
    ap.add_argument("--assigntfcesettings", nargs="+", type=int, help="TFCE setting index per surface")
    return ap


def setup_from_tmi(tmifile, tmitemp="tmi_temp", covariates=None, subset=None, noweight=False, setadjacencyobjs=None):
    """The memory-mapping step of the reference's mmr-lr front end (tm_mmr_rand_low_ram.py:138-199): split a TMI
    container into per-surface {i}_data_temp.npy (float32, subjects x vertices; residualised on the covariates when
    given), {i}_mask_temp.npy, {i}_adjacency_temp.npy and {i}_vdensity_temp.npy (neighbour-count density weights,
    1 - d/max(d) + mean(d)/max(d), or [1] with noweight).  Returns the number of surfaces."""
    from ..tm_func import create_position_array
    from ..tm_io import read_tm_filetype
    _, image_array, masking_array, _, _, _, _, _, adjacency_array, _, _ = read_tm_filetype(tmifile, verbose=False)
    position_array = create_position_array(masking_array)
    if setadjacencyobjs:
        if len(setadjacencyobjs) != len(masking_array):
            raise SystemExit("Error: # of masking arrays (%d) must and list of matching adjacency (%d) must be equal."
                             % (len(masking_array), len(setadjacencyobjs)))
        adjacent_range = [int(a) for a in setadjacencyobjs]
    else:
        adjacent_range = list(range(len(adjacency_array)))
    os.makedirs(tmitemp, exist_ok=True)
    for i in range(len(masking_array)):
        if not noweight:
            adj = adjacency_array[adjacent_range[i]]
            temp_vdensity = np.array([len(adj[j]) for j in range(adj.shape[0])], dtype=np.float64)
            if masking_array[i].shape[2] == 1:
                temp_vdensity = temp_vdensity[masking_array[i][:, 0, 0] == True]  # noqa: E712
        else:
            temp_vdensity = np.array([1])
        vdensity = np.array((1 - (temp_vdensity / temp_vdensity.max()) + (temp_vdensity.mean() / temp_vdensity.max())),
                            dtype=np.float32)
        np.save("%s/%s_vdensity_temp.npy" % (tmitemp, i), vdensity)
        if masking_array[i].shape[2] == 1:          # vertex image: mask of shape [V, 1, 1]
            outmask = masking_array[i][:, 0, 0]
        else:
            outmask = masking_array[i][masking_array[i] == True]  # noqa: E712
        np.save("%s/%s_mask_temp.npy" % (tmitemp, i), outmask)
    for num, j in enumerate(adjacent_range):
        np.save("%s/%s_adjacency_temp.npy" % (tmitemp, num), np.copy(adjacency_array[j]), allow_pickle=True)
    x_covars = None
    if covariates is not None:
        covars = np.genfromtxt(covariates, delimiter=',')
        x_covars = np.column_stack([np.ones(len(covars)), covars])
    keep = np.isfinite(np.genfromtxt(str(subset), delimiter=',')) if subset is not None else None
    for data_count in range(len(masking_array)):
        data_array = image_array[0][position_array[data_count]:position_array[data_count + 1], :]
        if keep is not None:
            data_array = data_array[:, keep]
        if x_covars is not None:
            from ..cynumstats import resid_covars
            merge_y = np.asarray(resid_covars(x_covars, data_array))
        else:
            merge_y = data_array.T
        np.save("%s/%s_data_temp.npy" % (tmitemp, data_count), merge_y.astype(np.float32, order="C"))
    return len(masking_array)


def load_setup(opts):
    tmp = opts.tmitemp
    sopts = None
    if os.path.exists("%s/opts.npy" % tmp):
        try:
            sopts = C.load("%s/opts.npy" % tmp).tolist()
        except Exception:
            sopts = None
    med = opts.inputmediation or (getattr(sopts, "inputmediation", None) if sopts is not None else None)
    inputs = opts.input or (getattr(sopts, "input", None) if sopts is not None else None)
    if med and not opts.input:
        pred_x = (str(med[0]), np.genfromtxt(med[1], delimiter=','), np.genfromtxt(med[2], delimiter=','))
    else:
        if not inputs:
            raise SystemExit("no predictor files: pass -i / -im or provide tmi_temp/opts.npy")
        pred_x = None
        for f in inputs:
            col = np.genfromtxt(f, delimiter=',')
            pred_x = col if pred_x is None else np.column_stack([pred_x, col])
    tfce = opts.tfce or (getattr(sopts, "tfce", None) if sopts is not None else None) or [2, 0.67]
    assign = opts.assigntfcesettings or (getattr(sopts, "assigntfcesettings", None) if sopts is not None else None)
    nsurf = 0
    while os.path.exists("%s/%d_data_temp.npy" % (tmp, nsurf)):
        nsurf += 1
    surfaces = opts.surfacenumber if opts.surfacenumber else list(range(nsurf))
    return pred_x, [float(t) for t in tfce], assign, surfaces


def run(opts):
    start_time = time()
    C.setup()            # bind cuda:LOCAL_RANK and join the process group before any device state exists
    np.seterr(divide="ignore", invalid="ignore")
    from ..engine import PermutationEngine
    tmp = opts.tmitemp
    if opts.tmifile and not os.path.exists("%s/0_data_temp.npy" % tmp):
        if parallel.world()[0] == 0:
            setup_from_tmi(opts.tmifile[0], tmp, opts.covariates[0] if opts.covariates else None,
                           opts.subset[0] if opts.subset else None, opts.noweight, opts.setadjacencyobjs)
        if parallel.world()[1] > 1:
            parallel.init_process_group()
            import torch.distributed as dist
            dist.barrier()
    pred_x, tfce, assign, surfaces = load_setup(opts)
    datas, surfs, off = [], [], 0
    for sn in surfaces:
        data = C.load("%s/%d_data_temp.npy" % (tmp, sn))
        mask = C.load("%s/%d_mask_temp.npy" % (tmp, sn))
        adjacency = C.load("%s/%d_adjacency_temp.npy" % (tmp, sn))
        vdensity = C.load("%s/%d_vdensity_temp.npy" % (tmp, sn))
        ptr = int(assign[sn] * 2) if assign else 0
        H, E = tfce[ptr], tfce[ptr + 1]
        keep = np.asarray(mask == 1)
        if keep.shape[0] != len(adjacency):        # voxel surfaces store an all-True flattened mask
            keep = None
        w = None if vdensity.shape[0] == 1 else vdensity     # --noweight stores [1]
        s = C.masked_surface(adjacency, H, E, keep, None, off)
        if w is not None:
            s.weight = np.ascontiguousarray(w, dtype=np.float32)
        surfs.append(s)
        datas.append(np.ascontiguousarray(data, dtype=np.float32))
        off += data.shape[1]
    y = np.ascontiguousarray(np.hstack(datas), dtype=np.float32)
    n = y.shape[0]
    mediation = isinstance(pred_x, tuple)
    if mediation:
        return run_mediation(opts, y, surfs, surfaces, pred_x, start_time)
    X = np.column_stack([np.ones(n), pred_x])
    k = X.shape[1]
    eng = PermutationEngine(y, surfs, two_sided=True)
    first, last = int(opts.permutationrange[0]), int(opts.permutationrange[1])
    seed = int(opts.seed[0])
    rank, ws, a, b = C.shard(first, last)
    outdir = str(opts.path[0])
    if rank == 0:
        os.makedirs(outdir, exist_ok=True)
    results = []
    idx = []
    for perm_number in range(a, b + 1):
        np.random.seed(perm_number + seed)                           # tm_func.py:147-148
        idx.append(np.random.permutation(list(range(n))))
    if idx:
        results.append(eng.regression_blocks(X, np.stack(idx), block=C.BLOCK))  # [P, C, S, 2]
    local = np.concatenate(results, axis=0) if results else np.zeros((0, k - 1, len(surfs), 2), dtype=np.float32)
    allrows = C.gather(local)
    if rank == 0:
        for si, sn in enumerate(surfaces):
            for c in range(k - 1):
                C.append_rows("%s/perm_maxTFCE_surf%d_tcon%d.csv" % (outdir, sn, c + 1),
                              allrows[:, c, si, :].reshape(-1), "%f")
        print("Surfaces %s, permutations %d -> %d took %i seconds." % (surfaces, first, last, int(time() - start_time)))


def run_mediation(opts, y, surfs, surfaces, med, start_time):
    """mmr-lr mediation (tm_mmr_rand_low_ram_parallel.py:190-206 -> tm_func.py:269-305): Sobel z of every shuffle on
    all surfaces at once, one-sided TFCE, '%f' rows in perm_maxTFCE_surf{i}_{medtype}_zstat.csv."""
    from ..engine import PermutationEngine
    medtype, pred_x, depend_y = med
    n = y.shape[0]
    eng = PermutationEngine(y, surfs, two_sided=False)
    first, last = int(opts.permutationrange[0]), int(opts.permutationrange[1])
    seed = int(opts.seed[0])
    rank, ws, a, b = C.shard(first, last)
    outdir = str(opts.path[0])
    if rank == 0:
        os.makedirs(outdir, exist_ok=True)
    idx = []
    for perm_number in range(a, b + 1):
        np.random.seed(perm_number + seed)                           # tm_func.py:281-283
        idx.append(np.random.permutation(list(range(n))))
    rows = [eng.mediation_block(medtype, pred_x, depend_y, np.stack(idx[i:i + C.BLOCK]))
            for i in range(0, len(idx), C.BLOCK)]
    local = np.concatenate(rows, axis=0) if rows else np.zeros((0, len(surfs)), dtype=np.float32)
    allrows = C.gather(local.reshape(local.shape[0], 1, -1))
    if rank == 0:
        for si, sn in enumerate(surfaces):
            C.append_rows("%s/perm_maxTFCE_surf%d_%s_zstat.csv" % (outdir, sn, medtype), allrows[:, 0, si], "%f")
        print("Surfaces %s, permutations %d -> %d took %i seconds." % (surfaces, first, last, int(time() - start_time)))


if __name__ == "__main__":
    parser = getArgumentParser()
    run(parser.parse_args())
