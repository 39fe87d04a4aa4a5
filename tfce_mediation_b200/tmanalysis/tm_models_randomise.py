#!/usr/bin/env python
"""Permutation testing for tm-models -- drop-in for the reference's tmanalysis/tm_models_randomise.py:70-677, all six
model families: GLM (-glm), mediation (-med), cosinor (-cos), cosinor mediation (-mcos) and the repeated-measures ANCOVA
with one or two between-subject factors (-ofa, -tfa): same options, same
tmtemp_<model>_<surface|volume>/ inputs, same output_<model>_*/perm_*/perm_<stat>_TFCE_max{Vertex,Voxel}.csv rows
('%.4f'; for t statistics +t then -t per shuffle).  Every shuffle permutes whole rows of the design [1, exog..., covariates] (pyfunc.py:2317-2321); blocks
of shuffles run through the batched GPU engine (one fused fit per block: partial F of every variable from the extra
sum of squares, t of the variables' columns), then TFCE and the scaled maximum.  Under torchrun the permutation range
is sharded across ranks and rank 0 writes the rows in order.  -med is run_mediation, -cos run_cosinor, -mcos run_cosinor_mediation and -ofa / -tfa
run_rm_ancova below."""
import argparse as ap
import os
from time import time

import numpy as np

from . import _common as C
from .. import parallel
from ..pyfunc import check_blocks, rand_blocks, typeI_design

DESCRIPTION = "Permutation testing for tm-models"


def getArgumentParser(ap=ap.ArgumentParser(description=DESCRIPTION)):
    ap.add_argument("-r", "--range", nargs=2, type=int, help="permutation [start] [stop]", metavar=('INT', 'INT'),
                    required=True)
    modality = ap.add_mutually_exclusive_group(required=True)
    modality.add_argument("-s", "--surface", nargs=1, help="Randomise surface analysis. -s {surface}", metavar=('STR'))
    modality.add_argument("-v", "--voxel", action='store_true', help="Randomise volumetric analysis")
    modality.add_argument("-t", "--tmi", action='store_true', help="Randomise multimodal TMI analysis")
    stat = ap.add_mutually_exclusive_group(required=False)
    stat.add_argument("-glm", "--generalizedlinearmodel", action='store_true')
    stat.add_argument("-med", "--mediation", action='store_true')
    stat.add_argument("-ofa", "--onebetweenssubjectfactor", action='store_true')
    stat.add_argument("-tfa", "--twobetweenssubjectfactor", action='store_true')
    stat.add_argument("-cos", "--cosinor", action='store_true')
    stat.add_argument("-mcos", "--cosinormediation", action='store_true')
    ap.add_argument("-e", "--exchangeblock", nargs=1, help="Exchangability blocks", metavar=('*.csv'), required=False)
    ap.add_argument("--seed", type=int, default=None,
                    help="Reproducible stream: np.random.seed(iter_perm*1000 + SEED) before each draw (the reference "
                         "draws from the unseeded global stream)")
    return ap


def run(opts):
    start_time = time()
    C.setup()            # bind cuda:LOCAL_RANK and join the process group before any device state exists
    np.seterr(divide="ignore", invalid="ignore")
    from ..engine import PermutationEngine
    if opts.mediation:
        return run_mediation(opts, start_time)
    if opts.cosinor:
        return run_cosinor(opts, start_time)
    if opts.cosinormediation:
        return run_cosinor_mediation(opts, start_time)
    if opts.onebetweenssubjectfactor or opts.twobetweenssubjectfactor:
        return run_rm_ancova(opts, start_time)
    if not opts.generalizedlinearmodel:
        raise SystemExit("tm_models_randomise: choose one of -glm, -med, -cos, -mcos, -ofa, -tfa")
    if opts.tmi:
        raise NotImplementedError("tm_models_randomise: TMI input (-t) is not built on the B200 path")
    first, last = int(opts.range[0]), int(opts.range[1])
    if opts.surface:
        surface = str(opts.surface[0])
        tempdir, outdir = "tmtemp_GLM_%s" % surface, "output_GLM_%s/perm_GLM" % surface
    else:
        tempdir, outdir = "tmtemp_GLM_volume", "output_GLM_volume/perm_GLM"
    exog_flat = C.load("%s/exog_flat.npy" % tempdir)
    exog_shape = C.load("%s/exog_shape.npy" % tempdir)
    exog, count = [], 0
    for nc in exog_shape:
        exog.append(exog_flat[:, count:(count + nc)])
        count += nc
    varnames = C.load("%s/varnames.npy" % tempdir)
    gstat = str(np.asarray(C.load("%s/gstat.npy" % tempdir)).reshape(-1)[0])
    data = C.load("%s/data.npy" % tempdir)
    optstfce = C.load("%s/optstfce.npy" % tempdir)
    dmy_covariates = C.load("%s/dmy_covariates.npy" % tempdir)
    if np.all(dmy_covariates) is None or dmy_covariates.ndim == 0:
        dmy_covariates = None
    H, E = float(optstfce[0]), float(optstfce[1])
    if opts.surface:
        num_vertex_lh = int(C.load("%s/num_vertex_lh.npy" % tempdir))
        mask_lh = C.load("%s/mask_lh.npy" % tempdir)
        mask_rh = C.load("%s/mask_rh.npy" % tempdir)
        surfs = [C.masked_surface(C.load("%s/adjac_lh.npy" % tempdir), H, E, mask_lh,
                                  C.load("%s/vdensity_lh.npy" % tempdir), 0),
                 C.masked_surface(C.load("%s/adjac_rh.npy" % tempdir), H, E, mask_rh,
                                  C.load("%s/vdensity_rh.npy" % tempdir), num_vertex_lh)]
        suffix = "maxVertex"
    else:
        surfs = [C.masked_surface(C.load("%s/adjac.npy" % tempdir), H, E)]
        suffix = "maxVoxel"
    if opts.exchangeblock:
        block_list = np.genfromtxt(opts.exchangeblock[0], dtype=str)
        is_equal_sizes = check_blocks(block_list)

    n = data.shape[0]
    exog_vars, kvars = typeI_design(exog, dmy_covariates, n)
    eng = PermutationEngine(data, surfs, two_sided=True)
    rank, ws, a, b = C.shard(first, last)
    if rank == 0:
        os.makedirs(outdir, exist_ok=True)
    stat = "f" if gstat == "f" else "t" if gstat == "t" else "both"
    res_f, res_t = [], []
    for p0, p1 in C.chunks(a, b):
        idx = []
        for iter_perm in range(p0, p1 + 1):
            if opts.seed is not None:
                np.random.seed(int(iter_perm * 1000 + opts.seed))
            idx.append(rand_blocks(block_list, is_equal_sizes) if opts.exchangeblock else C.draw_row_permutation(n))
        out = eng.glm_typeI_block(exog_vars, kvars, np.stack(idx), stat=stat)
        f, t = (out, None) if stat == "f" else (None, out) if stat == "t" else out
        if f is not None:
            res_f.append(f.max(axis=2))                 # max over the surfaces -> [P, nvar]
        if t is not None:
            res_t.append(t.max(axis=2))                 # [P, ncon, 2]
    nvar, ncon = len(kvars), int(sum(kvars))
    all_f = C.gather(np.concatenate(res_f, axis=0) if res_f else np.zeros((0, nvar), dtype=np.float32)) \
        if stat != "t" else None
    all_t = C.gather(np.concatenate(res_t, axis=0) if res_t else np.zeros((0, ncon, 2), dtype=np.float32)) \
        if stat != "f" else None
    if rank == 0:
        if all_t is not None:
            for j in range(ncon):                       # tm_models_randomise.py:228-256: +T then -T per shuffle
                C.append_rows("%s/perm_Tstat_con%d_TFCE_%s.csv" % (outdir, j + 1, suffix), all_t[:, j, :].reshape(-1), "%.4f")
        if all_f is not None:
            for j in range(nvar):                       # :258-272
                C.append_rows("%s/perm_Fstat_%s_TFCE_%s.csv" % (outdir, varnames[j], suffix), all_f[:, j], "%.4f")
        print("Finished. Randomization took %.1f seconds" % (time() - start_time))


def _surfaces(opts, tempdir, H, E):
    if opts.surface:
        num_vertex_lh = int(C.load("%s/num_vertex_lh.npy" % tempdir))
        return [C.masked_surface(C.load("%s/adjac_lh.npy" % tempdir), H, E, C.load("%s/mask_lh.npy" % tempdir),
                                 C.load("%s/vdensity_lh.npy" % tempdir), 0),
                C.masked_surface(C.load("%s/adjac_rh.npy" % tempdir), H, E, C.load("%s/mask_rh.npy" % tempdir),
                                 C.load("%s/vdensity_rh.npy" % tempdir), num_vertex_lh)], "maxVertex"
    return [C.masked_surface(C.load("%s/adjac.npy" % tempdir), H, E)], "maxVoxel"


def run_mediation(opts, start_time):
    """The mediation branch (-med), tm_models_randomise.py:92-100,430-520: perm_Zstat_<medtype>_TFCE_max{Vertex,Voxel}.csv,
    one '%.4f' row per shuffle.  The reference permutes `dmy_leftvar = dmy_leftvar[rand_array]` IN PLACE, so shuffle i
    sees the composition of all draws since the start of the range; the same composition is built here (every rank
    replays the draws from the first permutation of the range)."""
    from ..engine import PermutationEngine
    if opts.tmi:
        raise NotImplementedError("tm_models_randomise: TMI input (-t) is not built on the B200 path")
    first, last = int(opts.range[0]), int(opts.range[1])
    if opts.surface:
        surface = str(opts.surface[0])
        tempdir, outdir = "tmtemp_mediation_%s" % surface, "output_mediation_%s/perm_mediation" % surface
    else:
        tempdir, outdir = "tmtemp_mediation_volume", "output_mediation_volume/perm_mediation"
    dmy_leftvar = C.load("%s/dmy_leftvar.npy" % tempdir)
    dmy_rightvar = C.load("%s/dmy_rightvar.npy" % tempdir)
    medtype = str(np.asarray(C.load("%s/medtype.npy" % tempdir)).reshape(-1)[0])
    data = C.load("%s/data.npy" % tempdir)
    optstfce = C.load("%s/optstfce.npy" % tempdir)
    dmy_covariates = C.load("%s/dmy_covariates.npy" % tempdir)
    if np.all(dmy_covariates) is None or dmy_covariates.ndim == 0:
        dmy_covariates = None
    surfs, suffix = _surfaces(opts, tempdir, float(optstfce[0]), float(optstfce[1]))
    if opts.exchangeblock:
        block_list = np.genfromtxt(opts.exchangeblock[0], dtype=str)
        is_equal_sizes = check_blocks(block_list)
    n = data.shape[0]
    eng = PermutationEngine(data, surfs, two_sided=False)
    rank, ws, a, b = C.shard(first, last)
    if rank == 0:
        os.makedirs(outdir, exist_ok=True)
    composed = np.arange(n)
    idx = []
    for iter_perm in range(first, b + 1):
        if opts.seed is not None:
            np.random.seed(int(iter_perm * 1000 + opts.seed))
        rand_array = rand_blocks(block_list, is_equal_sizes) if opts.exchangeblock else C.draw_row_permutation(n)
        composed = composed[rand_array]              # x[r1][r2] == x[r1[r2]]
        if iter_perm >= a:
            idx.append(composed)
    res = []
    for c0 in range(0, len(idx), C.BLOCK):
        res.append(eng.tm_models_mediation_block(medtype, dmy_leftvar, dmy_rightvar, dmy_covariates,
                                                 np.stack(idx[c0:c0 + C.BLOCK])).max(axis=1))
    local = np.concatenate(res, axis=0) if res else np.zeros((0,), dtype=np.float32)
    allrows = C.gather(local)
    if rank == 0:
        C.append_rows("%s/perm_Zstat_%s_TFCE_%s.csv" % (outdir, medtype, suffix), allrows, "%.4f")
        print("Finished. Randomization took %.1f seconds" % (time() - start_time))


def _common_inputs(opts, model, permdir):
    """tempdir / outdir names and the inputs every branch loads (tm_models_randomise.py:101-190)."""
    if opts.tmi:
        raise NotImplementedError("tm_models_randomise: TMI input (-t) is not built on the B200 path")
    where = str(opts.surface[0]) if opts.surface else "volume"
    tempdir, outdir = "tmtemp_%s_%s" % (model, where), "output_%s_%s/%s" % (model, where, permdir)
    data = C.load("%s/data.npy" % tempdir)
    optstfce = C.load("%s/optstfce.npy" % tempdir)
    dmy_covariates = C.load("%s/dmy_covariates.npy" % tempdir)
    if np.all(dmy_covariates) is None or dmy_covariates.ndim == 0:
        dmy_covariates = None
    surfs, suffix = _surfaces(opts, tempdir, float(optstfce[0]), float(optstfce[1]))
    blocks = None
    if opts.exchangeblock:
        block_list = np.genfromtxt(opts.exchangeblock[0], dtype=str)
        blocks = (block_list, check_blocks(block_list))
    return tempdir, outdir, data, dmy_covariates, surfs, suffix, blocks


def _draws(opts, a, b, n, blocks):
    """One whole-row permutation per shuffle of [a, b], drawn as the reference does (:276-279)."""
    idx = []
    for iter_perm in range(a, b + 1):
        if opts.seed is not None:
            np.random.seed(int(iter_perm * 1000 + opts.seed))
        idx.append(rand_blocks(*blocks) if blocks is not None else C.draw_row_permutation(n))
    return np.stack(idx)


def run_cosinor(opts, start_time):
    """The cosinor branch (-cos), tm_models_randomise.py:103-123,274-381: per shuffle one row each in
    perm_Fstat_model_*, perm_Tstat_amplitude_<period>_*, perm_Tstat_acrophase_<period>_*, and +t then -t in
    perm_Tstat_<name | con<j>>_* for every tested column."""
    from ..engine import PermutationEngine
    tempdir, outdir, data, dmy_covariates, surfs, suffix, blocks = _common_inputs(opts, "cosinor", "perm_cosinor")
    first, last = int(opts.range[0]), int(opts.range[1])
    time_var = C.load("%s/time_var.npy" % tempdir)
    period = [float(x) for x in np.asarray(C.load("%s/period.npy" % tempdir)).reshape(-1)]
    exog_flat = C.load("%s/exog_flat.npy" % tempdir)
    exog, varnames = None, []
    if not (exog_flat.ndim == 0 or np.all(exog_flat) is None):
        exog, count = [], 0
        for nc in C.load("%s/exog_shape.npy" % tempdir):
            exog.append(exog_flat[:, count:(count + nc)])
            count += nc
        varnames = C.load("%s/varnames.npy" % tempdir)
    n = data.shape[0]
    eng = PermutationEngine(data, surfs, two_sided=True)
    rank, ws, a, b = C.shard(first, last)
    if rank == 0:
        os.makedirs(outdir, exist_ok=True)
    nper = len(period)
    numcon = int(np.concatenate(exog, 1).shape[1]) if exog is not None else 0
    res_pos, res_t = [], []
    for p0, p1 in C.chunks(a, b):
        pos, tex = eng.cosinor_block(time_var, period, exog, dmy_covariates, _draws(opts, p0, p1, n, blocks))
        res_pos.append(pos.max(axis=2))                  # max over the surfaces -> [P, 1 + 2*nper]
        if tex is not None:
            res_t.append(tex.max(axis=2))                # [P, numcon, 2]
    all_pos = C.gather(np.concatenate(res_pos, axis=0) if res_pos else np.zeros((0, 1 + 2 * nper), dtype=np.float32))
    all_t = C.gather(np.concatenate(res_t, axis=0) if res_t else np.zeros((0, numcon, 2), dtype=np.float32)) \
        if numcon else None
    if rank == 0:
        C.append_rows("%s/perm_Fstat_model_TFCE_%s.csv" % (outdir, suffix), all_pos[:, 0], "%.4f")
        for i, per in enumerate(period):
            C.append_rows("%s/perm_Tstat_amplitude_%2.2f_TFCE_%s.csv" % (outdir, per, suffix), all_pos[:, 1 + 2 * i], "%.4f")
            C.append_rows("%s/perm_Tstat_acrophase_%2.2f_TFCE_%s.csv" % (outdir, per, suffix), all_pos[:, 2 + 2 * i], "%.4f")
        for j in range(numcon):                          # :325-350: names when there is one per column, else con<j>
            name = "Tstat_%s" % varnames[j] if len(varnames) == numcon else "Tstat_con%d" % (j + 1)
            C.append_rows("%s/perm_%s_TFCE_%s.csv" % (outdir, name, suffix), all_t[:, j, :].reshape(-1), "%.4f")
        print("Finished. Randomization took %.1f seconds" % (time() - start_time))


def run_cosinor_mediation(opts, start_time):
    """The cosinor mediation branch (-mcos), tm_models_randomise.py:124-134,383-426: perm_Zstat_<medtype>_TFCE_*.csv, one
    row per shuffle.  (Without -e the reference draws np.random.permutation(range(dmy_leftvar.shape[0])) with
    dmy_leftvar never loaded in this branch and stops with a NameError; the number of subjects is used here.)"""
    from ..engine import PermutationEngine
    tempdir, outdir, data, _, surfs, suffix, blocks = _common_inputs(opts, "medcosinor", "perm_cosinor")
    first, last = int(opts.range[0]), int(opts.range[1])
    time_var = C.load("%s/time_var.npy" % tempdir)
    period = [float(x) for x in np.asarray(C.load("%s/period.npy" % tempdir)).reshape(-1)]
    dmy_mediator = C.load("%s/dmy_mediator.npy" % tempdir)
    medtype = str(np.asarray(C.load("%s/medtype.npy" % tempdir)).reshape(-1)[0])
    n = data.shape[0]
    eng = PermutationEngine(data, surfs, two_sided=False)
    rank, ws, a, b = C.shard(first, last)
    if rank == 0:
        os.makedirs(outdir, exist_ok=True)
    res = []
    for p0, p1 in C.chunks(a, b):
        res.append(eng.cosinor_mediation_block(time_var, period, dmy_mediator, _draws(opts, p0, p1, n, blocks)).max(axis=1))
    allrows = C.gather(np.concatenate(res, axis=0) if res else np.zeros((0,), dtype=np.float32))
    if rank == 0:
        C.append_rows("%s/perm_Zstat_%s_TFCE_%s.csv" % (outdir, medtype, suffix), allrows, "%.4f")
        print("Finished. Randomization took %.1f seconds" % (time() - start_time))


def run_rm_ancova(opts, start_time):
    """The repeated-measures ANCOVA branches (-ofa / -tfa), tm_models_randomise.py:135-158,522-677: per shuffle one row in
    perm_Fstat_<factor>_*, perm_Fstat_time_*, perm_Fstat_<factor>.X.time_* (and for two factors the second factor, the
    factors' interaction and their interactions with time).  The reference shuffles the rows of the long-format data IN
    PLACE every iteration (pyfunc.py:1826,2148), so shuffle i sees the composition of all shuffles since the start of the
    range; the same composition of row orders is replayed here (every rank from the first permutation of the range)
    while the data stay put in HBM.  ('long' data: the reference computes the number of intervals as a float and stops
    at range(s-1) under Python 3; the integer is used here.)"""
    from ..engine import PermutationEngine
    from ..rmancova import RmAncovaModel
    two = bool(opts.twobetweenssubjectfactor)
    model_name = "rmANCOVA2BS" if two else "rmANCOVA1BS"
    tempdir, outdir, data, dmy_covariates, surfs, suffix, blocks = _common_inputs(opts, model_name, "perm_" + model_name)
    first, last = int(opts.range[0]), int(opts.range[1])
    dmy_factor1 = C.load("%s/dmy_factor1.npy" % tempdir)
    dmy_factor2 = C.load("%s/dmy_factor2.npy" % tempdir) if two else None
    factors = [str(f) for f in np.asarray(C.load("%s/factors.npy" % tempdir)).reshape(-1)]
    dmy_subjects = C.load("%s/dmy_subjects.npy" % tempdir)
    dformat = str(np.asarray(C.load("%s/dformat.npy" % tempdir)).reshape(-1)[0])
    n = len(dmy_factor1)
    if dformat == "short":
        if data.ndim == 2:
            data = data[:, :, np.newaxis]
        s = data.shape[0]
        data = data.reshape(s * n, data.shape[2])
    elif dformat == "long":
        s = int(len(data) // n)
    else:
        raise ValueError("data format must be short or long")
    N = s * n
    model = RmAncovaModel(n, s, [dmy_factor1, dmy_factor2] if two else [dmy_factor1], dmy_subjects, dmy_covariates)
    eng = PermutationEngine(data, surfs, two_sided=False)
    rank, ws, a, b = C.shard(first, last)
    if rank == 0:
        os.makedirs(outdir, exist_ok=True)
    pi = np.arange(N)
    shuffles, rands = [], []
    for iter_perm in range(first, b + 1):
        if opts.seed is not None:
            np.random.seed(int(iter_perm * 1000 + opts.seed))
        rand_array = rand_blocks(*blocks) if blocks is not None else C.draw_row_permutation(n)
        np.random.shuffle(pi)                        # the draws of shuffling the data rows themselves
        if iter_perm >= a:
            shuffles.append(pi.copy())
            rands.append(rand_array)
    res = []
    for c0 in range(0, len(shuffles), C.BLOCK):
        res.append(eng.rm_ancova_block(model, shuffles[c0:c0 + C.BLOCK], rands[c0:c0 + C.BLOCK]).max(axis=2))
    allrows = C.gather(np.concatenate(res, axis=0) if res else np.zeros((0, model.nout), dtype=np.float32))
    if rank == 0:
        if two:          # :604-677, the reference's return order F_a, F_b, F_ab, F_s, F_sa, F_sb, F_sab; factors.npy holds
            f1, f2 = factors[0], factors[2]          # [name1, type1, name2, type2]
            names = [f1, f2, "%s.X.%s" % (f1, f2), "time", "%s.X.time" % f1, "%s.X.time" % f2, "%s.X.%s.X.time" % (f1, f2)]
        else:            # :537-574
            names = [factors[0], "time", "%s.X.time" % factors[0]]
        for j, name in enumerate(names):
            C.append_rows("%s/perm_Fstat_%s_TFCE_%s.csv" % (outdir, name, suffix), allrows[:, j], "%.4f")
        print("Finished. Randomization took %.1f seconds" % (time() - start_time))


if __name__ == "__main__":
    parser = getArgumentParser()
    run(parser.parse_args())
