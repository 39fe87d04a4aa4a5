#!/usr/bin/env python
"""Permutation testing for tm-models -- drop-in for the GLM branch (-glm) of the reference's
tmanalysis/tm_models_randomise.py:70-272: same options, same tmtemp_GLM_<surface|volume>/ inputs, same
output_GLM_*/perm_GLM/perm_{Tstat_con<j>,Fstat_<name>}_TFCE_max{Vertex,Voxel}.csv rows ('%.4f'; for t statistics +t then
-t per shuffle).  Every shuffle permutes whole rows of the design [1, exog..., covariates] (pyfunc.py:2317-2321); blocks
of shuffles run through the batched GPU engine (one fused fit per block: partial F of every variable from the extra
sum of squares, t of the variables' columns), then TFCE and the scaled maximum.  Under torchrun the permutation range
is sharded across ranks and rank 0 writes the rows in order.  The other model families of the reference script
(cosinor, repeated-measures ANCOVA: tm_models_randomise.py:274-427,522-677) are not built yet and exit loudly; the
mediation branch (-med, :430-520) is run_mediation below."""
import argparse as ap
import os
from time import time

import numpy as np

from . import _common as C
from .. import parallel
from ..pyfunc import check_blocks, rand_blocks, typeI_design

DESCRIPTION = "Permutation testing for tm-models (GLM)"


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
    if not opts.generalizedlinearmodel:
        raise NotImplementedError("tm_models_randomise: only the GLM (-glm) and mediation (-med) branches are built on the "
                                  "B200 path")
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


if __name__ == "__main__":
    parser = getArgumentParser()
    run(parser.parse_args())
