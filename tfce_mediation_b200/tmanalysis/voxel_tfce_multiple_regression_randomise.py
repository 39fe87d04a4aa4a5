#!/usr/bin/env python
"""Permutation testing for voxel-wise multiple regression with TFCE -- drop-in for the reference's
tmanalysis/voxel_tfce_multiple_regression_randomise.py (same options, python_temp/ inputs and
output/perm_Tstat/perm_tstat_con{j}_TFCE_maxVoxel.csv rows, '%1.4f', +t then -t per shuffle; the
ANCOVA branch writes perm_fstat_TFCE_maxVoxel.csv)."""
import argparse as ap
import os
from time import time

import numpy as np

from . import _common as C
from .. import parallel

DESCRIPTION = "Permutation testing for voxel-wise multiple regression with TFCE"


def getArgumentParser(ap=ap.ArgumentParser(description=DESCRIPTION)):
    ap.add_argument("-r", "--range", nargs=2, type=int, help="permutation [start] [stop]", metavar=('INT', 'INT'),
                    required=True)
    ap.add_argument("-v", "--specifyvars", nargs=2, type=int,
                    help="Optional. Specify which regressors are permuted [first] [last]. For one variable, first=last.",
                    metavar=('INT', 'INT'))
    ap.add_argument("-e", "--exchangeblock", nargs=1, help="Exchangability blocks", metavar=('*.csv'), required=False)
    ap.add_argument("--seed", type=int, default=None,
                    help="Reproducible stream: seed = iter_perm*1000 + SEED instead of the reference's time()")
    return ap


def run(opts):
    start_time = time()
    C.setup()            # bind cuda:LOCAL_RANK and join the process group before any device state exists
    np.seterr(divide="ignore", invalid="ignore")
    from ..engine import PermutationEngine
    first, last = int(opts.range[0]), int(opts.range[1])
    if opts.exchangeblock:
        block_list = np.genfromtxt(opts.exchangeblock[0], dtype=str)
        indexer = np.array(range(len(block_list)))

    n = int(C.load('python_temp/num_subjects.npy'))
    ny = C.load('python_temp/raw_nonzero_corr.npy').T           # stored V x n (voxel_..._randomise.py:63)
    pred_x = C.load('python_temp/pred_x.npy')
    adjac = C.load('python_temp/adjac.npy')
    ancova = int(C.load('python_temp/ancova.npy'))
    optstfce = C.load('python_temp/optstfce.npy')
    H, E = float(optstfce[0]), float(optstfce[1])

    X = np.column_stack([np.ones(n), pred_x])
    k = X.shape[1]
    surf = [C.masked_surface(adjac, H, E)]
    outdir = 'output/perm_Tstat'

    if ancova == 1:
        # F-statistic branch (voxel_..._randomise.py:79-89): twice the range, one-sided sqrt(F) maps.
        from ..cynumstats import calcF
        from ..engine import TfcePlan
        import torch
        last = int(last * 2)
        rank, ws, a, b = C.shard(first, last)
        if rank == 0:
            os.makedirs(outdir, exist_ok=True)
        plan = TfcePlan(surf)
        rows = []
        for iter_perm in range(a, b + 1):
            np.random.seed(C.reference_seed(iter_perm, opts.seed))
            nx = X[C.draw_row_permutation(n)]
            f = calcF(nx, ny, n, k)
            f[f < 0] = 0
            stat = torch.from_numpy(np.sqrt(f).astype(np.float32)[None, :]).cuda()
            mx, _, _ = plan.run(stat, two_sided=False)
            rows.append(mx[0, 0, 0].item())
        allrows = C.gather(np.asarray(rows, dtype=np.float32).reshape(-1, 1))
        if rank == 0:
            C.append_rows("%s/perm_fstat_TFCE_maxVoxel.csv" % outdir, allrows.reshape(-1), "%1.4f")
            print("Finished. Randomization took %.1f seconds" % (time() - start_time))
        return

    eng = PermutationEngine(np.ascontiguousarray(ny, dtype=np.float32) if ny.dtype == np.float32 else ny, surf,
                            two_sided=True, nan_to_zero=True)     # perm_tvalues[isnan] = 0 (:109)
    rank, ws, a, b = C.shard(first, last)
    if rank == 0:
        os.makedirs(outdir, exist_ok=True)
    ncon = (opts.specifyvars[1] + 1 - opts.specifyvars[0]) if opts.specifyvars else k - 1
    results = []
    if not opts.specifyvars:
        idx = []
        for iter_perm in range(a, b + 1):
            np.random.seed(C.reference_seed(iter_perm, opts.seed))
            idx.append(C.draw_block_permutation(block_list, indexer) if opts.exchangeblock
                       else C.draw_row_permutation(n))
        if idx:
            results.append(eng.regression_blocks(X, np.stack(idx), block=C.block_for(eng))[:, :, 0, :])
    else:
        # the reference permutes the chosen columns of X IN PLACE, so shuffle i sees the composition of all draws since
        # the start of the range (:93-97); a rank whose slice starts later replays the draws it skipped (RNG calls and
        # an index gather only), so the rows are the same for any number of ranks
        s0, s1 = opts.specifyvars[0], opts.specifyvars[1] + 1
        for iter_perm in range(first, a):
            np.random.seed(C.reference_seed(iter_perm, opts.seed))
            X[:, s0:s1] = X[:, s0:s1][C.draw_row_permutation(n)]
        for p0, p1 in C.chunks(a, b):
            designs = []
            for iter_perm in range(p0, p1 + 1):
                np.random.seed(C.reference_seed(iter_perm, opts.seed))
                s0, s1 = opts.specifyvars[0], opts.specifyvars[1] + 1
                X[:, s0:s1] = X[:, s0:s1][C.draw_row_permutation(n)]
                designs.append(X.copy())
            results.append(eng.regression_block(None, designs=np.stack(designs))[:, :, 0, :])   # [P, C, 2]
    local = np.concatenate(results, axis=0) if results else np.zeros((0, k - 1, 2), dtype=np.float32)
    allrows = C.gather(local)
    if rank == 0:
        for j in range(ncon):
            C.append_rows("%s/perm_tstat_con%d_TFCE_maxVoxel.csv" % (outdir, j + 1), allrows[:, j, :].reshape(-1),
                          "%1.4f")
        print("Finished. Randomization took %.1f seconds" % (time() - start_time))


if __name__ == "__main__":
    parser = getArgumentParser()
    run(parser.parse_args())
