#!/usr/bin/env python
"""Permutation testing for vertex-wise multiple regression with TFCE -- drop-in for the reference's
tmanalysis/vertex_tfce_multiple_regression_randomise.py (same options, same python_temp_<surface>/
inputs, same output_<surface>/perm_Tstat_<surface>/perm_tstat_con{j}_TFCE_maxVertex.csv rows:
+t then -t per shuffle, '%.4f').  Blocks of shuffles run through the batched GPU engine; under
torchrun the permutation range is sharded across ranks and rank 0 writes the rows in order."""
import argparse as ap
import os
from time import time

import numpy as np

from . import _common as C
from .. import parallel

DESCRIPTION = "Permutation testing for vertex-wise multiple regression with TFCE"


def getArgumentParser(ap=ap.ArgumentParser(description=DESCRIPTION)):
    ap.add_argument("-r", "--range", nargs=2, type=int, help="permutation [start] [stop]", metavar=('INT', 'INT'),
                    required=True)
    ap.add_argument("-s", "--surface", nargs=1, help="surface (area or thickness)", metavar=('STR'), required=True)
    ap.add_argument("-v", "--specifyvars", nargs=2, type=int,
                    help="Optional. Specify which regressors are permuted [first] [last]. For one variable, first=last.",
                    metavar=('INT', 'INT'))
    ap.add_argument("-e", "--exchangeblock", nargs=1, help="Exchangability blocks", metavar=('*.csv'), required=False)
    ap.add_argument("--seed", type=int, default=None,
                    help="Reproducible stream: seed = iter_perm*1000 + SEED instead of the reference's time()")
    return ap


def run(opts):
    start_time = time()
    C.tick()
    C.setup()            # bind cuda:LOCAL_RANK and join the process group before any device state exists
    C.tick("init (torch, CUDA context, process group)")
    np.seterr(divide="ignore", invalid="ignore")
    from ..engine import PermutationEngine
    first, last = int(opts.range[0]), int(opts.range[1])
    surface = str(opts.surface[0])
    tmp = "python_temp_%s" % surface
    if opts.exchangeblock:
        block_list = np.genfromtxt(opts.exchangeblock[0], dtype=str)
        indexer = np.array(range(len(block_list)))

    ny = C.load("%s/merge_y.npy" % tmp)
    num_vertex_lh = int(C.load("%s/num_vertex_lh.npy" % tmp))
    bin_mask_lh = C.load("%s/bin_mask_lh.npy" % tmp)
    bin_mask_rh = C.load("%s/bin_mask_rh.npy" % tmp)
    n = int(C.load("%s/num_subjects.npy" % tmp))
    pred_x = C.load("%s/pred_x.npy" % tmp)
    adjac_lh = C.load("%s/adjac_lh.npy" % tmp)
    adjac_rh = C.load("%s/adjac_rh.npy" % tmp)
    optstfce = C.load("%s/optstfce.npy" % tmp)
    vdensity_lh = C.load("%s/vdensity_lh.npy" % tmp)
    vdensity_rh = C.load("%s/vdensity_rh.npy" % tmp)
    H, E = float(optstfce[0]), float(optstfce[1])
    C.tick("load python_temp state")

    surfs = [C.masked_surface(adjac_lh, H, E, bin_mask_lh, vdensity_lh, 0),
             C.masked_surface(adjac_rh, H, E, bin_mask_rh, vdensity_rh, num_vertex_lh)]
    C.tick("graphs (CSR, locality order, upload)")
    eng = PermutationEngine(ny, surfs, two_sided=True)
    C.tick("engine (data upload, column order)")

    outdir = "output_%s/perm_Tstat_%s" % (surface, surface)
    rank, ws, a, b = C.shard(first, last)
    if rank == 0:
        os.makedirs(outdir, exist_ok=True)

    X = np.column_stack([np.ones(n), pred_x])
    k = X.shape[1]
    ncon = (opts.specifyvars[1] + 1 - opts.specifyvars[0]) if opts.specifyvars else k - 1
    results = []
    if not opts.specifyvars:
        # whole-row permutations: draw the index stream with the reference's RNG calls, then run the whole
        # slice through the pipelined engine
        idx = []
        for iter_perm in range(a, b + 1):
            np.random.seed(C.reference_seed(iter_perm, opts.seed))
            idx.append(C.draw_block_permutation(block_list, indexer) if opts.exchangeblock
                       else C.draw_row_permutation(n))
        C.tick("permutation index stream (numpy RNG)")
        if idx:
            results.append(eng.regression_blocks(X, np.stack(idx), block=C.block_for(eng)).max(axis=2))
        C.tick("shuffles (fit + TFCE + max)")
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
                X[:, s0:s1] = X[:, s0:s1][C.draw_row_permutation(n)]   # cumulative, like the reference (:93-97)
                designs.append(X.copy())
            mx = eng.regression_block(None, designs=np.stack(designs))
            results.append(mx.max(axis=2))             # max over the two hemispheres -> [P, C, 2]
    local = np.concatenate(results, axis=0) if results else np.zeros((0, k - 1, 2), dtype=np.float32)
    allrows = C.gather(local)
    C.tick("all-gather of the maxima")
    if rank == 0:
        for j in range(ncon):                          # the reference writes contrasts 1..ncon (:108-117)
            rows = allrows[:, j, :].reshape(-1)        # +t then -t per shuffle
            C.append_rows("%s/perm_tstat_con%d_TFCE_maxVertex.csv" % (outdir, j + 1), rows, "%.4f")
        C.tick("CSV rows")
        print("Finished. Randomization took %.1f seconds" % (time() - start_time))
    return eng


if __name__ == "__main__":
    parser = getArgumentParser()
    run(parser.parse_args())
