#!/usr/bin/env python
"""Permutation testing for voxel-wise mediation with TFCE -- drop-in for the reference's
tmanalysis/voxel_tfce_mediation_randomise.py (python_temp/ inputs,
output_med_<type>/perm_SobelZ/perm_Zstat_<type>_TFCE_maxVoxel.csv, '%1.4f')."""
import argparse as ap
import os
from time import time

import numpy as np

from . import _common as C
from .. import parallel

DESCRIPTION = "Voxel-wise mediation with TFCE"


def getArgumentParser(ap=ap.ArgumentParser(description=DESCRIPTION)):
    ap.add_argument("-r", "--range", nargs=2, type=int, help="permutation [start] [stop]", metavar=('INT', 'INT'),
                    required=True)
    ap.add_argument("-m", "--medtype", nargs=1, help="mediation type [M or Y or I].", choices=['M', 'Y', 'I'],
                    required=True)
    ap.add_argument("--seed", type=int, default=None,
                    help="Reproducible stream: seed = iter_perm*1000 + SEED instead of the reference's time()")
    return ap


def run(opts):
    start_time = time()
    C.setup()            # bind cuda:LOCAL_RANK and join the process group before any device state exists
    np.seterr(divide="ignore", invalid="ignore")
    from ..engine import PermutationEngine
    first, last = int(opts.range[0]), int(opts.range[1])
    medtype = str(opts.medtype[0])
    n = int(C.load('python_temp/num_subjects.npy'))
    ny = C.load('python_temp/raw_nonzero_corr.npy').T
    pred_x = C.load('python_temp/pred_x.npy')
    depend_y = C.load('python_temp/depend_y.npy')
    adjac = C.load('python_temp/adjac.npy')
    optstfce = C.load('python_temp/optstfce.npy')
    H, E = float(optstfce[0]), float(optstfce[1])
    eng = PermutationEngine(ny, [C.masked_surface(adjac, H, E)], two_sided=False)
    outdir = "output_med_%s/perm_SobelZ" % medtype
    rank, ws, a, b = C.shard(first, last)
    if rank == 0:
        os.makedirs(outdir, exist_ok=True)
    # the index stream with the reference's RNG calls, then the whole slice through the pipelined engine
    idx = []
    for iter_perm in range(a, b + 1):
        np.random.seed(C.reference_seed(iter_perm, opts.seed))
        idx.append(C.draw_row_permutation(n))
    local = (eng.mediation_blocks(medtype, pred_x, depend_y, np.stack(idx), block=C.block_for(eng))[:, 0]
             if idx else np.zeros((0,), dtype=np.float32))
    allrows = C.gather(local.reshape(-1, 1))
    if rank == 0:
        C.append_rows("%s/perm_Zstat_%s_TFCE_maxVoxel.csv" % (outdir, medtype), allrows.reshape(-1), "%1.4f")
        print("Finished. Randomization took %.1f seconds" % (time() - start_time))


if __name__ == "__main__":
    parser = getArgumentParser()
    run(parser.parse_args())
