#!/usr/bin/env python
"""Permutation testing for vertex-wise mediation with TFCE -- drop-in for the reference's
tmanalysis/vertex_tfce_mediation_randomise.py (python_temp_med_<surface>/ inputs,
output_med_<surface>/perm_SobelZ_<M|I|Y>/perm_Zstat_<type>_TFCE_maxVertex.csv, '%.4f', one row per shuffle)."""
import argparse as ap
import os
from time import time

import numpy as np

from . import _common as C
from .. import parallel

DESCRIPTION = "Vertex-wise mediation with TFCE"


def getArgumentParser(ap=ap.ArgumentParser(description=DESCRIPTION)):
    ap.add_argument("-r", "--range", nargs=2, type=int, help="permutation [start] [stop]", metavar=('INT', 'INT'),
                    required=True)
    ap.add_argument("-s", "--surface", nargs=1, help="surface (area or thickness)", metavar=('STR'), required=True)
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
    surface = str(opts.surface[0])
    tmp = "python_temp_med_%s" % surface
    y = C.load("%s/merge_y.npy" % tmp)
    num_vertex_lh = int(C.load("%s/num_vertex_lh.npy" % tmp))
    bin_mask_lh = C.load("%s/bin_mask_lh.npy" % tmp)
    bin_mask_rh = C.load("%s/bin_mask_rh.npy" % tmp)
    n = int(C.load("%s/num_subjects.npy" % tmp))
    pred_x = C.load("%s/pred_x.npy" % tmp)
    depend_y = C.load("%s/depend_y.npy" % tmp)
    adjac_lh = C.load("%s/adjac_lh.npy" % tmp)
    adjac_rh = C.load("%s/adjac_rh.npy" % tmp)
    optstfce = C.load("%s/optstfce.npy" % tmp)
    vdensity_lh = C.load("%s/vdensity_lh.npy" % tmp)
    vdensity_rh = C.load("%s/vdensity_rh.npy" % tmp)
    H, E = float(optstfce[0]), float(optstfce[1])
    surfs = [C.masked_surface(adjac_lh, H, E, bin_mask_lh, vdensity_lh, 0),
             C.masked_surface(adjac_rh, H, E, bin_mask_rh, vdensity_rh, num_vertex_lh)]
    eng = PermutationEngine(y, surfs, two_sided=False)
    outdir = "output_med_%s/perm_SobelZ_%s" % (surface, medtype)
    rank, ws, a, b = C.shard(first, last)
    if rank == 0:
        os.makedirs(outdir, exist_ok=True)
    # the index stream with the reference's RNG calls, then the whole slice through the pipelined engine
    idx = []
    for iter_perm in range(a, b + 1):
        np.random.seed(C.reference_seed(iter_perm, opts.seed))
        idx.append(C.draw_row_permutation(n))
    local = (eng.mediation_blocks(medtype, pred_x, depend_y, np.stack(idx), block=C.block_for(eng)).max(axis=1)   # [P, S]
             if idx else np.zeros((0,), dtype=np.float32))
    allrows = C.gather(local.reshape(-1, 1))
    if rank == 0:
        C.append_rows("%s/perm_Zstat_%s_TFCE_maxVertex.csv" % (outdir, medtype), allrows.reshape(-1), "%.4f")
        print("Finished. Randomization took %.1f seconds" % (time() - start_time))


if __name__ == "__main__":
    parser = getArgumentParser()
    run(parser.parse_args())
