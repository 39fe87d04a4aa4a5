#!/usr/bin/env python
"""Non-low-RAM `tm_multimodal mmr ... -p a b` randomisation and its fan-out `mmr-parallel -n N`
(tm_multimodality_multisurface_regression.py:405-572, tm_mmr_randomise_parallel.py:60-153).

The reference reads the whole TMI container, merges all surfaces' adjacency into ONE block-diagonal graph
(merge_adjacency_array, tm_func.py:521-540) and, per shuffle, calls calculate_tfce / calculate_mediation_tfce /
calc_mixed_tfce (tm_func.py:54-123,207-249,327-378): one TFCE over the merged image -- so the threshold step is the
maximum over ALL surfaces (of a TFCE-setting group) / 100 -- then a per-surface rescale and one '%f' row per surface,
contrast and sign in output_<name>/output_stats_<name>.tmi/perm_maxTFCE_surf{s}_tcon{c}.csv.  `mmr-parallel` writes one
`mmr -p a b` command per 100 shuffles for GNU parallel / HTCondor / fsl_sub.

Here the whole range is one batched run: every surface is a surface of ONE TFCE plan whose threshold tables are built
from the group's maximum (engine.TfcePlan threshold_groups), shuffles go through the fused fit + TFCE pipeline in blocks,
and under torchrun the range is sharded over the ranks (rank 0 writes the rows in permutation order, which is also the
row alignment apply_mfwer assumes, tm_func.py:422,453-463).  Seeds: the reference uses perm_number + sub-second clock
digits (tm_func.py:57); --seed S replaces the clock digits by S so that runs are reproducible."""
import argparse as ap
import os
from time import time

import numpy as np

from .. import parallel
from ..tmanalysis import _common as C

DESCRIPTION = "mmr randomisation on the GPU (all surfaces and a whole permutation range at once)"


def getArgumentParser(ap=ap.ArgumentParser(description=DESCRIPTION)):
    ap.add_argument("-i_tmi", "--tmifile", nargs=1, metavar=('*.tmi'), required=True, help="Input the *.tmi file for analysis.")
    group = ap.add_mutually_exclusive_group(required=True)
    group.add_argument("-i", "--input", nargs='+', metavar=('*.csv'), help="[Predictor(s)]")
    group.add_argument("-im", "--inputmediation", nargs=3, metavar=('{I|M|Y}', '*.csv', '*.csv'),
                       help="[Mediation Type {I,M,Y}] [Predictor] [Dependent]")
    group.add_argument("-r", "--regressors", nargs=1, metavar=('*.csv'), help="Single step regression")
    ap.add_argument("-c", "--covariates", nargs=1, metavar=('*.csv'), help="[Covariate(s)]")
    rng = ap.add_mutually_exclusive_group(required=True)
    rng.add_argument("-p", "--randomise", nargs=2, type=int, metavar=('INT', 'INT'),
                     help="Specify the range of permutations. e.g, -p 1 200")
    rng.add_argument("-n", "--numperm", nargs=1, type=int, metavar=('INT'),
                     help="mmr-parallel: # of permutations, rounded like the reference to round(N / 200) * 100 shuffles")
    ap.add_argument("-i_name", "--analysisname", nargs=1, help="Analysis name (output folder)")
    ap.add_argument("--tfce", nargs='+', default=[2.0, 0.67], type=float, help="H E [H E ...] (default: 2 0.67)")
    ap.add_argument("-sa", "--setadjacencyobjs", nargs='+', type=int, metavar=('INT'),
                    help="Adjacency object of every mask, e.g. -sa 0 1 0 1")
    ap.add_argument("-st", "--assigntfcesettings", nargs='+', type=int, metavar=('INT'),
                    help="TFCE setting (pair of --tfce values) of every mask, e.g. -st 0 0 0 0 1 1")
    ap.add_argument("--noweight", action="store_true", help="No vertex-density weighting")
    ap.add_argument("--subset", nargs=1, metavar=('*.csv'), help="Keep the subjects whose entry is finite")
    ap.add_argument("--seed", type=int, default=None, help="Reproducible stream: seed = perm_number + SEED")
    for flag, name in (("-pl", "--gnuparallel"), ("-cd", "--condor"), ("-f", "--fslsub"), ("-t", "--cmdtext")):
        ap.add_argument(flag, name, nargs='?', const=True, default=None, help="mmr-parallel scheduler option: accepted, ignored")
    return ap


def rounded_shuffles(numperm):
    """tm_mmr_randomise_parallel.py:129-131."""
    return int(np.round(numperm / 200.0) * 100.0)


def command_blocks(numperm):
    """The `-p a b` ranges the reference's fan-out writes (tm_mmr_randomise_parallel.py:134-136)."""
    return [(i * 100 + 1, i * 100 + 100) for i in range(int(rounded_shuffles(numperm) / 100))]


def density_weights(masking_array, adjacency_array, adjacent_range):
    """tm_multimodality_multisurface_regression.py:449-459: neighbour-count density per surface, float32 values in a
    float64 array (np.hstack from [])."""
    vdensity = []
    for i, m in enumerate(masking_array):
        adj = adjacency_array[adjacent_range[i]]
        d = np.array([len(adj[j]) for j in range(adj.shape[0])], dtype=np.float64)
        if m.shape[2] == 1:
            d = d[m[:, 0, 0] == True]  # noqa: E712
        vdensity = np.hstack((vdensity, np.array((1 - (d / d.max()) + (d.mean() / d.max())), dtype=np.float32)))
    return vdensity


def build(image_array, masking_array, adjacency_array, opts):
    """(engine(s), design inputs) for the arrays of one TMI container and the driver's options: a list of
    (engine, surface labels) -- one entry per TFCE-setting group, like calc_mixed_tfce -- plus merge_y's subject count."""
    from ..cynumstats import resid_covars
    from ..engine import PermutationEngine, Surface
    from ..tfce import CreateAdjSet
    from ..tm_func import _mask_piece, create_position_array
    from .._graph import adjacency_to_csr, induced_subgraph
    position_array = create_position_array(masking_array)
    nsurf = len(masking_array)
    adjacent_range = [int(a) for a in opts.setadjacencyobjs] if opts.setadjacencyobjs else list(range(len(adjacency_array)))
    if opts.setadjacencyobjs and len(adjacent_range) != nsurf:
        raise SystemExit("Error: # of masking arrays (%d) must and list of matching adjacency (%d) must be equal."
                         % (nsurf, len(adjacent_range)))
    tfce = [float(t) for t in opts.tfce]
    assign = [int(a) for a in opts.assigntfcesettings] if opts.assigntfcesettings else [0] * nsurf
    if len(assign) != nsurf:
        raise SystemExit("Error: # of masking arrays (%d) must and list of matching tfce setting (%d) must be equal."
                         % (nsurf, len(assign)))
    if len(tfce) % 2:
        raise SystemExit("Error. The must be an even number of input for --tfce")
    vdensity = 1 if opts.noweight else density_weights(masking_array, adjacency_array, adjacent_range)
    data = image_array[0]
    keep_subj = None
    if opts.subset and not opts.inputmediation:
        keep_subj = np.isfinite(np.genfromtxt(str(opts.subset[0]), delimiter=','))
        data = data[:, keep_subj]
    if opts.covariates and not opts.regressors:
        covars = np.genfromtxt(opts.covariates[0], delimiter=',')
        merge_y = resid_covars(np.column_stack([np.ones(len(covars)), covars]), data)
    else:
        merge_y = data.T
    merge_y = np.ascontiguousarray(merge_y, dtype=np.float32)                # mapped_y (:525)
    two_sided = not opts.inputmediation
    engines = []
    for gi in sorted(set(assign)):
        members = [s for s in range(nsurf) if assign[s] == gi]
        H, E = tfce[2 * gi], tfce[2 * gi + 1]
        # Which adjacency object every surface of the group runs on -- the reference's own selection, quirks included
        # (tm_multimodality_multisurface_regression.py:431-441 + tm_func.py:521-540): without -st the merge walks
        # adjacent_range over the whole adjacency list; with -st it is handed the group's SUB-lists (indexed by the
        # surface mask, so one adjacency object per surface is assumed) and adjacent_range's values index that sub-list.
        # In both cases the first block is element 0 of the list it was given.
        if opts.assigntfcesettings:
            sub_adj = [adjacency_array[s] for s in members]
            sub_range = [adjacent_range[s] for s in members]
        else:
            sub_adj, sub_range = list(adjacency_array), adjacent_range
        try:
            group_adj = [sub_adj[0]] + [sub_adj[e] for e in sub_range[1:]]
        except IndexError:
            raise SystemExit("Error: adjacency object %s does not exist for TFCE setting %d" % (sub_range, gi))
        surfs, cols, off = [], [], 0
        for pos, s in enumerate(members):
            adj = group_adj[pos]
            ip, ix = adjacency_to_csr(list(adj))
            g = induced_subgraph(ip, ix, np.asarray(_mask_piece(masking_array[s])) == 1)
            a, b = position_array[s], position_array[s + 1]
            w = None if np.ndim(vdensity) == 0 else vdensity[a:b]
            surfs.append(Surface(CreateAdjSet(H, E, g), off, w))
            cols.append(np.arange(a, b))
            off += b - a
        sub = merge_y if len(engines) == 0 and len(members) == nsurf else np.ascontiguousarray(merge_y[:, np.concatenate(cols)])
        engines.append((PermutationEngine(sub, surfs, two_sided=two_sided, threshold_groups=[list(range(len(surfs)))]), members))
    return engines, merge_y.shape[0]


def run(opts):
    start_time = time()
    C.setup()
    np.seterr(divide="ignore", invalid="ignore")
    from ..tm_io import read_tm_filetype
    _, image_array, masking_array, _, _, _, _, _, adjacency_array, _, _ = read_tm_filetype(opts.tmifile[0], verbose=False)
    engines, n = build(image_array, masking_array, adjacency_array, opts)
    if opts.numperm:
        first, last = 1, rounded_shuffles(opts.numperm[0])
        print("Evaluating %d permuations" % (last * 2))
    else:
        first, last = int(opts.randomise[0]), int(opts.randomise[1])
    if opts.inputmediation:
        medtype = str(opts.inputmediation[0])
        pred_x = np.genfromtxt(opts.inputmediation[1], delimiter=',')
        depend_y = np.genfromtxt(opts.inputmediation[2], delimiter=',')
    else:
        files = opts.input if opts.input else opts.regressors
        pred_x = None
        for f in files:
            col = np.genfromtxt(f, delimiter=',')
            pred_x = col if pred_x is None else np.column_stack([pred_x, col])
    outname = opts.analysisname[0] if (opts.input and opts.analysisname) else opts.tmifile[0][:-4]
    inner = outname if outname.endswith('tmi') else outname + '.tmi'
    inner = ('med_stats_' if opts.inputmediation else 'stats_') + inner
    outdir = os.path.join("output_%s" % outname, "output_%s" % inner)
    rank, ws, a, b = C.shard(first, last)
    if rank == 0:
        os.makedirs(outdir, exist_ok=True)
    from ..tm_func import _time_seed
    idx = []
    for perm_number in range(a, b + 1):
        np.random.seed(perm_number + opts.seed if opts.seed is not None else _time_seed(perm_number))   # tm_func.py:57
        idx.append(np.random.permutation(list(range(n))))
    idx = np.stack(idx) if idx else np.zeros((0, n), dtype=np.int64)
    for eng, members in engines:
        if opts.inputmediation:
            local = (eng.mediation_blocks(medtype, pred_x, depend_y, idx, block=C.block_for(eng)) if len(idx)
                     else np.zeros((0, len(members)), dtype=np.float32))
            allrows = C.gather(local.reshape(local.shape[0], 1, -1))
            if rank == 0:
                for si, s in enumerate(members):
                    C.append_rows("%s/perm_maxTFCE_surf%d_%s_zstat.csv" % (outdir, s, medtype), allrows[:, 0, si], "%f")
        else:
            X = np.column_stack([np.ones(n), pred_x])
            local = (eng.regression_blocks(X, idx, block=C.block_for(eng)) if len(idx)
                     else np.zeros((0, X.shape[1] - 1, len(members), 2), dtype=np.float32))
            allrows = C.gather(local)
            if rank == 0:
                for si, s in enumerate(members):
                    for c in range(X.shape[1] - 1):
                        C.append_rows("%s/perm_maxTFCE_surf%d_tcon%d.csv" % (outdir, s, c + 1),
                                      allrows[:, c, si, :].reshape(-1), "%f")
    if rank == 0:
        print("Randomization took %.1f seconds" % (time() - start_time))
    return outdir


if __name__ == "__main__":
    parser = getArgumentParser()
    run(parser.parse_args())
