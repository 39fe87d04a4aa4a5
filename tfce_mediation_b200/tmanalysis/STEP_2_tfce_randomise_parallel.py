#!/usr/bin/env python
"""Drop-in for the reference's tmanalysis/STEP_2_tfce_randomise_parallel.py (:93-162).

The reference rounds `-n N` to round(N/200)*100 shuffles (x2 for mediation), writes one
`tfce_mediation <driver> -r a b` command per block of 100 shuffles and hands the file to GNU parallel /
HTCondor / fsl_sub (one CPU process per block).  Here the blocks go through the batched GPU drivers of
this package instead: the whole range 1..roundperm is ONE call of the driver's run(); under torchrun
(one rank per GPU) the driver shards the range over the ranks (parallel.shard_range, contiguous
blocks like the reference's) and rank 0 writes the rows in permutation order.  `command_blocks`
returns the reference's own (start, stop) list, so that the two decompositions can be compared.
The scheduler options (-p/-c/-f) are accepted and ignored; the tm-models statistics
(-glm/-ofa/-tfa/-cos/-med/-mcos) go to tm_models_randomise with the reference's flags (:104-133).  (The families whose
shuffles compose in place -- tm-models mediation and repeated-measures ANCOVA -- compose over the whole range here,
where the reference's fan-out restarts the composition in every block of 100.)"""
import argparse as ap

import numpy as np

DESCRIPTION = "Wrapper for the randomise drivers: -n permutations on the local GPUs (torchrun shards them)."


def getArgumentParser(ap=ap.ArgumentParser(description=DESCRIPTION)):
    group = ap.add_mutually_exclusive_group(required=True)
    group.add_argument("--voxel", help="Voxel analysis", action="store_true")
    group.add_argument("--vertex", help="Vertex analysis. Specify surface: area or thickness", nargs=1, metavar=('surface'))
    stat = ap.add_mutually_exclusive_group(required=False)
    stat.add_argument("-m", "--mediation", nargs=1, help="Mediation type: M, Y, I", metavar=('STR'))
    for flag, name in (("-glm", "--generalizedlinearmodel"), ("-ofa", "--onebetweenssubjectfactor"),
                       ("-tfa", "--twobetweenssubjectfactor"), ("-cos", "--cosinor"), ("-med", "--modelmediation"),
                       ("-mcos", "--cosinormediation")):
        stat.add_argument(flag, name, action="store_true")
    ap.add_argument("-n", "--numperm", nargs=1, type=int, help="# of permutations", metavar=('INT'), required=True)
    ap.add_argument("-v", "--specifyvars", nargs=2, type=int, metavar=('INT', 'INT'),
                    help="Optional for multiple regression. Specify which regressors are permuted [first] [last].")
    ap.add_argument("-e", "--exchangeblock", nargs=1, help="Exchangability blocks", metavar=('*.csv'), required=False)
    sched = ap.add_mutually_exclusive_group(required=False)
    sched.add_argument("-p", "--gnuparallel", nargs=1, type=int, metavar=('INT'), help="accepted, ignored (GPU batches)")
    sched.add_argument("-c", "--condor", action="store_true", help="accepted, ignored")
    sched.add_argument("-f", "--fslsub", action="store_true", help="accepted, ignored")
    ap.add_argument("--seed", type=int, default=None, help="Reproducible stream (passed to the driver)")
    return ap


def rounded_shuffles(numperm, doubled):
    """STEP_2_tfce_randomise_parallel.py:139-143: shuffles actually run for `-n numperm`."""
    roundperm = int(np.round(numperm / 200.0) * 100.0)
    return roundperm * 2 if doubled else roundperm


def command_blocks(numperm, doubled):
    """The reference's `-r start stop` arguments, one per command line (:144-148)."""
    roundperm = rounded_shuffles(numperm, doubled)
    forperm = int(roundperm / 100) - 1
    return [(i * 100 + 1, i * 100 + 100) for i in range(forperm + 1)]


def driver_call(opts):
    """(driver module name, argv) equivalent to the reference's `whichScript -r 1 roundperm`."""
    models = [flag for flag, k in (("-glm", "generalizedlinearmodel"), ("-ofa", "onebetweenssubjectfactor"),
                                   ("-tfa", "twobetweenssubjectfactor"), ("-cos", "cosinor"), ("-med", "modelmediation"),
                                   ("-mcos", "cosinormediation")) if getattr(opts, k)]
    if models:
        # :142: the shuffle count doubles for -ofa, -tfa, -cos, -mcos (and -m), not for -glm / -med
        last = rounded_shuffles(opts.numperm[0], models[0] in ("-ofa", "-tfa", "-cos", "-mcos"))
        argv = ["-r", "1", str(last)] + (["-v"] if opts.voxel else ["-s", opts.vertex[0]]) + models
        if opts.exchangeblock:
            argv += ["-e", opts.exchangeblock[0]]
        if opts.seed is not None:
            argv += ["--seed", str(opts.seed)]
        return "tm_models_randomise", argv
    last = rounded_shuffles(opts.numperm[0], bool(opts.mediation))
    argv = ["-r", "1", str(last)]
    if opts.voxel:
        mod = "voxel_tfce_mediation_randomise" if opts.mediation else "voxel_tfce_multiple_regression_randomise"
    else:
        mod = "vertex_tfce_mediation_randomise" if opts.mediation else "vertex_tfce_multiple_regression_randomise"
        argv += ["-s", opts.vertex[0]]
    if opts.mediation:
        argv += ["-m", opts.mediation[0]]
    elif opts.specifyvars:
        argv += ["-v", str(opts.specifyvars[0]), str(opts.specifyvars[1])]
    if opts.exchangeblock:
        argv += ["-e", opts.exchangeblock[0]]
    if opts.seed is not None:
        argv += ["--seed", str(opts.seed)]
    return mod, argv


def run(opts):
    import importlib
    mod, argv = driver_call(opts)
    print("Evaluating %d permuations" % (rounded_shuffles(opts.numperm[0], False) * 2))
    drv = importlib.import_module("tfce_mediation_b200.tmanalysis." + mod)
    drv.run(drv.getArgumentParser(ap.ArgumentParser(description=drv.DESCRIPTION)).parse_args(argv))
    if opts.voxel:
        print("Run: tfce_mediation voxel-calculate-fwep to calculate (1-P[FWE]) image (after randomisation is finished).")
    else:
        print("Run: tfce_mediation vertex-calculate-fwep to calculate (1-P[FWE]) image (after randomisation is finished).")


if __name__ == "__main__":
    parser = getArgumentParser()
    run(parser.parse_args())
