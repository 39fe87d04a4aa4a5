"""Randomise drivers: same sub-command modules as the reference's tfce_mediation.tmanalysis
(getArgumentParser(ap) + run(opts)), running blocks of shuffles through the batched GPU engine."""
