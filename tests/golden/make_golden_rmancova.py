#!/usr/bin/env python
"""Golden fixture for the tm-models repeated-measures ANCOVA statistics: runs the REAL reference
pyfunc.reg_rm_ancova_one_bs_factor / reg_rm_ancova_two_bs_factor (/root/reference/tfce_mediation/pyfunc.py:1712-2280, with
the compiled cynumstats from oracle/_ref) un-permuted and the way the permutation driver calls them
(tm_models_randomise.py:522-677: one np.random.permutation draw, then the function shuffles the data rows IN PLACE with
the global stream, cumulatively over the iterations), and stores inputs and outputs in tests/golden/rmancova.npz.
Build container only (needs /root/reference); see make_golden.py for the import shims.

Usage:  python tests/golden/make_golden_rmancova.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import load_reference  # noqa: E402

SEED, ITERS = 8181, 4


def main():
    _, _, pyfunc, _, _ = load_reference()
    rs = np.random.RandomState(626262)
    s, n, V = 3, 30, 260
    grp = rs.randint(0, 3, n)
    f1 = pyfunc.dummy_code(grp, demean=True)                          # 3 levels -> [n, 2]
    f2 = pyfunc.dummy_code(rs.standard_normal(n), iscontinous=True, demean=True)     # continuous -> [n]
    subjects = pyfunc.dummy_code(np.arange(n), demean=False)          # [n, n-1]
    cov = np.column_stack([rs.standard_normal(n), (rs.rand(n) > 0.5) * 1.0])
    cov = cov - cov.mean(0)
    base = rs.standard_normal((n, V))                                 # subject effect
    data = np.stack([base * 0.8 + rs.standard_normal((n, V)) + 0.3 * t for t in range(s)]).astype(np.float32)
    data[:, :, :50] += (0.9 * (grp == 2))[None, :, None].astype(np.float32)
    data[2, :, 50:90] += (0.8 * (grp == 1))[:, None].astype(np.float32)
    out = {}
    for tag, cv in (("cov", cov), ("nocov", None)):
        out["one_%s" % tag] = np.stack(pyfunc.reg_rm_ancova_one_bs_factor(data.copy(), f1, subjects, dmy_covariates=cv,
                                                                          output_sig=False, verbose=False))
        out["two_%s" % tag] = np.stack(pyfunc.reg_rm_ancova_two_bs_factor(data.copy(), f1, f2, subjects, dmy_covariates=cv,
                                                                          output_sig=False, verbose=False))
        for kind in ("one", "two"):
            work = data.copy()                                        # shuffled in place, cumulatively
            np.random.seed(SEED)
            rows = []
            for _ in range(ITERS):
                rand_array = np.random.permutation(list(range(n)))
                if kind == "one":
                    rows.append(np.stack(pyfunc.reg_rm_ancova_one_bs_factor(
                        work, f1, subjects, dmy_covariates=cv, data_format="short", output_sig=False, verbose=False,
                        rand_array=rand_array)))
                else:
                    rows.append(np.stack(pyfunc.reg_rm_ancova_two_bs_factor(
                        work, f1, f2, subjects, dmy_covariates=cv, data_format="short", output_sig=False, verbose=False,
                        rand_array=rand_array)))
            out["perm_%s_%s" % (kind, tag)] = np.stack(rows)
    np.savez_compressed(os.path.join(HERE, "rmancova.npz"), data=data, f1=f1, f2=f2, subjects=subjects, cov=cov,
                        seed=SEED, **out)
    print("rmancova.npz written:", {k: np.shape(v) for k, v in out.items()})


if __name__ == "__main__":
    main()
