#!/usr/bin/env python
"""Golden fixture for the tm-models GLM statistics: runs the REAL reference pyfunc.glm_typeI
(/root/reference/tfce_mediation/pyfunc.py:2282-2401, with the compiled cynumstats from oracle/_ref) on a small
seeded problem -- un-permuted and with `rand_array` permutations -- and stores inputs and outputs in
tests/golden/glm_typeI.npz.  Build container only (needs /root/reference); see make_golden.py for the import shims.

Usage:  python tests/golden/make_golden_glm.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import load_reference  # noqa: E402


def main():
    _, _, pyfunc, _, _ = load_reference()
    rs = np.random.RandomState(424242)
    n, V = 48, 320
    group = rs.randint(0, 3, n)                                   # a 3-level factor -> two dummy columns
    exog = [rs.standard_normal(n), np.column_stack([(group == 1) * 1.0, (group == 2) * 1.0]), rs.standard_normal((n, 1))]
    cov = np.column_stack([rs.standard_normal(n), (rs.rand(n) > 0.5) * 1.0])
    data = rs.standard_normal((n, V)).astype(np.float32)
    data[:, :40] += np.float32(0.8) * exog[0][:, None].astype(np.float32)          # some real effects
    data[:, 40:80] += np.float32(1.1) * exog[1][:, 1][:, None].astype(np.float32)
    out = {}
    F, Fvar, T = pyfunc.glm_typeI(data, exog, dmy_covariates=cov, output_tvalues=True, verbose=False)
    out["F"], out["Fvar"], out["T"] = F, Fvar, T
    F0, Fvar0, T0 = pyfunc.glm_typeI(data, exog, dmy_covariates=None, output_tvalues=True, verbose=False)
    out["F_nocov"], out["Fvar_nocov"], out["T_nocov"] = F0, Fvar0, T0
    d64 = data.astype(np.float64)           # float64 data: SS_Total is then accumulated in float64 too (pyfunc.py:2331)
    out["F_f64"], out["Fvar_f64"] = pyfunc.glm_typeI(d64, exog, dmy_covariates=cov, verbose=False)
    perms = np.stack([np.random.RandomState(7000 + p).permutation(n) for p in range(6)])
    out["perm_Fvar"] = np.stack([pyfunc.glm_typeI(data, exog, dmy_covariates=cov, verbose=False, rand_array=r)[1]
                                 for r in perms])
    out["perm_T"] = np.stack([pyfunc.glm_typeI(data, exog, dmy_covariates=cov, output_fvalues=False,
                                               output_tvalues=True, verbose=False, rand_array=r) for r in perms])
    # tm-models mediation statistic (tm_models_randomise.py:435-503): reference glm_typeI t-values + calc_indirect
    left = rs.standard_normal((n, 1))
    right = 0.6 * left + rs.standard_normal((n, 1))
    for medtype in ("I", "M", "Y"):
        zs = []
        for r in perms[:3]:
            lv = left[r]
            rv = right[r] if medtype == "Y" else right
            tv = lambda endog, ex: pyfunc.glm_typeI(endog, ex, dmy_covariates=cov, output_fvalues=False,  # noqa: E731
                                                    output_tvalues=True, verbose=False)[1]
            if medtype == "I":
                ta, tb = tv(data, [lv]), tv(data, [lv, rv])
            elif medtype == "M":
                ta, tb = tv(data, [lv]), tv(data, [rv, lv])
            else:
                ta, tb = tv(rv, [lv]), tv(data, [rv, lv])
            zs.append(pyfunc.calc_indirect(ta, tb, alg="aroian"))
        out["med_%s" % medtype] = np.stack(zs)
    out["med_left"], out["med_right"] = left, right
    np.savez_compressed(os.path.join(HERE, "glm_typeI.npz"), data=data, exog0=exog[0], exog1=exog[1], exog2=exog[2],
                        cov=cov, perms=perms, **out)
    print("glm_typeI.npz written:", {k: np.shape(v) for k, v in out.items()})


if __name__ == "__main__":
    main()
