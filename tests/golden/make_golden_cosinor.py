#!/usr/bin/env python
"""Golden fixture for the tm-models cosinor statistics: runs the REAL reference pyfunc.glm_cosinor
(/root/reference/tfce_mediation/pyfunc.py:2406-2563, with the compiled cynumstats from oracle/_ref) the way the
permutation driver calls it (tm_models_randomise.py:274-412: calc_MESOR=False, rand_array per shuffle; for the
cosinor mediation path A un-permuted on the 1-D mediator and path B with the mediator as the one tested column) and
stores inputs and outputs in tests/golden/cosinor.npz.  Build container only (needs /root/reference); see
make_golden.py for the import shims.

Usage:  python tests/golden/make_golden_cosinor.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import load_reference  # noqa: E402


def main():
    _, _, pyfunc, _, _ = load_reference()
    rs = np.random.RandomState(515151)
    n, V = 52, 300
    time_var = rs.uniform(0.0, 24.0, n)
    period = [24.0, 12.0]
    group = rs.randint(0, 3, n)
    exog = [rs.standard_normal((n, 1)), np.column_stack([(group == 1) * 1.0, (group == 2) * 1.0])]
    cov = np.column_stack([rs.standard_normal(n), (rs.rand(n) > 0.5) * 1.0])
    data = rs.standard_normal((n, V)).astype(np.float32)
    data[:, :60] += (np.float32(1.2) * np.cos(2 * np.pi * (time_var - 5.0) / 24.0))[:, None].astype(np.float32)
    data[:, 60:100] += np.float32(0.9) * exog[0].astype(np.float32)
    perms = np.stack([np.random.RandomState(9100 + p).permutation(n) for p in range(6)])
    out = {}

    def stats(res):
        return res[7], res[9], res[10], res[11]          # Fmodel, |tAMPLITUDE|, |tACROPHASE|, tEXOG

    for tag, ex, cv, per in (("full", exog, cov, period), ("noexog", None, None, [24.0]), ("exog1", exog[:1], None, [24.0])):
        F, TA, TC, TE = [], [], [], []
        for r in perms:
            f, ta, tc, te = stats(pyfunc.glm_cosinor(endog=data, time_var=time_var, exog=ex, dmy_covariates=cv,
                                                     rand_array=r, period=per, calc_MESOR=False))
            F.append(f), TA.append(ta), TC.append(tc)
            if ex is not None:
                TE.append(te)
        out["F_" + tag], out["tamp_" + tag], out["tacr_" + tag] = np.stack(F), np.stack(TA), np.stack(TC)
        if ex is not None:
            out["texog_" + tag] = np.stack(TE)
    # the un-permuted call of STEP_1_tm_models.py (all twelve outputs) and the fit-only form
    names = ("R2", "MESOR", "SE_MESOR", "AMPLITUDE", "SE_AMPLITUDE", "ACROPHASE", "SE_ACROPHASE", "Fmodel", "tMESOR",
             "tAMPLITUDE", "tACROPHASE", "tEXOG")
    full = pyfunc.glm_cosinor(endog=data, time_var=time_var, exog=exog, dmy_covariates=cov, rand_array=None, period=period)
    for nm, val in zip(names, full):
        out["obs_" + nm] = np.asarray(val, dtype=np.float64)
    fit = pyfunc.glm_cosinor(endog=data, time_var=time_var, exog=exog, dmy_covariates=cov, period=period, output_fit_only=True)
    out["fit_MESOR"], out["fit_AMPLITUDE"], out["fit_ACROPHASE"] = (np.asarray(v, dtype=np.float64) for v in fit)
    # cosinor mediation (tm_models_randomise.py:383-412)
    mediator = 0.8 * np.cos(2 * np.pi * (time_var - 3.0) / 24.0) + 0.5 * rs.standard_normal(n)
    mediator = mediator - mediator.mean()
    ta = pyfunc.glm_cosinor(endog=mediator, time_var=time_var, exog=None, dmy_covariates=None, rand_array=None,
                            period=[24.0])[9]
    out["med_ta"] = np.asarray(ta, dtype=np.float64)
    zs = []
    for r in perms:
        tb = pyfunc.glm_cosinor(endog=data, time_var=time_var, exog=[mediator], dmy_covariates=None, rand_array=r,
                                period=[24.0])[11]
        zs.append(pyfunc.calc_indirect(ta[0], tb[0], alg="aroian"))
    out["med_z"] = np.stack(zs)
    np.savez_compressed(os.path.join(HERE, "cosinor.npz"), data=data, time_var=time_var, exog0=exog[0], exog1=exog[1],
                        cov=cov, perms=perms, mediator=mediator, period=np.array(period), **out)
    print("cosinor.npz written:", {k: np.shape(v) for k, v in out.items()})


if __name__ == "__main__":
    main()
