#!/usr/bin/env python
"""Golden fixture for the small design-coding helpers of the tm-models scripts: runs the REAL reference
pyfunc.dummy_code / dummy_code_cosine / column_product / stack_ones / calc_indirect
(/root/reference/tfce_mediation/pyfunc.py:2565-2709) on seeded inputs and stores inputs and outputs in
tests/golden/tm_models_helpers.npz.  Build container only (needs /root/reference); see make_golden.py for the shims.

Usage:  python tests/golden/make_golden_helpers.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import load_reference  # noqa: E402


def main():
    _, _, pyfunc, _, _ = load_reference()
    rs = np.random.RandomState(99)
    n = 17
    grp, two, cont, t = rs.randint(0, 4, n), rs.randint(0, 2, n), rs.standard_normal(n), rs.uniform(0, 24, n)
    a2, b3 = rs.standard_normal((n, 2)), rs.standard_normal((n, 3))
    ta, tb = rs.standard_normal(50) * 3, rs.standard_normal(50) * 3
    with np.errstate(invalid="ignore"):
        out = dict(grp=grp, two=two, cont=cont, t=t, a2=a2, b3=b3, ta=ta, tb=tb,
                   dc_grp=pyfunc.dummy_code(grp), dc_grp_raw=pyfunc.dummy_code(grp, demean=False), dc_two=pyfunc.dummy_code(two),
                   dc_cont=pyfunc.dummy_code(cont, iscontinous=True),
                   dc_cont_raw=pyfunc.dummy_code(cont, iscontinous=True, demean=False),
                   dcc=pyfunc.dummy_code_cosine(t, 12.0), cp_22=pyfunc.column_product(a2, b3),
                   cp_12=pyfunc.column_product(cont, b3), cp_21=pyfunc.column_product(a2, cont),
                   cp_11=pyfunc.column_product(cont, t), so=pyfunc.stack_ones(a2),
                   ci_a=pyfunc.calc_indirect(ta, tb, alg="aroian"), ci_s=pyfunc.calc_indirect(ta, tb, alg="sobel"),
                   ci_g=pyfunc.calc_indirect(ta, tb, alg="goodman"))
    np.savez_compressed(os.path.join(HERE, "tm_models_helpers.npz"), **out)
    print("tm_models_helpers.npz written:", {k: np.shape(v) for k, v in out.items()})


if __name__ == "__main__":
    main()
