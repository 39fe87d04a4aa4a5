"""Fit-kernel timing by design width: t statistics of all regressors for k = 2..9 columns (float32 data, fp64 tensor cores),
and the partial-F epilogue; config-2 size."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tfce_mediation_b200.engine import PermutationEngine
n, V, P = 300, 299881, 512
rs = np.random.RandomState(0)
y = rs.standard_normal((n, V)).astype(np.float32)
eng = PermutationEngine(y, None)
idx = np.stack([rs.permutation(n) for _ in range(P)])


def timed(f):
    f(); torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); f(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts)


for k in (2, 3, 4, 5, 7, 9):
    X = np.column_stack([np.ones(n), rs.standard_normal((n, k - 1))])
    ms = timed(lambda: eng.tstat_rowperm(X, idx))
    r = k - 1
    print("k=%d t of %d regressors: %.3f ms for %d designs -> %.2f fp64 TFLOP/s (%.3f ms per regressor row-block)" % (
        k, r, ms, P, 2.0 * P * r * n * V / ms / 1e9, ms / r), flush=True)
    if k >= 3:
        ms = timed(lambda: eng.fstat_rowperm(X, [0], [r], idx))
        print("      F of all %d: %.3f ms -> %.2f TFLOP/s" % (r, ms, 2.0 * P * r * n * V / ms / 1e9), flush=True)
