"""Fit-kernel timing probe: fp32 vs fp64 resident data, DMMA vs DFMA."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tfce_mediation_b200.engine import PermutationEngine, row_permuted_stack
n, V, P = 300, 327684, 512
rs = np.random.RandomState(0)
y = rs.standard_normal((n, V)).astype(np.float32)
X = np.column_stack([np.ones(n), rs.standard_normal(n)])
idx = np.stack([rs.permutation(n) for _ in range(P)])
stack = row_permuted_stack(X, idx)
for dt in (np.float32, np.float64):
    eng = PermutationEngine(y.astype(dt), None)
    f = lambda: eng.tstat(stack, caller_order=False)
    f(); torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); f(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ms = min(ts)
    print("%s data: %.3f ms for %d designs -> %.2f fp64 TFLOP/s" % (np.dtype(dt).name, ms, P, 2.0 * P * n * V / ms / 1e9), flush=True)
