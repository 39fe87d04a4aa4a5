"""The drivers' -v mode at config-2 size: k = 5 design columns, one of them permuted; whole-design fit against the
cross-product path (fit only, CUDA events, 512 designs)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tfce_mediation_b200.engine import PermutationEngine, design_stack
n, V, P, k = 300, 299881, 512, 5
rs = np.random.RandomState(0)
y = rs.standard_normal((n, V)).astype(np.float32)
eng = PermutationEngine(y, None)
X = np.column_stack([np.ones(n), rs.standard_normal((n, k - 1))])
designs = []
for p in range(P):
    X[:, 1:2] = X[rs.permutation(n), 1:2]
    designs.append(X.copy())
designs = np.stack(designs)


def timed(f):
    f(); torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); f(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts)


ch = eng._partial_columns(designs)
print("changing columns", ch)
print("whole designs (4 contraction rows per shuffle, host algebra included): %.2f ms" % timed(lambda: eng.tstat(design_stack(designs), caller_order=False)))
print("cross-products (1 contraction row per shuffle, host algebra included): %.2f ms" % timed(lambda: eng.tstat_partial(designs, ch)))
