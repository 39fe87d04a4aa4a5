"""Quick device-side timing probe (development aid): fit and TFCE on config-2-shaped synthetic input."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from tfce_mediation_b200 import synth, _lib
from tfce_mediation_b200.engine import PermutationEngine, Surface, row_permuted_stack
from tfce_mediation_b200.tfce import CreateAdjSet

level = int(sys.argv[1]) if len(sys.argv) > 1 else 7
n = int(sys.argv[2]) if len(sys.argv) > 2 else 300
P = int(sys.argv[3]) if len(sys.argv) > 3 else 128
rounds = int(sys.argv[4]) if len(sys.argv) > 4 else 6
t0 = time.time()
v, f = synth.icosphere(level)
csr = synth.faces_to_csr(v.shape[0], f)
V = v.shape[0]
print("mesh", V, csr[1].shape[0], "%.1fs" % (time.time() - t0), flush=True)
y = np.concatenate([synth.subject_data(n, csr, 1, rounds), synth.subject_data(n, csr, 2, rounds)], axis=1)
print("data", y.shape, "%.1fs" % (time.time() - t0), flush=True)
rs = np.random.RandomState(0)
X = np.column_stack([np.ones(n), rs.standard_normal(n)])
surfs = [Surface(CreateAdjSet(2, 0.67, csr), 0), Surface(CreateAdjSet(2, 0.67, csr), V)]
eng = PermutationEngine(y, surfs)
idx = np.stack([rs.permutation(n) for _ in range(P)])
stack = row_permuted_stack(X, idx)

def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); r = fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts), r

ms_fit, t32 = timed(lambda: eng.tstat(stack, caller_order=False))
print("fit: %.3f ms for %d shuffles -> %.1f us/shuffle ; fp64 TFLOP/s %.2f" % (ms_fit, P, ms_fit * 1e3 / P, 2.0 * P * n * 2 * V / ms_fit / 1e9), flush=True)
Pc, C, ld = t32.shape
stat = t32.view(Pc * C, ld)
ms_tfce, _ = timed(lambda: eng.plan.run(stat, two_sided=True))
print("tfce: %.3f ms for %d maps x 2 hemis x 2 signs -> %.1f us/shuffle" % (ms_tfce, Pc * C, ms_tfce * 1e3 / P), flush=True)
ms_all, _ = timed(lambda: eng.regression_block(X, perm_idx=idx))
print("block e2e (host algebra + h2d + fit + tfce + d2h): %.3f ms -> %.1f shuffles/s" % (ms_all, P / ms_all * 1e3), flush=True)
print("launches", _lib.launch_count())
eng.plan.__del__()   # prints the TMB_PHASE_TIMING summary (if enabled)
