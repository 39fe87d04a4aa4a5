"""Device-side timing of the TFCE stage on k-ring ('3 mm'-like) adjacency, with and without vertex weights, against the
one-kernel fallback (development aid; results are checked for equality between the paths, not against the oracle).
usage: probe_wide.py [level=7] [rings=4] [n=300] [P=256] [modes=pipe,pipe_w,basin,basin_w,ring1]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from tfce_mediation_b200 import synth
from tfce_mediation_b200.engine import PermutationEngine, Surface, TfcePlan, row_permuted_stack
from tfce_mediation_b200.tfce import CreateAdjSet

level = int(sys.argv[1]) if len(sys.argv) > 1 else 7
rings = int(sys.argv[2]) if len(sys.argv) > 2 else 4
n = int(sys.argv[3]) if len(sys.argv) > 3 else 300
P = int(sys.argv[4]) if len(sys.argv) > 4 else 256
modes = (sys.argv[5] if len(sys.argv) > 5 else "ring1,pipe,pipe_w,basin,basin_w").split(",")
reps = int(os.environ.get("PROBE_REPS", "3"))

t0 = time.time()
v, f = synth.icosphere(level)
csr1 = synth.faces_to_csr(v.shape[0], f)
csrk = synth.kring_csr(csr1, rings)
V = v.shape[0]
dens = synth.vertex_density(csrk)
print("mesh V=%d nnz1=%d nnzk=%d maxdeg=%d  %.1fs" % (V, csr1[1].shape[0], csrk[1].shape[0], np.diff(csrk[0]).max(), time.time() - t0), flush=True)
y = np.concatenate([synth.subject_data(n, csr1, 1, 6), synth.subject_data(n, csr1, 2, 6)], axis=1)
rs = np.random.RandomState(0)
X = np.column_stack([np.ones(n), rs.standard_normal(n)])
eng = PermutationEngine(y, None)
idx = np.stack([rs.permutation(n) for _ in range(P)])
t32 = eng.tstat(row_permuted_stack(X, idx), caller_order=False)
stat = t32.view(P, t32.shape[2])
print("t-maps ready %.1fs" % (time.time() - t0), flush=True)


def timed(fn):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); r = fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts), r


results = {}
for mode in modes:
    os.environ.pop("TMB_TFCE", None)
    if mode.startswith("basin"):
        os.environ["TMB_TFCE"] = "basin"
    csr = csr1 if mode == "ring1" else csrk
    w = dens if mode.endswith("_w") else None
    surfs = [Surface(CreateAdjSet(2, 0.67, csr), 0, w), Surface(CreateAdjSet(2, 0.67, csr), V, w)]
    plan = TfcePlan(surfs)
    Pm = P if not mode.startswith("basin") else min(P, 64)
    ms, out = timed(lambda: plan.run(stat[:Pm], two_sided=True, exact_pow=False))
    results[mode] = out[0].cpu().numpy()
    edges = 2 * csr[1].shape[0]
    print("%-8s %8.3f ms for %4d shuffles (x2 hemis x2 signs) -> %7.1f us/shuffle, %.2f G edge-visits/s"
          % (mode, ms, Pm, ms * 1e3 / Pm, edges * 2 * Pm / ms / 1e6), flush=True)
    del plan, surfs
for a, b in (("pipe", "basin"), ("pipe_w", "basin_w")):
    if a in results and b in results:
        m = results[b].shape[0]
        print("%s == %s: %s" % (a, b, np.array_equal(results[a][:m], results[b])))
