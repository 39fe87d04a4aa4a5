"""Throughput of the tm-models families at config-2 size: fsaverage-topology lh+rh, 299,881 vertices,
300 rows (GLM / cosinor: 300 subjects; rmANCOVA: 3 intervals x 100 subjects), 1-ring adjacency.  Device-timed blocks."""
import sys
import time
import numpy as np
import torch
sys.path.insert(0, ".")
import bench
from tfce_mediation_b200.rmancova import RmAncovaModel

P = int(sys.argv[1]) if len(sys.argv) > 1 else 64
w = bench.build_workload("config2")
eng, _, _ = bench.make_engine(w, torch.device("cuda", 0))
n = w["n"]
rs = np.random.RandomState(0)
perms = np.stack([rs.permutation(n) for _ in range(P)])


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


grp = rs.randint(0, 3, n)
exog = [rs.standard_normal((n, 1)), np.column_stack([(grp == 1) * 1.0, (grp == 2) * 1.0])]
cov = rs.standard_normal((n, 2))
from tfce_mediation_b200.pyfunc import typeI_design
X, kvars = typeI_design(exog, cov, n)
t = timed(lambda: eng.glm_typeI_block(X, kvars, perms, stat="f"))
print("GLM F (2 variables, 5 regressors): %.1f ms per %d shuffles = %.0f shuffles/s" % (t * 1e3, P, P / t))
t = timed(lambda: eng.glm_typeI_block(X, kvars, perms, stat="t"))
print("GLM t (3 contrasts, both signs):   %.1f ms per %d shuffles = %.0f shuffles/s" % (t * 1e3, P, P / t))
time_var = rs.uniform(0, 24, n)
t = timed(lambda: eng.cosinor_block(time_var, [24.0], exog[:1], cov, perms))
print("cosinor (1 period, 1 tested, 2 cov): %.1f ms per %d shuffles = %.0f shuffles/s" % (t * 1e3, P, P / t))
med = rs.standard_normal(n)
t = timed(lambda: eng.cosinor_mediation_block(time_var, [24.0], med - med.mean(), perms))
print("cosinor mediation:                 %.1f ms per %d shuffles = %.0f shuffles/s" % (t * 1e3, P, P / t))
left = rs.standard_normal((n, 1))
right = 0.5 * left + rs.standard_normal((n, 1))
for mt in ("I", "Y"):
    t = timed(lambda: eng.tm_models_mediation_block(mt, left, right, cov, perms))
    print("tm-models mediation %s (2 covariates): %.1f ms per %d shuffles = %.0f shuffles/s" % (mt, t * 1e3, P, P / t))
import os
os.environ["TMB_SOBEL"] = "designs"
t = timed(lambda: eng.tm_models_mediation_block("I", left, right, cov, perms))
print("   the same from two whole designs (7 rows per shuffle): %.1f ms" % (t * 1e3))
del os.environ["TMB_SOBEL"]
s, ns = 3, n // 3
g2 = np.arange(ns) % 2
f1 = (g2 - g2.mean()).astype(np.float64)
f2 = rs.standard_normal(ns)
f2 -= f2.mean()
cv = rs.standard_normal((ns, 2))
cv -= cv.mean(0)
subj = np.eye(ns)[:, 1:]
shuffles = np.stack([rs.permutation(s * ns) for _ in range(P)])
rands = np.stack([rs.permutation(ns) for _ in range(P)])
for name, model in (("rmANCOVA one factor", RmAncovaModel(ns, s, [f1], subj, cv)),
                    ("rmANCOVA two factors", RmAncovaModel(ns, s, [f1, f2], subj, cv))):
    t = timed(lambda: eng.rm_ancova_block(model, shuffles, rands))
    print("%s (%d columns, %d F maps): %.1f ms per %d shuffles = %.0f shuffles/s" % (name, model.rU, model.nout, t * 1e3, P, P / t))
    t0 = time.perf_counter(); model.operands(shuffles, rands); print("   host operands: %.1f ms" % ((time.perf_counter() - t0) * 1e3))
