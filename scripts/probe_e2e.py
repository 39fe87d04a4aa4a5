"""Host-side profile of the pipelined public path (PermutationEngine.regression_blocks) on config-2-shaped input."""
import sys, os, time, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from tfce_mediation_b200._graph import induced_subgraph
from tfce_mediation_b200.engine import PermutationEngine, Surface
from tfce_mediation_b200.tfce import CreateAdjSet
w = bench.build_workload("config2")
surfs, off = [], 0
for h in range(2):
    ip, ix = induced_subgraph(w["csr"][0], w["csr"][1], w["masks"][h])
    g = CreateAdjSet(w["H"], w["E"], (ip, ix)); surfs.append(Surface(g, off)); off += g.num_vertices
eng = PermutationEngine(torch.from_numpy(w["y"]).pin_memory(), surfs, two_sided=True)
P, K = 512, 12
idx = np.concatenate([bench.perm_rows(w, s * P, P) for s in range(K)], axis=0)
eng.regression_blocks(w["X"], idx[:2 * P], block=P)
torch.cuda.synchronize()
t0 = time.perf_counter(); eng.regression_blocks(w["X"], idx, block=P); torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("regression_blocks: %.2f ms per block of %d" % (dt * 1e3 / K, P))
pr = cProfile.Profile(); pr.enable(); eng.regression_blocks(w["X"], idx, block=P); torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
