"""How local are the neighbour references of the wide (default) adjacency in the graphs' internal vertex order?
For the count / ascent kernels: fraction of the neighbour ids of a 1,024-vertex chunk that fall inside the chunk extended
by a halo of H vertices on either side."""
import sys
import numpy as np
sys.path.insert(0, ".")
import bench
from tfce_mediation_b200 import _lib
from tfce_mediation_b200.tfce import CreateAdjSet

name = sys.argv[1] if len(sys.argv) > 1 else "config2_3mm"
w = bench.build_workload(name)
s = w["surfaces"][0]
adj = CreateAdjSet(s["H"], s["E"], s["csr"])
V = adj.num_vertices
vm = np.empty(V, dtype=np.int32)
_lib.check(_lib.lib().tmb_graph_vmap(adj._handle, _lib.ptr(vm)))      # internal position j holds caller vertex vm[j]
pos = np.empty(V, dtype=np.int64)
pos[vm] = np.arange(V)
indptr, indices = s["csr"]
deg = np.diff(indptr)
src = np.repeat(pos, deg)                  # internal index of the row vertex
dst = pos[indices]
chunk = src // 1024
lo, hi = chunk * 1024, chunk * 1024 + 1024
print(name, "V", V, "nnz", indices.shape[0], "mean degree %.1f" % deg.mean())
for H in (0, 512, 1024, 2048, 4096, 8192, 16384):
    inside = (dst >= lo - H) & (dst < hi + H)
    print("halo %5d: %.4f of neighbour references inside" % (H, inside.mean()))
d = np.abs(dst - src)
print("|offset| quantiles 50/90/99/max:", np.percentile(d, [50, 90, 99]), d.max())
