import sys, os
sys.path.insert(0, "/root/repo")
import numpy as np
from tfce_mediation_b200 import synth, _lib
from tfce_mediation_b200.tfce import CreateAdjSet
v, f = synth.icosphere(7); csr = synth.faces_to_csr(v.shape[0], f); V = v.shape[0]
g = CreateAdjSet(2, 0.67, csr)
vm = np.empty(V, dtype=np.int32); _lib.check(_lib.lib().tmb_graph_vmap(g._handle, _lib.ptr(vm)))
inv = np.empty(V, dtype=np.int64); inv[vm] = np.arange(V)
ip, ix = csr
src = np.repeat(np.arange(V), np.diff(ip))
a = inv[src]; b = inv[ix]
blk = (a // 256) * 256
for W in (0, 128, 256, 512, 1024, 2048, 4096):
    inside = (b >= blk - W) & (b < blk + 256 + W)
    print("window +-%d: %.1f%% of neighbour lookups inside" % (W, 100 * inside.mean()))
d = np.abs(a - b); print("median |du| %d, p90 %d, p99 %d, max %d" % (np.median(d), np.percentile(d, 90), np.percentile(d, 99), d.max()))
