import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, oracle
from tests import helpers
from tfce_mediation_b200.tfce import CreateAdjSet
adj = helpers.grid_csr(8, 8)
c = CreateAdjSet(2, 0.67, adj)
img = np.random.RandomState(1).standard_normal(64).astype(np.float32)
img[5] = np.nan
want = oracle.tfce_run(2, 0.67, oracle.adjacency_to_csr(adj), img)
got = np.zeros(64, dtype=np.float32)
c.run(img, got)
bad = np.nonzero(~((got == want)))[0]
print("status", c.last_status, "bad idx", bad, "got", got[bad], "want", want[bad], "img", img[bad])
