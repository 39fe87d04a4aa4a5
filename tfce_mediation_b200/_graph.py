"""Host-side graph plumbing: Python adjacency containers <-> CSR, masked sub-graphs."""
import itertools

import numpy as np


def adjacency_to_csr(py_adjacency):
    """Any length-V sequence of int iterables (lists, sets, arrays; 0-based; may be empty) ->
    (indptr int64[V+1], indices int32[nnz]).  Item iteration order is preserved, which is what
    CreateAdjSet.__init__ does when it copies each item into a vector<int> (tfce.pyx:36-40)."""
    V = len(py_adjacency)
    counts = np.fromiter((len(a) for a in py_adjacency), dtype=np.int64, count=V)
    indptr = np.zeros(V + 1, dtype=np.int64)
    np.cumsum(counts, out=indptr[1:])
    nnz = int(indptr[-1])
    flat = np.fromiter(itertools.chain.from_iterable(py_adjacency), dtype=np.int64, count=nnz)
    if nnz and (flat.min() < 0 or flat.max() >= V):
        raise ValueError("adjacency index out of range [0, %d)" % V)
    return indptr, flat.astype(np.int32)


def csr_to_lists(indptr, indices):
    return [indices[indptr[i]:indptr[i + 1]].tolist() for i in range(len(indptr) - 1)]


def induced_subgraph(indptr, indices, keep):
    """CSR of the sub-graph induced by boolean mask `keep`, vertices renumbered in mask order.
    Masked-out vertices carry statistic 0 in the reference (pyfunc.py:108-113) and can never
    activate, so dropping them leaves every TFCE value of the kept vertices unchanged."""
    keep = np.asarray(keep, dtype=bool)
    V = keep.shape[0]
    new_id = np.full(V, -1, dtype=np.int64)
    new_id[keep] = np.arange(int(keep.sum()))
    rows = np.repeat(np.arange(V), np.diff(indptr))
    sel = keep[rows] & keep[indices]
    r = new_id[rows[sel]]
    c = new_id[indices[sel]]
    Vn = int(keep.sum())
    out_ptr = np.zeros(Vn + 1, dtype=np.int64)
    np.add.at(out_ptr, r + 1, 1)
    np.cumsum(out_ptr, out=out_ptr)
    return out_ptr, c.astype(np.int32)  # rows stay grouped and ordered because `rows` is sorted
