"""Shared pieces of the randomise drivers."""
import os
from time import time

import numpy as np

from .. import parallel
from .._graph import adjacency_to_csr, induced_subgraph

BLOCK = 512      # shuffles per engine call when the data size is unknown (block_for adapts it); tests override it
_DEFAULT_BLOCK = BLOCK


def block_for(eng):
    """Shuffles per pipelined engine call for this engine's data: about 6e8 vertex-maps per call, a power of two in
    [64, 8192] -- small surfaces need many shuffles per launch to fill the GPU (BASELINE config 1: 8,192), large
    multi-surface rows few (config 5: 128 at most).  An explicitly changed C.BLOCK wins."""
    if BLOCK != _DEFAULT_BLOCK:
        return BLOCK
    per = max(1.0, 6e8 / max(1, int(eng.Y.V)))
    return int(min(8192, max(64, 2 ** int(np.floor(np.log2(per))))))


# wall-clock marks of the last driver run: [(label, seconds since the previous mark)], read by tmanalysis/job.py
TIMINGS = []
_last_tick = [None]


def tick(label=None):
    """Start (label None) or extend the phase log of a driver run."""
    now = time()
    if label is None:
        del TIMINGS[:]
    elif _last_tick[0] is not None:
        TIMINGS.append((label, now - _last_tick[0]))
    _last_tick[0] = now


def load(path):
    return np.load(path, allow_pickle=True)


def append_rows(path, rows, fmt):
    """One formatted value per line, appended -- what the reference's `echo ... >> file` produces."""
    with open(path, "a") as f:
        for r in rows:
            f.write((fmt % r) + "\n")


def reference_seed(iter_perm, seed):
    """The reference seeds with int(iter_perm*1000 + time()) (e.g. vertex_..._randomise.py:91), which is not
    reproducible; `--seed S` replaces time() by S so two runs (and the CPU oracle) see the same stream."""
    return int(iter_perm * 1000 + (time() if seed is None else seed))


def draw_row_permutation(n):
    return np.random.permutation(list(range(n)))


def draw_block_permutation(block_list, indexer):
    """Exchangeability blocks (vertex_..._randomise.py:98-103): shuffle the block order, then within blocks."""
    randindex = []
    for block in np.random.permutation(list(np.unique(block_list))):
        randindex.append(np.random.permutation(indexer[block_list == block]))
    return np.concatenate(randindex)


def masked_surface(adjacency, H, E, keep=None, weight=None, col_offset=0):
    """CreateAdjSet + Surface for the vertices selected by boolean mask `keep` (None = all)."""
    from ..engine import Surface
    from ..tfce import CreateAdjSet
    indptr, indices = adjacency_to_csr(adjacency)
    if keep is not None:
        keep = np.asarray(keep, dtype=bool)
        indptr, indices = induced_subgraph(indptr, indices, keep)
    w = None
    if weight is not None and np.ndim(weight) > 0:
        w = np.asarray(weight, dtype=np.float32)
        if keep is not None:
            w = w[keep]
    elif weight is not None and float(weight) != 1.0:
        w = float(weight)
    return Surface(CreateAdjSet(H, E, (indptr, indices)), col_offset, w)


def setup():
    """Bind this rank's GPU and join the process group -- before ANY device state exists."""
    return parallel.setup()


_SHARD_COUNTS = [None]


def shard(first, last):
    rank, ws, _ = parallel.world()
    if ws > 1:
        parallel.setup()
    a, b = parallel.shard_range(first, last, rank, ws)
    _SHARD_COUNTS[0] = parallel.shard_counts(first, last, ws)      # every rank's row count, known without communication
    return rank, ws, a, b


def gather(local_rows):
    """All-gather of the per-shuffle rows of the last shard(): ONE collective per job (tmb_allgather_max behind the C ABI
    on NCCL; torch.distributed on gloo)."""
    return parallel.gather_rows(np.ascontiguousarray(local_rows, dtype=np.float32), counts=_SHARD_COUNTS[0])


def chunks(a, b, size=BLOCK):
    p = a
    while p <= b:
        q = min(b, p + size - 1)
        yield p, q
        p = q + 1
