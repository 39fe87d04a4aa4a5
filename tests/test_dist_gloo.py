"""CPU: the permutation sharding + all-gather of per-shuffle maxima (the N>1 path) on a world_size-2
gloo group.  The per-shuffle "compute" is a deterministic stand-in; what is tested is that shard_range
covers the reference's `-r a b` range exactly once, in order, and that gather_rows returns the rows in
permutation order on every rank -- also when the shards have different sizes."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _fake_rows(first, last):
    p = np.arange(first, last + 1, dtype=np.float64)
    return np.stack([np.sin(p) * 100, np.cos(p) * 100], axis=1).astype(np.float32).reshape(-1, 1, 2)


def _worker(rank, world, port, first, last, out_dir):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from tfce_mediation_b200 import parallel
    parallel.init_process_group("gloo")
    a, b = parallel.shard_range(first, last, rank, world)
    local = _fake_rows(a, b) if b >= a else np.zeros((0, 1, 2), dtype=np.float32)
    allrows = parallel.gather_rows(local)
    np.save(os.path.join(out_dir, "rank%d.npy" % rank), allrows)
    import torch.distributed as dist
    dist.destroy_process_group()


@pytest.mark.parametrize("first,last", [(1, 100), (1, 7), (5, 5)])
def test_shard_and_gather_world2(tmp_path, first, last):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(2, port, first, last, str(tmp_path)), nprocs=2, join=True)
    want = _fake_rows(first, last)
    for r in range(2):
        got = np.load(tmp_path / ("rank%d.npy" % r))
        assert np.array_equal(got, want)


def test_shard_range_partitions_exactly():
    from tfce_mediation_b200 import parallel
    for first, last, ws in [(1, 100, 8), (1, 10000, 8), (3, 9, 4), (1, 3, 8), (1, 0, 2)]:
        seen = []
        for r in range(ws):
            a, b = parallel.shard_range(first, last, r, ws)
            seen += list(range(a, b + 1))
        assert seen == list(range(first, last + 1))
    assert parallel.world() == (int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)),
                                int(os.environ.get("LOCAL_RANK", 0)))


def _worker_1d(rank, world, port, first, last, out_dir):
    """1-D rows (one value per shuffle: the mediation drivers) and the composed permutation stream of the tm-models
    mediation branch, replayed on every rank from the first permutation of the range (tm_models_randomise.py:430-436)."""
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from tfce_mediation_b200 import parallel
    parallel.init_process_group("gloo")
    a, b = parallel.shard_range(first, last, rank, world)
    n = 11
    composed, local = np.arange(n), []
    for it in range(first, b + 1):
        np.random.seed(it * 1000 + 3)
        composed = composed[np.random.permutation(list(range(n)))]
        if it >= a:
            local.append(float(np.dot(composed, np.arange(n))))       # stand-in statistic of the composed order
    allrows = parallel.gather_rows(np.asarray(local, dtype=np.float32))
    np.save(os.path.join(out_dir, "rank%d.npy" % rank), allrows)
    import torch.distributed as dist
    dist.destroy_process_group()


def test_composed_permutation_stream_world2(tmp_path):
    import torch.multiprocessing as mp
    port = _free_port()
    first, last, n = 1, 9, 11
    mp.spawn(_worker_1d, args=(2, port, first, last, str(tmp_path)), nprocs=2, join=True)
    x, want = np.arange(n), []
    for it in range(first, last + 1):              # the reference's in-place permutation, one process
        np.random.seed(it * 1000 + 3)
        x = x[np.random.permutation(list(range(n)))]
        want.append(float(np.dot(x, np.arange(n))))
    for r in range(2):
        assert np.array_equal(np.load(tmp_path / ("rank%d.npy" % r)), np.asarray(want, dtype=np.float32))
