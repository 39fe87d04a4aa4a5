"""Permutation sharding across GPUs (one process per GPU, torch.distributed).

The reference's only parallelism is splitting the permutation range into blocks of 100 handed to
independent OS processes that append to shared CSV files (STEP_2_tfce_randomise_parallel.py:139-153).
Here every rank owns a contiguous slice of the permutation range, holds the full data in its own HBM,
and the per-shuffle maxima are collected with ONE all-gather (NCCL over NVLink on GPUs, gloo in the
CPU tests).  There is no exchange step inside a shuffle.
"""
import os

import numpy as np


def world():
    """(rank, world_size, local_rank) from the torchrun environment (1 process -> (0, 1, 0))."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def shard_range(first, last, rank, world_size):
    """Contiguous slice [a, b] (inclusive, like the reference's `-r a b`) of [first, last] for `rank`.
    Sizes differ by at most one; a > b means the rank has no work."""
    total = last - first + 1
    if total <= 0:
        return first, first - 1
    base, rem = divmod(total, world_size)
    a = first + rank * base + min(rank, rem)
    b = a + base + (1 if rank < rem else 0) - 1
    return a, b


def _shared_gpu():
    """TMB_ALLOW_SHARED_GPU=1 (tests on a one-GPU box): ranks beyond the GPU count share devices, collectives use gloo."""
    import torch
    _, ws, _ = world()
    return bool(os.environ.get("TMB_ALLOW_SHARED_GPU")) and torch.cuda.is_available() and torch.cuda.device_count() < ws


def local_device_index():
    import torch
    _, _, local = world()
    return local % max(1, torch.cuda.device_count()) if _shared_gpu() else local


def bind_device():
    """One process per GPU: make cuda:LOCAL_RANK the current device of this process.  Must run before any graph, plan
    or engine is created (they live on the current device); every driver calls it first through setup()."""
    import torch
    _, ws, local = world()
    if torch.cuda.is_available() and ws > 1:
        torch.cuda.set_device(local_device_index())


def setup(backend=None):
    """First call of every driver's run(): bind the GPU, join the process group.  Returns (rank, world_size)."""
    bind_device()
    init_process_group(backend)
    rank, ws, _ = world()
    return rank, ws


def check_device(device_index):
    """Refuse to build device state on another rank's GPU (all ranks piling onto cuda:0 was the round-1 driver bug)."""
    _, ws, local = world()
    if ws > 1:
        local = local_device_index()
    if ws > 1 and device_index is not None and int(device_index) != local:
        raise RuntimeError("rank with LOCAL_RANK=%d is about to allocate on cuda:%d: call parallel.setup() "
                           "(or torch.cuda.set_device(LOCAL_RANK)) before creating graphs or engines" % (local, device_index))


def init_process_group(backend=None):
    import torch
    import torch.distributed as dist
    rank, ws, local = world()
    if ws == 1 or dist.is_initialized():
        return
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() and not _shared_gpu() else "gloo"
    if backend == "nccl":
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        dist.init_process_group(backend)


def finalize():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.barrier()
        dist.destroy_process_group()


def gather_rows(local_rows, counts=None):
    """All-gather per-shuffle result rows.  local_rows: float array [P_local, ...]; every rank may hold a
    different P_local (shard_range).  Returns the concatenation in rank order on every rank."""
    import torch
    import torch.distributed as dist
    local_rows = np.ascontiguousarray(local_rows)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local_rows
    ws = dist.get_world_size()
    use_cuda = dist.get_backend() == "nccl"
    dev = torch.device("cuda", torch.cuda.current_device()) if use_cuda else torch.device("cpu")
    n_local = torch.tensor([local_rows.shape[0]], dtype=torch.int64, device=dev)
    all_n = [torch.zeros_like(n_local) for _ in range(ws)]
    dist.all_gather(all_n, n_local)
    all_n = [int(t.item()) for t in all_n]
    pmax = max(all_n)
    tail = local_rows.shape[1:]
    pad = np.zeros((pmax,) + tail, dtype=local_rows.dtype)
    pad[:local_rows.shape[0]] = local_rows
    t = torch.from_numpy(pad).to(dev)
    out = [torch.empty_like(t) for _ in range(ws)]
    dist.all_gather(out, t)
    return np.concatenate([o.cpu().numpy()[:k] for o, k in zip(out, all_n)], axis=0)
