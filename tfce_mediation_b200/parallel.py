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
    MaxComm.start_async()
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
        # control plane only (barriers, the 128-byte NCCL id): gloo is up in milliseconds, a torch NCCL process group
        # takes seconds to create -- the data path has its own communicator (MaxComm / tmb_allgather_max)
        backend = os.environ.get("TMB_DIST_BACKEND", "gloo")
    if backend == "nccl":
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        dist.init_process_group(backend)


def shard_counts(first, last, world_size):
    """Number of permutations every rank owns under shard_range (known to all ranks without communication)."""
    out = []
    for r in range(world_size):
        a, b = shard_range(first, last, r, world_size)
        out.append(max(0, b - a + 1))
    return out


class MaxComm(object):
    """tmb_comm handle (include/tfce_b200.h): the all-gather of the per-shuffle maxima behind the C ABI, NCCL over
    NVLink.  The 128-byte NCCL id is broadcast through the torch.distributed process group torchrun set up (gloo in the
    drivers: a control plane that is up in milliseconds).  Creating an NCCL communicator takes seconds -- longer than a
    whole 10,000-permutation job -- so setup() starts it on a side thread and the job only joins it at its one gather."""

    _instance = None
    _thread = None
    _error = None

    def __init__(self, defer_create=False):
        import ctypes
        import torch
        import torch.distributed as dist
        from . import _lib
        rank, ws, _ = world()
        self.rank, self.ws = rank, ws
        self.device = torch.device("cuda", local_device_index() if ws > 1 else torch.cuda.current_device())
        on_gpu = dist.get_backend() == "nccl"
        idt = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            buf = (ctypes.c_char * 128)()
            _lib.check(_lib.lib().tmb_comm_unique_id(buf))
            idt = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
        if on_gpu:
            idt = idt.to(self.device)
        dist.broadcast(idt, src=0)            # on the calling thread: collectives of one group must keep one order
        self._id = bytes(idt.cpu().numpy().tobytes())
        self._handle = None
        if not defer_create:
            self._create()

    def _create(self):
        import ctypes
        from . import _lib
        h = ctypes.c_void_p()
        _lib.check(_lib.lib().tmb_comm_create(self._id, self.rank, self.ws, self.device.index, ctypes.byref(h)))
        self._handle = h

    @classmethod
    def available(cls):
        import torch
        import torch.distributed as dist
        return (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1 and torch.cuda.is_available()
                and not _shared_gpu() and not os.environ.get("TMB_NO_NCCL"))

    @classmethod
    def start_async(cls):
        """Create the communicator on a side thread (no other collective may run on the process group meanwhile)."""
        import threading
        if cls._instance is not None or cls._thread is not None or not cls.available():
            return

        inst = cls(defer_create=True)          # id exchange here; ncclCommInitRank (seconds) on the side thread

        def work():
            try:
                inst._create()
                cls._instance = inst
            except Exception as exc:          # noqa: BLE001  (re-raised by get())
                cls._error = exc

        cls._thread = threading.Thread(target=work, daemon=True)
        cls._thread.start()

    @classmethod
    def get(cls):
        """The process-wide communicator (joining the side thread that creates it), or None when it does not apply."""
        if cls._thread is not None:
            cls._thread.join()
            cls._thread = None
            if cls._error is not None:
                err, cls._error = cls._error, None
                raise err
        if cls._instance is None and cls.available():
            cls._instance = cls()
        return cls._instance

    def allgather(self, local):
        """local: CUDA float32 tensor (contiguous), the same number of elements on every rank.  Returns [ws, ...]."""
        import torch
        from . import _lib
        local = local.contiguous()
        out = torch.empty((self.ws,) + tuple(local.shape), dtype=torch.float32, device=local.device)
        _lib.check(_lib.lib().tmb_allgather_max(self._handle, _lib.ptr(local), local.numel(), _lib.ptr(out),
                                                _lib.current_stream()))
        return out

    def close(self):
        from . import _lib
        if self._handle is not None:
            _lib.lib().tmb_comm_destroy(self._handle)
            self._handle = None
        MaxComm._instance = None


def finalize():
    if MaxComm._thread is not None:
        MaxComm._thread.join()
        MaxComm._thread = None
    if MaxComm._instance is not None:
        MaxComm._instance.close()
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
    comm = MaxComm.get() if (counts is not None and local_rows.dtype == np.float32) else None
    if comm is not None:
        dev = torch.device("cuda", torch.cuda.current_device())
        # every rank's row count follows from the shard rule: ONE collective, no size exchange, no host sync before it
        pmax = max(counts)
        tail = local_rows.shape[1:]
        pad = np.zeros((pmax,) + tail, dtype=np.float32)
        pad[:local_rows.shape[0]] = local_rows
        out = comm.allgather(torch.from_numpy(pad).to(dev)).cpu().numpy()
        return np.concatenate([out[r, :k] for r, k in enumerate(counts)], axis=0)
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
