"""Device-buffer plumbing (PyTorch owns HBM allocations and streams; no arithmetic here)."""
import numpy as np

from . import _lib

TILE_V = 128   # vertex tile of the fit kernel: rows of Y are padded to a multiple of this
TILE_M = 128   # design-row tile


def round_up(x, m):
    return (int(x) + m - 1) // m * m


def dev_empty(shape, dtype, device):
    import torch
    return torch.empty(shape, dtype=dtype, device=device)


def to_host(t):
    return t.contiguous().cpu().numpy()


def default_device():
    import torch
    _lib.require_device()
    return torch.device("cuda", torch.cuda.current_device())


class DeviceMatrix(object):
    """Subject-by-vertex data [n, V] resident in HBM as [n, ld] (ld = V rounded up to 128, pad = 0).

    Accepts a host numpy array (float32 or float64; anything else is promoted to float64, as numpy
    does inside the reference's dot products) or a CUDA torch tensor."""

    def __init__(self, y, device=None, pinned_stage=None):
        import torch
        if device is None:
            device = default_device()
        if isinstance(y, torch.Tensor):
            src = y
            if src.dtype not in (torch.float32, torch.float64):
                src = src.to(torch.float64)
        else:
            a = np.asarray(y)
            if a.dtype not in (np.float32, np.float64):
                a = a.astype(np.float64)
            src = torch.from_numpy(np.ascontiguousarray(a))
        if src.dim() != 2:
            raise ValueError("data must be 2-D [subjects, vertices]")
        self.n, self.V = int(src.shape[0]), int(src.shape[1])
        self.ld = round_up(self.V, TILE_V)
        self.t = torch.zeros((self.n, self.ld), dtype=src.dtype, device=device)
        self.t[:, :self.V].copy_(src, non_blocking=True)
        self.dtype_code = _lib.F64 if src.dtype == torch.float64 else _lib.F32
        self.h2d_bytes = self.n * self.V * src.element_size() if not src.is_cuda else 0
        self._yy = {}

    def permuted_copy(self, colperm):
        """A second DeviceMatrix whose column j holds this one's column colperm[j] (padding untouched)."""
        import torch
        out = object.__new__(DeviceMatrix)
        out.n, out.V, out.ld, out.dtype_code, out.h2d_bytes = self.n, self.V, self.ld, self.dtype_code, 0
        out.t = torch.zeros_like(self.t)
        out.t[:, :self.V] = self.t[:, :self.V].index_select(1, colperm)
        out._yy = {}
        return out

    def sumsq(self, center):
        """float64 [V] (centred) sum of squares per vertex, cached."""
        import torch
        key = bool(center)
        if key not in self._yy:
            yy = torch.empty((self.V,), dtype=torch.float64, device=self.t.device)
            _lib.check(_lib.lib().tmb_glm_sumsq(_lib.ptr(self.t), self.dtype_code, self.n, self.V, self.ld,
                                                1 if center else 0, _lib.ptr(yy), None, _lib.current_stream()))
            self._yy[key] = yy
        return self._yy[key]

    def sstotal_reference(self):
        """float64 [ld]: SS_Total as the reference forms it, np.sum((endog - np.mean(endog, 0))**2, 0) (pyfunc.py:2331,
        :2492): numpy reduces axis 0 row after row in the data's own type, so float32 data give a float32-accumulated
        value (widened).  The model F statistics subtract the float64 residual SS from THIS number.  Cached."""
        import torch
        if "ref" not in self._yy:
            tot = torch.empty((self.ld,), dtype=torch.float64, device=self.t.device)
            _lib.check(_lib.lib().tmb_rm_totals(_lib.ptr(self.t), self.dtype_code, self.n, self.V, self.ld, None, None, None,
                                                0, 1, _lib.ptr(tot), None, self.ld, _lib.current_stream()))
            self._yy["ref"] = tot
        return self._yy["ref"]
