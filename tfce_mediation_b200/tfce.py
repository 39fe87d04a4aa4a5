"""Drop-in for the reference extension module ``tfce_mediation.tfce`` (tfce.pyx:24-45).

``CreateAdjSet(H, E, pyAdjacency).run(image, enhn)`` keeps the reference's signature, buffer
checks and ``enhn += TFCE(image)`` semantics; the work is done by the sm_100a kernels behind
``tmb_graph_create`` / ``tmb_tfce_run`` (include/tfce_b200.h).
"""
import ctypes

import numpy as np

from . import _lib
from ._graph import adjacency_to_csr

MAP_MAX_IS_ZERO = 1


def _check_buffer(name, a):
    # mirrors the errors raised by Cython's ndarray[float, ndim=1, mode="c"] buffer acquisition
    if not isinstance(a, np.ndarray):
        raise TypeError("Argument '%s' has incorrect type (expected numpy.ndarray, got %s)" % (name, type(a).__name__))
    if a.ndim != 1:
        raise ValueError("Buffer has wrong number of dimensions (expected 1, got %d)" % a.ndim)
    if a.dtype != np.float32:
        raise ValueError("Buffer dtype mismatch, expected 'float' but got '%s'" % a.dtype.name)
    if not a.flags.c_contiguous:
        raise ValueError("ndarray is not C-contiguous")


class CreateAdjSet(object):
    """TFCE over a fixed adjacency; H and E are stored as C floats (tfce.pyx:27-33)."""

    def __init__(self, H, E, pyAdjacency, device=None):
        _lib.require_device()
        if isinstance(pyAdjacency, tuple) and len(pyAdjacency) == 2 and isinstance(pyAdjacency[0], np.ndarray) \
                and pyAdjacency[0].dtype == np.int64 and pyAdjacency[1].dtype == np.int32:
            indptr, indices = pyAdjacency          # already CSR (engine-internal fast path)
        else:
            indptr, indices = adjacency_to_csr(pyAdjacency)
        self.H = float(np.float32(H))
        self.E = float(np.float32(E))
        self.indptr = np.ascontiguousarray(indptr, dtype=np.int64)
        self.indices = np.ascontiguousarray(indices, dtype=np.int32)
        self.num_vertices = int(self.indptr.shape[0] - 1)
        if device is None:
            import torch
            device = torch.cuda.current_device()
        self.device = int(device)
        from . import parallel
        parallel.check_device(self.device)
        h = ctypes.c_void_p()
        _lib.check(_lib.lib().tmb_graph_create(self.device, self.num_vertices, _lib.ptr(self.indptr),
                                               _lib.ptr(self.indices), self.H, self.E, ctypes.byref(h)))
        self._handle = h

    def __del__(self):
        h = getattr(self, "_handle", None)
        if h is not None and _lib._lib is not None:
            _lib._lib.tmb_graph_destroy(h)
            self._handle = None

    def run(self, image, enhn):
        """enhn += TFCE(image); both float32, 1-D, C-contiguous, length V (tfce.pyx:44-45).

        Deviation: for ``image.max() == 0`` the reference never returns;
        here ``enhn`` is left unchanged and ``self.last_status`` is MAP_MAX_IS_ZERO."""
        _check_buffer("image", image)
        _check_buffer("enhn", enhn)
        if image.shape[0] != self.num_vertices or enhn.shape[0] != self.num_vertices:
            raise ValueError("image/enhn length %d/%d != number of vertices %d"
                             % (image.shape[0], enhn.shape[0], self.num_vertices))
        st = ctypes.c_int(0)
        _lib.check(_lib.lib().tmb_tfce_run(self._handle, _lib.ptr(image), _lib.ptr(enhn), ctypes.byref(st)))
        self.last_status = st.value

    def components(self, image, level):
        """Inspection: canonical labels / extents of {image > T_level} (tmb_tfce_components)."""
        _check_buffer("image", image)
        labels = np.empty(self.num_vertices, dtype=np.int32)
        extents = np.empty(self.num_vertices, dtype=np.int32)
        thr = ctypes.c_float(0)
        _lib.check(_lib.lib().tmb_tfce_components(self._handle, _lib.ptr(image), int(level), _lib.ptr(labels),
                                                  _lib.ptr(extents), ctypes.byref(thr)))
        return labels, extents, thr.value
