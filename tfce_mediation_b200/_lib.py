"""ctypes binding of libtfce_b200.so (the C ABI declared in include/tfce_b200.h).

There is no CPU fallback: if the shared library is missing, or no CUDA device is visible when a
compute entry point is called, the call raises.  PyTorch is used by the callers of this module
only to own device buffers and streams; nothing here imports it at module import time.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "libtfce_b200.so")
F32, F64 = 0, 1
ABI_VERSION = 3


class TmbError(RuntimeError):
    """A libtfce_b200 entry point returned non-zero."""


_lib = None

_c = ctypes
_vp, _i32, _i64, _f32, _f64 = _c.c_void_p, _c.c_int32, _c.c_int64, _c.c_float, _c.c_double
_int = _c.c_int

# name -> (restype, argtypes); every symbol declared in include/tfce_b200.h
SIGNATURES = {
    "tmb_last_error": (_c.c_char_p, []),
    "tmb_abi_version": (_int, []),
    "tmb_device_count": (_int, []),
    "tmb_launch_count": (_i64, []),
    "tmb_graph_create": (_int, [_int, _i32, _vp, _vp, _f32, _f32, _c.POINTER(_vp)]),
    "tmb_graph_destroy": (_int, [_vp]),
    "tmb_graph_num_vertices": (_int, [_vp, _c.POINTER(_i32), _c.POINTER(_i64)]),
    "tmb_graph_vmap": (_int, [_vp, _vp]),
    "tmb_plan_set_internal_order": (_int, [_vp, _int]),
    "tmb_tfce_run": (_int, [_vp, _vp, _vp, _c.POINTER(_int)]),
    "tmb_tfce_components": (_int, [_vp, _vp, _int, _vp, _vp, _c.POINTER(_f32)]),
    "tmb_plan_create": (_int, [_int, _int, _c.POINTER(_vp), _c.POINTER(_i64), _c.POINTER(_vp), _c.POINTER(_vp), _int,
                               _c.POINTER(_vp)]),
    "tmb_plan_destroy": (_int, [_vp]),
    "tmb_plan_run": (_int, [_vp, _vp, _i64, _int, _int, _vp, _vp, _vp, _vp, _vp]),
    "tmb_plan_maxima": (_int, [_vp, _vp, _i64, _int, _vp, _vp]),
    "tmb_threshold_tables": (_int, [_vp, _vp, _int, _vp, _vp, _vp, _vp, _vp]),
    "tmb_plan_run_tables": (_int, [_vp, _vp, _i64, _int, _int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "tmb_glm_sumsq": (_int, [_vp, _int, _int, _i64, _i64, _int, _vp, _vp, _vp]),
    "tmb_glm_tstat": (_int, [_vp, _int, _int, _i64, _i64, _vp, _i64, _vp, _vp, _int, _int, _int, _int, _int, _f64,
                             _vp, _vp, _vp, _i64, _int, _int, _vp]),
    "tmb_glm_layout": (_int, [_int, _int]),
    "tmb_glm_rp": (_int, [_int, _int]),
    "tmb_glm_packed_columns": (_i64, [_int, _int, _int]),
    "tmb_glm_tstat_beta": (_int, [_vp, _i64, _i64, _vp, _vp, _int, _int, _int, _int, _f64, _vp, _vp, _vp, _i64, _int, _vp]),
    "tmb_glm_fstat_beta": (_int, [_vp, _i64, _i64, _vp, _vp, _int, _int, _int, _vp, _vp, _int, _f64, _vp, _vp, _vp, _vp,
                                  _i64, _int, _vp]),
    "tmb_sobelz_beta": (_int, [_vp, _i64, _i64, _vp, _vp, _int, _int, _f64, _vp, _vp, _int, _int, _f64, _vp, _vp, _int,
                               _int, _vp, _vp, _i64, _vp]),
    "tmb_glm_cosinor_beta": (_int, [_vp, _i64, _i64, _vp, _vp, _int, _int, _int, _int, _f64, _vp, _vp, _int, _f64, _int, _vp, _vp,
                                    _i64, _int, _vp]),
    "tmb_rm_totals": (_int, [_vp, _int, _int, _i64, _i64, _vp, _vp, _vp, _int, _int, _vp, _vp, _i64, _vp]),
    "tmb_rm_ancova_stats": (_int, [_vp, _i64, _i64, _int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _i64, _int, _vp]),
    "tmb_glm_fstat": (_int, [_vp, _int, _int, _i64, _i64, _vp, _i64, _vp, _vp, _int, _int, _int, _int, _vp, _vp, _int,
                             _f64, _vp, _vp, _vp, _vp, _i64, _int, _int, _vp]),
    "tmb_glm_pack_rowperm": (_int, [_vp, _int, _int, _vp, _int, _int, _vp, _i64, _int, _vp]),
    "tmb_glm_beta": (_int, [_vp, _int, _int, _i64, _i64, _vp, _i64, _int, _vp, _i64, _vp]),
    "tmb_glm_direct": (_int, [_vp, _int, _int, _i64, _i64, _vp, _vp, _int, _vp, _f64, _f64, _vp, _vp, _vp, _i64,
                              _vp, _vp, _i64, _vp, _vp, _vp]),
    "tmb_se_of_slope": (_int, [_vp, _i64, _vp, _int, _vp, _i64, _vp]),
    "tmb_sobelz": (_int, [_vp, _int, _int, _i64, _i64, _vp, _i64, _int, _vp, _vp, _int, _int, _f64, _vp, _vp, _int,
                          _int, _f64, _vp, _vp, _int, _int, _vp, _vp, _i64, _int, _vp]),
    "tmb_sobelz_cross": (_int, [_vp, _int, _int, _i64, _i64, _vp, _i64, _vp, _f64, _f64, _vp, _int, _int, _f64, _vp, _int, _int,
                                _vp, _vp, _i64, _vp]),
    "tmb_sobelz_cross_rows": (_int, [_vp, _i64, _int, _vp, _i64, _int, _i64, _vp, _int, _int, _f64, _vp, _int, _int, _f64, _vp,
                                     _vp, _vp, _vp, _int, _int, _vp, _vp, _i64, _vp]),
    "tmb_glm_tstat_cross_rows": (_int, [_vp, _i64, _int, _vp, _i64, _int, _i64, _vp, _int, _vp, _vp, _int, _int, _f64, _vp, _int,
                                        _vp, _vp, _i64, _int, _vp]),
    "tmb_fwe_lookup": (_int, [_vp, _int, _vp, _i64, _vp, _vp]),
    "tmb_comm_unique_id": (_int, [_vp]),
    "tmb_comm_create": (_int, [_vp, _int, _int, _int, _c.POINTER(_vp)]),
    "tmb_comm_destroy": (_int, [_vp]),
    "tmb_allgather_max": (_int, [_vp, _vp, _i64, _vp, _vp]),
    "tmb_voxel_adjacency": (_int, [_int, _vp, _int, _int, _int, _int, _int, _c.POINTER(_i32), _c.POINTER(_i64),
                                   _vp, _vp]),
}


def lib():
    """Load the shared library once; raise loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise TmbError(
                "libtfce_b200.so is missing (%s). Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C tfce_mediation_b200/csrc`. There is no CPU fallback." % SO_PATH)
        L = ctypes.CDLL(SO_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        if L.tmb_abi_version() != ABI_VERSION:
            raise TmbError("libtfce_b200.so ABI %d != expected %d: rebuild" % (L.tmb_abi_version(), ABI_VERSION))
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise TmbError(lib().tmb_last_error().decode("utf-8", "replace"))


def require_device():
    if lib().tmb_device_count() < 1:
        raise TmbError("no CUDA device visible: tfce_mediation_b200 has no CPU fallback")


def launch_count():
    return int(lib().tmb_launch_count())


def ptr(t):
    """Raw device/host pointer of a torch tensor or numpy array (None -> NULL)."""
    if t is None:
        return None
    if hasattr(t, "data_ptr"):
        return ctypes.c_void_p(t.data_ptr())
    return ctypes.c_void_p(t.ctypes.data)


def current_stream():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
