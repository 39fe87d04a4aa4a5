"""Drop-in for the hot-path subset of the reference's ``tfce_mediation.pyfunc``.

    create_adjac_vertex(vertices, faces)                         pyfunc.py:37-46
    create_adjac_voxel(data_index, data_mask, num_voxel, dirtype) pyfunc.py:48-76
    write_perm_maxTFCE_vertex(...)                                pyfunc.py:107-119
    write_perm_maxTFCE_voxel(...)                                 pyfunc.py:121-126
    calc_sobelz(medtype, pred_x, depend_y, merge_y, n, num_vertex, alg) pyfunc.py:130-162

Same names, arguments and side effects (one formatted line appended to a CSV in the current
directory).  These are the single-map forms; the randomise drivers use the batched engine
(engine.PermutationEngine) which produces identical rows for whole blocks of shuffles.
"""
import ctypes

import numpy as np

from . import _lib
from .cynumstats import calc_beta_se  # noqa: F401  (re-exported like the reference does, pyfunc.py:34)


def _append_line(path, text):
    # the reference shells out to `echo ... >> file` (pyfunc.py:119,126); same bytes, no fork
    with open(path, "a") as f:
        f.write(text + "\n")


def create_adjac_vertex(vertices, faces):
    """1-ring neighbour sets from triangle faces (pyfunc.py:37-46).  Host-side integer set building."""
    adjacency = [set([]) for _ in range(vertices.shape[0])]
    f = np.asarray(faces)
    for a, b in ((0, 1), (0, 2), (1, 0), (1, 2), (2, 0), (2, 1)):
        for u, v in zip(f[:, a].tolist(), f[:, b].tolist()):
            adjacency[u].add(v)
    return adjacency


def _voxel_csr(mask, dirtype, variant):
    _lib.require_device()
    import torch
    m = np.ascontiguousarray(np.asarray(mask) != 0, dtype=np.uint8)
    if m.ndim != 3:
        raise ValueError("mask must be 3-D")
    nx, ny, nz = m.shape
    nv, nnz = ctypes.c_int32(0), ctypes.c_int64(0)
    dev = torch.cuda.current_device()
    L = _lib.lib()
    _lib.check(L.tmb_voxel_adjacency(dev, _lib.ptr(m), nx, ny, nz, int(dirtype), variant, ctypes.byref(nv),
                                     ctypes.byref(nnz), None, None))
    indptr = np.zeros(nv.value + 1, dtype=np.int64)
    indices = np.zeros(max(nnz.value, 1), dtype=np.int32)
    _lib.check(L.tmb_voxel_adjacency(dev, _lib.ptr(m), nx, ny, nz, int(dirtype), variant, ctypes.byref(nv),
                                     ctypes.byref(nnz), _lib.ptr(indptr), _lib.ptr(indices)))
    return indptr, indices[:nnz.value]


def create_adjac_voxel(data_index, data_mask, num_voxel, dirtype=26):
    """26- or 6-connectivity voxel adjacency (pyfunc.py:48-76): object array of sorted lists,
    self excluded, ``adjacency[0] = []`` and label 0 never listed.  Built by the GPU kernel.
    (data_mask only supplies the volume shape in the reference; it does the same here.)"""
    data_index = np.asarray(data_index)
    if data_index.shape != np.asarray(data_mask).shape:
        raise ValueError("data_index and data_mask must have the same shape")
    indptr, indices = _voxel_csr(data_index, dirtype, 0)
    if indptr.shape[0] - 1 != int(num_voxel):
        raise ValueError("num_voxel=%d but the mask holds %d voxels" % (int(num_voxel), indptr.shape[0] - 1))
    out = np.empty(indptr.shape[0] - 1, dtype=object)
    for i in range(out.shape[0]):
        out[i] = indices[indptr[i]:indptr[i + 1]].tolist()
    return out


def create_adjac_voxel_tools(data_index, dirtype=26):
    """tools/tm_mulitmodality_adjacency.py:40-66 variant: list of sets, self kept, voxel 0 asymmetric."""
    indptr, indices = _voxel_csr(data_index, dirtype, 1)
    return [set(indices[indptr[i]:indptr[i + 1]].tolist()) for i in range(indptr.shape[0] - 1)]


def write_perm_maxTFCE_vertex(statname, vertStat, num_vertex, bin_mask_lh, bin_mask_rh, calcTFCE_lh, calcTFCE_rh,
                              density_corr_lh=1, density_corr_rh=1):
    """pyfunc.py:107-119: scatter into full-length hemispheres, TFCE each, scaled max, append '%.4f'."""
    vertStat_out_lh = np.zeros(bin_mask_lh.shape[0]).astype(np.float32, order="C")
    vertStat_out_rh = np.zeros(bin_mask_rh.shape[0]).astype(np.float32, order="C")
    vertStat_TFCE_lh = np.zeros_like(vertStat_out_lh).astype(np.float32, order="C")
    vertStat_TFCE_rh = np.zeros_like(vertStat_out_rh).astype(np.float32, order="C")
    vertStat_out_lh[bin_mask_lh] = vertStat[:num_vertex]
    vertStat_out_rh[bin_mask_rh] = vertStat[num_vertex:]
    calcTFCE_lh.run(vertStat_out_lh, vertStat_TFCE_lh)
    calcTFCE_rh.run(vertStat_out_rh, vertStat_TFCE_rh)
    max_lh = vertStat_TFCE_lh[np.isfinite(vertStat_TFCE_lh)] * (vertStat_out_lh[np.isfinite(vertStat_out_lh)].max() / 100) * density_corr_lh
    max_rh = vertStat_TFCE_rh[np.isfinite(vertStat_TFCE_rh)] * (vertStat_out_rh[np.isfinite(vertStat_out_rh)].max() / 100) * density_corr_rh
    maxTFCE = np.array([max_lh.max(), max_rh.max()]).max()
    _append_line("perm_%s_TFCE_maxVertex.csv" % statname, "%.4f" % maxTFCE)


def write_perm_maxTFCE_voxel(statname, voxelStat, TFCEfunc):
    """pyfunc.py:121-126."""
    voxelStat_out = voxelStat.astype(np.float32, order="C")
    voxelStat_TFCE = np.zeros_like(voxelStat_out).astype(np.float32, order="C")
    TFCEfunc.run(voxelStat_out, voxelStat_TFCE)
    maxval = voxelStat_TFCE.max() * (voxelStat_out.max() / 100)
    _append_line("perm_%s_TFCE_maxVoxel.csv" % statname, "%1.4f" % maxval)


def calc_sobelz(medtype, pred_x, depend_y, merge_y, n, num_vertex, alg="aroian"):
    """pyfunc.py:130-162: Sobel-family z from two fits; float64 [V].  Runs the fused GPU kernel
    (tmb_sobelz) through a one-shuffle engine call with the identity permutation."""
    from .engine import sobelz_single
    return sobelz_single(medtype, pred_x, depend_y, merge_y, alg)
