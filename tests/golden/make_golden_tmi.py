#!/usr/bin/env python
"""Golden fixture for the TMI container reader and writer: writes tests/golden/sample.tmi with the product writer
(tfce_mediation_b200.tm_io.write_tm_filetype, header grammar of tm_io.py:158-228), checks that the REAL reference reader
returns the inputs for it, and stores what the reference reader
(/root/reference/tfce_mediation/tm_io.py:284-444 read_tm_filetype) returns for it in tests/golden/tmi_reader.npz.
Build container only (needs /root/reference); import shims as in make_golden.py.

Usage:  python tests/golden/make_golden_tmi.py
"""
import importlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
from make_golden import load_reference  # noqa: E402
from tfce_mediation_b200 import synth  # noqa: E402


def sample_inputs():
    """Two surfaces (a mesh with a vertex mask of shape [V,1,1] and a small voxel volume), 9 subjects."""
    rs = np.random.RandomState(31)
    v, f = synth.icosphere(2)                                   # 162 vertices
    csr = synth.faces_to_csr(v.shape[0], f)
    adj_mesh = np.empty(v.shape[0], dtype=object)
    for i, l in enumerate(synth.csr_to_lists(csr)):
        adj_mesh[i] = list(l)
    mask_mesh = np.zeros((v.shape[0], 1, 1), dtype=bool)
    mask_mesh[rs.rand(v.shape[0]) > 0.2, 0, 0] = True
    vol = np.zeros((6, 7, 5), dtype=bool)
    vol[1:5, 2:6, 1:4] = rs.rand(4, 4, 3) > 0.3
    nvox = int(vol.sum())
    adj_vol = np.empty(nvox, dtype=object)
    for i in range(nvox):
        adj_vol[i] = [j for j in (i - 1, i + 1) if 0 <= j < nvox]
    nrow = int(mask_mesh.sum()) + nvox
    data = rs.standard_normal((nrow, 9)).astype(np.float32)
    return dict(data=data, masks=[mask_mesh, vol], masknames=["lh.area.mgh", "skeleton.nii.gz"], adjacency=[adj_mesh, adj_vol],
                vertices=[v], faces=[f], surfnames=["lh.sphere"], affines=[np.eye(4) * 2.0],
                column_ids=np.array(["s%02d" % i for i in range(9)]))


def main():
    load_reference()
    tm_io = importlib.import_module("tfce_mediation.tm_io")
    from tfce_mediation_b200.tm_io import write_tm_filetype
    path = os.path.join(HERE, "sample.tmi")
    inp = sample_inputs()
    # the file is written by the PRODUCT writer (deterministic: fixed history line) and read back by the REFERENCE reader
    write_tm_filetype(path, columnids=inp["column_ids"], checkname=False, image_array=inp["data"], masking_array=inp["masks"],
                      maskname=inp["masknames"], affine_array=inp["affines"], vertex_array=inp["vertices"],
                      face_array=inp["faces"], surfname=inp["surfnames"], adjacency_array=inp["adjacency"],
                      tmi_history=["history mode_add 20261017000000 1 2 1 1 2"], append_history=False)
    el, img, masks, masknames, aff, vert, face, surfnames, adj, hist, cols = tm_io.read_tm_filetype(path, verbose=False)
    assert np.array_equal(img[0], inp["data"]) and all(np.array_equal(a, b) for a, b in zip(masks, inp["masks"]))
    assert np.array_equal(vert[0], inp["vertices"][0].astype(np.float32)) and np.array_equal(face[0], inp["faces"][0])
    assert np.array_equal(cols[0], inp["column_ids"]) and np.array_equal(aff[0], inp["affines"][0].astype(np.float32))
    assert all(list(x) == list(y) for a, b in zip(adj, inp["adjacency"]) for x, y in zip(a, b))
    out = dict(elements=np.array(el), image=img[0], masknames=np.array(masknames), surfnames=np.array(surfnames),
               history=np.array(hist), affine0=aff[0], vertex0=vert[0], face0=face[0], columns=cols[0])
    for i, m in enumerate(masks):
        out["mask%d" % i] = m
    for i, a in enumerate(adj):
        out["adj%d_indptr" % i] = np.concatenate([[0], np.cumsum([len(x) for x in a])])
        out["adj%d_indices" % i] = np.concatenate([np.asarray(list(x), dtype=np.int64) for x in a])
    np.savez_compressed(os.path.join(HERE, "tmi_reader.npz"), **out)
    print("sample.tmi (%d bytes) and tmi_reader.npz written" % os.path.getsize(path), {k: np.shape(v) for k, v in out.items()})


if __name__ == "__main__":
    main()
