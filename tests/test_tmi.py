"""TMI container reader and the mmr-lr memory-mapping step: tm_io.py:284-444
read_tm_filetype and tm_mmr_rand_low_ram.py:138-199.  The reader is pinned to what the REAL reference reader returned for
tests/golden/sample.tmi (tests/golden/make_golden_tmi.py -> tmi_reader.npz)."""
import argparse
import os

import numpy as np
import pytest

import oracle
from tests import helpers

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _csr_of(obj_array):
    indptr = np.concatenate([[0], np.cumsum([len(x) for x in obj_array])])
    indices = np.concatenate([np.asarray(list(x), dtype=np.int64) for x in obj_array])
    return indptr, indices


def test_reader_matches_reference_reader_golden():
    from tfce_mediation_b200.tm_io import read_tm_filetype
    g = np.load(os.path.join(G, "tmi_reader.npz"))
    el, img, masks, masknames, aff, vert, face, surfnames, adj, hist, cols = read_tm_filetype(
        os.path.join(G, "sample.tmi"), verbose=False)
    assert list(el) == list(g["elements"]) and list(masknames) == list(g["masknames"])
    assert list(surfnames) == list(g["surfnames"]) and list(hist) == list(g["history"])
    assert len(img) == 1 and img[0].dtype == g["image"].dtype and np.array_equal(img[0], g["image"])
    assert len(masks) == 2
    for i, m in enumerate(masks):
        assert m.dtype == bool and np.array_equal(m, g["mask%d" % i])
    assert np.array_equal(aff[0], g["affine0"]) and np.array_equal(vert[0], g["vertex0"])
    assert face[0].dtype == g["face0"].dtype and np.array_equal(face[0], g["face0"])
    assert np.array_equal(cols[0], g["columns"])
    for i, a in enumerate(adj):
        ip, ix = _csr_of(a)
        assert np.array_equal(ip, g["adj%d_indptr" % i]) and np.array_equal(ix, g["adj%d_indices" % i])


def test_reader_rejects_foreign_files(tmp_path):
    from tfce_mediation_b200.tm_io import read_tm_filetype
    p = tmp_path / "x.tmi"
    p.write_bytes(b"ply\nformat ascii 1.0\nend_header\n")
    with pytest.raises(ValueError):
        read_tm_filetype(str(p), verbose=False)
    p.write_bytes(b"tmi\nformat binary_big_endian 0.1\nend_header\n")
    with pytest.raises(ValueError):
        read_tm_filetype(str(p), verbose=False)


def test_setup_from_tmi_writes_reference_state(tmp_path):
    """tm_mmr_rand_low_ram.py:155-199 without covariates (host only): masks, adjacency, density weights, data."""
    from tfce_mediation_b200.tm_io import read_tm_filetype
    from tfce_mediation_b200.tm_multisurface.mmr_lr_randomise import setup_from_tmi
    tmi = os.path.join(G, "sample.tmi")
    tmp = str(tmp_path / "tmi_temp")
    assert setup_from_tmi(tmi, tmp) == 2
    _, img, masks, _, _, _, _, _, adj, _, _ = read_tm_filetype(tmi, verbose=False)
    nmesh = int(masks[0].sum())
    d0, d1 = np.load(tmp + "/0_data_temp.npy"), np.load(tmp + "/1_data_temp.npy")
    assert d0.dtype == np.float32 and d0.flags.c_contiguous
    assert np.array_equal(d0, img[0][:nmesh].T) and np.array_equal(d1, img[0][nmesh:].T)
    assert np.array_equal(np.load(tmp + "/0_mask_temp.npy"), masks[0][:, 0, 0])
    assert np.load(tmp + "/1_mask_temp.npy").all() and np.load(tmp + "/1_mask_temp.npy").shape == (int(masks[1].sum()),)
    deg = np.array([len(x) for x in adj[0]], dtype=np.float64)[masks[0][:, 0, 0]]
    want = (1 - deg / deg.max() + deg.mean() / deg.max()).astype(np.float32)
    assert np.array_equal(np.load(tmp + "/0_vdensity_temp.npy"), want)
    a0 = np.load(tmp + "/0_adjacency_temp.npy", allow_pickle=True)
    assert len(a0) == len(adj[0]) and all(list(x) == list(y) for x, y in zip(a0, adj[0]))
    tmp2 = str(tmp_path / "nw")
    setup_from_tmi(tmi, tmp2, noweight=True)
    assert np.array_equal(np.load(tmp2 + "/0_vdensity_temp.npy"), np.array([1], dtype=np.float32))


@pytest.mark.gpu
def test_mmr_lr_from_tmi_rows_match_oracle(tmp_path, monkeypatch):
    """mmr-lr straight from a TMI container (with covariates): per-surface '%f' rows against the oracle's restatement of
    low_ram_calculate_tfce (tm_func.py:144-185) on the same tmi_temp state."""
    from tfce_mediation_b200.tm_multisurface import mmr_lr_randomise as drv
    from tfce_mediation_b200 import synth
    tmi = os.path.join(G, "sample.tmi")
    monkeypatch.chdir(tmp_path)
    rs = np.random.RandomState(2)
    n = 9
    np.savetxt("pred.csv", rs.standard_normal(n), delimiter=",")
    np.savetxt("cov.csv", rs.standard_normal(n), delimiter=",")
    opts = drv.getArgumentParser(argparse.ArgumentParser()).parse_args(
        ["-pr", "0", "3", "--path", "out", "--seed", "40", "-i", "pred.csv", "--tmifile", tmi, "-c", "cov.csv",
         "--tfce", "2", "0.67", "2", "1", "--assigntfcesettings", "0", "1"])
    drv.run(opts)
    pred = np.genfromtxt("pred.csv", delimiter=",")
    for sn, (H, E) in enumerate([(2, 0.67), (2, 1.0)]):
        data = np.load("tmi_temp/%d_data_temp.npy" % sn)
        mask = np.load("tmi_temp/%d_mask_temp.npy" % sn)
        adjacency = np.load("tmi_temp/%d_adjacency_temp.npy" % sn, allow_pickle=True)
        vdensity = np.load("tmi_temp/%d_vdensity_temp.npy" % sn)
        csr = oracle.adjacency_to_csr(adjacency)
        run = helpers.oracle_run(H, E, csr)
        got = np.array([float(l) for l in open("out/perm_maxTFCE_surf%d_tcon1.csv" % sn)])
        want = []
        for p in range(4):
            rows = oracle.low_ram_max(data, mask, pred, run, vdensity, p, 40)
            want += [rows[0][0], rows[0][1]]
        assert np.allclose(got, np.array(want, dtype=np.float64), rtol=1e-5, atol=2e-6)


def test_writer_reproduces_golden_container(tmp_path):
    """write_tm_filetype (tm_io.py:72-282): the file the REFERENCE reader was verified to read back to the inputs
    (tests/golden/make_golden_tmi.py) is reproduced byte for byte; an existing name is never overwritten silently."""
    import sys
    from tfce_mediation_b200.tm_io import read_tm_filetype, write_tm_filetype
    sys.path.insert(0, G)
    try:
        from make_golden_tmi import sample_inputs
    finally:
        sys.path.remove(G)
    inp = sample_inputs()
    kw = dict(columnids=inp["column_ids"], image_array=inp["data"], masking_array=inp["masks"], maskname=inp["masknames"],
              affine_array=inp["affines"], vertex_array=inp["vertices"], face_array=inp["faces"], surfname=inp["surfnames"],
              adjacency_array=inp["adjacency"])
    name = write_tm_filetype(str(tmp_path / "out"), tmi_history=["history mode_add 20261017000000 1 2 1 1 2"],
                             append_history=False, **kw)
    assert name.endswith("out.tmi")
    assert open(name, "rb").read() == open(os.path.join(G, "sample.tmi"), "rb").read()
    # appended history: counts net of what the passed history already records; existing file -> new_<name>
    name2 = write_tm_filetype(name, tmi_history=["history mode_add 20261017000000 1 1 1 1 1"], **kw)
    assert os.path.basename(name2) == "new_out.tmi"
    hist = read_tm_filetype(name2, verbose=False)[9]
    assert len(hist) == 2 and hist[1].split()[1] == "mode_add" and hist[1].split()[3:] == ["1", "1", "0", "0", "1"]
    with pytest.raises(NotImplementedError):
        write_tm_filetype(str(tmp_path / "a"), output_binary=False, **kw)
