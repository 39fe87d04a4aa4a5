"""CPU: host-side logic of the product (no GPU compute) and the C-ABI surface."""
import ctypes
import os
import re

import numpy as np
import pytest

import oracle
from tests import helpers

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_loads_and_exports_every_declared_symbol():
    from tfce_mediation_b200 import _lib
    L = _lib.lib()
    header = open(os.path.join(ROOT, "include", "tfce_b200.h")).read()
    declared = set(re.findall(r"\b(tmb_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert hasattr(L, name), name
    assert L.tmb_abi_version() == _lib.ABI_VERSION == 3
    assert isinstance(L.tmb_device_count(), int)
    assert L.tmb_last_error() is not None


def test_no_cpu_fallback_without_device():
    from tfce_mediation_b200 import _lib
    if _lib.lib().tmb_device_count() > 0:
        pytest.skip("a GPU is present")
    from tfce_mediation_b200.tfce import CreateAdjSet
    with pytest.raises(_lib.TmbError, match="no CPU fallback"):
        CreateAdjSet(2, 1, [[1], [0]])
    from tfce_mediation_b200 import cynumstats
    with pytest.raises(_lib.TmbError):
        cynumstats.tval_int(np.ones((3, 1)), np.ones((1, 1)), np.ones((3, 2), dtype=np.float32), 3, 1, 2)


def test_graph_create_validates_on_host():
    from tfce_mediation_b200 import _lib
    L = _lib.lib()
    h = ctypes.c_void_p()
    indptr = np.array([0, 1, 2], dtype=np.int64)
    bad = np.array([1, 5], dtype=np.int32)
    rc = L.tmb_graph_create(0, 2, _lib.ptr(indptr), _lib.ptr(bad), 2.0, 1.0, ctypes.byref(h))
    assert rc != 0 and b"out of range" in L.tmb_last_error()
    bad_ptr = np.array([0, 2, 1], dtype=np.int64)
    rc = L.tmb_graph_create(0, 2, _lib.ptr(bad_ptr), _lib.ptr(bad), 2.0, 1.0, ctypes.byref(h))
    assert rc != 0 and b"monotone" in L.tmb_last_error()


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "tfce_mediation_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", src, re.M), f
                assert "libtfce_oracle" not in src, f


def test_adjacency_to_csr_accepts_lists_sets_arrays():
    from tfce_mediation_b200._graph import adjacency_to_csr, csr_to_lists
    adj = helpers.grid_csr(5, 4)
    p, i = adjacency_to_csr(adj)
    op, oi = oracle.adjacency_to_csr(adj)
    assert np.array_equal(p, op) and np.array_equal(i, oi)
    assert csr_to_lists(p, i) == adj
    p2, i2 = adjacency_to_csr([set(a) for a in adj])
    assert np.array_equal(p2, p) and sorted(i2.tolist()) == sorted(i.tolist())
    arr = np.empty(len(adj), dtype=object)
    for j, a in enumerate(adj):
        arr[j] = np.array(a)
    p3, i3 = adjacency_to_csr(arr)
    assert np.array_equal(i3, i)
    pe, ie = adjacency_to_csr([[], [], []])
    assert pe.tolist() == [0, 0, 0, 0] and ie.shape == (0,)
    with pytest.raises(ValueError):
        adjacency_to_csr([[3], [0]])


def test_induced_subgraph_preserves_tfce_of_kept_vertices():
    from tfce_mediation_b200._graph import induced_subgraph
    _, _, csr = helpers.ico(3)
    V = csr[0].shape[0] - 1
    rs = np.random.RandomState(0)
    keep = rs.rand(V) < 0.8
    sub = induced_subgraph(csr[0], csr[1], keep)
    img = helpers.smooth_map(csr, 3, 2)
    full = np.where(keep, img, 0).astype(np.float32)        # masked-out vertices carry 0 (pyfunc.py:108-113)
    want = oracle.tfce_run(2, 0.67, csr, full)[keep]
    got = oracle.tfce_run(2, 0.67, sub, np.ascontiguousarray(img[keep]))
    assert np.array_equal(got, want)


def test_design_stack_centering_is_exact_to_fp64_noise():
    from tfce_mediation_b200.engine import design_stack, pack_At, row_permuted_stack
    rs = np.random.RandomState(1)
    n, V, k, P = 50, 400, 4, 5
    y = (rs.standard_normal((n, V)) + 3.0).astype(np.float32)
    X = np.column_stack([np.ones(n), rs.standard_normal((n, k - 1))])
    idx = np.stack([rs.permutation(n) for _ in range(P)])
    a = row_permuted_stack(X, idx)
    b = design_stack(np.stack([X[i] for i in idx]))
    np.testing.assert_allclose(a["pinv"], b["pinv"], rtol=1e-10, atol=1e-14)
    yc = y.astype(np.float64)
    yy = ((yc - yc.mean(0)) ** 2).sum(0)
    for p in range(P):
        nx = X[idx[p]]
        want = oracle.tval_int(nx, np.linalg.inv(nx.T @ nx), y, n, k, V)[1:]
        beta = a["pinv"][p] @ yc
        sse = yy - np.einsum("iv,ij,jv->v", beta, a["G"][p], beta)
        se = np.sqrt(sse / a["dof"] * a["d"][p][:, None]).astype(np.float32)
        t = beta / se
        assert np.all(np.abs(t - want) <= 1e-10 * np.maximum(1, np.abs(want)))
        assert np.mean(t.astype(np.float32) != want.astype(np.float32)) < 1e-3
    At, ldA = pack_At(a["pinv"], 4)
    assert ldA % 64 == 0 and At.shape == (n, ldA)
    assert np.array_equal(At[:, 4 * 2 + 1], a["pinv"][2, 1]) and not At[:, 4 * 2 + 3].any()


def test_synth_shapes():
    from tfce_mediation_b200 import synth
    v, f = synth.icosphere(3)
    assert v.shape == (642, 3) and f.shape == (1280, 3)
    csr = synth.faces_to_csr(642, f)
    deg = np.diff(csr[0])
    assert deg.min() == 5 and deg.max() == 6 and csr[1].shape[0] == 3840
    ref = oracle.vertex_adjacency(642, f)
    assert all(set(csr[1][csr[0][i]:csr[0][i + 1]].tolist()) == ref[i] for i in range(642))
    k2 = synth.kring_csr(csr, 2)
    assert np.diff(k2[0]).min() > 6
    assert synth.cap_mask(v, 600).sum() == 600


def test_step2_wrapper_rounding_and_blocks_follow_the_reference():
    """STEP_2_tfce_randomise_parallel.py:139-148: round(N/200)*100 shuffles (x2 for mediation), blocks of 100."""
    import argparse
    from tfce_mediation_b200.tmanalysis import STEP_2_tfce_randomise_parallel as s2
    for N, doubled in [(10000, False), (10000, True), (250, False), (300, False), (1000, True), (200, False)]:
        want = int(np.round(N / 200.0) * 100.0) * (2 if doubled else 1)           # the reference's own expression
        assert s2.rounded_shuffles(N, doubled) == want
        blocks = s2.command_blocks(N, doubled)
        assert blocks == [(i * 100 + 1, i * 100 + 100) for i in range(int(want / 100))]
        if blocks:
            assert blocks[0][0] == 1 and blocks[-1][1] == want
    p = s2.getArgumentParser(argparse.ArgumentParser())
    mod, argv = s2.driver_call(p.parse_args(["--vertex", "area", "-n", "10000", "-v", "1", "2", "--seed", "7"]))
    assert mod == "vertex_tfce_multiple_regression_randomise"
    assert argv == ["-r", "1", "5000", "-s", "area", "-v", "1", "2", "--seed", "7"]
    mod, argv = s2.driver_call(p.parse_args(["--voxel", "-n", "1000", "-m", "M", "-p", "8"]))
    assert mod == "voxel_tfce_mediation_randomise" and argv == ["-r", "1", "1000", "-m", "M"]
    # tm-models families (:104-133): the reference's flags; the count doubles for -ofa/-tfa/-cos/-mcos only (:142)
    assert s2.driver_call(p.parse_args(["--voxel", "-n", "1000", "-glm"])) == \
        ("tm_models_randomise", ["-r", "1", "500", "-v", "-glm"])
    assert s2.driver_call(p.parse_args(["--vertex", "area", "-n", "1000", "-med", "-e", "blocks.csv"])) == \
        ("tm_models_randomise", ["-r", "1", "500", "-s", "area", "-med", "-e", "blocks.csv"])
    for flag in ("-ofa", "-tfa", "-cos", "-mcos"):
        assert s2.driver_call(p.parse_args(["--vertex", "area", "-n", "1000", flag, "--seed", "3"])) == \
            ("tm_models_randomise", ["-r", "1", "1000", "-s", "area", flag, "--seed", "3"])


def test_bench_reads_measured_hbm_peak_tolerantly(tmp_path, monkeypatch):
    """bench.py's roofline denominator: MEASURED_PEAKS.json is driver-written, its key names are not ours."""
    import importlib.util
    import json
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    monkeypatch.setattr(bench, "ROOT", str(tmp_path))
    assert bench.measured_hbm_peak()[0] == 6650.0
    for doc, want in (({"hbm_gbs": 6555.5}, 6555.5), ({"hbm": {"burst_gbs": 7000, "sustained_gbs": 6500}}, 6500.0),
                      ({"HBM_TBps": 6.6, "bf16_tflops": 1800}, 6600.0), ({"bf16_tflops": 1800}, 6650.0)):
        (tmp_path / "MEASURED_PEAKS.json").write_text(json.dumps(doc))
        assert abs(bench.measured_hbm_peak()[0] - want) < 1e-6
    (tmp_path / "MEASURED_PEAKS.json").write_text("not json")
    val, src = bench.measured_hbm_peak()
    assert val == 6650.0 and "unreadable" in src


def test_linregress_t_vectorised_equals_scipy_per_shuffle():
    """The scalar path A of medtype 'Y' (pyfunc.py:142 scipy.stats.linregress per shuffle) is evaluated for all shuffles
    at once with scipy's own formulas; a few ulp from the per-call values."""
    from scipy.stats import linregress
    from tfce_mediation_b200.engine import linregress_t
    rs = np.random.RandomState(3)
    x = rs.standard_normal((40, 120))
    y = 0.3 * x + rs.standard_normal((40, 120))
    want = np.array([linregress(x[p], y[p])[0] / linregress(x[p], y[p])[4] for p in range(40)])
    assert np.all(np.abs(linregress_t(x, y) - want) <= 1e-13 * np.abs(want))


def test_tm_models_design_helpers_equal_reference_golden():
    """pyfunc.dummy_code / dummy_code_cosine / column_product / stack_ones / calc_indirect (pyfunc.py:2565-2709): host
    helpers of the tm-models scripts, against outputs of the real reference (tests/golden/make_golden_helpers.py)."""
    import os
    from tfce_mediation_b200 import pyfunc
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tm_models_helpers.npz"))
    eq = lambda a, b: np.array_equal(np.asarray(a), b, equal_nan=True) and np.asarray(a).shape == b.shape  # noqa: E731
    assert eq(pyfunc.dummy_code(g["grp"]), g["dc_grp"]) and eq(pyfunc.dummy_code(g["grp"], demean=False), g["dc_grp_raw"])
    assert eq(pyfunc.dummy_code(g["two"]), g["dc_two"])
    assert eq(pyfunc.dummy_code(g["cont"], iscontinous=True), g["dc_cont"])
    assert eq(pyfunc.dummy_code(g["cont"], iscontinous=True, demean=False), g["dc_cont_raw"])
    assert eq(pyfunc.dummy_code_cosine(g["t"], 12.0), g["dcc"])
    assert eq(pyfunc.column_product(g["a2"], g["b3"]), g["cp_22"]) and eq(pyfunc.column_product(g["cont"], g["b3"]), g["cp_12"])
    assert eq(pyfunc.column_product(g["a2"], g["cont"]), g["cp_21"]) and eq(pyfunc.column_product(g["cont"], g["t"]), g["cp_11"])
    assert eq(pyfunc.stack_ones(g["a2"]), g["so"])
    with np.errstate(invalid="ignore"):
        for alg, key in (("aroian", "ci_a"), ("sobel", "ci_s"), ("goodman", "ci_g")):
            assert eq(pyfunc.calc_indirect(g["ta"], g["tb"], alg=alg), g[key])


def test_block_for_scales_with_the_data_size():
    """tmanalysis/_common.block_for: about 6e8 vertex-maps per engine call, a power of two in [64, 8192]; an explicit C.BLOCK wins."""
    from tfce_mediation_b200.tmanalysis import _common as C

    class Eng(object):
        class Y(object):
            V = 0

    def blk(V):
        Eng.Y.V = V
        return C.block_for(Eng)

    assert blk(10242) == 8192 and blk(299881) == 1024 and blk(130781) == 4096 and blk(7000000) == 64 and blk(50000000) == 64
    old = C.BLOCK
    try:
        C.BLOCK = 2
        assert blk(299881) == 2
    finally:
        C.BLOCK = old
