"""GPU parity: TFCE kernels (through the C ABI / drop-in CreateAdjSet) against the CPU oracle.
Integer results (labels, extents) and fp32 TFCE values are required to be BIT-EXACT."""
import numpy as np
import pytest

import oracle
from tests import helpers

pytestmark = pytest.mark.gpu


def _adjset(H, E, csr_or_lists):
    from tfce_mediation_b200.tfce import CreateAdjSet
    return CreateAdjSet(H, E, csr_or_lists)


@pytest.mark.parametrize("H,E", [(2, 0.67), (2, 1), (2, 0.5)])
@pytest.mark.parametrize("kind", ["smooth", "white", "scaled"])
def test_run_bitexact_ico5(H, E, kind):
    _, _, csr = helpers.ico(5)
    img = {"smooth": helpers.smooth_map(csr, 1, 3), "white": helpers.smooth_map(csr, 2, 0),
           "scaled": helpers.smooth_map(csr, 3, 6, scale=37.5)}[kind]
    want = oracle.tfce_run(H, E, csr, img)
    got = np.zeros_like(img)
    c = _adjset(H, E, csr)
    c.run(img, got)
    assert np.array_equal(got, want)
    assert c.last_status == 0


def test_run_accumulates_into_enhn():
    _, _, csr = helpers.ico(4)
    img = helpers.smooth_map(csr, 5, 2)
    start = np.abs(helpers.smooth_map(csr, 6, 0)).astype(np.float32)
    want = oracle.tfce_run(2, 0.67, csr, img, start.copy())
    got = start.copy()
    _adjset(2, 0.67, csr).run(img, got)
    assert np.array_equal(got, want)


def test_run_general_H_within_tolerance():
    # H != 2: powf(T, H) is evaluated on the device; tolerance 1e-5 relative (north_star), bit-exact in practice
    _, _, csr = helpers.ico(4)
    img = helpers.smooth_map(csr, 7, 2)
    want = oracle.tfce_run(1.5, 0.8, csr, img)
    got = np.zeros_like(img)
    _adjset(1.5, 0.8, csr).run(img, got)
    np.testing.assert_allclose(got, want, rtol=1e-5, atol=0)


def test_worked_example_ring():
    ring = [[(i - 1) % 12, (i + 1) % 12] for i in range(12)]
    img = np.array([0, 1, 2, 3, 2, 1, 0, -1, 4, 4, 0, 0], dtype=np.float32)
    got = np.zeros_like(img)
    _adjset(2, 0.67, ring).run(img, got)
    assert np.array_equal(got, oracle.tfce_run(2, 0.67, oracle.adjacency_to_csr(ring), img))
    assert got[8] == got[9] and got[0] == 0 and got[7] == 0


def test_accepts_sets_and_object_arrays():
    adj_lists = helpers.grid_csr(12, 9)
    adj_sets = [set(a) for a in adj_lists]
    arr = np.empty(len(adj_lists), dtype=object)
    for i, a in enumerate(adj_lists):
        arr[i] = a
    img = np.random.RandomState(0).standard_normal(len(adj_lists)).astype(np.float32)
    ref = oracle.tfce_run(2, 1, oracle.adjacency_to_csr(adj_lists), img)
    for adj in (adj_lists, adj_sets, arr):
        got = np.zeros_like(img)
        _adjset(2, 1, adj).run(img, got)
        assert np.array_equal(got, ref)


def test_degenerate_maps():
    adj = helpers.grid_csr(8, 8)
    c = _adjset(2, 0.67, adj)
    V = 64
    # all negative: nothing happens (fast_tfce.hpp:39 loop never entered)
    got = np.zeros(V, dtype=np.float32)
    c.run(-np.ones(V, dtype=np.float32), got)
    assert not got.any() and c.last_status == 0
    # maximum exactly zero: the reference never returns; we return zeros and flag it
    c.run(np.zeros(V, dtype=np.float32), got)
    assert not got.any() and c.last_status == 1
    # NaN entries never activate
    img = np.random.RandomState(1).standard_normal(V).astype(np.float32)
    img[5] = np.nan
    want = oracle.tfce_run(2, 0.67, oracle.adjacency_to_csr(adj), img)
    got[:] = 0
    c.run(img, got)
    assert np.array_equal(got, want)
    # +inf maximum: thresholds degenerate, output stays zero like the reference
    img2 = img.copy(); img2[5] = np.inf
    got[:] = 0
    c.run(img2, got)
    assert not got.any()


def test_empty_and_isolated_vertices():
    adj = [[] for _ in range(10)]
    adj[2] = [3]; adj[3] = [2]
    img = np.arange(10, dtype=np.float32) - 3
    got = np.zeros_like(img)
    _adjset(2, 0.5, adj).run(img, got)
    assert np.array_equal(got, oracle.tfce_run(2, 0.5, oracle.adjacency_to_csr(adj), img))


def test_asymmetric_adjacency_follows_directed_rule():
    # voxel 0 lists its neighbours but nobody lists 0 (tools builder).
    rs = np.random.RandomState(3)
    adj = helpers.grid_csr(10, 10)
    adj = [[a for a in lst if a != 0] for lst in adj]       # nobody lists 0, 0 keeps its list
    csr = oracle.adjacency_to_csr(adj)
    for seed in range(4):
        img = rs.standard_normal(100).astype(np.float32)
        got = np.zeros_like(img)
        _adjset(2, 0.67, adj).run(img, got)
        assert np.array_equal(got, oracle.tfce_run(2, 0.67, csr, img))


def test_run_rejects_bad_buffers():
    adj = helpers.grid_csr(4, 4)
    c = _adjset(2, 1, adj)
    ok = np.zeros(16, dtype=np.float32)
    with pytest.raises(ValueError, match="Buffer dtype mismatch"):
        c.run(np.zeros(16, dtype=np.float64), ok)
    with pytest.raises(ValueError, match="not C-contiguous"):
        c.run(np.zeros(32, dtype=np.float32)[::2], ok)
    with pytest.raises(ValueError, match="wrong number of dimensions"):
        c.run(np.zeros((4, 4), dtype=np.float32), ok)
    with pytest.raises(TypeError):
        c.run([0.0] * 16, ok)
    with pytest.raises(ValueError):
        _adjset(2, 1, [[1], [7]])          # neighbour index out of range


@pytest.mark.parametrize("level", [5, 30, 60, 85, 99])
def test_component_labels_and_extents_bitexact(level):
    _, _, csr = helpers.ico(5)
    img = helpers.smooth_map(csr, 11, 2)
    c = _adjset(2, 0.67, csr)
    labels, extents, thr = c.components(img, level)
    want_l, want_e = oracle.tfce_components(csr, img, level)
    assert np.array_equal(labels, want_l)
    assert np.array_equal(extents, want_e)
    assert np.float32(thr) == oracle.tfce_thresholds(img.max())[level]


def test_plan_two_sided_max_and_maps_bitexact():
    import torch
    from tfce_mediation_b200.engine import Surface, TfcePlan
    _, _, csr5 = helpers.ico(5)
    _, _, csr4 = helpers.ico(4)
    V5, V4 = csr5[0].shape[0] - 1, csr4[0].shape[0] - 1
    rs = np.random.RandomState(4)
    w5 = (0.5 + rs.rand(V5)).astype(np.float32)
    s0 = Surface(_adjset(2, 0.67, csr5), 0, w5)
    s1 = Surface(_adjset(2, 1.0, csr4), V5)           # mixed (H, E) per surface like mmr-lr
    plan = TfcePlan([s0, s1])
    B = 7
    ld = V5 + V4 + 3
    stat = np.zeros((B, ld), dtype=np.float32)
    for b in range(B):
        stat[b, :V5] = helpers.smooth_map(csr5, 100 + b, b % 4)
        stat[b, V5:V5 + V4] = helpers.smooth_map(csr4, 200 + b, 2, scale=1 + b)
    mx, status, (pos, neg) = plan.run(torch.from_numpy(stat).cuda(), two_sided=True, want_maps=True)
    mx, pos, neg = mx.cpu().numpy(), pos.cpu().numpy(), neg.cpu().numpy()
    assert int(status.abs().sum()) == 0
    for b in range(B):
        for s, (csr, off, V, H, E, w) in enumerate([(csr5, 0, V5, 2, 0.67, w5), (csr4, V5, V4, 2, 1.0, None)]):
            x = stat[b, off:off + V]
            want = helpers.oracle_signed_max(H, E, csr, x, w)
            assert mx[b, s, 0] == want[0] and mx[b, s, 1] == want[1]
            assert np.array_equal(pos[b, off:off + V], oracle.tfce_run(H, E, csr, x))
            assert np.array_equal(neg[b, off:off + V], oracle.tfce_run(H, E, csr, -x))


def test_plan_device_tables_match_correctly_rounded_oracle():
    """exact_pow=False builds the threshold tables on the device with a correctly rounded height term;
    checked bit-exactly against the oracle's correctly-rounded variant, and against the libm oracle within
    the north_star tolerance (1e-5 relative; measured <= 2e-7)."""
    import torch
    from tfce_mediation_b200.engine import Surface, TfcePlan
    _, _, csr = helpers.ico(5)
    V = csr[0].shape[0] - 1
    plan = TfcePlan([Surface(_adjset(2, 0.67, csr), 0)])
    B = 6
    stat = np.stack([helpers.smooth_map(csr, 100 + b, b % 4) for b in range(B)])
    mx, status, (pos, neg) = plan.run(torch.from_numpy(stat).cuda(), two_sided=True, want_maps=True, exact_pow=False)
    pos, neg = pos.cpu().numpy(), neg.cpu().numpy()
    for b in range(B):
        assert np.array_equal(pos[b], oracle.tfce_run(2, 0.67, csr, stat[b], correctly_rounded_height=True))
        assert np.array_equal(neg[b], oracle.tfce_run(2, 0.67, csr, -stat[b], correctly_rounded_height=True))
        np.testing.assert_allclose(pos[b], oracle.tfce_run(2, 0.67, csr, stat[b]), rtol=1e-5)
