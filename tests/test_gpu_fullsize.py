"""GPU parity at BASELINE.json's full sizes (fsaverage hemisphere = icosphere-7, 163,842 vertices):
direct comparison with the oracle on a few maps, plus size-independent properties that need no oracle
(determinism, sign mirror, exact power-of-two scaling, invariance under vertex relabelling)."""
import numpy as np
import pytest

import oracle
from tests import helpers
from tfce_mediation_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ico7():
    v, f, csr = helpers.ico(7)
    return v, csr


@pytest.fixture(scope="module")
def plan7(ico7):
    from tfce_mediation_b200.engine import Surface, TfcePlan
    from tfce_mediation_b200.tfce import CreateAdjSet
    _, csr = ico7
    c = CreateAdjSet(2, 0.67, csr)
    return c, TfcePlan([Surface(c, 0)])


def test_fullsize_run_bitexact_vs_oracle(ico7, plan7):
    _, csr = ico7
    c, _ = plan7
    for seed, rounds in ((1, 6), (2, 0)):
        img = helpers.smooth_map(csr, seed, rounds)
        got = np.zeros_like(img)
        c.run(img, got)
        assert np.array_equal(got, oracle.tfce_run(2, 0.67, csr, img))


def test_fullsize_batch_properties(ico7, plan7):
    import torch
    _, csr = ico7
    _, plan = plan7
    V = csr[0].shape[0] - 1
    B = 6
    stat = np.stack([helpers.smooth_map(csr, 40 + b, 6 if b % 2 else 2) for b in range(B)])
    dev = torch.from_numpy(stat).cuda()
    mx1, st1, (p1, n1) = plan.run(dev, two_sided=True, want_maps=True)
    mx2, st2, (p2, n2) = plan.run(dev, two_sided=True, want_maps=True)
    # determinism: the lock-free union order must not leak into any output bit
    assert torch.equal(mx1, mx2) and torch.equal(p1, p2) and torch.equal(n1, n2)
    assert int(st1.abs().sum()) == 0
    # sign mirror: TFCE of -x on the negative side == TFCE of x on the positive side
    mxm, _, (pm, nm) = plan.run(-dev, two_sided=True, want_maps=True)
    assert torch.equal(pm, n1) and torch.equal(nm, p1) and torch.equal(mxm[:, :, 0], mx1[:, :, 1])
    # exact scaling by a power of two for H = 2: thresholds scale by c, height terms by c^2, all fp32-exact
    mxs, _, (ps, _) = plan.run(dev * 4.0, two_sided=True, want_maps=True)
    assert torch.equal(ps, p1 * 16.0)
    assert torch.equal(mxs, mx1 * 64.0)                 # value * (max/100): 16 * 4
    # the scaled maximum agrees with the maps
    d = torch.from_numpy((stat.max(axis=1) / np.float32(100)).astype(np.float32)).cuda()
    assert torch.equal((p1[:, :V] * d[:, None]).max(dim=1).values, mx1[:, 0, 0])


def test_fullsize_relabelling_invariance(ico7):
    """TFCE does not depend on vertex numbering: relabel graph and map with a random permutation."""
    import torch
    from tfce_mediation_b200.engine import Surface, TfcePlan
    from tfce_mediation_b200.tfce import CreateAdjSet
    import scipy.sparse as sp
    _, csr = ico7
    V = csr[0].shape[0] - 1
    rs = np.random.RandomState(0)
    perm = rs.permutation(V)                      # new index i holds old vertex perm[i]
    inv = np.empty(V, dtype=np.int64); inv[perm] = np.arange(V)
    m = sp.csr_matrix((np.ones(csr[1].shape[0], dtype=np.int8), csr[1], csr[0]), shape=(V, V))
    m2 = m[perm][:, perm].tocsr(); m2.sort_indices()
    csr2 = (m2.indptr.astype(np.int64), m2.indices.astype(np.int32))
    img = helpers.smooth_map(csr, 77, 4)
    a = np.zeros_like(img); b = np.zeros_like(img)
    CreateAdjSet(2, 0.67, csr).run(img, a)
    CreateAdjSet(2, 0.67, csr2).run(np.ascontiguousarray(img[perm]), b)
    assert np.array_equal(b, a[perm])


def test_fullsize_engine_rows_vs_oracle(ico7):
    """Two hemispheres of fsaverage size through fit -> TFCE -> max, compared with the oracle pipeline."""
    from tfce_mediation_b200.engine import PermutationEngine, Surface
    from tfce_mediation_b200.tfce import CreateAdjSet
    v, csr = ico7
    V = csr[0].shape[0] - 1
    n, P = 40, 3
    y = np.concatenate([synth.subject_data(n, csr, 5, 4), synth.subject_data(n, csr, 6, 4)], axis=1)
    rs = np.random.RandomState(3)
    X = np.column_stack([np.ones(n), rs.standard_normal(n)])
    eng = PermutationEngine(y, [Surface(CreateAdjSet(2, 0.67, csr), 0), Surface(CreateAdjSet(2, 0.67, csr), V)])
    idx = np.stack([oracle.permutation_indices(2000 + p, n) for p in range(P)])
    got = eng.regression_block(X, perm_idx=idx)
    run = helpers.oracle_run(2, 0.67, csr)
    mask = np.ones(V, dtype=bool)
    for p in range(P):
        nx = X[idx[p]]
        t = oracle.tval_int(nx, np.linalg.inv(nx.T @ nx), y, n, 2, 2 * V)
        for sg, sign in enumerate((1.0, -1.0)):
            want = oracle.perm_max_vertex(t[1] * sign, V, mask, mask, run, run)
            assert "%.4f" % want == "%.4f" % max(got[p, 0, 0, sg], got[p, 0, 1, sg])


def test_high_degree_graph_csr_path_bitexact():
    """Degree > 32 (3-ring neighbourhoods, like the geodesic '3 mm' adjacency sets): no fixed-width rows, the
    kernels walk CSR rows in chunks of 8."""
    import torch
    from tfce_mediation_b200.engine import Surface, TfcePlan
    from tfce_mediation_b200.tfce import CreateAdjSet
    _, _, csr = helpers.ico(5)
    k3 = synth.kring_csr(csr, 3)
    assert np.diff(k3[0]).max() > 32
    V = csr[0].shape[0] - 1
    dens = synth.vertex_density(k3)
    c = CreateAdjSet(2, 0.67, k3)
    plan = TfcePlan([Surface(c, 0, dens)])
    stat = np.stack([helpers.smooth_map(csr, 300 + b, b % 3) for b in range(5)])
    mx, status, (pos, neg) = plan.run(torch.from_numpy(stat).cuda(), two_sided=True, want_maps=True)
    pos, neg, mx = pos.cpu().numpy(), neg.cpu().numpy(), mx.cpu().numpy()
    for b in range(5):
        assert np.array_equal(pos[b], oracle.tfce_run(2, 0.67, k3, stat[b]))
        assert np.array_equal(neg[b], oracle.tfce_run(2, 0.67, k3, -stat[b]))
        want = helpers.oracle_signed_max(2, 0.67, k3, stat[b], dens)
        assert mx[b, 0, 0] == want[0] and mx[b, 0, 1] == want[1]


def test_many_basins_32bit_ids_bitexact():
    """White noise on a 655,362-vertex sphere: > 65,535 basins, so basin ids take the 32-bit path and the
    per-basin arrays no longer fit in shared memory (global fallback)."""
    from tfce_mediation_b200.tfce import CreateAdjSet
    _, _, csr = helpers.ico(8)
    V = csr[0].shape[0] - 1
    img = np.random.RandomState(5).standard_normal(V).astype(np.float32)
    got = np.zeros_like(img)
    c = CreateAdjSet(2, 0.67, csr)
    c.run(img, got)
    assert np.array_equal(got, oracle.tfce_run(2, 0.67, csr, img))
    labels, extents, _ = c.components(img, 60)
    wl, we = oracle.tfce_components(csr, img, 60)
    assert np.array_equal(labels, wl) and np.array_equal(extents, we)
