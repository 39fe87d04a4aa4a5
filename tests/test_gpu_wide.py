"""Sliced adjacency rows (SELL-32-4) in the streaming TFCE pipeline: graphs with more than 32 neighbours per vertex --
the reference's DEFAULT geodesic adjacency sets, ~60 per vertex, with vertex-density weights
(STEP_1_vertex_tfce_multiple_regression.py:71-76,155-175; pyfunc.py:107-119) -- and, forced with TMB_PIPE_ROWS=sell,
the narrow graphs as well.  Everything against the CPU oracle (lib/fast_tfce.hpp:11-95), bit-exact."""
import numpy as np
import pytest

import oracle
from tests import helpers
from tests.test_gpu_pipeline import _adjset, _check_max, _maps
from tfce_mediation_b200 import synth

pytestmark = pytest.mark.gpu


def _kring(level, rings):
    _, _, csr = helpers.ico(level)
    return csr, synth.kring_csr(csr, rings)


@pytest.mark.parametrize("level,rings,words", [(5, 4, 2), (4, 6, 4), (4, 8, 8), (5, 2, 1)])
def test_wide_rows_max_only_and_weighted(level, rings, words):
    """k-ring neighbourhoods: 60 (two mask words), ~126 (four), ~216 (eight) and 18 (one word, uneven fill) neighbours."""
    from tfce_mediation_b200.engine import Surface, TfcePlan
    csr1, csr = _kring(level, rings)
    deg = int(np.diff(csr[0]).max())
    assert (deg + 31) // 32 <= words and (words == 1 or deg > 16 * words)
    V = csr[0].shape[0] - 1
    dens = synth.vertex_density(csr)
    stat = _maps(csr1, 9, 1100 + rings)
    plan = TfcePlan([Surface(_adjset(2, 0.67, csr), 0)])
    _check_max(plan, stat, [(csr, 0, V, 2, 0.67, None)], two_sided=True)
    _check_max(plan, stat, [(csr, 0, V, 2, 0.67, None)], two_sided=False)
    w = (dens * (0.5 + np.random.RandomState(rings).rand(V))).astype(np.float32)    # spread-out weights
    planw = TfcePlan([Surface(_adjset(2, 0.67, csr), 0, w)])
    _check_max(planw, stat, [(csr, 0, V, 2, 0.67, w)], two_sided=True)


def test_wide_rows_maps_bitexact():
    import torch
    from tfce_mediation_b200.engine import Surface, TfcePlan
    csr1, csr = _kring(5, 4)
    plan = TfcePlan([Surface(_adjset(2, 0.67, csr), 0, synth.vertex_density(csr))])
    stat = _maps(csr1, 7, 1200)
    _, _, (pos, neg) = plan.run(torch.from_numpy(stat).cuda(), two_sided=True, want_maps=True)
    pos, neg = pos.cpu().numpy(), neg.cpu().numpy()
    for b in range(stat.shape[0]):
        wp = oracle.tfce_run(2, 0.67, csr, stat[b]) if stat[b].max() > 0 else np.zeros_like(stat[b])
        wn = oracle.tfce_run(2, 0.67, csr, -stat[b]) if (-stat[b]).max() > 0 else np.zeros_like(stat[b])
        assert np.array_equal(pos[b], wp) and np.array_equal(neg[b], wn)


@pytest.mark.parametrize("mixed", ["1", "0"])
def test_mixed_plan_narrow_and_wide_surfaces(monkeypatch, mixed):
    """One plan with two 1-ring meshes, a 4-ring mesh and different (H, E).  Default: the meshes keep their fixed-width
    rows and only the 4-ring surface runs on sliced rows, each group in its own launch over its surface-slot range (the
    slots are issued meshes first); TMB_PIPE_MIXED=0: every surface on sliced rows.  Maxima and full maps against the oracle."""
    import torch
    from tfce_mediation_b200.engine import Surface, TfcePlan
    monkeypatch.setenv("TMB_PIPE_MIXED", mixed)
    csr1, csr4 = _kring(5, 4)
    _, _, csr_small = helpers.ico(4)
    V5, V4 = csr1[0].shape[0] - 1, csr_small[0].shape[0] - 1
    w4 = synth.vertex_density(csr4)
    plan = TfcePlan([Surface(_adjset(2, 0.67, csr4), 0, w4), Surface(_adjset(2, 1.0, csr_small), V5),
                     Surface(_adjset(2, 0.5, csr1), V5 + V4)])
    B = 8
    stat = np.zeros((B, 2 * V5 + V4 + 3), dtype=np.float32)
    stat[:, :V5] = _maps(csr1, B, 1300)
    stat[:, V5:V5 + V4] = _maps(csr_small, B, 1310)
    stat[:, V5 + V4:2 * V5 + V4] = _maps(csr1, B, 1320)[::-1]
    spec = [(csr4, 0, V5, 2, 0.67, w4), (csr_small, V5, V4, 2, 1.0, None), (csr1, V5 + V4, V5, 2, 0.5, None)]
    _check_max(plan, stat, spec, two_sided=True)
    _, _, (pos, neg) = plan.run(torch.from_numpy(stat).cuda(), two_sided=True, want_maps=True)
    pos, neg = pos.cpu().numpy(), neg.cpu().numpy()
    for csr, off, V, H, E, _ in spec:
        for b in range(3):
            seg = np.ascontiguousarray(stat[b, off:off + V])
            wp = oracle.tfce_run(H, E, csr, seg) if seg.max() > 0 else np.zeros_like(seg)
            wn = oracle.tfce_run(H, E, csr, -seg) if (-seg).max() > 0 else np.zeros_like(seg)
            assert np.array_equal(pos[b, off:off + V], wp) and np.array_equal(neg[b, off:off + V], wn)


def test_forced_sliced_rows_equal_fixed_width_rows(monkeypatch):
    """TMB_PIPE_ROWS=sell on a 1-ring mesh and on a 26-connectivity-like graph: same maxima as the fixed-width kernels."""
    import torch
    from tfce_mediation_b200.engine import Surface, TfcePlan
    csr1, csr2 = _kring(5, 2)
    stat = torch.from_numpy(_maps(csr1, 10, 1400)).cuda()
    for csr in (csr1, csr2):
        out = []
        for rows in ("ell", "sell"):
            monkeypatch.setenv("TMB_PIPE_ROWS", rows)
            plan = TfcePlan([Surface(_adjset(2, 0.67, csr), 0)])
            mx, st, _ = plan.run(stat, two_sided=True)
            out.append((mx.cpu().numpy(), st.cpu().numpy()))
        assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])


@pytest.mark.parametrize("env", [{"TMB_PIPE_NBCAP": "16"}, {"TMB_PIPE_PAIRCAP": "8"}])
def test_wide_rows_over_capacity(monkeypatch, env):
    from tfce_mediation_b200.engine import Surface, TfcePlan
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    csr1, csr = _kring(5, 4)
    V = csr[0].shape[0] - 1
    plan = TfcePlan([Surface(_adjset(2, 0.67, csr), 0)])
    _check_max(plan, _maps(csr1, 6, 1500), [(csr, 0, V, 2, 0.67, None)])


def test_fullsize_reference_default_configuration():
    """BASELINE config 2(ii): fsaverage-size hemisphere (icosphere 7, cortex-mask-sized cut), '3 mm'-like 4-ring adjacency
    (~60 neighbours), vertex-density weights, H=2 E=0.67 -- the reference's default vertex configuration -- through the
    engine (fit + TFCE + weighted max) against tval_int + the oracle, rows compared as the reference prints them."""
    from tfce_mediation_b200._graph import induced_subgraph
    from tfce_mediation_b200.engine import PermutationEngine, Surface
    from tfce_mediation_b200.tfce import CreateAdjSet
    v, f, csr1 = helpers.ico(7)
    k4 = synth.kring_csr(csr1, 4)
    dens_full = synth.vertex_density(k4)
    keep = synth.cap_mask(v, 149955)
    sub = induced_subgraph(k4[0], k4[1], keep)
    dens = np.ascontiguousarray(dens_full[keep])
    n, P = 40, 3
    y = synth.subject_data(n, csr1, 11, 6)[:, keep]
    X = np.column_stack([np.ones(n), np.random.RandomState(4).standard_normal(n)])
    eng = PermutationEngine(y, [Surface(CreateAdjSet(2, 0.67, sub), 0, dens)])
    idx = np.stack([oracle.permutation_indices(2100 + p, n) for p in range(P)])
    got = eng.regression_block(X, perm_idx=idx)
    V = y.shape[1]
    for p in range(P):
        nx = X[idx[p]]
        t = oracle.tval_int(nx, np.linalg.inv(nx.T @ nx), y, n, 2, V)[1].astype(np.float32)
        want = helpers.oracle_signed_max(2, 0.67, sub, t, dens)
        for sg in range(2):
            assert got[p, 0, 0, sg] == want[sg], (p, sg, got[p, 0, 0, sg], want[sg])
