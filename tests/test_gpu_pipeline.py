"""The streaming TFCE pipeline (csrc/tfce_pipeline.cu) against the oracle and against the one-kernel basin sweep:
max-only maps (leader accumulators), the class path (weights / maps), over-capacity maps redone by the sweep
kernel, voxel adjacency with 32-slot rows, degenerate rows.  Reference: lib/fast_tfce.hpp:11-95 via oracle/."""
import os

import numpy as np
import pytest

import oracle
from tests import helpers

pytestmark = pytest.mark.gpu


def _adjset(H, E, csr):
    from tfce_mediation_b200.tfce import CreateAdjSet
    return CreateAdjSet(H, E, csr)


def _maps(csr, B, seed0):
    """Smooth, rough (white noise: one basin per ~7 vertices) and degenerate rows."""
    V = csr[0].shape[0] - 1
    rows = [helpers.smooth_map(csr, seed0 + b, b % 5) for b in range(B - 4)]
    rows.append(np.zeros(V, dtype=np.float32))                                  # nothing active
    rows.append(-np.abs(helpers.smooth_map(csr, seed0 + 50, 2)) - np.float32(0.1))  # only the negative side
    c = np.full(V, 1.5, dtype=np.float32); c[::7] = 0.25                        # plateaus: ties broken by index
    rows.append(c)
    r = helpers.smooth_map(csr, seed0 + 51, 1); r[r < 0.3] = 0                   # exact zeros (masked vertices)
    rows.append(r)
    return np.ascontiguousarray(np.stack(rows), dtype=np.float32)


def _check_max(plan, stat, surfaces_spec, two_sided=True):
    import torch
    mx, status, _ = plan.run(torch.from_numpy(stat).cuda(), two_sided=two_sided)
    mx = mx.cpu().numpy()
    for b in range(stat.shape[0]):
        for s, (csr, off, V, H, E, w) in enumerate(surfaces_spec):
            want = helpers.oracle_signed_max(H, E, csr, stat[b, off:off + V], w)
            assert mx[b, s, 0] == want[0], (b, s, mx[b, s, 0], want[0])
            if two_sided:
                assert mx[b, s, 1] == want[1], (b, s, mx[b, s, 1], want[1])
    return mx


def test_max_only_bitexact_two_surfaces():
    from tfce_mediation_b200.engine import Surface, TfcePlan
    _, _, csr5 = helpers.ico(5)
    _, _, csr4 = helpers.ico(4)
    V5, V4 = csr5[0].shape[0] - 1, csr4[0].shape[0] - 1
    plan = TfcePlan([Surface(_adjset(2, 0.67, csr5), 0), Surface(_adjset(2, 1.0, csr4), V5)])
    B = 12
    stat = np.zeros((B, V5 + V4 + 5), dtype=np.float32)
    stat[:, :V5] = _maps(csr5, B, 300)
    stat[:, V5:V5 + V4] = _maps(csr4, B, 400)[::-1]
    spec = [(csr5, 0, V5, 2, 0.67, None), (csr4, V5, V4, 2, 1.0, None)]
    _check_max(plan, stat, spec, two_sided=True)
    _check_max(plan, stat, spec, two_sided=False)


def test_max_only_equals_maximum_of_the_maps():
    """The leader shortcut (one accumulator per live root) returns exactly the maximum of the full TFCE map."""
    import torch
    from tfce_mediation_b200.engine import Surface, TfcePlan
    _, _, csr = helpers.ico(5)
    plan = TfcePlan([Surface(_adjset(2, 0.67, csr), 0)])
    stat = _maps(csr, 10, 500)
    d = torch.from_numpy(stat).cuda()
    mx_only, _, _ = plan.run(d, two_sided=True)
    mx_maps, _, (pos, neg) = plan.run(d, two_sided=True, want_maps=True)
    assert torch.equal(mx_only, mx_maps)
    pos, neg = pos.cpu().numpy(), neg.cpu().numpy()
    for b in range(stat.shape[0]):
        assert np.array_equal(pos[b], oracle.tfce_run(2, 0.67, csr, stat[b]) if stat[b].max() > 0 else np.zeros_like(stat[b]))


def test_pipeline_equals_one_kernel_sweep(monkeypatch):
    import torch
    from tfce_mediation_b200.engine import Surface, TfcePlan
    _, _, csr = helpers.ico(5)
    V = csr[0].shape[0] - 1
    w = (0.5 + np.random.RandomState(3).rand(V)).astype(np.float32)
    stat = torch.from_numpy(_maps(csr, 9, 600)).cuda()
    out = {}
    for mode in ("pipeline", "basin"):
        if mode == "basin":
            monkeypatch.setenv("TMB_TFCE", "basin")
        for weighted in (False, True):
            plan = TfcePlan([Surface(_adjset(2, 0.67, csr), 0, w if weighted else None)])
            mx, st, _ = plan.run(stat, two_sided=True)
            out[(mode, weighted)] = (mx.cpu().numpy(), st.cpu().numpy())
    for weighted in (False, True):
        assert np.array_equal(out[("pipeline", weighted)][0], out[("basin", weighted)][0])
        assert np.array_equal(out[("pipeline", weighted)][1], out[("basin", weighted)][1])


@pytest.mark.parametrize("env", [{"TMB_PIPE_NBCAP": "16"}, {"TMB_PIPE_PAIRCAP": "64"}])
def test_over_capacity_maps_are_redone_by_the_sweep_kernel(monkeypatch, env):
    from tfce_mediation_b200.engine import Surface, TfcePlan
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    _, _, csr = helpers.ico(5)
    V = csr[0].shape[0] - 1
    plan = TfcePlan([Surface(_adjset(2, 0.67, csr), 0)])
    stat = _maps(csr, 8, 700)
    _check_max(plan, stat, [(csr, 0, V, 2, 0.67, None)])


def test_voxel_adjacency_wide_rows():
    """26-connectivity (up to 26 neighbours: 32-slot fixed-width rows, voxel 0 dropped -> symmetric here)."""
    from tfce_mediation_b200.engine import Surface, TfcePlan
    rs = np.random.RandomState(5)
    mask = rs.rand(14, 15, 13) < 0.55
    mask[0, 0, 0] = False
    idx = -np.ones(mask.shape, dtype=np.int64)
    idx[mask] = np.arange(mask.sum())
    adj = [[] for _ in range(int(mask.sum()))]
    for x, y, z in zip(*np.where(mask)):
        for dx in (-1, 0, 1):
            for dy in (-1, 0, 1):
                for dz in (-1, 0, 1):
                    if dx or dy or dz:
                        xx, yy, zz = x + dx, y + dy, z + dz
                        if 0 <= xx < mask.shape[0] and 0 <= yy < mask.shape[1] and 0 <= zz < mask.shape[2] and mask[xx, yy, zz]:
                            adj[idx[x, y, z]].append(int(idx[xx, yy, zz]))
    V = len(adj)
    indptr = np.zeros(V + 1, dtype=np.int64)
    indptr[1:] = np.cumsum([len(a) for a in adj])
    csr = (indptr, np.concatenate([np.asarray(a, dtype=np.int32) for a in adj]))
    plan = TfcePlan([Surface(_adjset(2, 0.5, csr), 0)])
    stat = np.stack([rs.standard_normal(V).astype(np.float32) for _ in range(6)])
    _check_max(plan, stat, [(csr, 0, V, 2, 0.5, None)])


@pytest.mark.parametrize("level,geom", [(6, "1"), (6, "2"), (5, "1"), (5, "2"), (5, "3"), (5, "4"), (5, "0"), (4, "0")])
def test_max_only_sweep_geometries(monkeypatch, level, geom):
    """TMB_PIPE_GEOM: one 1,024-thread sweep CTA per SM (1), two of 512 (2), four of 256 (3), eight of 128 (4), the large
    geometry taking the maps a smaller one cannot hold (rough maps: thousands of basins); 0 = chosen by surface size (two
    CTAs per SM above 40,000 vertices, four up to that, eight up to 16,000).  All bit-exact against the oracle, with and
    without vertex weights."""
    from tfce_mediation_b200.engine import Surface, TfcePlan
    monkeypatch.setenv("TMB_PIPE_GEOM", geom)
    _, _, csr = helpers.ico(level)
    V = csr[0].shape[0] - 1
    plan = TfcePlan([Surface(_adjset(2, 0.67, csr), 0)])
    rs = np.random.RandomState(77)
    rows = [helpers.smooth_map(csr, 900 + b, b) for b in range(3)]            # white noise, then smoother
    rows.append(rs.standard_normal(V).astype(np.float32))                     # ~ one basin per 7 vertices
    rows.append(np.abs(rs.standard_normal(V)).astype(np.float32))             # one sign only
    stat = np.ascontiguousarray(np.stack(rows), dtype=np.float32)
    _check_max(plan, stat, [(csr, 0, V, 2, 0.67, None)], two_sided=True)
    if level == 5:
        w = (0.5 + rs.rand(V)).astype(np.float32)
        planw = TfcePlan([Surface(_adjset(2, 0.67, csr), 0, w)])
        _check_max(planw, stat, [(csr, 0, V, 2, 0.67, w)], two_sided=True)
