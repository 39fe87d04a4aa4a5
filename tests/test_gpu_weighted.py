"""Weighted maxima without the per-vertex pass: the sweep keeps, per live root, the Pareto front of its classes'
(TFCE sum, largest weight) pairs (tfce_pipeline.cu, pipe_sweep_max_kernel<.., kW = true>).  The reference multiplies
per vertex -- max(tfce * (max(stat)/100) * vdensity), pyfunc.py:116-117, tm_func.py:173-174 -- so every test compares
with the oracle's per-vertex product, bit-exact, and with the class path (TMB_PIPE_WEIGHTS=class)."""
import numpy as np
import pytest

from tests import helpers
from tests.test_gpu_pipeline import _adjset, _check_max, _maps
from tfce_mediation_b200 import synth

pytestmark = pytest.mark.gpu


def _weights(kind, V, csr, seed):
    rs = np.random.RandomState(seed)
    if kind == "random":                       # every vertex its own weight: long fronts
        return (0.25 + rs.rand(V)).astype(np.float32)
    if kind == "few":                          # density-like: a handful of distinct values
        return rs.choice(np.array([0.5, 0.8, 1.0, 1.7], dtype=np.float32), V)
    if kind == "zeros":                        # zero weights are legal (product 0)
        w = (rs.rand(V) * 2).astype(np.float32)
        w[rs.rand(V) < 0.3] = 0
        return w
    if kind == "smooth":                       # spatially smooth weights (geodesic density looks like this)
        w = helpers.smooth_map(csr, seed, 4)
        return (1.0 + 0.4 * w / np.abs(w).max()).astype(np.float32)
    raise ValueError(kind)


@pytest.mark.parametrize("kind", ["random", "few", "zeros", "smooth"])
@pytest.mark.parametrize("level", [5, 6])
def test_weighted_leader_bitexact(kind, level):
    from tfce_mediation_b200.engine import Surface, TfcePlan
    _, _, csr = helpers.ico(level)
    V = csr[0].shape[0] - 1
    w = _weights(kind, V, csr, 31 + level)
    plan = TfcePlan([Surface(_adjset(2, 0.67, csr), 0, w)])
    stat = _maps(csr, 9 if level == 5 else 6, 2000 + level)
    _check_max(plan, stat, [(csr, 0, V, 2, 0.67, w)], two_sided=True)
    _check_max(plan, stat, [(csr, 0, V, 2, 0.67, w)], two_sided=False)


def test_weighted_leader_adversarial_weights():
    """Weights that grow as |statistic| falls: late (low) classes carry the large weights, so fronts keep growing."""
    from tfce_mediation_b200.engine import Surface, TfcePlan
    _, _, csr = helpers.ico(5)
    V = csr[0].shape[0] - 1
    stat = _maps(csr, 8, 2100)
    w = (1.0 / (0.05 + np.abs(stat[1]))).astype(np.float32)
    plan = TfcePlan([Surface(_adjset(2, 0.67, csr), 0, w)])
    _check_max(plan, stat, [(csr, 0, V, 2, 0.67, w)], two_sided=True)


def test_weighted_leader_equals_class_path(monkeypatch):
    import torch
    from tfce_mediation_b200.engine import Surface, TfcePlan
    _, _, csr = helpers.ico(6)
    k2 = synth.kring_csr(csr, 2)
    V = csr[0].shape[0] - 1
    rs = np.random.RandomState(8)
    stat = torch.from_numpy(np.concatenate([_maps(csr, 8, 2200), rs.standard_normal((3, V)).astype(np.float32)])).cuda()
    for graph in (csr, k2):
        for kind in ("random", "few"):
            w = _weights(kind, V, graph, 5)
            out = []
            for mode in ("leader", "class"):
                monkeypatch.setenv("TMB_PIPE_WEIGHTS", mode)
                plan = TfcePlan([Surface(_adjset(2, 0.67, graph), 0, w)])
                mx, st, _ = plan.run(stat, two_sided=True)
                out.append(mx.cpu().numpy())
            assert np.array_equal(out[0], out[1])


def test_weighted_two_surfaces_one_unweighted():
    from tfce_mediation_b200.engine import Surface, TfcePlan
    _, _, csr5 = helpers.ico(5)
    _, _, csr4 = helpers.ico(4)
    V5, V4 = csr5[0].shape[0] - 1, csr4[0].shape[0] - 1
    w = synth.vertex_density(synth.kring_csr(csr5, 3)) * np.float32(1.3)
    plan = TfcePlan([Surface(_adjset(2, 0.67, csr5), 0, w), Surface(_adjset(2, 1.0, csr4), V5)])
    B = 9
    stat = np.zeros((B, V5 + V4), dtype=np.float32)
    stat[:, :V5] = _maps(csr5, B, 2300)
    stat[:, V5:] = _maps(csr4, B, 2310)
    _check_max(plan, stat, [(csr5, 0, V5, 2, 0.67, w), (csr4, V5, V4, 2, 1.0, None)], two_sided=True)


def test_negative_weights_take_the_class_path():
    """The leader argument needs weights >= 0; anything else goes through the per-vertex pass, still exact."""
    from tfce_mediation_b200.engine import Surface, TfcePlan
    _, _, csr = helpers.ico(5)
    V = csr[0].shape[0] - 1
    w = (np.random.RandomState(2).rand(V) - 0.3).astype(np.float32)
    plan = TfcePlan([Surface(_adjset(2, 0.67, csr), 0, w)])
    stat = _maps(csr, 6, 2400)
    import torch
    mx, _, _ = plan.run(torch.from_numpy(stat).cuda(), two_sided=True)
    mx = mx.cpu().numpy()
    for b in range(stat.shape[0]):
        want = helpers.oracle_signed_max(2, 0.67, csr, stat[b], w)
        # the reference's max over products may be negative-free here only where a non-negative product exists
        assert mx[b, 0, 0] == max(want[0], np.float32(0)) and mx[b, 0, 1] == max(want[1], np.float32(0))


@pytest.mark.parametrize("env", [{"TMB_PIPE_NBCAP": "16"}, {"TMB_PIPE_FRONT": "1"}])
def test_weighted_over_capacity(monkeypatch, env):
    """Maps the weighted sweep gives up on (too many basins; a front longer than allowed) are redone by the one-kernel
    sweep with the per-vertex product."""
    from tfce_mediation_b200.engine import Surface, TfcePlan
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    _, _, csr = helpers.ico(5)
    V = csr[0].shape[0] - 1
    w = _weights("random", V, csr, 77)
    plan = TfcePlan([Surface(_adjset(2, 0.67, csr), 0, w)])
    _check_max(plan, _maps(csr, 7, 2500), [(csr, 0, V, 2, 0.67, w)])
