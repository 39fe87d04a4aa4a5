"""tm-models repeated-measures ANCOVA: pyfunc.py:1712-2280 reg_rm_ancova_{one,two}_bs_factor
as the permutation driver calls them (tmanalysis/tm_models_randomise.py:522-677).  CPU: the oracle restatement against the
golden fixture produced by the real reference (tests/golden/make_golden_rmancova.py), the host model (designs, shuffle
replay).  GPU: the statistics kernels against the golden fixture, the batched block and the driver against the oracle
pipeline."""
import argparse
import os

import numpy as np
import pytest

import oracle
from tests import helpers
from tfce_mediation_b200 import synth

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
F64_TOL = 1e-10        # |delta| <= tol * max(1, |value|) for float64 statistics (BASELINE.json north_star)
ITERS = 4


def _golden():
    return np.load(os.path.join(G, "rmancova.npz"))


def _close64(got, want, tol=F64_TOL):
    return np.all(np.abs(got - want) <= tol * np.maximum(1.0, np.abs(want)))


def _oracle_call(kind, work, g, cov, rand_array):
    if kind == "one":
        return np.stack(oracle.reg_rm_ancova_one_bs_factor(work, g["f1"], g["subjects"], dmy_covariates=cov,
                                                           rand_array=rand_array))
    return np.stack(oracle.reg_rm_ancova_two_bs_factor(work, g["f1"], g["f2"], g["subjects"], dmy_covariates=cov,
                                                       rand_array=rand_array))


# ------------------------------------------------------------------------------------------- CPU
@pytest.mark.parametrize("kind", ["one", "two"])
@pytest.mark.parametrize("tag", ["cov", "nocov"])
def test_oracle_rm_ancova_matches_reference_golden(kind, tag):
    g = _golden()
    cov = g["cov"] if tag == "cov" else None
    assert np.array_equal(_oracle_call(kind, g["data"].copy(), g, cov, None), g["%s_%s" % (kind, tag)])
    work = g["data"].copy()
    np.random.seed(int(g["seed"]))
    for it in range(ITERS):
        rand_array = np.random.permutation(list(range(g["f1"].shape[0])))
        assert np.array_equal(_oracle_call(kind, work, g, cov, rand_array), g["perm_%s_%s" % (kind, tag)][it])


def _emulate(model, Y, A, order, grp_rows):
    """numpy emulation of csrc/rmancova_kernels.cu for one shuffle (test-only): checks the host model -- designs,
    inverse blocks, program, row permutations -- without a GPU."""
    from tfce_mediation_b200 import rmancova as rm
    meta, rU = model.meta, model.rU
    D, nops, nout = int(meta[0]), int(meta[1]), int(meta[2])
    Y64 = Y.astype(np.float64)
    c = A.T @ Y64
    yy = np.sum((Y64 - Y64.mean(0)) ** 2, 0)
    Ys = Y[order]
    reg = {0: np.sum((Ys - np.mean(Ys, 0)) ** 2, 0).astype(np.float64)}
    w, at = 0.0, 0
    for sz in model.group_sizes:
        blk = Y64[grp_rows[at:at + sz]]
        w = w + np.sum((blk - blk.mean(0)) ** 2, 0)
        at += sz
    reg[1] = w
    for d in range(D):
        base = 8 + d * (2 + rm.MAX_COLUMNS)
        k, off = int(meta[base]), int(meta[base + 1])
        S = meta[base + 2:base + 2 + k]
        M = model.mats[off:off + k * k].reshape(k, k)
        reg[2 + d] = yy - np.einsum("iv,ij,jv->v", c[S], M, c[S])
    prog = meta[8 + D * (2 + rm.MAX_COLUMNS):]
    for o in range(nops):
        op, dst, a, b = (int(x) for x in prog[4 * o:4 * o + 4])
        reg[dst] = (reg[a] - reg[b] if op == 0 else reg[a] + reg[b] if op == 1 else reg[a] / model.consts[b] if op == 2
                    else reg[a] / reg[b] if op == 3 else np.zeros_like(yy))
    return np.stack([reg[int(r)] for r in prog[4 * nops:4 * nops + nout]])


def _replay(seed, n, N, iters):
    """The reference driver's draws: per iteration np.random.permutation(n), then the in-place shuffle of the data
    rows (its draws are those of shuffling an index array of the same length)."""
    np.random.seed(seed)
    pi = np.arange(N)
    out = []
    for _ in range(iters):
        rand_array = np.random.permutation(list(range(n)))
        np.random.shuffle(pi)
        out.append((pi.copy(), rand_array))
    return out


@pytest.mark.parametrize("kind", ["one", "two"])
@pytest.mark.parametrize("tag", ["cov", "nocov"])
def test_host_model_reproduces_reference_golden(kind, tag):
    from tfce_mediation_b200.rmancova import RmAncovaModel
    g = _golden()
    s, n, V = g["data"].shape
    cov = g["cov"] if tag == "cov" else None
    model = RmAncovaModel(n, s, [g["f1"]] if kind == "one" else [g["f1"], g["f2"]], g["subjects"], cov)
    Y = g["data"].reshape(s * n, V)
    A, order, grp = model.operands(None, [np.arange(n)])
    assert _close64(_emulate(model, Y, A, order[0], grp[0]), g["%s_%s" % (kind, tag)], 1e-9)
    draws = _replay(int(g["seed"]), n, s * n, ITERS)
    A, order, grp = model.operands([d[0] for d in draws], [d[1] for d in draws])
    for it in range(ITERS):
        got = _emulate(model, Y, A[:, it * model.rU:(it + 1) * model.rU], order[it], grp[it])
        assert _close64(got, g["perm_%s_%s" % (kind, tag)][it], 1e-9)


# ------------------------------------------------------------------------------------------- GPU
def _line_engine(data):
    from tfce_mediation_b200.engine import PermutationEngine
    from tfce_mediation_b200.tmanalysis import _common as C
    V = data.shape[1]
    adj = [[j for j in (i - 1, i + 1) if 0 <= j < V] for i in range(V)]
    return PermutationEngine(data, [C.masked_surface(adj, 2.0, 0.67)], two_sided=False)


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["one", "two"])
@pytest.mark.parametrize("tag", ["cov", "nocov"])
def test_rm_ancova_kernels_match_reference_golden(kind, tag):
    from tfce_mediation_b200.rmancova import RmAncovaModel
    g = _golden()
    s, n, V = g["data"].shape
    cov = g["cov"] if tag == "cov" else None
    model = RmAncovaModel(n, s, [g["f1"]] if kind == "one" else [g["f1"], g["f2"]], g["subjects"], cov)
    eng = _line_engine(g["data"].reshape(s * n, V))
    f32, f64 = eng.rm_ancova_stats(model, None, [np.arange(n)], want_f64=True)
    assert _close64(f64.cpu().numpy()[0, :, :V], g["%s_%s" % (kind, tag)], 1e-9)
    draws = _replay(int(g["seed"]), n, s * n, ITERS)
    f32, f64 = eng.rm_ancova_stats(model, [d[0] for d in draws], [d[1] for d in draws], want_f64=True)
    f32, f64 = f32.cpu().numpy()[:, :, :V], f64.cpu().numpy()[:, :, :V]
    assert f64.shape == (ITERS, model.nout, V)
    # SS_Total follows numpy's float32 accumulation in the shuffled row order, so every row agrees to float64 accuracy
    assert _close64(f64, g["perm_%s_%s" % (kind, tag)], 1e-9)
    assert np.array_equal(f32, f64.astype(np.float32))
    # chunked evaluation (a budget that holds one shuffle per chunk: the unstaged totals kernel) gives the same bits, and
    # so does the unstaged kernel on the whole block
    g32 = eng.rm_ancova_stats(model, [d[0] for d in draws], [d[1] for d in draws], budget=1.0)
    assert np.array_equal(g32.cpu().numpy()[:, :, :V], f32)
    os.environ["TMB_RM_TOTALS"] = "global"
    try:
        h32 = eng.rm_ancova_stats(model, [d[0] for d in draws], [d[1] for d in draws])
    finally:
        del os.environ["TMB_RM_TOTALS"]
    assert np.array_equal(h32.cpu().numpy()[:, :, :V], f32)


def _state(n=24, s=3, seed=33):
    v, f, csr = helpers.ico(3)
    keep_lh, keep_rh = synth.cap_mask(v, 600), synth.cap_mask(-v, 590)
    dens = synth.vertex_density(synth.kring_csr(csr, 2))
    rs = np.random.RandomState(seed)
    grp = np.arange(n) % 2
    f1 = (grp - grp.mean()).astype(np.float64)                    # two levels, demeaned -> [n]
    f2 = rs.standard_normal(n)
    f2 = f2 - f2.mean()
    subjects = np.eye(n)[:, 1:]
    cov = rs.standard_normal((n, 1))
    cov = cov - cov.mean(0)
    y = []
    for t in range(s):
        yt = np.hstack([synth.subject_data(n, csr, seed + 2 * t, 2)[:, keep_lh],
                        synth.subject_data(n, csr, seed + 2 * t + 1, 2)[:, keep_rh]]).astype(np.float32)
        yt[:, :120] += (np.float32(0.8) * grp)[:, None].astype(np.float32) + np.float32(0.3 * t)
        y.append(yt)
    return dict(v=v, csr=csr, keep_lh=keep_lh, keep_rh=keep_rh, dens=dens, y=np.stack(y), f1=f1, f2=f2, subjects=subjects,
                cov=cov, n=n, s=s)


def _oracle_rows(st, kind, seed_of, iters, cov):
    """Rows of the reference's loop: per iteration the draw, the in-place data shuffle inside the function, TFCE, max."""
    run = helpers.oracle_run(2, 0.67, st["csr"])
    nlh = int(st["keep_lh"].sum())
    work = st["y"].copy()
    rows = []
    for it in iters:
        np.random.seed(seed_of(it))
        rand_array = np.random.permutation(st["n"])
        if kind == "one":
            F = oracle.reg_rm_ancova_one_bs_factor(work, st["f1"], st["subjects"], dmy_covariates=cov, rand_array=rand_array)
        else:
            F = oracle.reg_rm_ancova_two_bs_factor(work, st["f1"], st["f2"], st["subjects"], dmy_covariates=cov,
                                                   rand_array=rand_array)
        rows.append([oracle.perm_max_vertex(f, nlh, st["keep_lh"], st["keep_rh"], run, run, st["dens"], st["dens"]) for f in F])
    return np.array(rows)


def _obj(lists):
    a = np.empty(len(lists), dtype=object)
    for i, l in enumerate(lists):
        a[i] = list(l)
    return a


def _rows(path):
    return np.array([float(l) for l in open(path)])


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["one", "two"])
def test_tm_models_randomise_rm_ancova_driver_rows(tmp_path, monkeypatch, kind):
    from tfce_mediation_b200.tmanalysis import tm_models_randomise as drv
    st = _state()
    name = "rmANCOVA1BS" if kind == "one" else "rmANCOVA2BS"
    d = os.path.join(str(tmp_path), "tmtemp_%s_area" % name)
    os.makedirs(d)
    adj = synth.csr_to_lists(st["csr"])
    np.save(d + "/data.npy", st["y"]); np.save(d + "/optstfce.npy", np.array([2, 0.67]))
    np.save(d + "/num_vertex_lh.npy", int(st["keep_lh"].sum()))
    np.save(d + "/mask_lh.npy", st["keep_lh"]); np.save(d + "/mask_rh.npy", st["keep_rh"])
    np.save(d + "/adjac_lh.npy", _obj(adj), allow_pickle=True); np.save(d + "/adjac_rh.npy", _obj(adj), allow_pickle=True)
    np.save(d + "/vdensity_lh.npy", st["dens"]); np.save(d + "/vdensity_rh.npy", st["dens"])
    np.save(d + "/dmy_factor1.npy", st["f1"]); np.save(d + "/dmy_subjects.npy", st["subjects"])
    np.save(d + "/dformat.npy", np.array(["short"])); np.save(d + "/dmy_covariates.npy", st["cov"])
    if kind == "two":
        np.save(d + "/dmy_factor2.npy", st["f2"]); np.save(d + "/factors.npy", np.array(["sex", "d", "score", "c"]))
        names = ["sex", "score", "sex.X.score", "time", "sex.X.time", "score.X.time", "sex.X.score.X.time"]
    else:
        np.save(d + "/factors.npy", np.array(["sex", "d"]))
        names = ["sex", "time", "sex.X.time"]
    monkeypatch.chdir(tmp_path)
    opts = drv.getArgumentParser(argparse.ArgumentParser()).parse_args(
        ["-r", "1", "4", "-s", "area", "-ofa" if kind == "one" else "-tfa", "--seed", "9"])
    drv.run(opts)
    want = _oracle_rows(st, kind, lambda it: it * 1000 + 9, range(1, 5), st["cov"])
    out = "output_%s_area/perm_%s" % (name, name)
    for j, nm in enumerate(names):
        assert np.allclose(_rows("%s/perm_Fstat_%s_TFCE_maxVertex.csv" % (out, nm)), want[:, j], rtol=1e-5, atol=6e-5)


@pytest.mark.gpu
def test_rm_ancova_driver_under_torchrun_gives_the_single_process_rows(tmp_path):
    """Two ranks (torchrun; both on GPU 0 over gloo when the box has one GPU): the range is sharded, every rank replays
    the cumulative shuffles from the first permutation of the range, rank 0 writes the rows in order -- the same rows
    as one process, which the test above checks against the oracle pipeline."""
    import subprocess
    import sys
    import torch
    st = _state()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = []
    for ranks in (1, 2):
        wd = os.path.join(str(tmp_path), "r%d" % ranks)
        d = os.path.join(wd, "tmtemp_rmANCOVA1BS_area")
        os.makedirs(d)
        adj = synth.csr_to_lists(st["csr"])
        np.save(d + "/data.npy", st["y"]); np.save(d + "/optstfce.npy", np.array([2, 0.67]))
        np.save(d + "/num_vertex_lh.npy", int(st["keep_lh"].sum()))
        np.save(d + "/mask_lh.npy", st["keep_lh"]); np.save(d + "/mask_rh.npy", st["keep_rh"])
        np.save(d + "/adjac_lh.npy", _obj(adj), allow_pickle=True); np.save(d + "/adjac_rh.npy", _obj(adj), allow_pickle=True)
        np.save(d + "/vdensity_lh.npy", st["dens"]); np.save(d + "/vdensity_rh.npy", st["dens"])
        np.save(d + "/dmy_factor1.npy", st["f1"]); np.save(d + "/dmy_subjects.npy", st["subjects"])
        np.save(d + "/dformat.npy", np.array(["short"])); np.save(d + "/dmy_covariates.npy", st["cov"])
        np.save(d + "/factors.npy", np.array(["sex", "d"]))
        script = os.path.join(wd, "run.py")
        with open(script, "w") as f:
            f.write("import sys, argparse\nsys.path.insert(0, %r)\n"
                    "from tfce_mediation_b200.tmanalysis import tm_models_randomise as drv\n"
                    "from tfce_mediation_b200.tmanalysis import _common as C\n"
                    "C.BLOCK = 2\n"
                    "drv.run(drv.getArgumentParser(argparse.ArgumentParser()).parse_args("
                    "['-r', '1', '6', '-s', 'area', '-ofa', '--seed', '9']))\n" % root)
        env = dict(os.environ)
        if ranks == 1:
            cmd = [sys.executable, script]
        else:
            if torch.cuda.device_count() < 2:
                env["TMB_ALLOW_SHARED_GPU"] = "1"
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                   "--master-addr", "127.0.0.1", "--master-port", "29647", script]
        subprocess.run(cmd, cwd=wd, env=env, check=True, timeout=600)
        out = os.path.join(wd, "output_rmANCOVA1BS_area/perm_rmANCOVA1BS")
        outs.append([open("%s/perm_Fstat_%s_TFCE_maxVertex.csv" % (out, nm)).read() for nm in ("sex", "time", "sex.X.time")])
    assert outs[0] == outs[1] and len(outs[0][0].splitlines()) == 6


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["one", "two"])
@pytest.mark.parametrize("tag", ["cov", "nocov"])
def test_reg_rm_ancova_dropins_match_reference_golden(kind, tag):
    """pyfunc.reg_rm_ancova_{one,two}_bs_factor with the reference's signatures: un-permuted, p-values, and the permutation
    form with its side effect -- the caller's data are shuffled in place with the global numpy stream, cumulatively."""
    from scipy.stats import f as f_dist
    from tfce_mediation_b200 import pyfunc
    g = _golden()
    cov = g["cov"] if tag == "cov" else None
    n = g["f1"].shape[0]

    def call(data, **kw):
        if kind == "one":
            return pyfunc.reg_rm_ancova_one_bs_factor(data, g["f1"], g["subjects"], dmy_covariates=cov, verbose=False, **kw)
        return pyfunc.reg_rm_ancova_two_bs_factor(data, g["f1"], g["f2"], g["subjects"], dmy_covariates=cov, verbose=False, **kw)

    want = g["%s_%s" % (kind, tag)]
    got = call(g["data"].copy())
    assert len(got) == want.shape[0] and _close64(np.stack(got), want, 1e-9)
    sig = call(g["data"].copy(), output_sig=True)
    assert len(sig) == 2 * want.shape[0]
    s_, c = 3, (0 if cov is None else cov.shape[1])
    df_a, df_b = 2, 1
    df_w = (n - 1) - df_a - c - ((df_b + df_a * df_b) if kind == "two" else 0)
    assert np.allclose(sig[want.shape[0]], 1 - f_dist.cdf(want[0], df_a, df_w), rtol=0, atol=1e-9)       # P of the first factor
    assert np.allclose(sig[-1], 1 - f_dist.cdf(want[-1], (df_a * df_b if kind == "two" else df_a) * (s_ - 1), df_w * (s_ - 1)),
                       rtol=0, atol=1e-9)
    work = g["data"].copy()
    np.random.seed(int(g["seed"]))
    for it in range(ITERS):
        rand_array = np.random.permutation(list(range(n)))
        got = call(work, rand_array=rand_array)
        assert _close64(np.stack(got), g["perm_%s_%s" % (kind, tag)][it], 1e-9), it
    assert not np.array_equal(work, g["data"])           # shuffled in place, like the reference
