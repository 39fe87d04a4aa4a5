"""Functional check + timing of the other BASELINE.json configurations at (near) full size.
config3: voxel skeleton (~130k voxels, 26-connectivity built by the GPU adjacency kernel), H=2 E=0.5
config4: Sobel mediation (medtype M) on fsaverage lh+rh
config5: mmr-lr style multi-surface job (S surfaces of icosphere-7 + small voxel pieces, mixed H/E)
A few shuffles of each are compared with the oracle pipeline."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import oracle
from tfce_mediation_b200 import synth, pyfunc
from tfce_mediation_b200.engine import PermutationEngine, Surface
from tfce_mediation_b200.tfce import CreateAdjSet

which = sys.argv[1] if len(sys.argv) > 1 else "3"
P = int(sys.argv[2]) if len(sys.argv) > 2 else 128
rs = np.random.RandomState(0)

def timed(fn):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); r = fn(); b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b), r

if which == "3":
    n = 300
    t0 = time.time()
    mask = synth.skeleton_mask((91, 109, 91), 0.29, seed=2)
    adj = pyfunc.create_adjac_voxel(mask, mask.astype(np.float32), int(mask.sum()), 26)
    V = len(adj)
    csr = oracle.adjacency_to_csr(list(adj))
    print("config3: %d voxels, nnz %d, mean degree %.1f, adjacency %.2fs (GPU builder)" % (V, csr[1].shape[0], csr[1].shape[0] / V, time.time() - t0), flush=True)
    y = rs.standard_normal((n, V)).astype(np.float32)
    y = synth.smooth_columns(y, csr, 2); y = ((y - y.mean(0)) / y.std(0)).astype(np.float32)
    X = np.column_stack([np.ones(n), rs.standard_normal(n)])
    eng = PermutationEngine(y, [Surface(CreateAdjSet(2, 0.5, csr), 0)], two_sided=True, nan_to_zero=True)
    idx = np.stack([oracle.permutation_indices(3000 + p, n) for p in range(P)])
    ms, got = timed(lambda: eng.regression_block(X, perm_idx=idx))
    print("config3: %d shuffles in %.2f ms -> %.0f shuffles/s" % (P, ms, P / ms * 1e3), flush=True)
    run = lambda img, out: oracle.tfce_run(2, 0.5, csr, img, out)
    for p in range(2):
        nx = X[idx[p]]
        t = oracle.tval_int(nx, np.linalg.inv(nx.T @ nx), y, n, 2, V); t[np.isnan(t)] = 0
        for sg, sign in enumerate((1, -1)):
            want = oracle.perm_max_voxel(t[1] * sign, run)
            print("  shuffle %d sign %d: gpu %.4f oracle %.4f %s" % (p, sg, got[p, 0, 0, sg], want, "OK" if "%1.4f" % want == "%1.4f" % got[p, 0, 0, sg] else "MISMATCH"))
elif which == "4":
    n = 300
    v, f = synth.icosphere(7); csr = synth.faces_to_csr(v.shape[0], f); V = v.shape[0]
    px = rs.standard_normal(n); dep = 0.5 * px + rs.standard_normal(n)
    y = np.concatenate([synth.subject_data(n, csr, 1, 6), synth.subject_data(n, csr, 2, 6)], axis=1)
    y = (y + 0.2 * px[:, None] + 0.2 * dep[:, None]).astype(np.float32)
    eng = PermutationEngine(y, [Surface(CreateAdjSet(2, 0.67, csr), 0), Surface(CreateAdjSet(2, 0.67, csr), V)], two_sided=False)
    idx = np.stack([oracle.permutation_indices(4000 + p, n) for p in range(P)])
    ms, got = timed(lambda: eng.mediation_block("M", px, dep, idx))
    print("config4: %d shuffles in %.2f ms -> %.0f shuffles/s" % (P, ms, P / ms * 1e3), flush=True)
    run = lambda img, out: oracle.tfce_run(2, 0.67, csr, img, out)
    mask = np.ones(V, dtype=bool)
    for p in range(2):
        z = oracle.sobelz("M", px[idx[p]], dep, y, n, 2 * V)
        want = oracle.perm_max_vertex(z, V, mask, mask, run, run)
        have = max(got[p, 0], got[p, 1])
        print("  shuffle %d: gpu %.4f oracle %.4f rel %.2e" % (p, have, want, abs(have - want) / want))
elif which == "5":
    S_big = int(sys.argv[3]) if len(sys.argv) > 3 else 8
    n = 350
    v, f = synth.icosphere(7); csr = synth.faces_to_csr(v.shape[0], f); V = v.shape[0]
    v5, f5 = synth.icosphere(5); csr5 = synth.faces_to_csr(v5.shape[0], f5); V5 = v5.shape[0]
    g7, g5 = CreateAdjSet(2, 0.67, csr), CreateAdjSet(2, 1.0, csr5)
    dens7, dens5 = synth.vertex_density(synth.kring_csr(csr5, 2)), None
    surfs, off = [], 0
    cols = []
    base7 = synth.subject_data(n, csr, 1, 6); base5 = synth.subject_data(n, csr5, 2, 3)
    for s in range(S_big):
        surfs.append(Surface(g7, off)); off += V; cols.append(np.roll(base7, s, axis=0))
    for s in range(2):
        surfs.append(Surface(g5, off)); off += V5; cols.append(np.roll(base5, s, axis=0))
    y = np.ascontiguousarray(np.hstack(cols), dtype=np.float32)
    print("config5-lite: %d surfaces, %d vertices, data %.2f GB" % (len(surfs), off, y.nbytes / 1e9), flush=True)
    X = np.column_stack([np.ones(n), rs.standard_normal(n)])
    eng = PermutationEngine(y, surfs, two_sided=True)
    idx = np.stack([oracle.permutation_indices(5000 + p, n) for p in range(P)])
    ms, got = timed(lambda: eng.regression_block(X, perm_idx=idx))
    print("config5-lite: %d shuffles x %d surfaces in %.2f ms -> %.1f shuffles/s (%.0f surface-maps/s)" % (P, len(surfs), ms, P / ms * 1e3, P * len(surfs) / ms * 1e3), flush=True)
    run7 = lambda img, out: oracle.tfce_run(2, 0.67, csr, img, out)
    nx = X[idx[0]]
    sidx = S_big - 1
    ysub = y[:, sidx * V:(sidx + 1) * V]
    t = oracle.tval_int(nx, np.linalg.inv(nx.T @ nx), ysub, n, 2, V)[1].astype(np.float32)
    for sg, sign in enumerate((1, -1)):
        img = np.ascontiguousarray(t * np.float32(sign)); tf = np.zeros_like(img); run7(img, tf)
        want = np.nanmax(tf * (img.max() / 100))
        print("  surface %d sign %d: gpu %f oracle %f %s" % (sidx, sg, got[0, 0, sidx, sg], want, "OK" if "%f" % want == "%f" % got[0, 0, sidx, sg] else "MISMATCH"))
