"""Mismatch census of the float32 t-maps (VERDICT r1 'parity gaps' 1): one full block of the BASELINE config-2 workload --
`--shuffles` permuted designs x 299,881 vertices -- from the fused GPU fit (tmb_glm_tstat, fp64 tensor cores, fp32-seeded
epilogue) against the reference's own compiled cynumstats.tval_int (oracle/_ref) cast to float32 the way its callers do
(pyfunc.py:112-113), value by value, bitwise.  Also: are the FWER rows of the shuffles that contain a differing value
still identical?  Writes one JSON object (stdout and --out).  Test infrastructure: runs the checker, not the product."""
import argparse
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
_W = {}


def _worker(args):
    path, shape, lo, hi = args
    try:
        from threadpoolctl import threadpool_limits
        lim = threadpool_limits(limits=1)
    except Exception:
        lim = None
    import bench
    from oracle import build_ref
    if "w" not in _W:
        _W["w"] = bench.build_workload("config2")
        _W["tval"] = build_ref.load()[1].tval_int
    w = _W["w"]
    gpu = np.memmap(path, dtype=np.float32, mode="r", shape=shape)
    n, X, y = w["n"], w["X"], w["y"]
    out = []
    for p in range(lo, hi):
        np.random.seed(w["seed_base"] + p)
        nx = X[np.random.permutation(list(range(n)))]
        t = _W["tval"](nx, np.linalg.inv(np.dot(nx.T, nx)), y, n, 2, y.shape[1])[1].astype(np.float32)
        g = np.asarray(gpu[p])
        diff = np.flatnonzero(g.view(np.int32) != t.view(np.int32))
        ulps = np.abs(g.view(np.int32)[diff].astype(np.int64) - t.view(np.int32)[diff].astype(np.int64)) if diff.size else np.zeros(0, np.int64)
        out.append((p, int(diff.size), int(ulps.max()) if diff.size else 0, diff[:8].tolist()))
    del lim
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shuffles", type=int, default=1024)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    import torch
    import bench
    from tfce_mediation_b200 import _lib
    w = bench.build_workload("config2")
    eng, _, _ = bench.make_engine(w, torch.device("cuda", 0))
    P = args.shuffles
    idx = bench.perm_rows(w, 0, P)
    t32 = eng.tstat_rowperm(w["X"], idx)                         # [P, 1, ld], internal column order
    t32 = eng.to_caller_order(t32)[:, 0, :eng.Y.V].contiguous()
    path = "/dev/shm/tmb_census_%d.f32" % os.getpid()
    mm = np.memmap(path, dtype=np.float32, mode="w+", shape=(P, eng.Y.V))
    mm[...] = t32.cpu().numpy()
    mm.flush()
    cores = min(os.cpu_count() or 1, P)
    bounds = np.linspace(0, P, cores + 1).astype(int)
    t0 = time.time()
    with mp.get_context("spawn").Pool(cores) as pool:
        parts = pool.map(_worker, [(path, (P, eng.Y.V), int(bounds[i]), int(bounds[i + 1])) for i in range(cores)])
    os.unlink(path)
    rows = [r for part in parts for r in part]
    bad = [r for r in rows if r[1]]
    total = P * eng.Y.V
    res = {"workload": "config2", "shuffles": P, "vertices": int(eng.Y.V), "values_compared": int(total),
           "mismatching_values": int(sum(r[1] for r in rows)), "shuffles_with_a_mismatch": len(bad),
           "max_ulp_distance": int(max([r[2] for r in rows] + [0])), "cpu_seconds": round(time.time() - t0, 1), "cores": cores,
           "reference": "oracle/_ref cynumstats.tval_int (compiled from /root/reference unmodified), fp64 -> astype(float32)",
           "abi_launches": _lib.launch_count()}
    if bad:
        # FWER rows of the affected shuffles, GPU pipeline against the reference pipeline
        ctx = bench._cpu_context(w)
        sel = [r[0] for r in bad][:8]
        got = eng.regression_block(w["X"], perm_idx=idx[sel])
        same = True
        for j, p in enumerate(sel):
            ref = bench.cpu_shuffles(w, ctx, p, 1)[0]
            g = bench.gpu_rows(w, got, j)
            same = same and all("%.4f" % a == "%.4f" % b for a, b in zip(ref, g))
        res["rows_of_affected_shuffles_identical"] = bool(same)
        res["affected_shuffles_checked"] = sel
        res["examples"] = [{"shuffle": r[0], "count": r[1], "max_ulp": r[2], "first_vertices": r[3]} for r in bad[:8]]
    line = json.dumps(res)
    print(line)
    if args.out:
        with open(args.out, "w") as f:
            f.write(line + "\n")


if __name__ == "__main__":
    main()
