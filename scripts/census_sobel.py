"""Mismatch census of the float32 Sobel-z maps: one block of the BASELINE config-4 workload (medtype 'M') -- `--shuffles`
permuted predictors x 299,881 vertices -- from tmb_sobelz_cross (one contraction row per shuffle, float32-seeded epilogue)
against the reference's calc_sobelz arithmetic (pyfunc.py:130-162) evaluated with its own compiled cynumstats.calc_beta_se
(oracle/_ref) and cast to float32 the way write_perm_maxTFCE_vertex does (pyfunc.py:112-113), value by value, bitwise.
Writes one JSON object (stdout and --out).  Test infrastructure: runs the checker, not the product."""
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
        _W["w"] = bench.build_workload("config4")
        _W["cbs"] = build_ref.load()[1].calc_beta_se
    w, cbs = _W["w"], _W["cbs"]
    gpu = np.memmap(path, dtype=np.float32, mode="r", shape=shape)
    n, y, V = w["n"], w["y"], w["y"].shape[1]
    out = []
    for p in range(lo, hi):
        np.random.seed(w["seed_base"] + p)
        xp = w["pred_x"][np.random.permutation(list(range(n)))]
        a_beta, a_se = cbs(xp, y, n, V)                                            # medtype 'M', pyfunc.py:137-141
        b_beta, b_se = cbs(np.column_stack([w["depend_y"], xp]), y, n, V)
        ta, tb = a_beta / a_se[1], b_beta / b_se[1]
        with np.errstate(divide="ignore", invalid="ignore"):
            z = (1 / np.sqrt((1 / (tb ** 2)) + (1 / (ta ** 2)) + (1 / (ta ** 2 * tb ** 2)))).astype(np.float32)
        g = np.asarray(gpu[p])
        diff = np.flatnonzero(g.view(np.int32) != z.view(np.int32))
        ulps = np.abs(g.view(np.int32)[diff].astype(np.int64) - z.view(np.int32)[diff].astype(np.int64)) if diff.size else np.zeros(0, np.int64)
        out.append((p, int(diff.size), int(ulps.max()) if diff.size else 0))
    del lim
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shuffles", type=int, default=256)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    import torch
    import bench
    w = bench.build_workload("config4")
    eng, _, _ = bench.make_engine(w, torch.device("cuda", 0))
    P = args.shuffles
    idx = bench.perm_rows(w, 0, P)
    assert eng.sobelz_cross_ok(w["medtype"])
    z32 = eng.sobelz(w["medtype"], w["pred_x"], w["depend_y"], idx)[:, :eng.Y.V].contiguous()   # caller order
    path = "/dev/shm/tmb_census_sobel_%d.f32" % os.getpid()
    mm = np.memmap(path, dtype=np.float32, mode="w+", shape=(P, eng.Y.V))
    mm[...] = z32.cpu().numpy()
    mm.flush()
    cores = min(os.cpu_count() or 1, P)
    bounds = np.linspace(0, P, cores + 1).astype(int)
    t0 = time.time()
    with mp.get_context("spawn").Pool(cores) as pool:
        parts = pool.map(_worker, [(path, (P, eng.Y.V), int(bounds[i]), int(bounds[i + 1])) for i in range(cores)])
    os.unlink(path)
    rows = [r for part in parts for r in part]
    bad = [r for r in rows if r[1]]
    res = {"workload": "config4 (Sobel 'M')", "shuffles": P, "vertices": int(eng.Y.V), "values_compared": int(P * eng.Y.V),
           "mismatching_values": int(sum(r[1] for r in rows)), "shuffles_with_a_mismatch": len(bad),
           "max_ulp_distance": int(max([r[2] for r in rows] + [0])), "cpu_seconds": round(time.time() - t0, 1), "cores": cores,
           "gpu": "tmb_sobelz_cross: one contraction row per shuffle, float32-seeded epilogue",
           "reference": "calc_sobelz arithmetic on oracle/_ref cynumstats.calc_beta_se (compiled from /root/reference unmodified), fp64 -> astype(float32)"}
    line = json.dumps(res)
    print(line)
    if args.out:
        with open(args.out, "w") as f:
            f.write(line + "\n")


if __name__ == "__main__":
    main()
