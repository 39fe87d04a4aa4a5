#!/usr/bin/env python
"""bench.py -- permutations/sec of the TFCE_mediation hot path (fit + TFCE + max) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload NAME] [--block P]
    python bench.py --job 10000 [--gpus N]          whole-job wall clock of the north-star run (see run_job)

One *step* = one block of `--block` shuffles per GPU through the hot path
(permuted OLS fit + t  ->  TFCE of +t and -t on every surface  ->  scaled max per surface;
mediation: two fits -> Sobel z -> one-sided TFCE -> max).  One permutation = one shuffled design (the
reference's `-n` counts two per shuffle for two-sided tests).

Workloads (BASELINE.json configs, synthetic data of the named shape):
    config1      vertex regression, fsaverage5 lh (10,242 vertices), n=100, 3 covariates residualised, k=2
    config2      vertex regression, fsaverage lh+rh (149,955 + 149,926 cortex vertices), n=300, k=2, 1-ring  [default]
    config2_3mm  the same with the reference's DEFAULT adjacency and weights: '3 mm'-like 4-ring neighbourhoods
                 (~60 neighbours) and vertex-density weights (STEP_1_vertex_tfce_multiple_regression.py:71-76,155-175)
    config3      voxel regression, ~130 k-voxel skeleton, 26-connectivity, H=2 E=0.5
    config4      vertex Sobel mediation ('M') on config2's graphs, one-sided
    config5      mmr-lr: 42 fsaverage-size meshes + 8 voxel pieces (~7.0 M vertices), n=350, mixed (H,E), density weights
    tiny         smoke size

Prints ONE JSON line: value = whole-job shuffles/s with inputs resident in HBM, device-timed, max over ranks;
`e2e` = the same through the engine's public calls with host index rows in / host maxima out; `roofline` for the
dominant stage (the TFCE pipeline) with a `fit` sub-object against a cuBLAS DGEMM peak measured in the run;
`cpu_baseline` = the reference's own compiled kernels (oracle/_ref) on one host core with the rows compared.
`--impl reference` times the reference's CPU implementation with all host cores.
"""
import time as _time_mod

T_PROCESS_START = _time_mod.time()

import argparse  # noqa: E402
import json  # noqa: E402
import os  # noqa: E402
import subprocess  # noqa: E402
import sys  # noqa: E402
import threading  # noqa: E402
import time  # noqa: E402

import numpy as np  # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "permutations/sec (regression+TFCE+max)"
TFCE_STAGE = "tfce pipeline (pipe_levels + pipe_ascent + pipe_basin + pipe_count + pipe_sweep_max kernels)"
UNIT = "permutations/s"
DEFAULT_BLOCK = {"config1": 8192, "config2": 1024, "config2_3mm": 512, "config3": 1024, "config4": 1024,
                 "config5": 128, "tiny": 64}


# ----------------------------------------------------------------------------------------- workloads
def _surface(csr_full, mask, H, E, weight_full=None):
    """One TFCE surface: the reference runs TFCE on the FULL graph with zeros outside `mask` (pyfunc.py:108-113);
    the GPU side uses the induced sub-graph of the mask, which gives the same values on the kept vertices."""
    from tfce_mediation_b200._graph import induced_subgraph
    if mask is None:
        sub = csr_full
        w = weight_full
    else:
        sub = induced_subgraph(csr_full[0], csr_full[1], mask)
        w = None if weight_full is None else np.ascontiguousarray(weight_full[mask])
    return dict(csr_full=csr_full, mask=mask, csr=sub, V=int(sub[0].shape[0] - 1), H=H, E=E, weight_full=weight_full,
                weight=w)


def build_workload(name, host_data=True, sample_surfaces=None):
    """Synthetic inputs of the named BASELINE.json configuration (seeded).
    host_data=False (config5 on the GPU arm): the subject data are generated on the device by device_data().
    sample_surfaces: build host data only for these surfaces (CPU legs of config5)."""
    from tfce_mediation_b200 import synth
    w = dict(name=name, kind="regression", rowkind="vertex", two_sided=True, nan_to_zero=False, k=2)
    rs = np.random.RandomState(1)
    if name in ("config2", "config2_3mm", "config4", "tiny"):
        if name == "tiny":
            level, n, rounds, keep = 4, 40, 2, (2400, 2300)
        else:
            level, n, rounds, keep = 7, 300, 6, (149955, 149926)
        v, f = synth.icosphere(level)
        csr1 = synth.faces_to_csr(v.shape[0], f)
        masks = [synth.cap_mask(v, keep[0]), synth.cap_mask(-v, keep[1])]
        csr, dens = csr1, None
        if name == "config2_3mm":
            csr = synth.kring_csr(csr1, 4)
            dens = synth.vertex_density(csr)          # from the FULL adjacency, like the reference (:166-173)
        w["surfaces"] = [_surface(csr, masks[h], 2.0, 0.67, dens) for h in range(2)]
        ys = [synth.subject_data(n, csr1, 1 + h, rounds)[:, masks[h]] for h in range(2)]
        y = np.ascontiguousarray(np.hstack(ys), dtype=np.float32)
        w.update(n=n, seed_base=2000)
        w["X"] = np.column_stack([np.ones(n), rs.standard_normal(n)])
        if name == "config4":
            rs4 = np.random.RandomState(3)
            px = rs4.standard_normal(n)
            dep = 0.5 * px + rs4.standard_normal(n)
            y = (y + np.float32(0.2) * px[:, None].astype(np.float32) + np.float32(0.2) * dep[:, None].astype(np.float32))
            w.update(kind="mediation", medtype="M", pred_x=px, depend_y=dep, two_sided=False, seed_base=4000)
        w["y"] = np.ascontiguousarray(y, dtype=np.float32)
    elif name == "config1":
        n = 100
        v, f = synth.icosphere(5)
        csr1 = synth.faces_to_csr(v.shape[0], f)
        w["surfaces"] = [_surface(csr1, None, 2.0, 0.67)]
        y = synth.subject_data(n, csr1, 0, 3)
        rs0 = np.random.RandomState(0)
        xc = np.column_stack([np.ones(n), rs0.standard_normal((n, 3))])        # intercept + 3 covariates
        y = (y.astype(np.float64) - xc @ (np.linalg.pinv(xc) @ y.astype(np.float64))).astype(np.float32)  # resid_covars, once
        w.update(n=n, seed_base=1000, y=np.ascontiguousarray(y), X=np.column_stack([np.ones(n), rs0.standard_normal(n)]))
    elif name == "config3":
        n = 300
        mask = synth.skeleton_mask((91, 109, 91), 0.25, seed=2)
        csr = synth.voxel_csr(mask, 26)
        w["surfaces"] = [_surface(csr, None, 2.0, 0.5)]
        rs3 = np.random.RandomState(2)
        y = rs3.standard_normal((n, csr[0].shape[0] - 1)).astype(np.float32)
        y = synth.smooth_columns(y, csr, 2)
        y = ((y - y.mean(0)) / y.std(0)).astype(np.float32)
        w.update(n=n, seed_base=3000, y=np.ascontiguousarray(y), rowkind="voxel", nan_to_zero=True,
                 X=np.column_stack([np.ones(n), rs3.standard_normal(n)]))
    elif name == "config5":
        n = 350
        v, f = synth.icosphere(7)
        csr7 = synth.faces_to_csr(v.shape[0], f)
        dens7 = synth.vertex_density(csr7)
        surfaces = [_surface(csr7, None, 2.0, 0.67, dens7) for _ in range(42)]
        for i in range(8):
            m = synth.skeleton_mask((40, 48, 40), 0.37, seed=20 + i, margin=4)
            c = synth.voxel_csr(m, 26)
            surfaces.append(_surface(c, None, 2.0, 1.0, synth.vertex_density(c)))
        w["surfaces"] = surfaces
        w.update(n=n, seed_base=5000, rowkind="mmr", X=np.column_stack([np.ones(n), np.random.RandomState(4).standard_normal(n)]))
        w["y"] = None
        w["smooth"] = [6] * 42 + [2] * 8
        if host_data or sample_surfaces is not None:
            w["y_host"] = {}
            for s in (sample_surfaces if sample_surfaces is not None else range(len(surfaces))):
                w["y_host"][s] = synth.subject_data(n, surfaces[s]["csr"], 100 + s, w["smooth"][s])
    else:
        raise SystemExit("unknown workload %r" % name)
    off = 0
    for s in w["surfaces"]:
        s["col"] = off
        off += s["V"]
    w["V_total"] = off
    return w


def device_data(w, dev):
    """config5: [n, sum V] float32 generated ON THE DEVICE (white noise, `smooth` rounds of (self + neighbours)
    averaging, standardised per vertex) -- 9.8 GB that would take minutes to synthesise on the host."""
    import torch
    n = w["n"]
    y = torch.empty((n, w["V_total"]), dtype=torch.float32, device=dev)
    g = torch.Generator(device=dev)
    cache = {}
    for si, s in enumerate(w["surfaces"]):
        ip, ix = s["csr"]
        key = id(s["csr_full"])
        if key not in cache:
            V = s["V"]
            rows = np.repeat(np.arange(V), np.diff(ip))
            deg = np.diff(ip).astype(np.float32) + 1.0
            ii = torch.from_numpy(np.stack([np.concatenate([rows, np.arange(V)]),
                                            np.concatenate([ix.astype(np.int64), np.arange(V)])])).to(dev)
            vv = torch.from_numpy(np.concatenate([1.0 / deg[rows], 1.0 / deg]).astype(np.float32)).to(dev)
            cache[key] = torch.sparse_coo_tensor(ii, vv, (V, V)).coalesce().to_sparse_csr()
        W = cache[key]
        g.manual_seed(100 + si)
        yt = torch.randn((s["V"], n), generator=g, device=dev, dtype=torch.float32)
        for _ in range(w["smooth"][si]):
            yt = torch.sparse.mm(W, yt)
        yt = (yt - yt.mean(dim=1, keepdim=True)) / yt.std(dim=1, unbiased=False, keepdim=True)
        y[:, s["col"]:s["col"] + s["V"]] = yt.t()
    return y


def perm_rows(w, first, count):
    """Permutation index rows generated with the reference's own numpy RNG calls."""
    out = np.empty((count, w["n"]), dtype=np.int64)
    for i in range(count):
        np.random.seed(w["seed_base"] + first + i)
        out[i] = np.random.permutation(list(range(w["n"])))
    return out


# ----------------------------------------------------------------------------------------- clocks
class ClockSampler(object):
    def __init__(self, index):
        self.index = index
        self.samples = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) >= 6:
                self.samples.append((time.time(), parts))

    def stop(self, t0=None, t1=None):
        """Summary over the samples that arrived inside [t0, t1] (the timed regions); all samples if none did."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        mhz, mx, reasons = [], None, set()
        inside = [p for (ts, p) in self.samples if t0 is None or (t0 <= ts <= t1 + 0.15)]
        if not inside:
            inside = [p for (_, p) in self.samples]
        for p in inside:
            try:
                mhz.append(float(p[0])); mx = float(p[1])
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(mhz)) if mhz else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(mhz)}


# ----------------------------------------------------------------------------------------- CPU reference
def _cpu_context(w, surfaces=None):
    """Per surface a CreateAdjSet.run-like callable on the FULL graph: the reference's own compiled kernels
    (oracle/_ref) when present, else the oracle port."""
    import oracle
    from oracle import build_ref
    from tfce_mediation_b200 import synth
    mods = build_ref.load()
    runs, keep, made = {}, [], {}
    for si in (surfaces if surfaces is not None else range(len(w["surfaces"]))):
        s = w["surfaces"][si]
        key = (id(s["csr_full"]), s["H"], s["E"])
        if key not in made:
            if mods is not None:
                c = mods[0].CreateAdjSet(s["H"], s["E"], synth.csr_to_lists(s["csr_full"]))
                keep.append(c)
                made[key] = c.run
            else:
                made[key] = (lambda img, out, s=s: oracle.tfce_run(s["H"], s["E"], s["csr_full"], img, out))
        runs[si] = made[key]
    tval = mods[1].tval_int if mods is not None else oracle.tval_int
    return dict(kind="reference" if mods is not None else "port", tval=tval, runs=runs, keep=keep)


def _cpu_surface_values(w, si, stat_kept, run):
    """(scaled TFCE values over the full surface, their maximum) for one statistic on surface si, composed like
    pyfunc.py:107-119 (vertex), :121-126 (voxel) and tm_func.py:160-182 (mmr-lr)."""
    s = w["surfaces"][si]
    mask = s["mask"]
    Vf = s["csr_full"][0].shape[0] - 1
    full = np.zeros(Vf, dtype=np.float32)
    if mask is None:
        full[:] = stat_kept
    else:
        full[mask] = stat_kept
    tf = np.zeros_like(full)
    run(full, tf)
    if w["rowkind"] == "voxel":
        return tf.max() * (full.max() / 100)
    wt = 1 if s["weight_full"] is None else s["weight_full"]
    if w["rowkind"] == "mmr":
        return np.nanmax((tf * (full.max() / 100) * wt).astype(np.float32))
    vals = tf[np.isfinite(tf)] * (full[np.isfinite(full)].max() / 100) * wt
    return vals.max()


def cpu_shuffles(w, ctx, first, count, surfaces=None, y_of=None):
    """The reference call sequence for `count` shuffles -- vertex_tfce_multiple_regression_randomise.py:104-117 +
    pyfunc.py:107-119; voxel_...:91-119; vertex_tfce_mediation_randomise.py:80-91 + pyfunc.py:130-162;
    tm_func.py:144-185 -- returning per shuffle the rows the reference would append: vertex/voxel kinds one value per
    sign (max over surfaces), mmr kind a dict surface -> (pos, neg)."""
    import oracle
    n, k = w["n"], w["k"]
    sel = list(surfaces) if surfaces is not None else list(range(len(w["surfaces"])))
    rows = []
    for i in range(count):
        np.random.seed(w["seed_base"] + first + i)
        perm = np.random.permutation(list(range(n)))
        per_surface = {}
        for si in sel:
            s = w["surfaces"][si]
            ys = y_of(si) if y_of is not None else w["y"][:, s["col"]:s["col"] + s["V"]]
            if w["kind"] == "mediation":
                z = oracle.sobelz(w["medtype"], w["pred_x"][perm], w["depend_y"], ys, n, ys.shape[1]).astype(np.float32)
                per_surface[si] = (_cpu_surface_values(w, si, z, ctx["runs"][si]),)
            else:
                nx = w["X"][perm]
                invXX = np.linalg.inv(np.dot(nx.T, nx))
                t = ctx["tval"](nx, invXX, ys, n, k, ys.shape[1])[1]
                if w["nan_to_zero"]:
                    t[np.isnan(t)] = 0
                t = t.astype(np.float32)
                per_surface[si] = (_cpu_surface_values(w, si, t, ctx["runs"][si]),
                                   _cpu_surface_values(w, si, -t, ctx["runs"][si]))
        if w["rowkind"] == "mmr":
            rows.append(per_surface)
        else:
            nsign = 1 if w["kind"] == "mediation" else 2
            rows.append(tuple(np.array([per_surface[si][g] for si in sel]).max() for g in range(nsign)))
    return rows


def row_format(w):
    return {"vertex": "%.4f", "voxel": "%1.4f", "mmr": "%f"}[w["rowkind"]]


def cpu_sample_plan(w):
    """(surfaces timed on the CPU, multipliers to extrapolate to the whole shuffle, shuffles) -- config5 times one
    mesh and one voxel piece (a surface subset extrapolated) because one full shuffle takes > 1 min."""
    if w["name"] == "config5":
        return [0, 42], {0: 42.0, 42: 8.0}, 1
    counts = {"config1": 40, "config2": 3, "config2_3mm": 2, "config3": 4, "config4": 2, "tiny": 10}
    return None, None, counts.get(w["name"], 3)


_REF = {}


def _ref_worker(args):
    name, first, count = args
    try:
        from threadpoolctl import threadpool_limits
        limiter = threadpool_limits(limits=1)
    except Exception:
        limiter = None
    if _REF.get("name") != name:
        surf, mult, _ = cpu_sample_plan({"name": name})
        w = build_workload(name, host_data=True, sample_surfaces=surf)
        _REF.update(name=name, w=w, ctx=_cpu_context(w, surf), surf=surf, mult=mult)
    w, surf, mult = _REF["w"], _REF["surf"], _REF["mult"]
    if count == 0:
        return 0.0
    y_of = (lambda si: w["y_host"][si]) if w["y"] is None else None
    if surf is None:
        t0 = time.perf_counter()
        cpu_shuffles(w, _REF["ctx"], first, count, None, y_of)
        dt = time.perf_counter() - t0
    else:
        dt = 0.0
        for si in surf:                       # time each sampled surface, scale by the number of surfaces like it
            t0 = time.perf_counter()
            cpu_shuffles(w, _REF["ctx"], first, count, [si], y_of)
            dt += (time.perf_counter() - t0) * mult[si]
    del limiter
    return dt


def run_reference(args):
    """--impl reference: the reference's CPU implementation with all host cores (one single-threaded worker
    process per core over disjoint shuffles, mirroring `parallel -j N`, STEP_2_tfce_randomise_parallel.py:153)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    per_worker = 1
    surf, mult, _ = cpu_sample_plan({"name": args.workload})
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        pool.map(_ref_worker, [(args.workload, 10 ** 6 + c, 0) for c in range(cores)])     # build inputs, untimed
        step_ms = []
        for step in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            dts = pool.map(_ref_worker, [(args.workload, (step * cores + c) * per_worker, per_worker) for c in range(cores)])
            dt = time.perf_counter() - t0
            if surf is not None:
                dt = max(dts)                  # extrapolated per-shuffle time of the slowest worker
            if step >= args.warmup:
                step_ms.append(dt * 1e3)
    shuffles = cores * per_worker * args.steps
    total_s = sum(step_ms) / 1e3
    value = shuffles / total_s
    from oracle import build_ref
    kind = "reference" if build_ref.load() is not None else "port"
    sample = "%d shuffles per step (%d worker processes x %d), OPENBLAS threads=1 each" % (cores * per_worker, cores, per_worker)
    if surf is not None:
        sample += "; per shuffle surfaces %s timed and scaled by %s (50 surfaces)" % (surf, [mult[s] for s in surf])
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": float(np.mean(step_ms)), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64 fit / f32 TFCE", "data": "synthetic",
        "config": {"workload": args.workload, "shuffles_per_step": cores * per_worker},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


# ----------------------------------------------------------------------------------------- B200 arm
def algorithmic_bytes_per_shuffle(w, C, signs):
    """C*4*sum(V) (t-map write) + g*C*sum[4V + 4(nnz + V + 1)] (each TFCE call
    reads its statistic map and CSR once, max-only output).  The Y stream term is dropped (P_batch large)."""
    sv = sum(s["V"] for s in w["surfaces"])
    per_call = sum(4 * s["V"] + 4 * (int(s["csr"][1].shape[0]) + s["V"] + 1) for s in w["surfaces"])
    return C * 4 * sv + signs * C * per_call, signs * C * per_call


def measured_hbm_peak():
    """HBM GB/s from the driver-written MEASURED_PEAKS.json (the SUSTAINED figure when both are given: the TFCE stage is
    timed inside a long step), else the profiling guide's fallback.  The file's exact key names are the driver's; any
    numeric entry whose key path mentions "hbm" is accepted, and a malformed file falls back instead of failing the run."""
    fallback = (6650.0, "fallback (B200_PROFILING.md)")
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if not os.path.exists(path):
        return fallback
    try:
        found = []

        def walk(node, trail):
            if isinstance(node, dict):
                for k, v in node.items():
                    walk(v, trail + [str(k).lower()])
            elif isinstance(node, (int, float)) and not isinstance(node, bool):
                name = ".".join(trail)
                if "hbm" in name and node > 0:
                    found.append((name, float(node)))

        walk(json.load(open(path)), [])
        if not found:
            return fallback
        found.sort(key=lambda kv: (0 if "sustain" in kv[0] else 1 if "burst" not in kv[0] else 2))
        name, val = found[0]
        if val < 100:            # TB/s
            val *= 1000.0
        return val, "measured (MEASURED_PEAKS.json %s)" % name
    except Exception as exc:    # noqa: BLE001
        return fallback[0], "fallback (B200_PROFILING.md; MEASURED_PEAKS.json unreadable: %s)" % type(exc).__name__


def measure_fp64_peak(dev, seconds=0.25):
    """fp64 TFLOP/s of a cuBLAS DGEMM (torch.matmul, 4096^3) on this GPU, best of a few after warm-up: the measured
    denominator of the fit's roofline (MEASURED_PEAKS.json has no fp64 entry)."""
    import torch
    N = 4096
    a = torch.randn((N, N), dtype=torch.float64, device=dev)
    b = torch.randn((N, N), dtype=torch.float64, device=dev)
    torch.matmul(a, b)
    torch.cuda.synchronize()
    best = 0.0
    t_end = time.time() + seconds
    while True:
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize()
        best = max(best, 2.0 * N ** 3 / (e0.elapsed_time(e1) / 1e3) / 1e12)
        if time.time() > t_end:
            break
    return best


def make_engine(w, dev, pin=True):
    """Graphs, surfaces and the engine with the data resident in HBM.  Returns (engine, upload ms, graph-build ms)."""
    import torch
    from tfce_mediation_b200.engine import PermutationEngine, Surface
    from tfce_mediation_b200.tfce import CreateAdjSet
    t0 = time.perf_counter()
    made, surfs = {}, []
    for s in w["surfaces"]:
        key = (id(s["csr_full"]), id(s["mask"]), s["H"], s["E"])
        if key not in made:
            made[key] = CreateAdjSet(s["H"], s["E"], s["csr"])
        surfs.append(Surface(made[key], s["col"], s["weight"]))
    graph_ms = (time.perf_counter() - t0) * 1e3
    if w["y"] is None:
        data = device_data(w, dev)
    else:
        data = torch.from_numpy(w["y"]).pin_memory() if pin else w["y"]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    eng = PermutationEngine(data, surfs, two_sided=w["two_sided"], nan_to_zero=w["nan_to_zero"])
    torch.cuda.synchronize()
    return eng, (time.perf_counter() - t0) * 1e3, graph_ms


def gpu_rows(w, res, p):
    """The rows of shuffle p as the reference would print them, from the engine's result array."""
    if w["kind"] == "mediation":                                   # [P, S]
        return (res[p].max(),)
    if w["rowkind"] == "mmr":                                      # [P, C, S, 2]
        return {si: (res[p, 0, si, 0], res[p, 0, si, 1]) for si in range(res.shape[2])}
    return (res[p, 0, :, 0].max(), res[p, 0, :, 1].max())


def run_b200(args):
    import torch
    import torch.distributed as dist
    from tfce_mediation_b200 import _lib

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()          # nvidia-smi needs a moment to start: launched first, filtered to the timed regions
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus %d needs torchrun (one rank per GPU)" % args.gpus)
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    w = build_workload(args.workload, host_data=False)
    eng, data_upload_ms, graph_ms = make_engine(w, dev)
    P = args.block
    C = w["k"] - 1
    S = len(w["surfaces"])
    signs = 2 if w["two_sided"] else 1
    X = w.get("X")
    total_steps = args.warmup + args.steps
    # every rank owns a contiguous range of the permutation index stream
    idx_all = [perm_rows(w, (rank * total_steps + s) * P, P) for s in range(total_steps)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident arm: per-step operands already in HBM, maxima stay on the device ----------
    from tfce_mediation_b200 import engine as E
    ld = eng.Y.ld
    yy = eng.Y.sumsq(True)
    L = _lib.lib()
    stacks = []
    # medtype 'M' / 'I': one contraction row per shuffle (dep'y is fitted once), else path A's row + path B's two rows
    cross = w["kind"] == "mediation" and eng.sobelz_cross_ok(w["medtype"])
    for s in range(total_steps):
        if w["kind"] == "mediation" and cross:
            stacks.append(idx_all[s])       # the index rows: the predictor columns are gathered on the device
        elif w["kind"] == "mediation":
            XA, XB, ta = eng.mediation_designs(w["medtype"], w["pred_x"], w["depend_y"], idx_all[s])
            stacks.append(eng.sobelz_operands(XA, XB, ta, "aroian", resident=True))
        else:
            st = E.row_permuted_stack(X, idx_all[s])
            At, ldA = E.pack_At(st["pinv"], 1)
            stacks.append((torch.from_numpy(At).to(dev), ldA, torch.from_numpy(st["G"]).to(dev),
                           torch.from_numpy(st["d"]).to(dev), st["dof"]))
    fit_rows = 3 if (w["kind"] == "mediation" and not cross) else 1   # contraction rows per shuffle
    t32b = [torch.empty((P, C, ld), dtype=torch.float32, device=dev) for _ in range(2)]
    out_max = torch.empty((P * C, S, 2), dtype=torch.float32, device=dev)
    comm = None
    if world > 1:
        from tfce_mediation_b200 import parallel
        comm = parallel.MaxComm.get()          # tmb_allgather_max behind the C ABI (NCCL over NVLink)
    fit_events = []

    def fit(s, buf):
        fa = None
        if len(fit_events) < 64:
            fa = torch.cuda.Event(enable_timing=True); fb = torch.cuda.Event(enable_timing=True)
            fa.record()
        if w["kind"] == "mediation" and cross:
            eng.sobelz_cross(w["medtype"], w["pred_x"], w["depend_y"], stacks[s], "aroian", out=buf.view(P, ld))
        elif w["kind"] == "mediation":
            eng.sobelz_launch(stacks[s], out=buf.view(P, ld))
        else:
            At_d, ldA, G_d, d_d, dof = stacks[s]
            _lib.check(L.tmb_glm_tstat(_lib.ptr(eng.Y.t), eng.Y.dtype_code, eng.Y.n, eng.Y.V, ld, _lib.ptr(At_d), ldA,
                                       _lib.ptr(G_d), _lib.ptr(d_d), P, 1, 1, 0, 1, dof, _lib.ptr(yy), _lib.ptr(buf), None,
                                       ld, 1 if w["nan_to_zero"] else 0, 0, _lib.current_stream()))
        if fa is not None:
            fb.record(); fit_events.append((fa, fb))
        return eng.plan.prepare(buf.view(P * C, ld))          # maxima kernel + async copy to the host

    def run_steps(first, count, tfce_events=None):
        """`count` steps, software-pipelined on one stream exactly like PermutationEngine.regression_blocks:
        fit + maxima of step s+1 are queued before the sweep of step s, whose exact-libm threshold tables the
        host builds meanwhile.  Every step does the full work: fit, maxima, host tables, sweep, (all-gather)."""
        tk = fit(first, t32b[0])
        for i in range(count):
            s = first + i
            nxt = fit(s + 1, t32b[(i + 1) & 1]) if i + 1 < count else None
            if tfce_events is not None:
                a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
                a.record()
            eng.plan.finish(tk, t32b[i & 1].view(P * C, ld), two_sided=w["two_sided"], out_max=out_max)
            if tfce_events is not None:
                b.record(); tfce_events.append((a, b))
            if world > 1 and args.gather == "step":
                comm.allgather(out_max)              # the per-shuffle maxima, tiny (NCCL over NVLink)
            tk = nxt
        if world > 1 and args.gather != "step":
            comm.allgather(out_max)                  # one collective per job

    run_steps(0, args.warmup)
    barrier()
    fit_events.clear()
    t_load0 = time.time()
    launches0 = _lib.launch_count()
    ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
    tfce_events = []
    ev0.record()
    run_steps(args.warmup, args.steps, tfce_events)
    ev1.record()
    barrier()
    dev_ms = ev0.elapsed_time(ev1)
    launches = _lib.launch_count() - launches0
    tfce_ms = float(np.mean([a.elapsed_time(b) for a, b in tfce_events]))
    fit_ms = float(np.mean([a.elapsed_time(b) for a, b in fit_events])) if fit_events else None

    # ---- end-to-end arm: the public call, host index rows in (pinned staging), host maxima out ----
    def e2e_call(idx):
        if w["kind"] == "mediation":
            return eng.mediation_blocks(w["medtype"], w["pred_x"], w["depend_y"], idx, block=P)
        return eng.regression_blocks(X, idx, block=P)

    e2e_call(np.concatenate(idx_all[:min(args.warmup, 2)], axis=0))
    idx_timed = np.concatenate(idx_all[args.warmup:], axis=0)
    barrier()
    eng.h2d_bytes = eng.d2h_bytes = 0
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    res = e2e_call(idx_timed)                                        # numpy on the host
    if world > 1:
        comm.allgather(torch.from_numpy(np.ascontiguousarray(res, dtype=np.float32)).to(dev))   # the job's maxima, once
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1)
    clocks = sampler.stop(t_load0, time.time()) if rank == 0 else None   # sampled during both timed regions
    times = torch.tensor([dev_ms, e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms = float(times[0]), float(times[1])
    total_shuffles = world * P * args.steps
    value = total_shuffles / (dev_ms / 1e3)
    e2e_value = total_shuffles / (e2e_ms / 1e3)

    if rank == 0:
        peak, peak_src = measured_hbm_peak()
        bytes_shuffle, bytes_tfce = algorithmic_bytes_per_shuffle(w, C, signs)
        achieved = bytes_tfce * P / (tfce_ms / 1e3) / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "tfce_sweep_traffic.json")
        if os.path.exists(tpath):
            for tj in json.load(open(tpath)).get("entries", []):
                if tj.get("workload") == w["name"] and tj.get("block") == P:
                    traffic = tj.get("dram_bytes_per_launch")
        fp64_peak = measure_fp64_peak(dev)
        cpu = None
        if world == 1 and not args.no_cpu:
            try:
                from threadpoolctl import threadpool_limits
                lim = threadpool_limits(limits=1)
            except Exception:
                lim = None
            surf, mult, sample = cpu_sample_plan(w)
            ctx = _cpu_context(w, surf)
            first = args.warmup * P
            if w["y"] is None:                                      # config5: the sampled surfaces' data back from the GPU
                Yd = eng.to_caller_order(eng.Y.t)
                host = {si: Yd[:, w["surfaces"][si]["col"]:w["surfaces"][si]["col"] + w["surfaces"][si]["V"]].cpu().numpy()
                        for si in surf}
                del Yd
                y_of = lambda si: host[si]                          # noqa: E731
            else:
                y_of = None
            if surf is None:
                t0 = time.perf_counter()
                rows = cpu_shuffles(w, ctx, first, sample, None, y_of)
                dt = time.perf_counter() - t0
            else:
                dt, rows = 0.0, [dict() for _ in range(sample)]
                for si in surf:
                    t0 = time.perf_counter()
                    part = cpu_shuffles(w, ctx, first, sample, [si], y_of)
                    dt += (time.perf_counter() - t0) * mult[si]
                    for p in range(sample):
                        rows[p].update(part[p])
            del lim
            # the same shuffles on the GPU: the FWER rows must agree
            idx = idx_all[args.warmup][:sample]
            chk = (eng.mediation_block(w["medtype"], w["pred_x"], w["depend_y"], idx) if w["kind"] == "mediation"
                   else eng.regression_block(X, perm_idx=idx))
            fmt = row_format(w)
            agree = True
            for p in range(sample):
                g = gpu_rows(w, chk, p)
                if w["rowkind"] == "mmr":
                    agree = agree and all(fmt % a == fmt % b for si in rows[p] for a, b in zip(rows[p][si], g[si]))
                else:
                    agree = agree and all(fmt % a == fmt % b for a, b in zip(rows[p], g))
            cpu = {"value": sample / dt, "unit": UNIT, "cores": 1, "kind": ctx["kind"],
                   "sample": "%d shuffles of the same workload%s, 1 process, BLAS threads=1; rows identical to GPU: %s"
                             % (sample, "" if surf is None else " (surfaces %s timed, scaled to all 50)" % surf, agree)}
        fit_flops = 2.0 * fit_rows * P * w["n"] * eng.Y.V
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64 fit / f32 TFCE", "data": "synthetic",
            "config": {"workload": w["name"], "kind": w["kind"], "vertices": [s["V"] for s in w["surfaces"]][:4] +
                       (["... %d surfaces, %d vertices" % (S, w["V_total"])] if S > 4 else []),
                       "mean_degree": round(float(np.mean([s["csr"][1].shape[0] / s["V"] for s in w["surfaces"]])), 1),
                       "vertex_weights": any(s["weight"] is not None for s in w["surfaces"]),
                       "subjects": w["n"], "k": w["k"], "H": w["surfaces"][0]["H"],
                       "E": sorted(set(s["E"] for s in w["surfaces"])), "shuffles_per_step_per_gpu": P,
                       "ref_permutations_per_shuffle": signs,
                       "l2": "inputs larger than L2 (Y %.0f MB, statistic maps %.0f MB per step)"
                             % (eng.Y.t.numel() * eng.Y.t.element_size() / 1e6, t32b[0].numel() * 4 / 1e6),
                       "parallelism": "perm-shard x%d" % world,
                       "gather": "tmb_allgather_max (NCCL behind the C ABI), once per %s" % ("step" if args.gather == "step" else "timed region")},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": eng.h2d_bytes // args.steps,
                    "d2h_bytes_per_step": eng.d2h_bytes // args.steps, "ms_per_step": e2e_ms / args.steps,
                    "data_upload_once_bytes": int(eng.Y.n * eng.Y.V * 4), "data_upload_once_ms": data_upload_ms,
                    "graph_build_once_ms": graph_ms},
            "gpu_launches": int(launches),
            # the dominant stage is TFCE + max: a pipeline of five kernels launched back to back (levels, ascent,
            # basins, counts, basin sweep), timed as one unit with CUDA events on their stream
            "roofline": {"bound": "hbm", "kernel": TFCE_STAGE, "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": bytes_tfce * P, "kernel_ms_per_launch": tfce_ms,
                         "kernel_share_of_step": tfce_ms / (dev_ms / args.steps),
                         "fit": None if fit_ms is None else {
                             "kernel": ("tmb_sobelz_cross (one contraction row per shuffle + two-path Sobel epilogue)" if cross else
                                        "tmb_sobelz (glm fit x2 + Sobel epilogue)" if w["kind"] == "mediation" else "glm_dmma_kernel"),
                             "bound": "tensor (fp64 DMMA)", "ms_per_launch": fit_ms,
                             "achieved": fit_flops / (fit_ms / 1e3) / 1e12, "unit": "TFLOP/s",
                             "peak": fp64_peak, "frac": fit_flops / (fit_ms / 1e3) / 1e12 / fp64_peak,
                             "peak_source": "measured in this run: torch.matmul fp64 4096^3 (cuBLAS DGEMM), best of several",
                             "share_of_step": fit_ms / (dev_ms / args.steps)}},
            "cpu_baseline": cpu,
        }
        emit(line)
    if world > 1:
        comm.close()
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------- whole job
def run_job(args):
    """--job N: wall clock of the north-star run -- N permutations (N/2 shuffles, two-sided) of the vertex-wise
    regression + TFCE on fsaverage lh+rh with 300 subjects -- from process start to the FWER-corrected p-map, through the
    drop-in drivers: STEP_2_tfce_randomise_parallel (fan-out) -> vertex_tfce_multiple_regression_randomise (every rank
    its slice) -> CSV rows -> calculate_fweP.  The python_temp_<surface>/ state is synthesised first (untimed: it is
    the reference's step-1 output, STEP_1_vertex_tfce_multiple_regression.py:251-266)."""
    from tfce_mediation_b200.tmanalysis import job
    line = job.run_job(args.job, workload=args.workload, gpus=args.gpus, t_process_start=T_PROCESS_START,
                       check=args.job_check, keep=args.job_dir, checker=reference_rows)
    if line is not None:
        emit(line)


# ---- the job's checker (compiled reference; runs after the clock has stopped) ----------------------------------
_CHK = {}


def _ref_rows_worker(args):
    workdir, first, last, seed = args
    import oracle
    from oracle import build_ref
    try:
        from threadpoolctl import threadpool_limits
        lim = threadpool_limits(limits=1)
    except Exception:
        lim = None
    if _CHK.get("dir") != workdir:
        tmp = os.path.join(workdir, "python_temp_area")
        ld = lambda name: np.load(os.path.join(tmp, name), allow_pickle=True)   # noqa: E731
        mods = build_ref.load()
        adj_lh, adj_rh = list(ld("adjac_lh.npy")), list(ld("adjac_rh.npy"))
        if mods is not None:
            c_lh, c_rh = mods[0].CreateAdjSet(2.0, 0.67, adj_lh), mods[0].CreateAdjSet(2.0, 0.67, adj_rh)
            run_lh, run_rh, tval = c_lh.run, c_rh.run, mods[1].tval_int
            _CHK["keep"] = (c_lh, c_rh)
        else:
            csr = oracle.adjacency_to_csr(adj_lh)
            run_lh = run_rh = lambda img, out: oracle.tfce_run(2.0, 0.67, csr, img, out)   # noqa: E731
            tval = oracle.tval_int
        _CHK.update(dir=workdir, y=ld("merge_y.npy"), x=ld("pred_x.npy"), n=int(ld("num_subjects.npy")),
                    nlh=int(ld("num_vertex_lh.npy")), mlh=ld("bin_mask_lh.npy"), mrh=ld("bin_mask_rh.npy"),
                    dlh=ld("vdensity_lh.npy"), drh=ld("vdensity_rh.npy"), run_lh=run_lh, run_rh=run_rh, tval=tval)
    c = _CHK
    n = c["n"]
    X = np.column_stack([np.ones(n), c["x"]])
    rows = []
    for it in range(first, last + 1):
        np.random.seed(int(it * 1000 + seed))                       # vertex_..._randomise.py:91 with time() := seed
        nx = X[np.random.permutation(list(range(n)))]
        invXX = np.linalg.inv(np.dot(nx.T, nx))
        t = c["tval"](nx, invXX, c["y"], n, X.shape[1], c["y"].shape[1])
        for sign in (1, -1):
            rows.append(oracle.perm_max_vertex(t[1] * sign, c["nlh"], c["mlh"], c["mrh"], c["run_lh"], c["run_rh"],
                                               c["dlh"], c["drh"]))
    del lim
    return rows


def reference_rows(workdir, shuffles, seed):
    import multiprocessing as mp
    cores = min(os.cpu_count() or 1, shuffles)
    bounds = np.linspace(0, shuffles, cores + 1).astype(int)
    jobs = [(workdir, int(bounds[i]) + 1, int(bounds[i + 1]), seed) for i in range(cores) if bounds[i + 1] > bounds[i]]
    with mp.get_context("spawn").Pool(len(jobs)) as pool:
        parts = pool.map(_ref_rows_worker, jobs)
    return [r for part in parts for r in part]



_JSON_OUT = None


def emit(line):
    """The ONE JSON line goes to the process's original stdout; everything else printed to fd 1 during the run (e.g. NCCL's
    version banner under torchrun) was redirected to stderr in main()."""
    out = _JSON_OUT if _JSON_OUT is not None else sys.stdout
    out.write(json.dumps(line, default=float) + "\n")
    out.flush()


def main():
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)            # library chatter on stdout -> stderr; only emit() writes to the real stdout
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="config2")
    ap.add_argument("--block", type=int, default=0, help="shuffles per step per GPU (default: per workload)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--gather", default="job", choices=["job", "step"],
                    help="all-gather of the maxima once per timed region (default) or after every step")
    ap.add_argument("--job", type=int, default=0, help="whole-job mode: this many permutations through the drivers")
    ap.add_argument("--job-check", type=int, default=200, help="--job: permutations re-done by the compiled reference")
    ap.add_argument("--job-dir", default=None, help="--job: keep the working directory here")
    args = ap.parse_args()
    if args.block <= 0:
        args.block = DEFAULT_BLOCK.get(args.workload, 256)
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.job:
        run_job(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
