#!/usr/bin/env python
"""bench.py -- permutations/sec of the TFCE_mediation hot path (regression + TFCE + max) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload NAME]

One *step* = one block of `--block` shuffles per GPU through the hot path
(permuted OLS fit + t  ->  TFCE of +t and -t on every surface  ->  scaled max per surface).
One permutation = one shuffled design (the reference's `-n` counts two per shuffle, one per sign;
SURVEY.md section 8d).  Workload `config2` is BASELINE.json configs[1]: vertex-wise regression on
fsaverage lh+rh (2 x 163,842 vertices, icosphere-7 stand-in, cortex masks of 149,955/149,926
vertices), 300 subjects, k=2, H=2, E=0.67, synthetic data.

Prints ONE JSON line (see the task contract): value = whole-job shuffles/s with inputs resident in
HBM, device-timed, max over ranks; `e2e` = the same through PermutationEngine.regression_block with
host index rows in / host maxima out; `roofline` for the dominant stage (the TFCE pipeline);
`cpu_baseline` = the reference's own compiled kernels (oracle/_ref) on one host core.
`--impl reference` times the reference's CPU implementation with all host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "permutations/sec (regression+TFCE+max)"
TFCE_STAGE = "tfce pipeline (pipe_levels + pipe_ascent + pipe_basin + pipe_count + pipe_sweep_max kernels)"
UNIT = "permutations/s"


# ----------------------------------------------------------------------------------------- workloads
def build_workload(name):
    """Synthetic inputs of the named BASELINE.json configuration (seeded)."""
    from tfce_mediation_b200 import synth
    if name == "config2":
        level, n, rounds, keep = 7, 300, 6, (149955, 149926)
    elif name == "config1":
        level, n, rounds, keep = 5, 100, 3, (10242, 10242)
    elif name == "tiny":
        level, n, rounds, keep = 4, 40, 2, (2400, 2300)
    else:
        raise SystemExit("unknown workload %r" % name)
    v, f = synth.icosphere(level)
    csr = synth.faces_to_csr(v.shape[0], f)
    masks = [synth.cap_mask(v, keep[0]), synth.cap_mask(-v, keep[1])]
    ys = [synth.subject_data(n, csr, 1 + h, rounds)[:, masks[h]] for h in range(2)]
    y = np.ascontiguousarray(np.hstack(ys), dtype=np.float32)
    rs = np.random.RandomState(1)
    X = np.column_stack([np.ones(n), rs.standard_normal(n)])
    return dict(name=name, n=n, k=2, H=2.0, E=0.67, csr=csr, masks=masks, y=y, X=X, V_full=v.shape[0],
                seed_base=2000)


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
def _cpu_context(w):
    """The reference's own compiled kernels (oracle/_ref) when present, else the oracle port."""
    import oracle
    from oracle import build_ref
    from tfce_mediation_b200 import synth
    mods = build_ref.load()
    adj = synth.csr_to_lists(w["csr"])
    if mods is not None:
        ref_tfce, ref_stats = mods
        c_lh = ref_tfce.CreateAdjSet(w["H"], w["E"], adj)
        c_rh = ref_tfce.CreateAdjSet(w["H"], w["E"], adj)
        return dict(kind="reference", tval=ref_stats.tval_int, run_lh=c_lh.run, run_rh=c_rh.run, keep=(c_lh, c_rh))
    run = lambda img, out: oracle.tfce_run(w["H"], w["E"], w["csr"], img, out)  # noqa: E731
    return dict(kind="port", tval=oracle.tval_int, run_lh=run, run_rh=run)


def cpu_shuffles(w, ctx, first, count):
    """The reference call sequence of vertex_tfce_multiple_regression_randomise.py:104-117 +
    pyfunc.py:107-119 for `count` shuffles; returns the (+, -) rows."""
    import oracle
    n, k, X, y = w["n"], w["k"], w["X"], w["y"]
    nv_lh = int(w["masks"][0].sum())
    rows = []
    for i in range(count):
        np.random.seed(w["seed_base"] + first + i)
        nx = X[np.random.permutation(list(range(n)))]
        invXX = np.linalg.inv(np.dot(nx.T, nx))
        tvals = ctx["tval"](nx, invXX, y, n, k, y.shape[1])
        for j in range(1, k):
            for sign in (1, -1):
                rows.append(oracle.perm_max_vertex(tvals[j] * sign, nv_lh, w["masks"][0], w["masks"][1],
                                                   ctx["run_lh"], ctx["run_rh"]))
    return rows


def _ref_worker(args):
    name, first, count = args
    try:
        from threadpoolctl import threadpool_limits
        limiter = threadpool_limits(limits=1)
    except Exception:
        limiter = None
    global _REF_W, _REF_CTX
    if "_REF_W" not in globals() or _REF_W["name"] != name:
        _REF_W = build_workload(name)
        _REF_CTX = _cpu_context(_REF_W)
    t0 = time.perf_counter()
    cpu_shuffles(_REF_W, _REF_CTX, first, count)
    dt = time.perf_counter() - t0
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
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        pool.map(_ref_worker, [(args.workload, 10 ** 6 + c, 0) for c in range(cores)])     # build inputs, untimed
        step_ms = []
        for step in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            pool.map(_ref_worker, [(args.workload, (step * cores + c) * per_worker, per_worker) for c in range(cores)])
            dt = time.perf_counter() - t0
            if step >= args.warmup:
                step_ms.append(dt * 1e3)
    shuffles = cores * per_worker * args.steps
    total_s = sum(step_ms) / 1e3
    value = shuffles / total_s
    from oracle import build_ref
    kind = "reference" if build_ref.load() is not None else "port"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": float(np.mean(step_ms)), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "shuffles_per_step": cores * per_worker},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": "%d shuffles per step (%d worker processes x %d), OPENBLAS threads=1 each"
                                   % (cores * per_worker, cores, per_worker)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


# ----------------------------------------------------------------------------------------- B200 arm
def algorithmic_bytes_per_shuffle(surf_graphs, C, signs):
    """C*4*sum(V) (t-map write) + g*C*sum[4V + 4(nnz + V + 1)] (each TFCE call
    reads its statistic map and CSR once, max-only output).  The Y stream term is dropped (P_batch large)."""
    sv = sum(g.num_vertices for g in surf_graphs)
    per_call = sum(4 * g.num_vertices + 4 * (int(g.indices.shape[0]) + g.num_vertices + 1) for g in surf_graphs)
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


def run_b200(args):
    import torch
    import torch.distributed as dist
    from tfce_mediation_b200 import _lib
    from tfce_mediation_b200._graph import induced_subgraph
    from tfce_mediation_b200.engine import PermutationEngine, Surface
    from tfce_mediation_b200.tfce import CreateAdjSet

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

    w = build_workload(args.workload)
    graphs, off, surfs = [], 0, []
    for h in range(2):
        ip, ix = induced_subgraph(w["csr"][0], w["csr"][1], w["masks"][h])
        g = CreateAdjSet(w["H"], w["E"], (ip, ix))
        graphs.append(g)
        surfs.append(Surface(g, off))
        off += g.num_vertices
    pin_y = torch.from_numpy(w["y"]).pin_memory()
    t0 = time.perf_counter()
    eng = PermutationEngine(pin_y, surfs, two_sided=True)
    torch.cuda.synchronize()
    data_upload_ms = (time.perf_counter() - t0) * 1e3
    P = args.block
    C = w["k"] - 1
    X = w["X"]
    total_steps = args.warmup + args.steps
    # every rank owns a contiguous range of the permutation index stream
    idx_all = [perm_rows(w, (rank * total_steps + s) * P, P) for s in range(total_steps)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident arm: design stacks already in HBM, maxima stay on the device ------------
    from tfce_mediation_b200 import engine as E
    stacks = []
    for s in range(total_steps):
        st = E.row_permuted_stack(X, idx_all[s])
        At, ldA = E.pack_At(st["pinv"], 1)
        stacks.append((torch.from_numpy(At).to(dev), ldA, torch.from_numpy(st["G"]).to(dev),
                       torch.from_numpy(st["d"]).to(dev), st["dof"]))
    yy = eng.Y.sumsq(True)
    t32 = torch.empty((P, C, eng.Y.ld), dtype=torch.float32, device=dev)
    out_max = torch.empty((P * C, len(surfs), 2), dtype=torch.float32, device=dev)
    gathered = [torch.empty_like(out_max) for _ in range(world)] if world > 1 else None
    L = _lib.lib()

    t32b = [t32, torch.empty_like(t32)]

    fit_events = []

    def fit(s, buf):
        At_d, ldA, G_d, d_d, dof = stacks[s]
        if fit_events is not None and len(fit_events) < 64:
            fa = torch.cuda.Event(enable_timing=True); fb = torch.cuda.Event(enable_timing=True)
            fa.record()
        else:
            fa = None
        _lib.check(L.tmb_glm_tstat(_lib.ptr(eng.Y.t), eng.Y.dtype_code, eng.Y.n, eng.Y.V, eng.Y.ld, _lib.ptr(At_d), ldA,
                                   _lib.ptr(G_d), _lib.ptr(d_d), P, 1, 1, 0, 1, dof, _lib.ptr(yy), _lib.ptr(buf), None,
                                   eng.Y.ld, 0, _lib.current_stream()))
        if fa is not None:
            fb.record(); fit_events.append((fa, fb))
        return eng.plan.prepare(buf.view(P * C, eng.Y.ld))          # maxima kernel + async copy to the host

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
            eng.plan.finish(tk, t32b[i & 1].view(P * C, eng.Y.ld), two_sided=True, out_max=out_max)
            if tfce_events is not None:
                b.record(); tfce_events.append((a, b))
            if world > 1:
                dist.all_gather(gathered, out_max)   # the per-shuffle maxima, tiny (NCCL over NVLink)
            tk = nxt

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
    idx_warm = np.concatenate(idx_all[:min(args.warmup, 2)], axis=0)
    eng.regression_blocks(X, idx_warm, block=P)
    idx_timed = np.concatenate(idx_all[args.warmup:], axis=0)
    barrier()
    eng.h2d_bytes = eng.d2h_bytes = 0
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    res = eng.regression_blocks(X, idx_timed, block=P)              # numpy [K*P, C, S, 2] on the host
    if world > 1:
        dist.all_gather(gathered, torch.from_numpy(res[-P:]).to(dev).view_as(out_max))
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
        bytes_shuffle, bytes_tfce = algorithmic_bytes_per_shuffle(graphs, C, 2)
        achieved = bytes_tfce * P / (tfce_ms / 1e3) / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "tfce_sweep_traffic.json")
        if os.path.exists(tpath):
            tj = json.load(open(tpath))
            if tj.get("workload") == w["name"] and tj.get("block") == P and tj.get("kernel") == TFCE_STAGE:
                traffic = tj.get("dram_bytes_per_launch")
        cpu = None
        if world == 1 and not args.no_cpu:
            try:
                from threadpoolctl import threadpool_limits
                lim = threadpool_limits(limits=1)
            except Exception:
                lim = None
            ctx = _cpu_context(w)
            sample = {"config2": 3, "config1": 40}.get(w["name"], 10)
            t0 = time.perf_counter()
            rows = cpu_shuffles(w, ctx, args.warmup * P, sample)
            dt = time.perf_counter() - t0
            del lim
            # the same shuffles on the GPU: the FWER rows must agree
            chk = eng.regression_block(X, perm_idx=idx_all[args.warmup][:sample])
            gpu_rows = [max(chk[p, 0, 0, sg], chk[p, 0, 1, sg]) for p in range(sample) for sg in (0, 1)]
            agree = all("%.4f" % a == "%.4f" % b for a, b in zip(rows, gpu_rows))
            cpu = {"value": sample / dt, "unit": UNIT, "cores": 1, "kind": ctx["kind"],
                   "sample": "%d shuffles of the same workload, 1 process, BLAS threads=1; rows identical to GPU: %s"
                             % (sample, agree)}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64 fit / f32 TFCE", "data": "synthetic",
            "config": {"workload": w["name"], "vertices": [g.num_vertices for g in graphs], "subjects": w["n"],
                       "k": w["k"], "H": w["H"], "E": w["E"], "shuffles_per_step_per_gpu": P,
                       "ref_permutations_per_shuffle": 2, "l2": "inputs larger than L2 (Y %.0f MB, t-maps %.0f MB per step)"
                       % (w["y"].nbytes / 1e6, t32.numel() * 4 / 1e6), "parallelism": "perm-shard x%d" % world},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": eng.h2d_bytes // args.steps,
                    "d2h_bytes_per_step": eng.d2h_bytes // args.steps, "ms_per_step": e2e_ms / args.steps,
                    "data_upload_once_bytes": int(w["y"].nbytes), "data_upload_once_ms": data_upload_ms},
            "gpu_launches": int(launches),
            # the dominant stage is TFCE + max: since round 1 v7 a pipeline of five kernels launched back to back
            # (levels, ascent, basins, counts, basin sweep), timed as one unit with CUDA events on their stream
            "roofline": {"bound": "hbm", "kernel": TFCE_STAGE, "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": bytes_tfce * P, "kernel_ms_per_launch": tfce_ms,
                         "kernel_share_of_step": tfce_ms / (dev_ms / args.steps),
                         "fit": None if fit_ms is None else {
                             "kernel": "glm_dmma_kernel", "bound": "tensor (fp64 DMMA)", "ms_per_launch": fit_ms,
                             "achieved": 2.0 * P * w["n"] * eng.Y.V / (fit_ms / 1e3) / 1e12, "unit": "TFLOP/s",
                             "peak": 40.0, "peak_source": "nominal B200 fp64 (no fp64 entry in MEASURED_PEAKS.json)",
                             "share_of_step": fit_ms / (dev_ms / args.steps)}},
            "cpu_baseline": cpu,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


_JSON_OUT = None


def emit(line):
    """The ONE JSON line goes to the process's original stdout; everything else printed to fd 1 during the run (e.g. NCCL's
    version banner under torchrun) was redirected to stderr in main()."""
    out = _JSON_OUT if _JSON_OUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
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
    ap.add_argument("--block", type=int, default=1024, help="shuffles per step per GPU")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
