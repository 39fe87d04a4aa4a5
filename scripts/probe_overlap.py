"""Can the fit (fp64 tensor pipe, 29 % of the issue slots) run UNDER the flat TFCE kernels (integer / load-store pipes) of the
previous block?  Two streams: fit of block i+1 on stream B while TFCE of block i runs on stream A.  Config 2, 1,024 shuffles per
block, device threshold tables (no host round trip) so that only the GPU schedule is measured.
Measured (B200): serial 20.15 ms per block, two streams 20.11 ms; with the fit limited to one CTA per SM (an experimental hook
that requested 116 KB of shared memory, not kept) 21.71 serial / 21.65 two streams -- no gain from concurrency."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench

mode = sys.argv[1] if len(sys.argv) > 1 else "serial"
w = bench.build_workload("config2")
eng, _, _ = bench.make_engine(w, torch.device("cuda", 0))
P, NB = 1024, 12
idx = [bench.perm_rows(w, i * P, P) for i in range(3)]
X = w["X"]
sA = torch.cuda.current_stream()
sB = torch.cuda.Stream()
ld = eng.Y.ld
bufs = [torch.empty((P, 1, ld), dtype=torch.float32, device="cuda") for _ in range(3)]


def fit(i, stream):
    with torch.cuda.stream(stream):
        t32 = eng.tstat_rowperm(X, idx[i % 3])
        bufs[i % 3].copy_(t32)          # keep the probe simple: private buffers per slot
        ev = torch.cuda.Event(); ev.record(stream)
    return ev


def tfce(i):
    mx, st, _ = eng.plan.run(bufs[i % 3].view(P, ld), two_sided=True, exact_pow=False)
    ev = torch.cuda.Event(); ev.record(sA)
    return ev


def run(nblocks):
    two = mode != "serial"
    sf = sB if two else sA
    e_fit = {0: fit(0, sf)}
    e_tf = {}
    for i in range(nblocks):
        if i + 1 < nblocks:
            if two and i - 2 in e_tf:
                sB.wait_event(e_tf[i - 2])      # buffer slot (i+1) % 3 was read by TFCE of block i-2
            e_fit[i + 1] = fit(i + 1, sf)
        if two:
            sA.wait_event(e_fit[i])
        e_tf[i] = tfce(i)
    torch.cuda.synchronize()


run(4)
t0 = time.perf_counter()
run(NB)
dt = time.perf_counter() - t0
print("%s: %.2f ms per block of %d shuffles -> %.0f shuffles/s" % (mode, dt / NB * 1e3, P, NB * P / dt))
