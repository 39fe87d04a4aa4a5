// TFCE on sm_100a: batched, level-synchronous connected components over CSR adjacency.
//
// Replaces lib/fast_tfce.hpp:11-95 (reference) for many statistic maps at once.
//
// One CTA owns one work item = (statistic row b, surface s) and sweeps BOTH signs of the map in
// lock-step (the +t forest lives on the positive vertices, the -t forest on the negative ones, so
// the two never touch).  The reference's numerical contract is kept exactly:
//   * thresholds T_0 = max, T_{i+1} = fl32(T_i - fl32(max/100)) while T_i >= 0   (fast_tfce.hpp:32-39)
//   * vertex v is active at step i iff x_v > T_i (strict)                         (:41)
//   * a directed entry a of adjacency[u] joins u and a iff a was activated before u (:47-65);
//     here "before" = larger value, ties broken by smaller index
//   * every component c of step i adds fl32( pow((double)|c|,(double)E) * (double)fl32(T_i^H) )
//     to each member by a sequential fp32 add in descending-T order              (:70-84)
// pow(|c|, E) comes from a table computed on the host with the C library, so increments are
// bit-identical to the reference.
//
// Sweep (V1):  per level i
//   P1  newly activated vertices hook into earlier-activated neighbours (lock-free union-find,
//       larger root index under smaller; roots are therefore the smallest index of a component)
//   P2  sizes: every new vertex adds 1 to its final root; every root hooked in this level adds
//       its frozen size to its final root
//   P3  every active vertex adds the increment of its component
#include "common.cuh"

#include <cfloat>

namespace tmb {

static constexpr int kSweepThreads = 512;
static constexpr int kMaxSteps = 128;

__device__ __forceinline__ int ld_cg(const int *p) { return __ldcg(p); }

// union-find with path halving.  All reads go to L2 (ld.cg) so they are coherent with the
// atomicCAS hooks issued by other warps of the CTA.
__device__ __forceinline__ int uf_find(int *parent, int v) {
    int cur = v;
    int p = ld_cg(parent + cur);
    while (p != cur) {
        int gp = ld_cg(parent + p);
        if (gp != p) parent[cur] = gp;
        cur = p;
        p = gp;
    }
    return cur;
}

struct SlotWs {
    int *parent;
    int *size;
    float *acc;
    int *order;  // vertex | sign<<31, grouped by activation level
    int *mlist;  // roots hooked during the current level
};

__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

size_t tfce_slot_bytes(int32_t Vmax) {
    size_t per = align_up(sizeof(int) * (size_t)Vmax, 256);
    return per * 5;
}

__device__ __forceinline__ SlotWs carve(char *base, int32_t Vmax) {
    size_t per = align_up(sizeof(int) * (size_t)Vmax, 256);
    SlotWs w;
    w.parent = reinterpret_cast<int *>(base);
    w.size = reinterpret_cast<int *>(base + per);
    w.acc = reinterpret_cast<float *>(base + 2 * per);
    w.order = reinterpret_cast<int *>(base + 3 * per);
    w.mlist = reinterpret_cast<int *>(base + 4 * per);
    return w;
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

__global__ void __launch_bounds__(kSweepThreads, 2) tfce_sweep_kernel(SweepParams P) {
    __shared__ float sT[2][kMaxSteps];
    __shared__ float sHH[2][kMaxSteps];
    __shared__ int sNs[2];
    __shared__ float sDelta[2];
    __shared__ int sStatus[2];
    __shared__ int sCount[kMaxSteps];   // histogram, then running cursor
    __shared__ int sStart[kMaxSteps + 1];
    __shared__ float sRed[2][kSweepThreads / 32];
    __shared__ int sItem;
    __shared__ int sMcount[2];

    const int tid = threadIdx.x;
    const int nthr = blockDim.x;
    const int lane = tid & 31, wid = tid >> 5;
    SlotWs ws = carve(P.workspace + (size_t)blockIdx.x * P.slot_stride, P.Vmax);
    const int total_items = P.B * P.S;

    for (;;) {
        if (tid == 0) sItem = atomicAdd(P.work_counter, 1);
        __syncthreads();
        const int item = sItem;
        if (item >= total_items) break;
        const int s = P.surf_order[item / P.B];
        const int b = item % P.B;
        const SurfDesc sd = P.surfs[s];
        const int V = sd.V;
        const float *__restrict__ x = P.stat + (size_t)b * P.ld + sd.col_off;
        const int64_t *__restrict__ indptr = sd.indptr;
        const int32_t *__restrict__ indices = sd.indices;

        // ---- maxima of +x and -x (NaN ignored: fmaxf returns the non-NaN operand) -------------
        float mp = -INFINITY, mn = -INFINITY;
        for (int v = tid; v < V; v += nthr) {
            float xv = x[v];
            mp = fmaxf(mp, xv);
            mn = fmaxf(mn, -xv);
        }
        mp = warp_max(mp);
        mn = warp_max(mn);
        if (lane == 0) { sRed[0][wid] = mp; sRed[1][wid] = mn; }
        if (tid < kMaxSteps) sCount[tid] = 0;
        if (tid == 0) { sMcount[0] = 0; sMcount[1] = 0; }
        __syncthreads();
        // ---- threshold tables: one thread per sign, same fp32 ops as fast_tfce.hpp:32-39 -------
        if (tid == 0 || tid == 32) {
            const int sg = tid ? 1 : 0;
            float mx = -INFINITY;
            for (int w = 0; w < nthr / 32; ++w) mx = fmaxf(mx, sRed[sg][w]);
            int ns = 0, st = 0;
            float d = 0.f;
            if ((sg == 0 || P.two_sided) && mx >= 0.f) {
                d = __fdiv_rn(mx, 100.0f);
                if (d == 0.f) {
                    st = 1; // TMB_MAP_MAX_IS_ZERO: the reference would spin forever
                } else {
                    float T = mx;
                    while (T >= 0.f) {
                        if (ns == kMaxSteps) { st = 2; ns = 0; break; }
                        sT[sg][ns] = T;
                        sHH[sg][ns] = (sd.H == 2.0f) ? __fmul_rn(T, T)
                                                      : (float)pow((double)T, (double)sd.H);
                        ++ns;
                        T = __fsub_rn(T, d);
                    }
                }
            }
            sNs[sg] = ns;
            sDelta[sg] = d;
            sStatus[sg] = st;
        }
        __syncthreads();
        const int ns0 = sNs[0], ns1 = sNs[1];
        const int nlev = max(ns0, ns1); // levels 1 .. nlev-1 carry activations

        // ---- activation level per vertex + histogram ------------------------------------------
        // order[] temporarily holds the level; parent/size/acc are initialised for every vertex
        for (int v = tid; v < V; v += nthr) {
            float xv = x[v];
            int lev = 0;
            if (xv > 0.f || xv < 0.f) {
                const int sg = xv < 0.f;
                const int ns = sg ? ns1 : ns0;
                const float ax = fabsf(xv);
                const float *T = sT[sg];
                int lo = 1, hi = ns;
                while (lo < hi) {
                    int mid = (lo + hi) >> 1;
                    if (ax > T[mid]) hi = mid; else lo = mid + 1;
                }
                if (lo < ns) { lev = lo; atomicAdd(&sCount[lev], 1); }
            }
            ws.order[v] = lev; // staged; rewritten by the scatter below via mlist as scratch
            ws.parent[v] = v;
            ws.size[v] = 0;
        }
        __syncthreads();
        if (tid == 0) {
            int run = 0;
            sStart[0] = 0; sStart[1] = 0;
            for (int l = 1; l < kMaxSteps; ++l) { int c = sCount[l]; sStart[l] = run; sCount[l] = run; run += c; }
            sStart[kMaxSteps] = run;
        }
        __syncthreads();
        // scatter into mlist (scratch), then swap roles: mlist <-> order by pointer
        for (int v = tid; v < V; v += nthr) {
            int lev = ws.order[v];
            if (lev > 0) {
                int pos = atomicAdd(&sCount[lev], 1);
                ws.mlist[pos] = v | ((x[v] < 0.f) ? 0x80000000 : 0);
            }
        }
        __syncthreads();
        { int *t = ws.order; ws.order = ws.mlist; ws.mlist = t; }
        const int total_active = sStart[kMaxSteps];

        // acc init (CreateAdjSet.run accumulates into enhn)
        for (int idx = tid; idx < total_active; idx += nthr) {
            int e = ws.order[idx];
            int u = e & 0x7fffffff;
            float a0 = 0.f;
            if (P.accumulate) {
                float *dst = (e < 0) ? P.tfce_neg : P.tfce_pos;
                if (dst) a0 = dst[(size_t)b * P.ld + sd.col_off + u];
            }
            ws.acc[u] = a0;
        }
        __syncthreads();

        const int last_level = (P.stop_level >= 0) ? min(P.stop_level, nlev - 1) : nlev - 1;
        for (int lev = 1; lev <= last_level; ++lev) {
            const int beg = sStart[lev];
            const int end = sStart[lev + 1];
            // ---- P1: hook new vertices into earlier-activated neighbours -----------------------
            for (int idx = beg + tid; idx < end; idx += nthr) {
                const int e = ws.order[idx];
                const int u = e & 0x7fffffff;
                const bool neg = e < 0;
                const float xu = x[u];
                const int64_t r0 = indptr[u], r1 = indptr[u + 1];
                for (int64_t k = r0; k < r1; ++k) {
                    const int a = indices[k];
                    const float xa = x[a];
                    const bool earlier = neg ? (xa < xu || (xa == xu && a < u))
                                             : (xa > xu || (xa == xu && a < u));
                    if (!earlier) continue;
                    int ru = uf_find(ws.parent, u);
                    int ra = uf_find(ws.parent, a);
                    while (ru != ra) {
                        if (ru < ra) { int t = ru; ru = ra; ra = t; }
                        const int old = atomicCAS(ws.parent + ru, ru, ra);
                        if (old == ru) {
                            if (ld_cg(ws.size + ru) > 0) { // an older component lost its root
                                int m = atomicAdd(&sMcount[lev & 1], 1);
                                ws.mlist[m] = ru;
                            }
                            break;
                        }
                        ru = uf_find(ws.parent, old);
                        ra = uf_find(ws.parent, ra);
                    }
                }
            }
            __syncthreads();
            // ---- P2: sizes ---------------------------------------------------------------------
            const int mcount = sMcount[lev & 1];
            if (tid == 0) sMcount[(lev + 1) & 1] = 0; // next level's list (first used after two more barriers)
            for (int idx = beg + tid; idx < end; idx += nthr) {
                const int u = ws.order[idx] & 0x7fffffff;
                const int r = uf_find(ws.parent, u);
                atomicAdd(ws.size + r, 1);
            }
            for (int m = tid; m < mcount; m += nthr) {
                const int h = ws.mlist[m];
                const int r = uf_find(ws.parent, h);
                atomicAdd(ws.size + r, ld_cg(ws.size + h));
            }
            __syncthreads();
            if (P.stop_level >= 0) continue; // components only: no accumulation
            // ---- P3: every active vertex receives its component's increment --------------------
            const float hh0 = (lev < ns0) ? sHH[0][lev] : 0.f;
            const float hh1 = (lev < ns1) ? sHH[1][lev] : 0.f;
            for (int idx = tid; idx < end; idx += nthr) {
                const int e = ws.order[idx];
                const int u = e & 0x7fffffff;
                const bool neg = e < 0;
                if (lev >= (neg ? ns1 : ns0)) continue;
                const int r = uf_find(ws.parent, u);
                const int n = ld_cg(ws.size + r);
                const float inc = __double2float_rn(__dmul_rn(sd.powE[n], (double)(neg ? hh1 : hh0)));
                ws.acc[u] = __fadd_rn(ws.acc[u], inc);
            }
            __syncthreads();
        }

        // ---- outputs -----------------------------------------------------------------------------
        if (P.stop_level >= 0) {
            __syncthreads();
            for (int v = tid; v < V; v += nthr) { P.labels[v] = -1; P.extents[v] = 0; }
            __syncthreads();
            const int endl = (last_level >= 0) ? sStart[last_level + 1] : 0;
            for (int idx = tid; idx < endl; idx += nthr) {
                const int e = ws.order[idx];
                if (e < 0) continue; // inspection is one-sided (+ map)
                const int u = e & 0x7fffffff;
                const int r = uf_find(ws.parent, u);
                P.labels[u] = r;
                P.extents[u] = ld_cg(ws.size + r);
            }
            if (tid == 0 && P.threshold_out)
                *P.threshold_out = (P.stop_level < ns0) ? sT[0][P.stop_level] : NAN;
        } else {
            float m0 = 0.f, m1 = 0.f;
            const float d0 = sDelta[0], d1 = sDelta[1];
            const float *__restrict__ w = sd.weight;
            for (int idx = tid; idx < total_active; idx += nthr) {
                const int e = ws.order[idx];
                const int u = e & 0x7fffffff;
                const bool neg = e < 0;
                float val = __fmul_rn(ws.acc[u], neg ? d1 : d0);
                if (w) val = __fmul_rn(val, w[u]);
                if (neg) m1 = fmaxf(m1, val); else m0 = fmaxf(m0, val);
            }
            m0 = warp_max(m0);
            m1 = warp_max(m1);
            __syncthreads();
            if (lane == 0) { sRed[0][wid] = m0; sRed[1][wid] = m1; }
            __syncthreads();
            if (tid == 0) {
                float a = 0.f, c = 0.f;
                for (int w2 = 0; w2 < nthr / 32; ++w2) { a = fmaxf(a, sRed[0][w2]); c = fmaxf(c, sRed[1][w2]); }
                const size_t o = ((size_t)b * P.S + s) * 2;
                if (P.max_out) { P.max_out[o] = a; P.max_out[o + 1] = c; }
                if (P.status) { P.status[o] = sStatus[0]; P.status[o + 1] = sStatus[1]; }
            }
            // full maps (unscaled TFCE), inactive vertices are 0 (or untouched when accumulating)
            if (P.tfce_pos || P.tfce_neg) {
                if (!P.accumulate) {
                    for (int v = tid; v < V; v += nthr) {
                        const size_t o = (size_t)b * P.ld + sd.col_off + v;
                        if (P.tfce_pos) P.tfce_pos[o] = 0.f;
                        if (P.tfce_neg) P.tfce_neg[o] = 0.f;
                    }
                    __syncthreads();
                }
                for (int idx = tid; idx < total_active; idx += nthr) {
                    const int e = ws.order[idx];
                    const int u = e & 0x7fffffff;
                    float *dst = (e < 0) ? P.tfce_neg : P.tfce_pos;
                    if (dst) dst[(size_t)b * P.ld + sd.col_off + u] = ws.acc[u];
                }
            }
        }
        __syncthreads();
    }
}

int launch_tfce_sweep(const SweepParams &p, int num_slots, cudaStream_t stream) {
    const int items = p.B * p.S;
    if (items <= 0) return 0;
    int grid = items < num_slots ? items : num_slots;
    TMB_CUDA(cudaMemsetAsync(p.work_counter, 0, sizeof(int), stream));
    tfce_sweep_kernel<<<grid, kSweepThreads, 0, stream>>>(p);
    count_launch();
    TMB_CUDA(cudaGetLastError());
    return 0;
}

} // namespace tmb
