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
//     "before" = larger value, ties broken by smaller index.  For symmetric adjacency (checked on
//     the host at graph creation) only activation LEVELS matter, which is what the fast path uses.
//   * every component c of step i adds fl32( pow((double)|c|,(double)E) * (double)fl32(T_i^H) )
//     to each member by a sequential fp32 add in descending-T order              (:70-84)
// pow(|c|, E) comes from a table computed on the host with the C library, so increments are
// bit-identical to the reference.
//
// Algorithm (V2).  Vertices are bucketed by activation level (counting sort).  Levels are swept in
// order; per level only the NEWLY activated vertices do work:
//   P1   hook new vertices into earlier-activated neighbours (lock-free union-find; roots are
//        totally ordered by (activation level, index) and the later root goes under the earlier)
//   P2a  sizes + component-tree bookkeeping: every component whose membership changed in this
//        level gets a new tree node (size, level); warp-aggregated atomics
//   P2bc links the previous nodes of merged/grown components to the new node and records each new
//        vertex's leaf node.  It touches no union-find state, so it runs in the same barrier
//        interval as P1 of the NEXT level: two __syncthreads per level.
// The per-vertex TFCE value is then a walk from the vertex's leaf node to the root of the component
// tree, adding one increment per level in the reference's order.  All vertices that share a leaf
// share the value, so the walk is done once per node.  Nothing is visited per (vertex, level).
#include "common.cuh"

#include <cfloat>
#include <cstdio>
#include <cstdlib>

namespace tmb {

static constexpr int kSweepMaxThreads = 1024;
static constexpr int kMaxSteps = 128;
static constexpr int kNone = 0x00FFFFFF; // "no parent" in the 24-bit parent field of a node
static constexpr int kPending = -2;

__device__ __forceinline__ int ld_cg(const int *p) { return __ldcg(p); }

// software prefetch into L1 (generic address of a global object): the sweep is bound by the latency of
// dependent loads, and most addresses are known one barrier interval before they are needed
__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.L1 [%0];" ::"l"(p)); }
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.L2 [%0];" ::"l"(p)); }

// Union-find pointer load.  CACHED uses the default L1-allocating load.  That is safe here:
//  * only this CTA ever touches its slot, and __syncthreads() orders the block's earlier global
//    writes AND atomics before later loads of every thread of the block, so the phases that need the
//    exact forest (P2a, outputs) see it;
//  * inside P1, a stale pointer is still a pointer to an ancestor (pointers only ever move towards
//    smaller indices within the same tree), and every hook is decided by an atomicCAS at L2.
template <bool CACHED>
__device__ __forceinline__ int ld_uf(const int *p) { return CACHED ? *p : __ldcg(p); }

// union-find with path halving
template <bool CACHED>
__device__ __forceinline__ int uf_find(int *parent, int v) {
    int cur = v;
    int p = ld_uf<CACHED>(parent + cur);
    while (p != cur) {
        int gp = ld_uf<CACHED>(parent + p);
        if (gp != p) parent[cur] = gp;
        cur = p;
        p = gp;
    }
    return cur;
}

// find continuing from an already loaded parent pointer p == parent[v]
template <bool CACHED>
__device__ __forceinline__ int uf_find_from(int *parent, int v, int p) {
    int cur = v;
    while (p != cur) {
        int gp = ld_uf<CACHED>(parent + p);
        if (gp != p) parent[cur] = gp;
        cur = p;
        p = gp;
    }
    return cur;
}

struct SlotWs {
    int *parent;
    int *size;      // component size at its root; reused as float node values after the sweep
    int *curnode;   // root -> node id of its component's current tree node
    int *leaf;      // vertex -> leaf node (holds the root between P2a and P2bc)
    int *order;     // vertices grouped by activation level
    int2 *nodes;    // {size, (sign<<31)|(level<<24)|parent}
    int2 *mlist[2]; // {hooked older root, its final root}   (double buffered across levels)
    int2 *clist;    // {changed root, its previous node}
    unsigned char *lev8; // activation level (bits 0..6) | sign (bit 7); 0 = never active
};

__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

size_t tfce_slot_bytes(int32_t Vmax) {
    const size_t per4 = align_up(sizeof(int) * (size_t)Vmax, 256);
    const size_t per8 = align_up(sizeof(int2) * (size_t)Vmax, 256);
    return per4 * 5 + per8 * 4 + align_up((size_t)Vmax, 256);
}

__device__ __forceinline__ SlotWs carve(char *base, int32_t Vmax) {
    const size_t per4 = align_up(sizeof(int) * (size_t)Vmax, 256);
    const size_t per8 = align_up(sizeof(int2) * (size_t)Vmax, 256);
    SlotWs w;
    w.parent = reinterpret_cast<int *>(base);
    w.size = reinterpret_cast<int *>(base + per4);
    w.curnode = reinterpret_cast<int *>(base + 2 * per4);
    w.leaf = reinterpret_cast<int *>(base + 3 * per4);
    w.order = reinterpret_cast<int *>(base + 4 * per4);
    char *b8 = base + 5 * per4;
    w.nodes = reinterpret_cast<int2 *>(b8);
    w.mlist[0] = reinterpret_cast<int2 *>(b8 + per8);
    w.mlist[1] = reinterpret_cast<int2 *>(b8 + 2 * per8);
    w.clist = reinterpret_cast<int2 *>(b8 + 3 * per8);
    w.lev8 = reinterpret_cast<unsigned char *>(b8 + 4 * per8);
    return w;
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Sum of the increments along the component-tree path that starts at node j, in the reference's
// order (descending threshold), starting from acc.
__device__ __forceinline__ float walk_path(const int2 *__restrict__ nodes, const double *__restrict__ powE,
                                           const float (*sHH)[kMaxSteps], const int *sNs, int j, float acc) {
    int2 rec = nodes[j];
    for (;;) {
        const int pk = rec.y;
        const int par = pk & kNone;
        const int lv = (pk >> 24) & 0x7f;
        const int sg = (pk >> 31) & 1;
        int2 prec = make_int2(0, 0);
        int endl;
        if (par == kNone) {
            endl = sNs[sg];
        } else {
            prec = nodes[par];
            endl = (prec.y >> 24) & 0x7f;
        }
        const double pw = powE[rec.x];
        const float *hh = sHH[sg];
        for (int l = lv; l < endl; ++l) acc = __fadd_rn(acc, __double2float_rn(__dmul_rn(pw, (double)hh[l])));
        if (par == kNone) break;
        rec = prec;
    }
    return acc;
}

// LEV_SMEM: the per-vertex activation level/sign bytes live in dynamic shared memory (V bytes) instead
// of the slot workspace: the neighbour filter of P1, the hottest random read, then never leaves the SM.
template <bool LEV_SMEM, bool CACHED>
__global__ void __launch_bounds__(kSweepMaxThreads, 1) tfce_sweep_kernel(SweepParams P) {
    extern __shared__ __align__(16) unsigned char sLevDyn[];
    __shared__ float sT[2][kMaxSteps];
    __shared__ float sHH[2][kMaxSteps];
    __shared__ int sNs[2];
    __shared__ float sDelta[2];
    __shared__ float sScale[2];   // factor of the scaled maximum (the threshold step unless the caller supplies another)
    __shared__ int sStatus[2];
    __shared__ int sCount[kMaxSteps];   // histogram, then running cursor
    __shared__ int sStart[kMaxSteps + 1];
    __shared__ float sRed[2][kSweepMaxThreads / 32];
    __shared__ int sItem;
    __shared__ int sMcount[2];
    __shared__ int sCcount[2];

    const int tid = threadIdx.x;
    const int nthr = blockDim.x;
    const int lane = tid & 31, wid = tid >> 5;
    const SlotWs ws = carve(P.workspace + (size_t)blockIdx.x * P.slot_stride, P.Vmax);
    unsigned char *const lev8 = LEV_SMEM ? sLevDyn : ws.lev8;
    const int total_items = P.B * P.S;

    for (;;) {
        if (tid == 0) sItem = atomicAdd(P.work_counter, 1);
        __syncthreads();
        const int item = sItem;
        if (item >= total_items) break;
        const int s = P.surf_order[item / P.B];
        const int b = item % P.B;
        const SurfDesc sd = P.surfs[s];
        long long tk = P.timing ? clock64() : 0;
#define TMB_TICK(i)                                                                  \
        if (P.timing && tid == 0) {                                                  \
            const long long now = clock64();                                         \
            atomicAdd(P.timing + (i), (unsigned long long)(now - tk));               \
            tk = now;                                                                \
        }
        const int V = sd.V;
        const float *__restrict__ x = P.stat + (size_t)b * P.ld + sd.col_off;
        const int64_t *__restrict__ indptr = sd.indptr;
        const int32_t *__restrict__ indices = sd.indices;
        const int32_t *__restrict__ vmap = (P.flags & 4) ? nullptr : sd.vmap; // internal index -> caller's index, or null
        const bool directed = sd.directed != 0;

        if (tid < kMaxSteps) sCount[tid] = 0;
        if (tid == 0) { sMcount[0] = sMcount[1] = 0; sCcount[0] = sCcount[1] = 0; }
        if (P.tab_ns) {
            // ---- threshold tables computed on the host with the C library's powf (bit-identical to the
            //      reference's std::pow(float,float), which is not correctly rounded) ------------------
            const size_t e0 = ((size_t)b * P.S + s) * 2;
            for (int i = tid; i < 2 * kMaxSteps; i += nthr) {
                const int sg = i / kMaxSteps, l = i % kMaxSteps;
                sT[sg][l] = P.tab_T[(e0 + sg) * kMaxSteps + l];
                sHH[sg][l] = P.tab_HH[(e0 + sg) * kMaxSteps + l];
            }
            if (tid < 2) {
                const bool on = (tid == 0) || P.two_sided;
                sNs[tid] = on ? P.tab_ns[e0 + tid] : 0;
                sDelta[tid] = P.tab_delta[e0 + tid];
                sScale[tid] = P.tab_scale ? P.tab_scale[e0 + tid] : P.tab_delta[e0 + tid];
                sStatus[tid] = on ? P.tab_status[e0 + tid] : 0;
            }
            __syncthreads();
        } else {
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
            __syncthreads();
            // ---- threshold tables on the device: same fp32 ops as fast_tfce.hpp:32-39, with the height
            //      term correctly rounded (T*T for H == 2) ----------------------------------------------
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
                sScale[sg] = d;
                sStatus[sg] = st;
            }
            __syncthreads();
        }
        const int ns0 = sNs[0], ns1 = sNs[1];
        const int nlev = max(ns0, ns1); // levels 1 .. nlev-1 carry activations
        TMB_TICK(0)

        // ---- activation level per vertex + histogram ------------------------------------------
        for (int v = tid; v < V; v += nthr) {
            const float xv = x[vmap ? vmap[v] : v];
            int code = 0;
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
                if (lo < ns) { code = lo | (sg << 7); atomicAdd(&sCount[lo], 1); }
            }
            lev8[v] = (unsigned char)code;
            ws.parent[v] = v;
            ws.size[v] = 0;
            ws.curnode[v] = -1;
        }
        __syncthreads();
        if (tid == 0) {
            int run = 0;
            sStart[0] = 0;
            for (int l = 1; l < kMaxSteps; ++l) { int c = sCount[l]; sStart[l] = run; sCount[l] = run; run += c; }
            sStart[kMaxSteps] = run;
        }
        __syncthreads();
        for (int v = tid; v < V; v += nthr) {
            const int lev = lev8[v] & 0x7f;
            if (lev > 0) ws.order[atomicAdd(&sCount[lev], 1)] = v;
        }
        __syncthreads();
        const int total_active = sStart[kMaxSteps];
        TMB_TICK(1)

        const int last_level = (P.stop_level >= 0) ? min(P.stop_level, nlev - 1) : nlev - 1;
        // pending P2bc work of the previous non-empty level (uniform across the CTA)
        bool pend = false;
        int pend_beg = 0, pend_end = 0, pend_buf = 0, pend_nodebase = 0, pend_lev = 0;
        int node_base = 0;
        int buf = 0;

        for (int lev = 1; lev <= last_level + 1; ++lev) {
            const bool have_level = lev <= last_level;
            const int beg = have_level ? sStart[lev] : 0;
            const int end = have_level ? sStart[lev + 1] : 0;
            if (have_level && end == beg) continue; // nothing activates: partition unchanged
            // ================= interval X: P2bc(previous level)  +  P1(this level) ===============
            if (pend) {
                const int ccount = sCcount[pend_buf];
                const int mcount = sMcount[pend_buf];
                for (int c = tid; c < ccount; c += nthr) {
                    const int2 e = ws.clist[c];
                    const int r = e.x, old = e.y;
                    const int j = pend_nodebase + c;
                    const int sg = lev8[r] >> 7;
                    ws.nodes[j] = make_int2(ld_cg(ws.size + r), (sg << 31) | (pend_lev << 24) | kNone);
                    if (old >= 0) ws.nodes[old].y = (ws.nodes[old].y & 0xFF000000) | j;
                }
                const int2 *ml = ws.mlist[pend_buf];
                for (int m = tid; m < mcount; m += nthr) {
                    const int2 e = ml[m];
                    const int o = ld_cg(ws.curnode + e.x);
                    const int jn = ld_cg(ws.curnode + e.y);
                    ws.nodes[o].y = (ws.nodes[o].y & 0xFF000000) | jn;
                }
                for (int idx = pend_beg + tid; idx < pend_end; idx += nthr) {
                    const int u = ws.order[idx];
                    ws.leaf[u] = ld_cg(ws.curnode + ws.leaf[u]);
                }
            }
            if (!have_level) break;
            int2 *mcur = ws.mlist[buf];
            for (int idx = beg + tid; idx < end; idx += nthr) {
                const int u = ws.order[idx];
                const int cu = lev8[u];
                const float xu = directed ? x[u] : 0.f;
                const int64_t r0 = indptr[u], r1 = indptr[u + 1];
                int ru = u;
                for (int64_t k0 = r0; k0 < r1; k0 += 8) {
                    // neighbour ids, their level bytes and their first parent pointers are fetched as
                    // three batches of independent loads (memory-level parallelism per thread)
                    const int cnt = (int)min((int64_t)8, r1 - k0);
                    int nb[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) nb[j] = (j < cnt) ? indices[k0 + j] : -1;
                    unsigned early = 0;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        if (nb[j] < 0) continue;
                        const int a = nb[j];
                        const int ca = lev8[a];
                        if (ca == 0 || ((ca ^ cu) & 0x80)) continue; // inactive or other sign
                        bool e;
                        if (ca != cu) {
                            e = ca < cu; // same sign: smaller level == activated at a higher threshold
                        } else if (directed) {
                            const float xa = x[a];
                            e = (cu & 0x80) ? (xa < xu || (xa == xu && a < u)) : (xa > xu || (xa == xu && a < u));
                        } else {
                            e = a < u;   // symmetric graph: the pair is seen from both ends, join once
                        }
                        if (e) early |= 1u << j;
                    }
                    int pa[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) pa[j] = ((early >> j) & 1u) ? ld_uf<CACHED>(ws.parent + nb[j]) : -1;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        if (!((early >> j) & 1u)) continue;
                        const int a = nb[j];
                        if (pa[j] == ru || a == ru) continue; // a already hangs under u's (possibly former) root
                        int ra = uf_find_from<CACHED>(ws.parent, a, pa[j]);
                        ru = uf_find<CACHED>(ws.parent, ru);
                        while (ru != ra) {
                            // total order on roots: (activation level, index).  The root that activated at
                            // the later threshold goes under the earlier one, so a growing component keeps
                            // its root and new vertices attach directly to it (shallow trees).
                            const int kru = ((lev8[ru] & 0x7f) << 24) | ru, kra = ((lev8[ra] & 0x7f) << 24) | ra;
                            const int hi = kru > kra ? ru : ra, lo = kru > kra ? ra : ru;
                            const int old = atomicCAS(ws.parent + hi, hi, lo);
                            if (old == hi) {
                                if ((lev8[hi] & 0x7f) != lev) { // an older component lost its root
                                    const int m = atomicAdd(&sMcount[buf], 1);
                                    mcur[m] = make_int2(hi, -1);
                                }
                                ru = lo;
                                break;
                            }
                            // hi was hooked by someone else meanwhile: climb from its real parent
                            const int nh = uf_find<CACHED>(ws.parent, old);
                            if (hi == ru) { ru = nh; ra = uf_find<CACHED>(ws.parent, ra); }
                            else          { ra = nh; ru = uf_find<CACHED>(ws.parent, ru); }
                        }
                    }
                }
            }
            __syncthreads();
            TMB_TICK(2)
            // ================= interval Y: P2a(this level) ======================================
            {
                const int mcount = sMcount[buf];
                // the other buffer was consumed by P2bc in interval X: recycle it for the next level
                if (tid == 0) { sMcount[buf ^ 1] = 0; sCcount[buf ^ 1] = 0; }
                const int span = end - beg;
                const int iters = (span + nthr - 1) / nthr;
                for (int it = 0; it < iters; ++it) { // warp-uniform trip count (match_any below)
                    const int idx = beg + it * nthr + tid;
                    const bool valid = idx < end;
                    int r = -1 - lane; // distinct dummy keys for idle lanes
                    int u = 0;
                    if (valid) {
                        u = ws.order[idx];
                        r = uf_find<CACHED>(ws.parent, u);
                        ws.leaf[u] = r;
                    }
                    const unsigned peers = __match_any_sync(0xffffffffu, r);
                    if (valid && lane == (__ffs(peers) - 1)) {
                        atomicAdd(ws.size + r, __popc(peers));
                        int old = ld_cg(ws.curnode + r);
                        for (;;) {
                            if (old >= node_base || old == kPending) break; // claimed in this level
                            const int prev = atomicCAS(ws.curnode + r, old, kPending);
                            if (prev == old) {
                                const int pos = atomicAdd(&sCcount[buf], 1);
                                ws.clist[pos] = make_int2(r, old);
                                atomicExch(ws.curnode + r, node_base + pos);
                                break;
                            }
                            old = prev;
                        }
                    }
                }
                for (int m = tid; m < mcount; m += nthr) {
                    const int h = mcur[m].x;
                    const int r = uf_find<CACHED>(ws.parent, h);
                    mcur[m].y = r;
                    atomicAdd(ws.size + r, ld_cg(ws.size + h));
                }
            }
            __syncthreads();
            TMB_TICK(3)
            pend = true; pend_beg = beg; pend_end = end; pend_buf = buf; pend_nodebase = node_base; pend_lev = lev;
            node_base += sCcount[buf];
            buf ^= 1;
        }
        __syncthreads();
        TMB_TICK(4)
        const int num_nodes = node_base;

        // ---- outputs -----------------------------------------------------------------------------
        if (P.stop_level >= 0) {
            for (int v = tid; v < V; v += nthr) {
                const int o = vmap ? vmap[v] : v;
                int lab = -1, ext = 0;
                const int code = lev8[v];
                if (code != 0 && !(code & 0x80) && (code & 0x7f) <= last_level) { // inspection is one-sided (+ map)
                    const int r = uf_find<CACHED>(ws.parent, v);
                    lab = r;
                    ext = ld_cg(ws.size + r);
                }
                P.labels[o] = lab;   // internal root id; canonicalised (smallest caller index) on the host
                P.extents[o] = ext;
            }
            if (tid == 0 && P.threshold_out)
                *P.threshold_out = (P.stop_level < ns0) ? sT[0][P.stop_level] : NAN;
        } else {
            float *nodeval = reinterpret_cast<float *>(ws.size);
            const bool per_vertex_walk = P.accumulate != 0;
            if (!per_vertex_walk) {
                for (int j = tid; j < num_nodes; j += nthr)
                    nodeval[j] = walk_path(ws.nodes, sd.powE, sHH, sNs, j, 0.f);
                __syncthreads();
                TMB_TICK(5)
                if (P.timing && tid == 0) { atomicAdd(P.timing + 7, (unsigned long long)num_nodes); }
            }
            float m0 = 0.f, m1 = 0.f;
            const float d0 = sScale[0], d1 = sScale[1];
            const bool want_maps = (P.tfce_pos != nullptr) || (P.tfce_neg != nullptr);
            if (want_maps && !P.accumulate) {
                for (int v = tid; v < V; v += nthr) {
                    const size_t o = (size_t)b * P.ld + sd.col_off + v;
                    if (P.tfce_pos) P.tfce_pos[o] = 0.f;
                    if (P.tfce_neg) P.tfce_neg[o] = 0.f;
                }
                __syncthreads();
            }
            for (int idx = tid; idx < total_active; idx += nthr) {
                const int u = ws.order[idx];
                const bool neg = (lev8[u] & 0x80) != 0;
                const size_t o = (size_t)b * P.ld + sd.col_off + (vmap ? vmap[u] : u);
                float *dst = neg ? P.tfce_neg : P.tfce_pos;
                float val;
                if (per_vertex_walk) {
                    // CreateAdjSet.run semantics: enhn[v] += inc, one add per level, starting from enhn[v]
                    val = walk_path(ws.nodes, sd.powE, sHH, sNs, ws.leaf[u], dst ? dst[o] : 0.f);
                } else {
                    val = nodeval[ws.leaf[u]];
                }
                if (dst) dst[o] = val;
                const float sc = scaled_vertex_value(val, neg ? d1 : d0, sd, u);
                if (neg) m1 = fmaxf(m1, sc); else m0 = fmaxf(m0, sc);
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
        }
        __syncthreads();
        TMB_TICK(6)
#undef TMB_TICK
    }
}

// per (row, surface): max of +x and of -x, NaN ignored -- the inputs of the host threshold tables
__global__ void tfce_maxima_kernel(const SurfDesc *__restrict__ surfs, int S, const float *__restrict__ stat,
                                   int64_t ld, float *__restrict__ max_out) {
    __shared__ float sRed[2][32];
    const int b = blockIdx.x / S, s = blockIdx.x % S;
    const SurfDesc sd = surfs[s];
    const float *__restrict__ x = stat + (size_t)b * ld + sd.col_off;
    float mp = -INFINITY, mn = -INFINITY;
    const int V = sd.V, step = blockDim.x;
    int v = threadIdx.x;
    for (; v + 7 * step < V; v += 8 * step) { // eight independent loads in flight per thread: the map streams from HBM
        float xq[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) xq[q] = x[v + q * step];
#pragma unroll
        for (int q = 0; q < 8; ++q) { mp = fmaxf(mp, xq[q]); mn = fmaxf(mn, -xq[q]); }
    }
    for (; v < V; v += step) {
        const float xv = x[v];
        mp = fmaxf(mp, xv);
        mn = fmaxf(mn, -xv);
    }
    mp = warp_max(mp);
    mn = warp_max(mn);
    if ((threadIdx.x & 31) == 0) { sRed[0][threadIdx.x >> 5] = mp; sRed[1][threadIdx.x >> 5] = mn; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)blockDim.x / 32; ++w) { mp = fmaxf(mp, sRed[0][w]); mn = fmaxf(mn, sRed[1][w]); }
        max_out[(size_t)blockIdx.x * 2] = mp;
        max_out[(size_t)blockIdx.x * 2 + 1] = mn;
    }
}

int launch_tfce_maxima(const SurfDesc *surfs, int S, const float *stat, int64_t ld, int B, float *max_out,
                       cudaStream_t stream) {
    if (B * S <= 0) return 0;
    tfce_maxima_kernel<<<B * S, 512, 0, stream>>>(surfs, S, stat, ld, max_out);
    count_launch();
    TMB_CUDA(cudaGetLastError());
    return 0;
}

// =================================================================================================
// Basin sweep (symmetric adjacency) -- the production kernel.
//
// Observation: a vertex is connected, inside the superlevel set of its own activation level, to any
// neighbour that activated at a strictly earlier level -- and through it, step by step, to a vertex
// with no earlier neighbour (a "peak").  So every vertex can inherit a BASIN id from one earlier
// neighbour, and connectivity only has to be tracked between basins, of which there are 10-100x
// fewer than vertices: the lock-free union-find of the level sweep lives on the basin ids, in
// SHARED memory when it fits, and a vertex only issues a union when an earlier neighbour carries a
// different basin id.  The assignment itself needs no pointer chasing because it is done in level
// order: when level l+1 is assigned, every basin id of levels <= l is final.
//
//   P   one pass in vertex order: ascent pointer up[v] = earliest-level neighbour (ELL row + level
//       bytes in shared memory), peaks get compact basin ids
//   per level l (two __syncthreads):
//     I1  unions between the basins of level-l vertices and their earlier neighbours
//         (+ component-tree links of level l-1, which touch no union-find state)
//     I2  sizes and changed components of level l  (+ basin assignment of the next level)
//   then the per-node walk and the outputs exactly as in V2: values are bit-identical.
//
// Vertices are bucketed by level with a STABLE counting sort (per-warp chunks, warp-local cursors):
// bucket entries ascend in vertex id, so with the locality relabelling of the graph a warp's 32
// vertices and all their neighbours fall into a handful of cache lines (few wavefronts per request).
struct RowIter {
    const int32_t *ell;
    const int32_t *indices;
    int64_t r0, r1;
    int nchunks;
    __device__ __forceinline__ RowIter(const SurfDesc &sd, int u) {
        if (sd.ell) {
            ell = sd.ell + (size_t)u * sd.ell_width;
            indices = nullptr;
            r0 = r1 = 0;
            nchunks = sd.ell_width >> 3;
        } else {
            ell = nullptr;
            indices = sd.indices;
            r0 = sd.indptr[u];
            r1 = sd.indptr[u + 1];
            nchunks = (int)((r1 - r0 + 7) >> 3);
        }
    }
    // 8 neighbour ids (-1 = none); ELL rows come in with two 128-bit loads from one 32-byte sector
    __device__ __forceinline__ void load(int c, int (&nb)[8]) const {
        if (ell) {
            const int4 a = __ldg(reinterpret_cast<const int4 *>(ell) + 2 * c);
            const int4 b2 = __ldg(reinterpret_cast<const int4 *>(ell) + 2 * c + 1);
            nb[0] = a.x; nb[1] = a.y; nb[2] = a.z; nb[3] = a.w;
            nb[4] = b2.x; nb[5] = b2.y; nb[6] = b2.z; nb[7] = b2.w;
        } else {
            const int64_t k0 = r0 + 8 * (int64_t)c;
            const int cnt = (int)min((int64_t)8, r1 - k0);
#pragma unroll
            for (int j = 0; j < 8; ++j) nb[j] = (j < cnt) ? indices[k0 + j] : -1;
        }
    }
};

// component-tree node of the basin sweep: pow(size, E) is looked up once at creation, so a hop of the
// final walk is a single 128-bit load
struct __align__(16) TreeNode {
    double pw;
    int pack; // (sign<<31) | (level<<24) | parent
    int size;
};

struct BasinWs {
    int *up;            // ascent target per SORTED POSITION (the vertex itself for a peak)
    int *basin;         // compact basin id per VERTEX (the one array read at random neighbours)
    int *basinS;        // the same per SORTED POSITION (dense per level)
    int *leaf;          // leaf node per SORTED POSITION (same indexing as order[])
    int *order;         // vertices by (activation level, vertex id)
    TreeNode *nodes;    // {pow(size,E), (sign<<31)|(level<<24)|parent}
    int2 *mlist[2];
    int2 *clist;
    int *bparent_g, *bsize_g, *bcur_g; // global fallback of the per-basin arrays
    unsigned *emask;    // per SORTED POSITION: which ELL slots hold an earlier-activated same-sign neighbour
    float *nodeval;
    unsigned char *blev_g;
    unsigned char *lev8_g;
};

size_t tfce_basin_slot_bytes(int32_t Vmax) {
    const size_t per4 = align_up(sizeof(int) * (size_t)Vmax, 256);
    const size_t per8 = align_up(sizeof(int2) * (size_t)Vmax, 256);
    return per4 * 10 + per8 * 5 + 2 * align_up((size_t)Vmax, 256);
}

__device__ __forceinline__ BasinWs carve_basin(char *base, int32_t Vmax) {
    const size_t per4 = align_up(sizeof(int) * (size_t)Vmax, 256);
    const size_t per8 = align_up(sizeof(int2) * (size_t)Vmax, 256);
    const size_t per1 = align_up((size_t)Vmax, 256);
    BasinWs w;
    w.up = reinterpret_cast<int *>(base);
    w.basin = reinterpret_cast<int *>(base + per4);
    w.leaf = reinterpret_cast<int *>(base + 2 * per4);
    w.order = reinterpret_cast<int *>(base + 3 * per4);
    w.bparent_g = reinterpret_cast<int *>(base + 4 * per4);
    w.bsize_g = reinterpret_cast<int *>(base + 5 * per4);
    w.bcur_g = reinterpret_cast<int *>(base + 6 * per4);
    w.nodeval = reinterpret_cast<float *>(base + 7 * per4);
    w.emask = reinterpret_cast<unsigned *>(base + 8 * per4);
    w.basinS = reinterpret_cast<int *>(base + 9 * per4);
    char *b8 = base + 10 * per4;
    w.nodes = reinterpret_cast<TreeNode *>(b8); // 16 bytes per node: two per8 blocks
    w.mlist[0] = reinterpret_cast<int2 *>(b8 + 2 * per8);
    w.mlist[1] = reinterpret_cast<int2 *>(b8 + 3 * per8);
    w.clist = reinterpret_cast<int2 *>(b8 + 4 * per8);
    w.blev_g = reinterpret_cast<unsigned char *>(b8 + 5 * per8);
    w.lev8_g = w.blev_g + per1;
    return w;
}

// find with path halving on a plain (generic: shared or global) pointer array
__device__ __forceinline__ int bf_find(int *parent, int v) {
    int cur = v;
    int p = parent[cur];
    while (p != cur) {
        const int gp = parent[p];
        if (gp != p) parent[cur] = gp;
        cur = p;
        p = gp;
    }
    return cur;
}

// One step of a component-tree walk: consume node `rec` (levels [its level, parent's level)) and move up.
// Returns false when the root has been consumed.
__device__ __forceinline__ bool walk_step(const TreeNode *__restrict__ nodes, const double (*sHH)[kMaxSteps],
                                          const int *sNs, const double (*sMainPw)[kMaxSteps], TreeNode &rec,
                                          float &acc) {
    const int pk = rec.pack;
    const int par = pk & kNone;
    const int lv = (pk >> 24) & 0x7f;
    const int sg = (pk >> 31) & 1;
    const double *hh = sHH[sg];
    if (rec.size < 0) {
        // The node belongs to the component of the sign's first peak, whose root is never hooked: from here
        // on the path is that component's history, tabulated per level in shared memory -- no more hops.
        const double *mp = sMainPw[sg];
        const int endl = sNs[sg];
        for (int l = lv; l < endl; ++l) acc = __fadd_rn(acc, __double2float_rn(__dmul_rn(mp[l], hh[l])));
        return false;
    }
    TreeNode prec;
    int endl;
    if (par == kNone) {
        endl = sNs[sg];
    } else {
        prec = nodes[par];
        endl = (prec.pack >> 24) & 0x7f;
    }
    const double pw = rec.pw;
    for (int l = lv; l < endl; ++l) acc = __fadd_rn(acc, __double2float_rn(__dmul_rn(pw, hh[l])));
    if (par == kNone) return false;
    rec = prec;
    return true;
}

__device__ __forceinline__ float walk_path16(const TreeNode *__restrict__ nodes, const double (*sHH)[kMaxSteps],
                                             const int *sNs, const double (*sMainPw)[kMaxSteps], int j, float acc) {
    TreeNode rec = nodes[j];
    while (walk_step(nodes, sHH, sNs, sMainPw, rec, acc)) {
    }
    return acc;
}

template <int kThreads, int kMinBlocks>
__global__ void __launch_bounds__(kThreads, kMinBlocks) tfce_basin_kernel(SweepParams P, int smem_budget) {
    extern __shared__ __align__(16) unsigned char sDyn[];
    __shared__ float sT[2][kMaxSteps];
    __shared__ float sHH[2][kMaxSteps];
    __shared__ int sNs[2];
    __shared__ float sDelta[2];
    __shared__ float sScale[2];   // factor of the scaled maximum (the threshold step unless the caller supplies another)
    __shared__ int sStatus[2];
    __shared__ int sStart[kMaxSteps + 1];
    __shared__ float sRed[2][kSweepMaxThreads / 32];
    __shared__ int sItem;
    __shared__ int sMcount[2];
    __shared__ int sCcount[2];
    __shared__ int sNB;
    __shared__ int sStarKey[2];              // smallest (level, id) key per sign: the never-hooked root
    __shared__ double sMainPw[2][kMaxSteps]; // pow(size, E) of that root's component per level (0 = unset)
    __shared__ double sHHd[2][kMaxSteps];    // height terms widened once (the walk multiplies in double)
    __shared__ float sMainInc[2][kMaxSteps]; // per-level increment of the main chain

    const int tid = threadIdx.x;
    const int nthr = blockDim.x;
    const int lane = tid & 31, wid = tid >> 5;
    const int nwarps = nthr >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    const BasinWs ws = carve_basin(P.workspace + (size_t)blockIdx.x * P.slot_stride, P.Vmax);
    const int total_items = P.B * P.S;

    for (;;) {
        if (tid == 0) sItem = atomicAdd(P.work_counter, 1);
        __syncthreads();
        const int item = sItem;
        if (item >= total_items) break;
        if (P.only_flagged && P.only_flagged[(size_t)item * 4 + 2] == 0) { __syncthreads(); continue; }
        const int s = P.surf_order[item / P.B];
        const int b = item % P.B;
        const SurfDesc sd = P.surfs[s];
        const int V = sd.V;
        const float *__restrict__ x = P.stat + (size_t)b * P.ld + sd.col_off;
        const int32_t *__restrict__ vmap = (P.flags & 4) ? nullptr : sd.vmap;
        long long tk = P.timing ? clock64() : 0;
#define TMB_TICK(i)                                                                  \
        if (P.timing && tid == 0) {                                                  \
            const long long now = clock64();                                         \
            atomicAdd(P.timing + (i), (unsigned long long)(now - tk));               \
            tk = now;                                                                \
        }
        // shared-memory layout: [level bytes | scratch: sort histogram, later the per-basin arrays]
        const int hist_bytes = nwarps * kMaxSteps * (int)sizeof(int);
        const int lev_bytes = (int)align_up((size_t)V, 16);
        const bool lev_smem = lev_bytes + hist_bytes <= smem_budget;
        unsigned char *const lev8 = lev_smem ? sDyn : ws.lev8_g;
        const int scratch_off = lev_smem ? lev_bytes : 0;
        int *const sHist = reinterpret_cast<int *>(sDyn + scratch_off); // [nwarps][kMaxSteps]

        for (int i = tid; i < nwarps * kMaxSteps; i += nthr) sHist[i] = 0;
        if (tid == 0) { sMcount[0] = sMcount[1] = 0; sCcount[0] = sCcount[1] = 0; sNB = 0; sStarKey[0] = sStarKey[1] = 0x7fffffff; }
        for (int i = tid; i < 2 * kMaxSteps; i += nthr) sMainPw[i / kMaxSteps][i % kMaxSteps] = 0.0;
        if (P.tab_ns) {
            const size_t e0 = ((size_t)b * P.S + s) * 2;
            for (int i = tid; i < 2 * kMaxSteps; i += nthr) {
                const int sg = i / kMaxSteps, l = i % kMaxSteps;
                sT[sg][l] = P.tab_T[(e0 + sg) * kMaxSteps + l];
                sHH[sg][l] = P.tab_HH[(e0 + sg) * kMaxSteps + l];
            }
            if (tid < 2) {
                const bool on = (tid == 0) || P.two_sided;
                sNs[tid] = on ? P.tab_ns[e0 + tid] : 0;
                sDelta[tid] = P.tab_delta[e0 + tid];
                sScale[tid] = P.tab_scale ? P.tab_scale[e0 + tid] : P.tab_delta[e0 + tid];
                sStatus[tid] = on ? P.tab_status[e0 + tid] : 0;
            }
            __syncthreads();
        } else {
            float mp = -INFINITY, mn = -INFINITY;
            for (int v = tid; v < V; v += nthr) {
                float xv = x[v];
                mp = fmaxf(mp, xv);
                mn = fmaxf(mn, -xv);
            }
            mp = warp_max(mp);
            mn = warp_max(mn);
            if (lane == 0) { sRed[0][wid] = mp; sRed[1][wid] = mn; }
            __syncthreads();
            if (tid == 0 || tid == 32) {
                const int sg = tid ? 1 : 0;
                float mx = -INFINITY;
                for (int w = 0; w < nwarps; ++w) mx = fmaxf(mx, sRed[sg][w]);
                int ns = 0, st = 0;
                float d = 0.f;
                if ((sg == 0 || P.two_sided) && mx >= 0.f) {
                    d = __fdiv_rn(mx, 100.0f);
                    if (d == 0.f) {
                        st = 1;
                    } else {
                        float T = mx;
                        while (T >= 0.f) {
                            if (ns == kMaxSteps) { st = 2; ns = 0; break; }
                            sT[sg][ns] = T;
                            sHH[sg][ns] = (sd.H == 2.0f) ? __fmul_rn(T, T) : (float)pow((double)T, (double)sd.H);
                            ++ns;
                            T = __fsub_rn(T, d);
                        }
                    }
                }
                sNs[sg] = ns;
                sDelta[sg] = d;
                sScale[sg] = d;
                sStatus[sg] = st;
            }
            __syncthreads();
        }
        const int ns0 = sNs[0], ns1 = sNs[1];
        const int nlev = max(ns0, ns1);
        TMB_TICK(0)

        // ---- activation level per vertex + STABLE counting sort by level ----------------------------
        // warp w owns the contiguous vertex chunk [w*chunk, (w+1)*chunk) and its own histogram row
        const int chunk = (int)align_up((size_t)(V + nwarps - 1) / nwarps, 32);
        const int c_beg = wid * chunk, c_end = min(V, c_beg + chunk);
        int *const myHist = sHist + wid * kMaxSteps;
        // activation level = smallest i >= 1 with |x| > T[i]: thresholds are (almost) equally spaced, so the
        // index is guessed from (T[0] - |x|) / delta and corrected against the exact fp32 table (1-2 probes
        // instead of a 7-step binary search).  Four 32-vertex groups per iteration: the four statistic loads
        // (cold: t-maps stream from HBM) are in flight together.
        const float rdel0 = sDelta[0] > 0.f ? __frcp_rn(sDelta[0]) : 0.f;
        const float rdel1 = sDelta[1] > 0.f ? __frcp_rn(sDelta[1]) : 0.f;
        for (int v0 = c_beg; v0 < c_end; v0 += 128) {
            float xq[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int v = v0 + q * 32 + lane;
                xq[q] = (v < c_end) ? x[vmap ? vmap[v] : v] : 0.f;
            }
            int codes[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int v = v0 + q * 32 + lane;
                const float xv = xq[q];
                int code = 0;
                if (xv > 0.f || xv < 0.f) {
                    const int sg = xv < 0.f;
                    const int ns = sg ? ns1 : ns0;
                    if (ns > 1) {
                        const float ax = fabsf(xv);
                        const float *T = sT[sg];
                        int g = __float2int_rz((T[0] - ax) * (sg ? rdel1 : rdel0));
                        g = max(0, min(g, ns - 1)) + 1;
                        while (g > 1 && ax > T[g - 1]) --g;
                        while (g < ns && !(ax > T[g])) ++g;
                        if (g < ns) code = g | (sg << 7);
                    }
                }
                codes[q] = code;
                if (v < c_end) lev8[v] = (unsigned char)code;
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (v0 + q * 32 >= c_end) break; // warp-uniform
                const int lev = codes[q] & 0x7f;
                const unsigned peers = __match_any_sync(0xffffffffu, lev);
                if (lev > 0 && lane == (__ffs(peers) - 1)) myHist[lev] += __popc(peers);
                __syncwarp();
            }
        }
        __syncthreads();
        // level starts, then per-(warp, level) cursors = start[level] + counts of the earlier warps
        if (tid < kMaxSteps) {
            int tot = 0;
            for (int w = 0; w < nwarps; ++w) { const int c = sHist[w * kMaxSteps + tid]; sHist[w * kMaxSteps + tid] = tot; tot += c; }
            sStart[tid] = tot; // per-level totals for now
        }
        __syncthreads();
        if (tid == 0) {
            int run = 0;
            for (int l = 0; l < kMaxSteps; ++l) { const int c = (l == 0) ? 0 : sStart[l]; sStart[l] = run; run += c; }
            sStart[kMaxSteps] = run;
        }
        __syncthreads();
        // Second pass of the sort, fused with the ascent scan (P): in vertex order the fixed-width rows and the
        // level bytes of the neighbours are read coalesced / in a narrow window, and the sorted position is known
        // here, so everything a level owns -- ascent target, earlier-neighbour mask, basin id, later the leaf --
        // is written as a dense segment indexed by sorted position.  Only basin[] is looked up at random.
        const bool use_mask = sd.ell != nullptr && sd.ell_width <= 32;
        int *const peaklist = reinterpret_cast<int *>(ws.clist); // scratch until the level loop starts
        for (int v0 = c_beg; v0 < c_end; v0 += 32) {
            const int v = v0 + lane;
            if (sd.ell && v + 64 < c_end) prefetch_l1(sd.ell + (size_t)(v + 64) * sd.ell_width); // rows two groups ahead
            const int cv = (v < c_end) ? lev8[v] : 0;
            const int lev = cv & 0x7f;
            const unsigned peers = __match_any_sync(0xffffffffu, lev);
            int base = 0;
            if (lev > 0) base = myHist[lev] + sStart[lev];
            __syncwarp();
            if (lev > 0) {
                const int pos = base + __popc(peers & lt_mask);
                ws.order[pos] = v;
                if (lane == (__ffs(peers) - 1)) myHist[lev] += __popc(peers);
                int best = v, bestlev = lev;
                unsigned em = 0; // earlier-activated neighbours: lower level, or same level and smaller index
                const RowIter row(sd, v);
                for (int c = 0; c < row.nchunks; ++c) {
                    int nb[8];
                    row.load(c, nb);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        if (nb[j] < 0) continue;
                        const int ca = lev8[nb[j]];
                        if (ca == 0 || ((ca ^ cv) & 0x80)) continue;
                        const int la = ca & 0x7f;
                        if (la < bestlev) { best = nb[j]; bestlev = la; }
                        if (c < 4 && (la < lev || (la == lev && nb[j] < v))) em |= 1u << (c * 8 + j);
                    }
                }
                ws.up[pos] = best;
                if (use_mask) {
                    // a single earlier neighbour that is also the ascent target can never trigger a union
                    if (best != v && __popc(em) == 1) em = 0;
                    ws.emask[pos] = em;
                }
                if (best == v) { // a peak: new basin
                    const int pid = atomicAdd(&sNB, 1);
                    peaklist[pid] = v; // basin[] is written once the id width is known
                    ws.basinS[pos] = pid;
                    ws.blev_g[pid] = (unsigned char)cv;
                }
            }
            __syncwarp();
        }
        __syncthreads();
        const int total_active = sStart[kMaxSteps];
        TMB_TICK(1)
        const int NB = sNB;
        // basin ids by vertex -- the one array that is looked up at random neighbours -- are stored as 16-bit
        // values whenever they fit: half the sectors, and the sweeps of 148 items in flight then fit the L2
        const bool b16 = NB < 65536;
        unsigned short *const basin16 = reinterpret_cast<unsigned short *>(ws.basin);
#define RD_BASIN(v) (b16 ? (int)basin16[(v)] : ws.basin[(v)])
#define WR_BASIN(v, val) do { if (b16) basin16[(v)] = (unsigned short)(val); else ws.basin[(v)] = (val); } while (0)
        for (int i = tid; i < NB; i += nthr) WR_BASIN(peaklist[i], i);
        // per-basin arrays: parent + level byte first, then size, then current node, while they fit
        int avail = smem_budget - scratch_off;
        int off = scratch_off;
        int *bparent = ws.bparent_g, *bsize = ws.bsize_g, *bcur = ws.bcur_g;
        unsigned char *blev = ws.blev_g;
        unsigned *bbits = reinterpret_cast<unsigned *>(ws.nodeval); // "changed in this level" bitset
        const int nb_al = (int)align_up((size_t)NB, 4);
        const int bits_bytes = (int)align_up((size_t)(NB + 31) / 32 * 4, 16);
        if (avail >= bits_bytes) { bbits = reinterpret_cast<unsigned *>(sDyn + off); off += bits_bytes; avail -= bits_bytes; }
        if (avail >= nb_al * 5) {
            bparent = reinterpret_cast<int *>(sDyn + off); off += nb_al * 4;
            blev = sDyn + off; off += nb_al; avail -= nb_al * 5;
            if (avail >= nb_al * 4) { bsize = reinterpret_cast<int *>(sDyn + off); off += nb_al * 4; avail -= nb_al * 4; }
            if (avail >= nb_al * 4) { bcur = reinterpret_cast<int *>(sDyn + off); off += nb_al * 4; avail -= nb_al * 4; }
        }
        for (int i = tid; i < NB; i += nthr) {
            bparent[i] = i;
            bsize[i] = 0;
            bcur[i] = -1;
            const int bl = ws.blev_g[i];
            if (blev != ws.blev_g) blev[i] = (unsigned char)bl;
            atomicMin(&sStarKey[bl >> 7], ((bl & 0x7f) << 24) | i);
        }
        for (int i = tid; i < (NB + 31) / 32; i += nthr) bbits[i] = 0u;
        // basin ids of the first non-empty level (its vertices are all peaks or ... have no earlier level)
        __syncthreads();
        TMB_TICK(2)

        const int last_level = (P.stop_level >= 0) ? min(P.stop_level, nlev - 1) : nlev - 1;
        bool pend = false;
        int pend_beg = 0, pend_end = 0, pend_buf = 0, pend_nodebase = 0, pend_lev = 0;
        int node_base = 0;
        int buf = 0;
        // assignment of the first non-empty level: every vertex there is a peak (no earlier level exists)
        int assigned_upto = 0; // levels <= assigned_upto carry final basin ids
        for (int l = 1; l < nlev; ++l)
            if (sStart[l + 1] > sStart[l]) { assigned_upto = l; break; }

        for (int lev = 1; lev <= last_level + 1; ++lev) {
            const bool have_level = lev <= last_level;
            const int beg = have_level ? sStart[lev] : 0;
            const int end = have_level ? sStart[lev + 1] : 0;
            if (have_level && end == beg) continue;
            // ================= I1: tree links of the previous level + unions of this level ==============
            if (pend) {
                const int ccount = sCcount[pend_buf];
                const int mcount = sMcount[pend_buf];
                for (int c = tid; c < ccount; c += nthr) {
                    const int2 e = ws.clist[c];
                    const int r = e.x, old = e.y;
                    const int j = pend_nodebase + c;
                    const int sg = blev[r] >> 7;
                    const int sz = bsize[r];
                    TreeNode nd;
                    nd.pw = sd.powE[sz];
                    nd.pack = (sg << 31) | (pend_lev << 24) | kNone;
                    const bool main_chain = r == (sStarKey[sg] & kNone);
                    nd.size = main_chain ? (sz | 0x80000000) : sz;
                    if (main_chain) sMainPw[sg][pend_lev] = nd.pw;
                    ws.nodes[j] = nd;
                    if (old >= 0) ws.nodes[old].pack = (ws.nodes[old].pack & 0xFF000000) | j;
                    atomicAnd(bbits + (r >> 5), ~(1u << (r & 31))); // release this level's claim
                }
                const int2 *ml = ws.mlist[pend_buf];
                for (int m = tid; m < mcount; m += nthr) {
                    const int2 e = ml[m];
                    const int o = bcur[e.x];
                    const int jn = bcur[e.y];
                    ws.nodes[o].pack = (ws.nodes[o].pack & 0xFF000000) | jn;
                }
                for (int idx = pend_beg + tid; idx < pend_end; idx += nthr)
                    ws.leaf[idx] = bcur[ws.leaf[idx]];
            }
            if (!have_level) break;
            {
                int nl = lev + 1;
                while (nl <= last_level && sStart[nl + 1] == sStart[nl]) ++nl;
                if (nl <= last_level)
                    for (int idx = sStart[nl] + tid; idx < sStart[nl + 1]; idx += nthr) {
                        prefetch_l1(ws.order + idx);
                        prefetch_l1(ws.up + idx);
                    }
            }
            int2 *mcur = ws.mlist[buf];
            for (int idx = beg + tid; idx < end; idx += nthr) {
                const int u = ws.order[idx];
                unsigned em = 0xffffffffu;
                if (use_mask) {
                    em = ws.emask[idx];
                    if (em == 0) continue; // no earlier neighbour that could carry another basin
                }
                const int cu = use_mask ? 0 : lev8[u];
                const int bu = ws.basinS[idx];
                const RowIter row(sd, u);
                int ru = bu;
                for (int c = 0; c < row.nchunks; ++c) {
                    const unsigned cm = use_mask ? ((em >> (c * 8)) & 0xffu) : 0xffu;
                    if (cm == 0) continue;
                    int nb[8];
                    row.load(c, nb);
                    int ba[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        ba[j] = -1;
                        if (nb[j] < 0) continue;
                        const int a = nb[j];
                        if (use_mask) {
                            if ((cm >> j) & 1u) ba[j] = RD_BASIN(a);
                        } else {
                            const int ca = lev8[a];
                            if (ca == 0 || ((ca ^ cu) & 0x80)) continue;
                            // earlier-activated neighbour (same level: the pair is seen from both ends, join once)
                            if (ca < cu || (ca == cu && a < u)) ba[j] = RD_BASIN(a);
                        }
                    }
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        if (ba[j] < 0 || ba[j] == bu) continue;
                        int ra = bf_find(bparent, ba[j]);
                        ru = bf_find(bparent, ru);
                        while (ru != ra) {
                            const int kru = ((blev[ru] & 0x7f) << 24) | ru, kra = ((blev[ra] & 0x7f) << 24) | ra;
                            const int hi = kru > kra ? ru : ra, lo = kru > kra ? ra : ru;
                            const int old = atomicCAS(bparent + hi, hi, lo);
                            if (old == hi) {
                                if ((blev[hi] & 0x7f) != lev) { // an older component lost its root
                                    const int m = atomicAdd(&sMcount[buf], 1);
                                    mcur[m] = make_int2(hi, -1);
                                }
                                ru = lo;
                                break;
                            }
                            const int nh = bf_find(bparent, old);
                            if (hi == ru) { ru = nh; ra = bf_find(bparent, ra); }
                            else          { ra = nh; ru = bf_find(bparent, ru); }
                        }
                    }
                }
            }
            __syncthreads();
            if (P.timing && tid == 0) atomicAdd(P.timing + 16 + lev, (unsigned long long)(clock64() - tk));
            TMB_TICK(3)
            // ================= I2: sizes + changed components; basin ids of the next level ==============
            {
                const int mcount = sMcount[buf];
                if (tid == 0) { sMcount[buf ^ 1] = 0; sCcount[buf ^ 1] = 0; }
                const int span = end - beg;
                const int iters = (span + nthr - 1) / nthr;
                for (int it = 0; it < iters; it += 4) { // 4 vertices per thread in flight (independent loads)
                    int idxs[4], rr[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        idxs[q] = beg + (it + q) * nthr + tid;
                        rr[q] = (it + q < iters && idxs[q] < end) ? ws.basinS[idxs[q]] : -1;
                    }
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        if (it + q >= iters) break; // warp-uniform
                        const bool valid = rr[q] >= 0;
                        int r = -1 - lane; // distinct dummy keys for idle lanes
                        if (valid) {
                            r = bf_find(bparent, rr[q]);
                            ws.leaf[idxs[q]] = r;
                        }
                        const unsigned peers = __match_any_sync(0xffffffffu, r);
                        if (valid && lane == (__ffs(peers) - 1)) {
                            atomicAdd(bsize + r, __popc(peers));
                            const unsigned bit = 1u << (r & 31);
                            if (!(atomicOr(bbits + (r >> 5), bit) & bit)) { // first touch of this component in this level
                                const int pos = atomicAdd(&sCcount[buf], 1);
                                ws.clist[pos] = make_int2(r, bcur[r]);
                                bcur[r] = node_base + pos;
                            }
                        }
                    }
                }
                for (int m = tid; m < mcount; m += nthr) {
                    const int h = mcur[m].x;
                    const int r = bf_find(bparent, h);
                    mcur[m].y = r;
                    atomicAdd(bsize + r, bsize[h]);
                }
                // basin ids of the next non-empty level: inherit from the ascent target (an earlier level,
                // final by now).  Peaks already carry their own id.
                int nl = lev + 1;
                while (nl <= last_level && sStart[nl + 1] == sStart[nl]) ++nl;
                if (nl <= last_level && nl > assigned_upto) {
                    const int nbeg = sStart[nl], nend = sStart[nl + 1];
                    for (int idx0 = nbeg + tid; idx0 < nend; idx0 += 4 * nthr) { // 4 independent chains in flight
                        int vv[4], tt[4], bb[4];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const int idx = idx0 + q * nthr;
                            vv[q] = idx < nend ? ws.order[idx] : -1;
                            tt[q] = idx < nend ? ws.up[idx] : -1;
                            if (vv[q] >= 0) { // what I1 of that level reads for this vertex (cold since the scan)
                                if (use_mask) prefetch_l1(ws.emask + idx);
                                if (sd.ell) prefetch_l1(sd.ell + (size_t)vv[q] * sd.ell_width);
                            }
                        }
#pragma unroll
                        for (int q = 0; q < 4; ++q) bb[q] = (vv[q] >= 0 && tt[q] != vv[q]) ? RD_BASIN(tt[q]) : -1;
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            if (vv[q] >= 0 && tt[q] != vv[q]) {
                                WR_BASIN(vv[q], bb[q]);
                                ws.basinS[idx0 + q * nthr] = bb[q];
                            }
                    }
                }
            }
            __syncthreads();
            if (P.timing && tid == 0) atomicAdd(P.timing + 16 + 128 + lev, (unsigned long long)(clock64() - tk));
            TMB_TICK(4)
            pend = true; pend_beg = beg; pend_end = end; pend_buf = buf; pend_nodebase = node_base; pend_lev = lev;
            node_base += sCcount[buf];
            buf ^= 1;
        }
        __syncthreads();
        const int num_nodes = node_base;

        // ---- outputs (identical to V2) -------------------------------------------------------------
        if (P.stop_level >= 0) {
            for (int v = tid; v < V; v += nthr) {
                const int o = vmap ? vmap[v] : v;
                int lab = -1, ext = 0;
                const int code = lev8[v];
                if (code != 0 && !(code & 0x80) && (code & 0x7f) <= last_level) {
                    const int r = bf_find(bparent, RD_BASIN(v));
                    lab = r;
                    ext = bsize[r];
                }
                P.labels[o] = lab;
                P.extents[o] = ext;
            }
            if (tid == 0 && P.threshold_out)
                *P.threshold_out = (P.stop_level < ns0) ? sT[0][P.stop_level] : NAN;
        } else {
            float *nodeval = ws.nodeval;
            const bool per_vertex_walk = P.accumulate != 0;
            for (int i = tid; i < 2 * kMaxSteps; i += nthr) sHHd[i / kMaxSteps][i % kMaxSteps] = (double)sHH[i / kMaxSteps][i % kMaxSteps];
            if (tid == 0 || tid == 32) { // dense per-level table of the main chain (carry unchanged levels forward)
                double *mp = sMainPw[tid ? 1 : 0];
                for (int l = 1; l < kMaxSteps; ++l)
                    if (mp[l] == 0.0) mp[l] = mp[l - 1];
            }
            __syncthreads();
            for (int i = tid; i < 2 * kMaxSteps; i += nthr) {
                const int sg = i / kMaxSteps, l = i % kMaxSteps;
                sMainInc[sg][l] = __double2float_rn(__dmul_rn(sMainPw[sg][l], sHHd[sg][l]));
            }
            __syncthreads();
            if (!per_vertex_walk) {
                // Level-synchronous walk: node ids ascend with their creation level, so the 32 walkers of a warp
                // start at (almost) the same level and step through the levels in lock-step -- one fp32 add per
                // lane per level, hops (parent loads) only where a lane's node ends.  Two walkers per thread.
                for (int j0 = tid; j0 < num_nodes; j0 += 2 * nthr) {
                    TreeNode prec[2];
                    double pw[2];
                    float acc[2];
                    int lvl[2], endl[2], sgn[2], nsq[2];
                    bool mainc[2];
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        const int j = j0 + q * nthr;
                        acc[q] = 0.f; lvl[q] = 0; endl[q] = 0; sgn[q] = 0; nsq[q] = 0; mainc[q] = true; pw[q] = 0.0;
                        if (j < num_nodes) {
                            const TreeNode rec = ws.nodes[j];
                            const int par = rec.pack & kNone;
                            sgn[q] = (rec.pack >> 31) & 1;
                            lvl[q] = (rec.pack >> 24) & 0x7f;
                            nsq[q] = sNs[sgn[q]];
                            pw[q] = rec.pw;
                            mainc[q] = rec.size < 0;
                            endl[q] = nsq[q];
                            if (!mainc[q] && par != kNone) {
                                prec[q] = ws.nodes[par];
                                endl[q] = (prec[q].pack >> 24) & 0x7f;
                                if ((prec[q].pack & kNone) != kNone) prefetch_l1(ws.nodes + (prec[q].pack & kNone));
                            }
                        }
                    }
                    while (lvl[0] < nsq[0] || lvl[1] < nsq[1]) {
#pragma unroll
                        for (int q = 0; q < 2; ++q) {
                            if (lvl[q] >= nsq[q]) continue;
                            if (lvl[q] == endl[q]) { // this lane's node ends here: move to its parent
                                const TreeNode rec = prec[q];
                                const int par = rec.pack & kNone;
                                pw[q] = rec.pw;
                                mainc[q] = rec.size < 0;
                                endl[q] = nsq[q];
                                if (!mainc[q] && par != kNone) {
                                    prec[q] = ws.nodes[par];
                                    endl[q] = (prec[q].pack >> 24) & 0x7f;
                                    if ((prec[q].pack & kNone) != kNone) prefetch_l1(ws.nodes + (prec[q].pack & kNone));
                                }
                            }
                            const float inc = mainc[q] ? sMainInc[sgn[q]][lvl[q]]
                                                       : __double2float_rn(__dmul_rn(pw[q], sHHd[sgn[q]][lvl[q]]));
                            acc[q] = __fadd_rn(acc[q], inc);
                            ++lvl[q];
                        }
                    }
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        const int j = j0 + q * nthr;
                        if (j < num_nodes) nodeval[j] = acc[q];
                    }
                }
                __syncthreads();
            }
            TMB_TICK(5)
            if (P.timing && tid == 0) { atomicAdd(P.timing + 7, (unsigned long long)num_nodes); }
            float m0 = 0.f, m1 = 0.f;
            const float d0 = sScale[0], d1 = sScale[1];
            const bool want_maps = (P.tfce_pos != nullptr) || (P.tfce_neg != nullptr);
            if (want_maps && !P.accumulate) {
                for (int v = tid; v < V; v += nthr) {
                    const size_t o = (size_t)b * P.ld + sd.col_off + v;
                    if (P.tfce_pos) P.tfce_pos[o] = 0.f;
                    if (P.tfce_neg) P.tfce_neg[o] = 0.f;
                }
                __syncthreads();
            }
            for (int idx = tid; idx < total_active; idx += nthr) {
                const int u = ws.order[idx];
                const bool neg = (lev8[u] & 0x80) != 0;
                float val;
                if (per_vertex_walk || want_maps) {
                    const size_t o = (size_t)b * P.ld + sd.col_off + (vmap ? vmap[u] : u);
                    float *dst = neg ? P.tfce_neg : P.tfce_pos;
                    if (per_vertex_walk) val = walk_path16(ws.nodes, sHHd, sNs, sMainPw, ws.leaf[idx], dst ? dst[o] : 0.f);
                    else val = nodeval[ws.leaf[idx]];
                    if (dst) dst[o] = val;
                } else {
                    val = nodeval[ws.leaf[idx]];
                }
                const float sc = scaled_vertex_value(val, neg ? d1 : d0, sd, u);
                if (neg) m1 = fmaxf(m1, sc); else m0 = fmaxf(m0, sc);
            }
            m0 = warp_max(m0);
            m1 = warp_max(m1);
            __syncthreads();
            if (lane == 0) { sRed[0][wid] = m0; sRed[1][wid] = m1; }
            __syncthreads();
            if (tid == 0) {
                float a = 0.f, c = 0.f;
                for (int w2 = 0; w2 < nwarps; ++w2) { a = fmaxf(a, sRed[0][w2]); c = fmaxf(c, sRed[1][w2]); }
                const size_t o = ((size_t)b * P.S + s) * 2;
                if (P.max_out) { P.max_out[o] = a; P.max_out[o + 1] = c; }
                if (P.status) { P.status[o] = sStatus[0]; P.status[o + 1] = sStatus[1]; }
            }
        }
        __syncthreads();
        TMB_TICK(6)
#undef TMB_TICK
#undef RD_BASIN
#undef WR_BASIN
    }
}

// Launch geometry (measured on B200, profiles/README.md): the sweep is bound by the latency of dependent,
// data-dependent loads, and a large L1 serves them better than a large shared-memory carve-out -- keeping
// the 146 KB of level bytes of an fsaverage hemisphere in shared memory was 45% SLOWER than leaving them in
// global memory behind a ~128 KB L1.  So only the small per-basin union-find arrays (and the sort histogram)
// live in shared memory: ~100 KB per SM in total, split between the resident CTAs.
//   basin sweep: 1 CTA x 1024 threads per SM for large surfaces, 2 x 512 below 100k vertices,
//                4 x 256 below 40k (small surfaces are latency bound: more items in flight)
//   V2 (directed adjacency): 2 CTAs x 512 threads, no dynamic shared memory.
void tfce_sweep_geometry(int32_t Vmax, int use_basin, int *threads, int *ctas_per_sm, size_t *dyn_smem) {
    if (use_basin) {
        size_t total = 100 * 1024;
        if (Vmax <= 40000) { *threads = 256; *ctas_per_sm = 4; }
        else if (Vmax <= 100000) { *threads = 512; *ctas_per_sm = 2; }
        else { *threads = 1024; *ctas_per_sm = 1; }
        if (const char *g = getenv("TMB_GEOM")) { // "threads,ctas_per_sm" (development aid)
            int t = 0, c = 0;
            if (sscanf(g, "%d,%d", &t, &c) == 2 && t >= 64 && t <= 1024 && c >= 1 && c <= 8) { *threads = t; *ctas_per_sm = c; }
        }
        if (const char *g = getenv("TMB_SMEM_KB")) { int kb = atoi(g); if (kb >= 16 && kb <= 216) total = (size_t)kb * 1024; }
        *dyn_smem = total / *ctas_per_sm - 6 * 1024;
        return;
    }
    *threads = 512; *ctas_per_sm = 2; *dyn_smem = 0;
}

size_t tfce_slot_bytes_for(int32_t Vmax, int use_basin) {
    return use_basin ? tfce_basin_slot_bytes(Vmax) : tfce_slot_bytes(Vmax);
}

int launch_tfce_sweep(const SweepParams &p, int num_slots, cudaStream_t stream) {
    const int items = p.B * p.S;
    if (items <= 0) return 0;
    const int use_basin = (p.flags & 2) != 0;
    int threads, per_sm;
    size_t dyn;
    tfce_sweep_geometry(p.Vmax, use_basin, &threads, &per_sm, &dyn);
    int grid = items < num_slots ? items : num_slots;
    TMB_CUDA(cudaMemsetAsync(p.work_counter, 0, sizeof(int), stream));
    const bool cached = (p.flags & 1) != 0;
    if (use_basin) {
        // register budget follows the geometry: 1024 threads -> 64 regs, 512 x 1 -> 128, 512 x 2 / 256 x 4 -> 64
        if (threads > 512) {
            TMB_CUDA(cudaFuncSetAttribute(tfce_basin_kernel<1024, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
            tfce_basin_kernel<1024, 1><<<grid, threads, dyn, stream>>>(p, (int)dyn);
        } else if (threads > 256 && per_sm == 1) {
            TMB_CUDA(cudaFuncSetAttribute(tfce_basin_kernel<512, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
            tfce_basin_kernel<512, 1><<<grid, threads, dyn, stream>>>(p, (int)dyn);
        } else if (threads > 256) {
            TMB_CUDA(cudaFuncSetAttribute(tfce_basin_kernel<512, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
            tfce_basin_kernel<512, 2><<<grid, threads, dyn, stream>>>(p, (int)dyn);
        } else {
            TMB_CUDA(cudaFuncSetAttribute(tfce_basin_kernel<256, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
            tfce_basin_kernel<256, 4><<<grid, threads, dyn, stream>>>(p, (int)dyn);
        }
    } else if (dyn > 0) {
        if (cached) {
            TMB_CUDA(cudaFuncSetAttribute(tfce_sweep_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
            tfce_sweep_kernel<true, true><<<grid, threads, dyn, stream>>>(p);
        } else {
            TMB_CUDA(cudaFuncSetAttribute(tfce_sweep_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
            tfce_sweep_kernel<true, false><<<grid, threads, dyn, stream>>>(p);
        }
    } else {
        if (cached) tfce_sweep_kernel<false, true><<<grid, threads, 0, stream>>>(p);
        else tfce_sweep_kernel<false, false><<<grid, threads, 0, stream>>>(p);
    }
    count_launch();
    TMB_CUDA(cudaGetLastError());
    return 0;
}

} // namespace tmb
