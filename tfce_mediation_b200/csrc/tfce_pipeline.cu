// TFCE pipeline (sm_100a): the level sweep of lib/fast_tfce.hpp:11-95 moved from the VERTICES to the BASINS.
//
// tfce_basin_kernel (tfce_kernels.cu) walks every vertex once per barrier interval of a 100-level loop; it
// is bound by the latency of dependent loads inside those intervals.  Here everything that is per VERTEX
// is done by flat, barrier-free streaming kernels over all (statistic row, surface, vertex) triples of a
// block of maps -- full occupancy, coalesced rows, no level loop:
//
//   K_A  levels   activation level per vertex (fast_tfce.hpp:39-41: x > T_i, exact fp32 threshold table) and the
//                 map's histogram of the levels
//   K_B  ascent   per vertex: the neighbour of the earliest level ("up"), the set of earlier neighbours;
//                 vertices without an earlier neighbour are peaks and get compact basin ids
//   K_C  basins   basin id per vertex = peak at the end of its ascent chain (read-only pointer chase); max-only
//                 maps: the basin ids of the active vertices bucketed by activation level (the sweep's input)
//   K_D  counts   every (vertex, earlier neighbour) pair that straddles two basins becomes a candidate union
//                 (level, basin, basin); class path only: table[level][basin] += 1
//
// and the sequential part -- the threshold sweep with its union-find, component sizes and the fp32 sums
// in the reference's order -- runs on the few thousand basins of a map, entirely in shared memory:
//
//   K_S  sweep    one CTA per map: candidate unions bucketed by level; per level: unions, sizes from the level's
//                 vertex list (max-only maps: one leader accumulator per live root, pipe_sweep_max_kernel) or from
//                 the table row (class path: one accumulator per component that gains vertices), and every live
//                 root / class adds fl32(pow(size, E) * pow(T, H)) of its component (fast_tfce.hpp:70-84)
//   K_G  output   (only with vertex weights or when the maps are requested) per vertex value lookup
//
// Values are bit-identical to the reference: a vertex activated at level l receives, one fp32 add per level
// l, l+1, ... in descending-threshold order, the increment of the component it belongs to at that level.
// All vertices of one (level, component) share that sequence, hence one accumulator per such pair.
// Maps whose basin count / pair count exceed the fixed capacities are flagged and redone by
// tfce_basin_kernel (launch_tfce_sweep with only_flagged), so the result never depends on a capacity.
#include "common.cuh"

#include <climits>
#include <cstdlib>
#include <type_traits>

namespace tmb {

static constexpr int kLevels = 128;
static constexpr int kInactive = INT_MIN;

namespace {

__device__ __forceinline__ int pf_find(int *parent, int v) {
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

// A base pointer the compiler must keep in registers.  Without it nvcc re-derives `P.array + item * vstride + index` --
// a 64-bit multiply-add plus two constant-bank loads -- in front of EVERY predicated gather of the flat kernels (seen in
// the SASS of round 1: seven instructions per neighbour byte instead of three).
template <typename T>
__device__ __forceinline__ T *pin_ptr(T *p) {
    asm volatile("" : "+l"(p));
    return p;
}

// asynchronous 8-byte copy global -> shared (LDGSTS): no register, no scoreboard stall until cp_async_wait_all
__device__ __forceinline__ void cp_async8(void *smem_dst, const void *gmem_src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src)
                 : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// predicated 8-byte global load: issued without a dependent select, so the thread does not stall until first use
__device__ __forceinline__ unsigned long long ld_u64_if(const unsigned long long *p, bool pred) {
    unsigned long long v = 0ull;
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\t@q ld.global.u64 %0, [%1];\n\t}" : "+l"(v) : "l"(p), "r"((int)pred) : "memory");
    return v;
}

__device__ __forceinline__ float pwarp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

__device__ __forceinline__ void item_coords(const PipeParams &P, int item, int &s, int &b) {
    s = P.surf_order[item / P.B];
    b = item % P.B;
}

// streaming kernels run on a 3-D grid (chunk, statistic row, surface slot): no integer divisions per thread
__device__ __forceinline__ void grid_coords(const PipeParams &P, int &item, int &chunk, int &s, int &b) {
    chunk = blockIdx.x;
    b = blockIdx.y;
    const int z = blockIdx.z + P.z0;
    s = P.surf_order[z];
    item = z * P.B + b;
}

} // namespace

// ------------------------------------------------------------------------------------------- tables
// exact_pow = False: threshold tables on the device from the per-map maxima, same fp32 operations as
// fast_tfce.hpp:32-39 with a correctly rounded height term (see tfce_basin_kernel).
__global__ void pipe_tables_kernel(const SurfDesc *__restrict__ surfs, int S, int count, const float *__restrict__ maxima,
                                   int two_sided, int32_t *__restrict__ ns_out, float *__restrict__ delta_out,
                                   float *__restrict__ T_out, float *__restrict__ HH_out, int32_t *__restrict__ st_out) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= count) return;
    const int sg = e & 1;
    const int s = (e >> 1) % S;
    const float H = surfs[s].H;
    const float mx = maxima[e];
    int ns = 0, st = 0;
    float d = 0.f;
    if ((sg == 0 || two_sided) && mx >= 0.f) {
        d = __fdiv_rn(mx, 100.0f);
        if (d == 0.f) {
            st = 1;
        } else {
            float T = mx;
            while (T >= 0.f) {
                if (ns == kLevels) { st = 2; ns = 0; break; }
                T_out[(size_t)e * kLevels + ns] = T;
                HH_out[(size_t)e * kLevels + ns] = (H == 2.0f) ? __fmul_rn(T, T) : (float)pow((double)T, (double)H);
                ++ns;
                T = __fsub_rn(T, d);
            }
        }
    }
    ns_out[e] = ns;
    delta_out[e] = d;
    st_out[e] = st;
}

// ------------------------------------------------------------------------------------------- K_A
static constexpr int kChunkA = 2048; // vertices per CTA (8 per thread)

__global__ void __launch_bounds__(256) pipe_levels_kernel(PipeParams P, int chunks) {
    __shared__ float sT[2][kLevels];
    __shared__ int sNs[2];
    __shared__ float sRd[2];
    __shared__ int sHist[kLevels]; // vertices of this CTA per activation level -> the map's histogram (lhist, zeroed by the launcher)
    const int tid = threadIdx.x;
    int item, chunk, s, b;
    grid_coords(P, item, chunk, s, b);
    const SurfDesc sd = P.surfs[s];
    const int V = sd.V;
    const int v_beg = chunk * kChunkA;
    if (v_beg >= V) return;
    const size_t e0 = ((size_t)b * P.S + s) * 2;
    sT[tid >> 7][tid & 127] = P.tab_T[(e0 + (tid >> 7)) * kLevels + (tid & 127)];
    if (tid < 2) {
        const bool on = (tid == 0) || P.two_sided;
        sNs[tid] = on ? P.tab_ns[e0 + tid] : 0;
        const float d = P.tab_delta[e0 + tid];
        sRd[tid] = d > 0.f ? __frcp_rn(d) : 0.f;
    }
    if (chunk == 0 && tid < 4) P.meta[(size_t)item * 4 + tid] = 0; // npeaks, npairs, flag (K_B runs after this kernel)
    if (tid < kLevels) sHist[tid] = 0;
    __syncthreads();
    const float *__restrict__ x = P.stat + (size_t)b * P.ld + sd.col_off;
    const int32_t *__restrict__ vmap = (P.flags & 4) ? nullptr : sd.vmap;
    unsigned char *__restrict__ lev8 = P.lev8 + (size_t)item * P.vstride;
    const int ns0 = sNs[0], ns1 = sNs[1];
    float xq[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const int v = v_beg + q * 256 + tid;
        xq[q] = (v < V) ? x[vmap ? vmap[v] : v] : 0.f;
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const int v = v_beg + q * 256 + tid;
        if (v >= V) break;
        const float xv = xq[q];
        int code = 0;
        if (xv > 0.f || xv < 0.f) {
            const int sg = xv < 0.f;
            const int ns = sg ? ns1 : ns0;
            if (ns > 1) {
                // smallest i >= 1 with |x| > T[i]: guessed from the (almost) equal spacing, fixed against the exact table
                const float ax = fabsf(xv);
                const float *T = sT[sg];
                int g = __float2int_rz((T[0] - ax) * sRd[sg]);
                g = max(0, min(g, ns - 1)) + 1;
                while (g > 1 && ax > T[g - 1]) --g;
                while (g < ns && !(ax > T[g])) ++g;
                if (g < ns) code = g | (sg << 7);
            }
        }
        lev8[v] = (unsigned char)code;
        if (code) atomicAdd(&sHist[code & 0x7f], 1);
    }
    __syncthreads();
    if (tid < kLevels && sHist[tid]) atomicAdd(P.lhist + (size_t)item * 256 + tid, sHist[tid]);
}

// ------------------------------------------------------------------------------------------- K_B
// ascent target and earlier-neighbour mask of vertex v; returns the vertex's level code when it is a peak, else -1
// (Staging the CTA's own 256 level codes in shared memory -- 81% of the neighbour lookups fall into the vertex's own
//  block, scripts/probe_window.py -- was measured slower than the L1 gathers: 1.26 vs 1.17 ms.)
// kSix: every row has at most six neighbours in one 8-slot chunk (triangle meshes): slots 6 and 7 are padding and skipped.
// kSelf: the rows are padded with the vertex itself (SurfDesc::ell_self, width 8): a self-reference is never "earlier"
// and never an ascent target, so the gathers need no pad test.  Levels are compared as y = (level - 1) & 255 for a
// same-sign active neighbour (0..126) and >= 127 for everything else, which needs no select.
template <bool kSix, bool kSelf>
__device__ __forceinline__ int ascent_of_vertex(const PipeParams &P, const SurfDesc &sd, size_t base, int v) {
    if (v >= sd.V) return -1;
    const unsigned char *__restrict__ lev8 = pin_ptr(P.lev8 + base);
    const int cv = lev8[v];
    const int lev = cv & 0x7f;
    if (lev == 0) {
        P.up[base + v] = kInactive;
        P.emask[base + v] = 0u;
        return -1;
    }
    // per neighbour: y = its level - 1 when it is active with v's sign (0..126), else 127..255.  The ascent target is the
    // first neighbour of the smallest level below v's own (packed key: y << 8 | slot); "earlier" = (level, index) below
    // (lev, v), one packed unsigned compare.
    unsigned bestkey = ((unsigned)(lev - 1) << 8) | 0xffu; // no strictly earlier level found yet
    unsigned em = 0;
    const unsigned sign = (unsigned)cv & 0x80u;
    const unsigned mykey = ((unsigned)(lev - 1) << 24) | (unsigned)v;
    const int width = kSelf ? 8 : sd.ell_width;
    const int4 *__restrict__ row = pin_ptr(reinterpret_cast<const int4 *>((kSelf ? sd.ell_self : sd.ell) + (size_t)v * width));
    const int nch = (kSix || kSelf) ? 1 : (width >> 3);
    constexpr int kSlots = kSix ? 6 : 8;
    for (int c = 0; c < nch; ++c) {
        const int4 r0 = __ldg(row + 2 * c), r1 = __ldg(row + 2 * c + 1);
        const int nb[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
        unsigned ca[8];
#pragma unroll
        for (int j = 0; j < kSlots; ++j) ca[j] = (unsigned)__ldg(lev8 + (unsigned)(kSelf ? nb[j] : max(nb[j], 0)));
#pragma unroll
        for (int j = 0; j < kSlots; ++j) {
            unsigned y = ((ca[j] ^ sign) + 255u) & 255u;       // 0..126: active, same sign
            if (!kSelf && nb[j] < 0) y = 255u;                 // pad slot (-1)
            bestkey = min(bestkey, (y << 8) | (unsigned)(c * 8 + j));
            if (((y << 24) | (unsigned)nb[j]) < mykey) em |= 1u << (c * 8 + j); // nb < 2^24; y >= 127 never passes
        }
    }
    const int bestbit = (int)(bestkey & 0xffu);
    const bool has_up = (int)(bestkey >> 8) < lev - 1;
    int best = v;
    if (has_up) {
        best = reinterpret_cast<const int *>(row)[bestbit];
        em &= ~(1u << bestbit); // the ascent target lies in the same basin by construction
        P.up[base + v] = best;
    }
    P.emask[base + v] = em;
    return best == v ? cv : -1;
}

template <bool kSix, bool kSelf>
__global__ void __launch_bounds__(256) pipe_ascent_kernel(PipeParams P, int chunks) {
    __shared__ int sPeaks, sBase;
    int item, chunk, s, b;
    grid_coords(P, item, chunk, s, b);
    const SurfDesc sd = P.surfs[s];
    if (chunk * 256 >= sd.V) return; // CTA-uniform
    if (threadIdx.x == 0) sPeaks = 0;
    __syncthreads();
    const int v = chunk * 256 + threadIdx.x;
    const size_t base = (size_t)item * P.vstride;
    const int peak_code = ascent_of_vertex<kSix, kSelf>(P, sd, base, v);
    // peaks get compact basin ids: one returning atomic per CTA on the map's counter
    int local = -1;
    if (peak_code >= 0) local = atomicAdd(&sPeaks, 1);
    __syncthreads();
    if (threadIdx.x == 0 && sPeaks > 0) sBase = atomicAdd(P.meta + (size_t)item * 4, sPeaks);
    __syncthreads();
    if (local >= 0) {
        const int pid = sBase + local;
        if (pid < P.nbcap) P.blev[(size_t)item * P.nbcap + pid] = (unsigned char)peak_code;
        P.up[base + v] = -1 - pid;
    }
}


// ---- K_B, sliced rows (SELL-32-4): graphs with more than 32 neighbours per vertex (the reference's default geodesic
// adjacency sets, ~60 per vertex) or very uneven degrees.  One thread per vertex as above; the 32 lanes of a warp own
// the 32 vertices of one slice, so slot group j4 of the slice is ONE coalesced 512-byte load for the warp and the trip
// count is warp-uniform.  kWords 32-bit words of earlier-neighbour mask per vertex (slot j -> bit j & 31 of word j >> 5).
// Measured and rejected: taking each vertex through 2 or 4 maps per thread, so that every slot group is loaded once and
// serves 8 / 16 level gathers (the rows are 77 % of the kernel's L2 sectors, which run at 7.5 TB/s): correct, and not
// faster -- 24.2 / 24.2 / 24.6 ms per 1,024 maps for the stage with 1 / 2 / 4 maps per thread (config 3: 10.47 / 10.59).
// The byte gathers, not the row stream, set the pace.
// Also rejected: staging the map's level bytes of the chunk's neighbourhood in shared memory (1,024 vertices per CTA, a
// window of +-8,192 vertices = 17 KB, which holds 99.9 % of the neighbours; the rest read from global memory): correct,
// slower -- 26.3 against 24.3 ms (config 3: 11.17 against 10.48).  L1 serves these semi-coherent byte gathers better
// than shared memory plus the staging copy and the two-path select.
template <int kWords>
__global__ void __launch_bounds__(256) pipe_ascent_wide_kernel(PipeParams P, int chunks) {
    __shared__ int sPeaks, sBase;
    int item, chunk, s, b;
    grid_coords(P, item, chunk, s, b);
    const SurfDesc sd = P.surfs[s];
    if (chunk * 256 >= sd.V) return; // CTA-uniform
    if (threadIdx.x == 0) sPeaks = 0;
    __syncthreads();
    const int v = chunk * 256 + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const size_t base = (size_t)item * P.vstride;
    int peak_code = -1;
    if (v < sd.V) {
        const unsigned char *__restrict__ lev8 = pin_ptr(P.lev8 + base);
        const int cv = lev8[v];
        const int lev = cv & 0x7f;
        if (lev == 0) {
            P.up[base + v] = kInactive; // the mask words of an inactive vertex are never read
        } else {
            const int o0 = sd.sell_off[v >> 5];
            const int n4 = (sd.sell_off[(v >> 5) + 1] - o0) >> 5; // slot groups of this slice (warp-uniform)
            const int4 *__restrict__ rows = pin_ptr(sd.sell + o0 + lane);
            unsigned bestkey = ((unsigned)lev << 8) | 0xffu; // level << 8 | slot of the ascent target so far
            const unsigned sign = (unsigned)cv & 0x80u;
            const unsigned mykey = ((unsigned)lev << 24) | (unsigned)v;
            unsigned em[kWords];
#pragma unroll
            for (int w = 0; w < kWords; ++w) {
                unsigned e = 0u;
                if (w * 8 < n4) {
#pragma unroll 2
                    for (int jj = 0; jj < 8; ++jj) {
                        const int j4 = w * 8 + jj;
                        if (j4 >= n4) break;
                        const int4 r = __ldg(rows + j4 * 32);
                        const int nb[4] = {r.x, r.y, r.z, r.w};
                        unsigned ca[4];
#pragma unroll
                        for (int q = 0; q < 4; ++q) ca[q] = (unsigned)__ldg(lev8 + (unsigned)nb[q]); // pad slots hold the vertex itself
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const unsigned x = ca[q] ^ sign;                 // 1..127: active, same sign
                            const unsigned la = (x - 1u) < 127u ? x : 255u;
                            bestkey = min(bestkey, (la << 8) | (unsigned)(j4 * 4 + q));
                            if (((la << 24) | (unsigned)nb[q]) < mykey) e |= 1u << (jj * 4 + q);
                        }
                    }
                }
                em[w] = e;
            }
            const int bestbit = (int)(bestkey & 0xffu);
            const bool has_up = (int)(bestkey >> 8) < lev;
            if (has_up) {
                const int best = reinterpret_cast<const int *>(rows + (bestbit >> 2) * 32)[bestbit & 3];
#pragma unroll
                for (int w = 0; w < kWords; ++w)
                    if (w == (bestbit >> 5)) em[w] &= ~(1u << (bestbit & 31)); // the ascent target lies in the same basin
                P.up[base + v] = best;
            } else {
                peak_code = cv;
            }
#pragma unroll
            for (int w = 0; w < kWords; ++w) P.emask[((size_t)item * kWords + w) * P.vstride + v] = em[w];
        }
    }
    // peaks get compact basin ids: one returning atomic per CTA on the map's counter
    int local = -1;
    if (peak_code >= 0) local = atomicAdd(&sPeaks, 1);
    __syncthreads();
    if (threadIdx.x == 0 && sPeaks > 0) sBase = atomicAdd(P.meta + (size_t)item * 4, sPeaks);
    __syncthreads();
    if (local >= 0) {
        const int pid = sBase + local;
        if (pid < P.nbcap) P.blev[(size_t)item * P.nbcap + pid] = (unsigned char)peak_code;
        P.up[base + v] = -1 - pid;
    }
}

// ------------------------------------------------------------------------------------------- K_C
// Basin of every vertex (read-only pointer chase along `up`, four independent chains per thread), and -- for max-only
// maps -- the map's VERTEX LISTS bucketed by activation level: the basin id (2 bytes) of every active vertex, which the
// sweep reads level by level.  K_A left the level histogram; a CTA ranks its 1,024 vertices per level in shared memory,
// reserves its range of each level's list with one returning atomic on the level's cursor, and writes the runs from a
// shared-memory staging buffer (consecutive 2-byte stores instead of one 32-byte L2 sector per vertex).  Order inside a
// level is irrelevant.  The kernel waits on dependent loads most of the time (the chase), so the ranking rides along
// almost for free; the sweep did it itself until round-1 v9 (one CTA per SM: 19% of its time), the count kernel in v10
// (+1.1 ms: it is bound by its instruction count).
static constexpr int kBasinVPT = 4;                 // vertices per thread
static constexpr int kBasinChunk = 256 * kBasinVPT; // vertices per CTA

template <bool kWeighted>
__global__ void __launch_bounds__(256) pipe_basin_kernel(PipeParams P, int chunks) {
    using StageT = typename std::conditional<kWeighted, unsigned, unsigned short>::type;
    int item, chunk, s, b;
    grid_coords(P, item, chunk, s, b);
    const int V = P.surfs[s].V;
    const unsigned short *__restrict__ wrank = P.surfs[s].wrank;
    const int v_beg = chunk * kBasinChunk;
    if (v_beg >= V) return; // CTA-uniform
    const int tid = threadIdx.x, lane = tid & 31;
    const size_t base = (size_t)item * P.vstride;
    const int *__restrict__ up = pin_ptr(P.up + base);
    int t[kBasinVPT], levq[kBasinVPT];
#pragma unroll
    for (int q = 0; q < kBasinVPT; ++q) {
        const int v = v_beg + q * 256 + tid;
        t[q] = v < V ? up[v] : kInactive;
        levq[q] = v < V ? (int)(P.lev8[base + v] & 0x7f) : 0;
    }
    for (;;) { // kInactive and the peaks' -1 - basin are negative: a chain ends at the first negative entry
        bool more = false;
#pragma unroll
        for (int q = 0; q < kBasinVPT; ++q)
            if (t[q] >= 0) { t[q] = __ldg(up + (unsigned)t[q]); more |= t[q] >= 0; }
        if (!more) break;
    }
    int bas[kBasinVPT];
#pragma unroll
    for (int q = 0; q < kBasinVPT; ++q) {
        const int v = v_beg + q * 256 + tid;
        bas[q] = t[q] == kInactive ? -1 : -1 - t[q];
        if (v < V) P.basin[base + v] = bas[q];
    }
    const size_t e0 = ((size_t)b * P.S + s) * 2;
    const int NB = P.meta[(size_t)item * 4];
    const int nlev = max(P.tab_ns[e0], P.two_sided ? P.tab_ns[e0 + 1] : 0);
    const int64_t n = (int64_t)nlev * NB;
    if (NB > P.nbcap || (P.want_vertex_pass && n > P.tabcap)) { // CTA-uniform
        if (chunk == 0 && tid == 0) P.meta[(size_t)item * 4 + 2] = 1; // over capacity: redone by tfce_basin_kernel
        return;
    }
    if (P.want_vertex_pass) {
        // class path: zero this map's count table (rows 1 .. nlev-1 of NB entries), a slice per vertex slot
        unsigned *__restrict__ tab = P.table + (size_t)item * P.tabcap;
#pragma unroll
        for (int q = 0; q < kBasinVPT; ++q) {
            const int v = v_beg + q * 256 + tid;
            if (v < V)
                for (int64_t i = v; i < n; i += V) tab[i] = 0u;
        }
        return;
    }
    // ---- max-only maps: vertex lists by level
    __shared__ int sLcnt[kLevels], sLbase[kLevels], sLloc[kLevels];
    __shared__ unsigned long long sWtot[kLevels / 32];
    __shared__ StageT sStage[kBasinChunk];
    __shared__ int sStagePos[kBasinChunk];
    if (tid < kLevels) sLcnt[tid] = 0;
    __syncthreads();
    int rankq[kBasinVPT];
#pragma unroll
    for (int q = 0; q < kBasinVPT; ++q) rankq[q] = levq[q] ? atomicAdd(&sLcnt[levq[q]], 1) : 0;
    __syncthreads();
    {
        // thread t < 128 owns level t: one packed warp scan gives the exclusive prefix of the map's level histogram (low
        // word: where level t's list starts) and of this CTA's counts (high word: where level t starts in the staging buffer)
        const int cl = tid < kLevels ? sLcnt[tid] : 0;
        const int hl = tid < kLevels ? P.lhist[(size_t)item * 256 + tid] : 0;
        const unsigned long long mine = ((unsigned long long)(unsigned)cl << 32) | (unsigned)hl;
        unsigned long long incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long nb = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += nb;
        }
        if (tid < kLevels && lane == 31) sWtot[tid >> 5] = incl;
        __syncthreads();
        if (tid < kLevels) {
            unsigned long long excl = incl - mine;
            for (int w = 0; w < (tid >> 5); ++w) excl += sWtot[w];
            sLloc[tid] = (int)(excl >> 32);
            if (cl) sLbase[tid] = (int)(unsigned)excl + atomicAdd(P.lhist + (size_t)item * 256 + kLevels + tid, cl);
        }
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < kBasinVPT; ++q)
        if (levq[q]) {
            const int i = sLloc[levq[q]] + rankq[q];
            if (kWeighted) {
                const int v = v_beg + q * 256 + tid;
                sStage[i] = (StageT)((unsigned)bas[q] | ((unsigned)(wrank ? wrank[v] : 0) << 16));
            } else {
                sStage[i] = (StageT)bas[q];
            }
            sStagePos[i] = sLbase[levq[q]] + rankq[q];
        }
    __syncthreads();
    const int total = sLloc[kLevels - 1] + sLcnt[kLevels - 1];
    StageT *__restrict__ vlist = reinterpret_cast<StageT *>(P.vlist) + base;
    for (int i = tid; i < total; i += 256) vlist[sStagePos[i]] = sStage[i];
}

// ------------------------------------------------------------------------------------------- K_D
// kDense (class path): table[level][basin] += 1 in global memory.  Otherwise (max-only maps) the (level, basin)
// counts of the CTA's 256 vertices are aggregated in a shared-memory hash and appended to the map's entry list
// {level, basin, vertices}; the same key may come from several CTAs -- the sweep simply adds them up.
static constexpr int kCountVPT = 4;                 // vertices per thread
static constexpr int kCountChunk = 256 * kCountVPT; // vertices per CTA

template <bool kDense>
__global__ void __launch_bounds__(256) pipe_count_kernel(PipeParams P, int chunks) {
    int item, chunk, s, b;
    grid_coords(P, item, chunk, s, b);
    const SurfDesc sd = P.surfs[s];
    constexpr int kStage = 2 * kCountChunk;
    __shared__ unsigned long long sPairs[kStage];
    __shared__ int sCnt, sBase;
    int *meta = P.meta + (size_t)item * 4;
    const int v_beg = chunk * kCountChunk;
    if (v_beg >= sd.V || meta[2]) return; // CTA-uniform
    if (threadIdx.x == 0) sCnt = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const size_t base = (size_t)item * P.vstride;
    const int *__restrict__ basin = P.basin + base;   // (pinned base pointer + ld.global.nc measured SLOWER here: 3.21 vs 2.68 ms)
    const int NB = meta[0];
    __shared__ int sBasin[kCountChunk]; // basins of the CTA's own vertices: most neighbour lookups land here
    int buq[kCountVPT], levq[kCountVPT];
    unsigned emq[kCountVPT];
#pragma unroll
    for (int q = 0; q < kCountVPT; ++q) {
        const int v = v_beg + q * 256 + threadIdx.x;
        buq[q] = (v < sd.V) ? basin[v] : -1;
        sBasin[q * 256 + threadIdx.x] = buq[q];
        levq[q] = (v < sd.V) ? (int)(P.lev8[base + v] & 0x7f) : 0;
        emq[q] = (v < sd.V) ? P.emask[base + v] : 0u;
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < kCountVPT; ++q) {
        const int v = v_beg + q * 256 + threadIdx.x;
        const int bu = buq[q], lev = levq[q];
        const unsigned em = bu >= 0 ? emq[q] : 0u;
        if (kDense) {
            // table[level][basin] += 1, one atomic per distinct (level, basin) of the warp: neighbouring vertices
            // mostly share it, and same-address atomics of one warp serialise in the L2
            const int key = bu >= 0 ? lev * NB + bu : -1 - lane;
            const unsigned peers = __match_any_sync(0xffffffffu, key);
            if (!(P.flags & 256) && bu >= 0 && lane == (__ffs(peers) - 1))
                atomicAdd(P.table + (size_t)item * P.tabcap + key, (unsigned)__popc(peers));
        }
        // candidate unions: earlier neighbours lying in another basin (each distinct basin once per vertex, best
        // effort).  Staged in shared memory: one returning atomic per CTA reserves the output range.
        if (em && !(P.flags & 512)) {
            const int *__restrict__ row = sd.ell + (size_t)v * sd.ell_width;
            int s0 = -1, s1 = -1, s2 = -1, s3 = -1;
            unsigned m = em;
            while (m) {
                const int j = __ffs(m) - 1;
                m &= m - 1;
                const int a = row[j];
                const unsigned off = (unsigned)(a - v_beg);
                const int ba = off < (unsigned)kCountChunk ? sBasin[off] : basin[a];
                if (ba == bu || ba == s0 || ba == s1 || ba == s2 || ba == s3) continue;
                s3 = s2; s2 = s1; s1 = s0; s0 = ba;
                const unsigned long long pr = ((unsigned long long)lev << 48) | ((unsigned long long)bu << 24) | (unsigned long long)ba;
                const int pos = atomicAdd(&sCnt, 1);
                if (pos < kStage) {
                    sPairs[pos] = pr;
                } else { // staging buffer full (cannot happen on meshes; kept for correctness)
                    const int gp = atomicAdd(meta + 1, 1);
                    if (gp < P.paircap) P.pairs[(size_t)item * P.paircap + gp] = pr;
                }
            }
        }
    }
    __syncthreads();
    // flush: one returning atomic per CTA
    const int n = min(sCnt, kStage);
    if (n == 0) return;
    if (threadIdx.x == 0) sBase = atomicAdd(meta + 1, n);
    __syncthreads();
    const int gbase = sBase;
    if (gbase + n <= P.paircap) { // else: K_S sees npairs > paircap and flags the map
        unsigned long long *__restrict__ dst = P.pairs + (size_t)item * P.paircap + gbase;
        for (int i = threadIdx.x; i < n; i += 256) dst[i] = sPairs[i];
    }
}


// ---- K_D, sliced rows.  With ~60 neighbours most vertices lie within reach of another basin, and one map would emit
// more candidate unions than it has vertices.  Only the EARLIEST union of two basins matters (later ones find them in
// one component already), so the CTA keeps one entry per unordered basin pair in a shared-memory hash -- key = the two
// basin ids, value = the smallest level seen (one 64-bit atomicMin: the level sits in the low bits) -- and flushes the
// few dozen distinct pairs of its 1,024 vertices with one returning atomic.  A full probe sequence falls back to a
// direct append (duplicates are harmless).
static constexpr int kPairHashBits = 11;
static constexpr int kPairHash = 1 << kPairHashBits;

template <bool kDense, int kWords>
__global__ void __launch_bounds__(256) pipe_count_wide_kernel(PipeParams P, int chunks) {
    int item, chunk, s, b;
    grid_coords(P, item, chunk, s, b);
    const SurfDesc sd = P.surfs[s];
    __shared__ unsigned long long sHash[kPairHash];
    __shared__ int sBasin[kCountChunk]; // basins of the CTA's own vertices: most neighbour lookups land here
    __shared__ int sWarpTot[8], sBase;
    int *meta = P.meta + (size_t)item * 4;
    const int v_beg = chunk * kCountChunk;
    if (v_beg >= sd.V || meta[2]) return; // CTA-uniform
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    for (int i = tid; i < kPairHash; i += 256) sHash[i] = ~0ull;
    const size_t base = (size_t)item * P.vstride;
    const int *__restrict__ basin = pin_ptr(P.basin + base);
    const int NB = meta[0];
    const unsigned *emw[kWords];
#pragma unroll
    for (int w = 0; w < kWords; ++w) emw[w] = pin_ptr(P.emask + ((size_t)item * kWords + w) * P.vstride);
    int buq[kCountVPT], levq[kCountVPT];
#pragma unroll
    for (int q = 0; q < kCountVPT; ++q) {
        const int v = v_beg + q * 256 + tid;
        buq[q] = (v < sd.V) ? basin[v] : -1;
        sBasin[q * 256 + tid] = buq[q];
        levq[q] = (v < sd.V) ? (int)(P.lev8[base + v] & 0x7f) : 0;
    }
    __syncthreads();
#pragma unroll 1
    for (int q = 0; q < kCountVPT; ++q) {
        const int v = v_beg + q * 256 + tid;
        const int bu = buq[q], lev = levq[q];
        if (kDense) {
            const int key = bu >= 0 ? lev * NB + bu : -1 - lane;
            const unsigned peers = __match_any_sync(0xffffffffu, key);
            if (bu >= 0 && lane == (__ffs(peers) - 1))
                atomicAdd(P.table + (size_t)item * P.tabcap + key, (unsigned)__popc(peers));
        }
        if (bu < 0) continue;
        const int o0 = sd.sell_off[v >> 5];
        const int *__restrict__ row = pin_ptr(reinterpret_cast<const int *>(sd.sell + o0 + lane));
        int s0 = -1, s1 = -1, s2 = -1, s3 = -1;
        // Every lane walks the SET bits of its own mask.  Walking the slot groups of the slice in lockstep instead (one
        // coalesced 16-byte load per group and lane, basin lookups only where the bit is set) was measured twice and is
        // slower (10.9 against 6.7 ms per 512 maps): the kernel is bound by the number of warp-level memory instructions
        // with few active lanes, and lockstep issues one predicated shared AND one predicated global lookup per slot.
        // Also measured and rejected: queueing the candidate unions per warp (ballot-compacted) and inserting 32 at a
        // time with every lane busy -- 26.5 against 24.5 ms per 1,024 maps for the stage: the walk already costs only
        // ~12 warp instructions per step (ncu: 2,260 per warp over 4 vertices per lane and ~45 steps each), the hash
        // path is not where the time goes, and the votes cost more than they save.  Upper bound for staging a wider
        // window of basin ids in shared memory: with every out-of-chunk lookup skipped outright (wrong results, timing
        // only) the stage drops from 25.1 to 22.7 ms per 1,024 maps; a 16-bit window of +-4,096 vertices would cover 91 %
        // of them (scripts/probe_locality.py) at the price of the staging loads -- not built.
#pragma unroll
        for (int w = 0; w < kWords; ++w) {
            unsigned m = __ldg(emw[w] + v);
            while (m) {
                const int j = w * 32 + __ffs(m) - 1;
                m &= m - 1;
                const int a = __ldg(row + (j >> 2) * 128 + (j & 3));
                const unsigned off = (unsigned)(a - v_beg);
                const int ba = off < (unsigned)kCountChunk ? sBasin[off] : __ldg(basin + (unsigned)a);
                if (ba == bu || ba == s0 || ba == s1 || ba == s2 || ba == s3) continue;
                s3 = s2; s2 = s1; s1 = s0; s0 = ba;
                const unsigned lo = (unsigned)min(bu, ba), hi = (unsigned)max(bu, ba);
                const unsigned long long val = ((unsigned long long)lo << 32) | ((unsigned long long)hi << 8) | (unsigned long long)lev;
                unsigned h = ((lo * 0x9E3779B1u) ^ (hi * 0x85EBCA6Bu)) >> (32 - kPairHashBits);
                bool placed = false;
                for (int probe = 0; probe < 16; ++probe) {
                    // plain look first: neighbouring vertices mostly find their pair already there with an earlier level
                    unsigned long long old = *reinterpret_cast<volatile unsigned long long *>(&sHash[h]);
                    if (old == ~0ull) old = atomicCAS(&sHash[h], ~0ull, val);
                    if (old == ~0ull) { placed = true; break; }
                    if ((old >> 8) == (val >> 8)) {
                        if (val < old) atomicMin(&sHash[h], val);
                        placed = true;
                        break;
                    }
                    h = (h + 1) & (kPairHash - 1);
                }
                if (!placed) { // table crowded: append directly
                    const int gp = atomicAdd(meta + 1, 1);
                    if (gp < P.paircap)
                        P.pairs[(size_t)item * P.paircap + gp] = ((unsigned long long)lev << 48) | ((unsigned long long)lo << 24) | (unsigned long long)hi;
                }
            }
        }
    }
    __syncthreads();
    // flush the distinct pairs: thread t owns slots [t * 8, t * 8 + 8)
    constexpr int kPer = kPairHash / 256;
    unsigned long long mine[kPer];
    int cnt = 0;
#pragma unroll
    for (int i = 0; i < kPer; ++i) {
        mine[i] = sHash[tid * kPer + i];
        cnt += mine[i] != ~0ull;
    }
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int nbv = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += nbv;
    }
    if (lane == 31) sWarpTot[wid] = incl;
    __syncthreads();
    int before = incl - cnt, total = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
        if (w < wid) before += sWarpTot[w];
        total += sWarpTot[w];
    }
    if (total == 0) return;
    if (tid == 0) sBase = atomicAdd(meta + 1, total);
    __syncthreads();
    const int gbase = sBase;
    if (gbase + total <= P.paircap) { // else: K_S sees npairs > paircap and flags the map
        unsigned long long *__restrict__ dst = P.pairs + (size_t)item * P.paircap + gbase + before;
        int k = 0;
#pragma unroll
        for (int i = 0; i < kPer; ++i)
            if (mine[i] != ~0ull) {
                const unsigned long long lev = mine[i] & 0xffull, lo = mine[i] >> 32, hi = (mine[i] >> 8) & 0xFFFFFFull;
                dst[k++] = (lev << 48) | (lo << 24) | hi;
            }
    }
}

// ------------------------------------------------------------------------------------------- K_S
// Per-slot scratch in global memory.
static constexpr int kPowSmem = 2048;
static constexpr int kSweepBasinCap = 12288; // (< 65536: the fold packs root ids in 16 bits) // basins whose state fits the sweep's shared memory (18 B each)

struct SweepSlot {
    unsigned long long *pairs2; // [paircap] candidate unions bucketed by level
    uint2 *elist;               // [Vmax] non-empty table entries {basin, vertices} bucketed by level
    int2 *cls;                  // [Vmax] class = {root at creation, creation level} -> {root, fp32 value bits}
    float *incseq;              // [min(nbcap, kSweepBasinCap)][128] increment of root r's component at level l
};

static inline size_t pipe_al256(size_t x) { return (x + 255) / 256 * 256; }

size_t pipe_slot_bytes(int32_t Vmax, int nbcap, int paircap) {
    const size_t nb = (size_t)(nbcap < kSweepBasinCap ? nbcap : kSweepBasinCap);
    return pipe_al256(sizeof(unsigned long long) * (size_t)paircap) + 2 * pipe_al256(sizeof(int2) * (size_t)Vmax) +
           pipe_al256(sizeof(float) * kLevels * nb);
}

__device__ __forceinline__ SweepSlot carve_slot(char *base, int32_t Vmax, int nbcap, int paircap) {
    auto al = [](size_t x) { return (x + 255) / 256 * 256; };
    SweepSlot w;
    w.pairs2 = reinterpret_cast<unsigned long long *>(base);
    base += al(sizeof(unsigned long long) * (size_t)paircap);
    w.elist = reinterpret_cast<uint2 *>(base);
    base += al(sizeof(int2) * (size_t)Vmax);
    w.cls = reinterpret_cast<int2 *>(base);
    base += al(sizeof(int2) * (size_t)Vmax);
    w.incseq = reinterpret_cast<float *>(base);
    return w;
}

// bytes of dynamic shared memory the sweep needs for NB basins
__host__ __device__ inline size_t pipe_sweep_smem(int NB, bool max_only) {
    const size_t nba = ((size_t)NB + 3) / 4 * 4;
    const size_t bits = (((size_t)NB + 31) / 32 * 4 + 15) / 16 * 16;
    if (max_only) return nba * (8 + 4 + 4 + 4 + 1 + 1) + 16;   // pow(size,E), parent, size, leader accumulator, hook level, peak level
    return nba * (8 + 4 + 4 + 4 + 4 + 1 + 1 + 8) + bits + 16; // + class of the level, hook parent, two fp32 row buffers
}

// The sweep over the basins of one map.  Per level l (three barrier intervals, shared memory only on the
// critical path; the table row and the candidate unions of level l+1 are fetched into registers during F3):
//   F1  unions of the level's candidate pairs (lock-free union-find; roots ordered by (level, id); the hook of
//       a root is logged: hookpar/hooklev keep the uncompressed merge history)
//   F2  sizes: table row l (vertices per basin at this level) added to the components' roots, sizes of older
//       roots hooked in this level carried over; one class {root, level} per component that gained vertices
//   F3  every live root logs this level's increment fl32(pow(size, E) * pow(T, H)) in incseq[root][l]
// After the sweep a class is folded on its own: its vertices receive, one fp32 add per level in
// descending-threshold order (fast_tfce.hpp:70-84), the logged increments of the root their component
// had at each level -- which follows from the merge history alone.
//
// kMaxOnly (no vertex weights, no maps requested): only the per-map maximum leaves the kernel, and that needs
// ONE accumulator per live root instead of one per class.  All increments are >= 0 and fp32 round-to-nearest
// addition is monotone, so among the classes of a component the one with the largest sum so far keeps the
// largest sum for ever (they all add the same increments from here on): the "leader".  A root's leader is its
// own first class (born with the root, at its peak); when roots merge the leader of the union is the larger of
// the two.  max_v TFCE(v) is therefore max over the final roots of the leader sums -- bit-identical to the
// maximum of the full map, with no per-class state at all.
template <int kThreads, int kMinBlocks, bool kMaxOnly>
__global__ void __launch_bounds__(kThreads, kMinBlocks) pipe_sweep_kernel(PipeParams P, int smem_bytes) {
    extern __shared__ __align__(16) unsigned char sDyn[];
    __shared__ double sHHd[2][kLevels];
    __shared__ int sPstart[kLevels + 1];
    __shared__ int sCursor[kLevels];
    __shared__ int sEstart[kLevels + 1];
    __shared__ int sEcursor[kLevels];
    __shared__ double sPow[kPowSmem];     // pow(n, E) for small n: most components are small
    __shared__ int sNs[2];
    __shared__ float sDelta[2];
    __shared__ int sItem;
    __shared__ int sC;
    __shared__ float sRed[2][kThreads / 32];

    const int tid = threadIdx.x;
    constexpr int nthr = kThreads;
    const int lane = tid & 31, wid = tid >> 5;
    const SweepSlot ws = carve_slot(P.slot_ws + (size_t)blockIdx.x * P.slot_stride, P.Vmax, P.nbcap, P.paircap);
    const int total_items = P.B * P.S;

    for (;;) {
        __syncthreads();
        if (tid == 0) sItem = atomicAdd(P.work_counter, 1);
        __syncthreads();
        const int item = sItem;
        if (item >= total_items) break;
        int s, b;
        item_coords(P, item, s, b);
        const SurfDesc sd = P.surfs[s];
        int *meta = P.meta + (size_t)item * 4;
        const int NB = meta[0], NP = meta[1];
        const size_t e0 = ((size_t)b * P.S + s) * 2;
        long long tk = P.timing ? clock64() : 0;
#define PIPE_TICK(i)                                                                 \
        if (P.timing && tid == 0) {                                                  \
            const long long now = clock64();                                         \
            atomicAdd(P.timing + (i), (unsigned long long)(now - tk));               \
            tk = now;                                                                \
        }
        if (meta[2] || NB > P.nbcap || NB > kSweepBasinCap || NP > P.paircap ||
            pipe_sweep_smem(NB, kMaxOnly) > (size_t)smem_bytes) {
            if (tid == 0) {
                meta[2] = 1; // redone by tfce_basin_kernel
                if (P.timing) atomicAdd(P.timing + 10, 1ull);
            }
            continue;
        }
        if (P.timing && tid == 0) {
            atomicAdd(P.timing + 11, (unsigned long long)NB);
            atomicAdd(P.timing + 12, (unsigned long long)NP);
            atomicAdd(P.timing + 13, 1ull);
        }
        // ---- shared-memory layout of the per-basin state
        const int nba = (NB + 3) / 4 * 4;
        const int bits_words = (((NB + 31) / 32 * 4 + 15) / 16 * 16) / 4;
        double *bpw = reinterpret_cast<double *>(sDyn); // pow(size, E) of the live roots, fetched asynchronously
        int *bparent = reinterpret_cast<int *>(bpw + nba);
        int *bsize = bparent + nba;
        int *bcur = bsize + nba;       // root -> class created for it in this level; kMaxOnly: leader sum (fp32 bits)
        int *hookpar = kMaxOnly ? bcur : bcur + nba; // merge history: the root this one was hooked under ... (not kMaxOnly)
        unsigned *bitsC = reinterpret_cast<unsigned *>(hookpar + nba);  // root got a class in this level (not kMaxOnly)
        unsigned char *hooklev = reinterpret_cast<unsigned char *>(kMaxOnly ? (unsigned *)(bcur + nba) : bitsC + bits_words); // ... and the level (255: never)
        unsigned char *blev = hooklev + nba;
        float *rowbuf = reinterpret_cast<float *>(blev + nba); // [2][nba] increments of one level (the fold; not kMaxOnly)
        int *racc = bcur;
        for (int i = tid; i < 2 * kLevels; i += nthr)
            sHHd[i / kLevels][i % kLevels] = (double)P.tab_HH[(e0 + i / kLevels) * kLevels + i % kLevels];
        if (tid < 2) {
            const bool on = (tid == 0) || P.two_sided;
            sNs[tid] = on ? P.tab_ns[e0 + tid] : 0;
            sDelta[tid] = P.tab_scale ? P.tab_scale[e0 + tid] : P.tab_delta[e0 + tid]; // factor of the scaled maximum
            if (P.status) P.status[e0 + tid] = on ? P.tab_status[e0 + tid] : 0;
        }
        if (tid == 0) sC = 0;
        for (int i = tid; i < kLevels; i += nthr) { sCursor[i] = 0; sEcursor[i] = 0; }
        for (int i = tid; i < kPowSmem; i += nthr) sPow[i] = (i <= sd.V) ? sd.powE[i] : 0.0;
        const unsigned char *__restrict__ gblev = P.blev + (size_t)item * P.nbcap;
        for (int i = tid; i < NB; i += nthr) {
            bparent[i] = i;
            bsize[i] = 0;
            if (kMaxOnly) {
                racc[i] = 0; // +0.0f
            } else {
                bcur[i] = -1;
                hookpar[i] = i;
            }
            hooklev[i] = 255;
            blev[i] = gblev[i];
        }
        if (!kMaxOnly)
            for (int i = tid; i < bits_words; i += nthr) bitsC[i] = 0u;
        __syncthreads();
        const int ns0 = sNs[0], ns1 = sNs[1];
        const int nlev = max(ns0, ns1);
        // ---- candidate unions bucketed by level (counting sort; order inside a level is irrelevant)
        const unsigned long long *__restrict__ pairs = P.pairs + (size_t)item * P.paircap;
        for (int i = tid; i < NP; i += nthr) atomicAdd(&sCursor[(int)(pairs[i] >> 48)], 1);
        // ---- the count table [level][basin] is ~90% zeros: compact it once into per-level lists {basin, vertices}.
        //      Column-wise traversal: a warp looks at 32 neighbouring basins of ONE level at a time (coalesced, the
        //      level is the loop index), eight levels in flight per thread.
        unsigned *__restrict__ tab = P.table + (size_t)item * P.tabcap;
        const int ncols = (NB + 31) & ~31; // warp-uniform column loop
        for (int col = tid; col < ncols; col += nthr) {
            for (int l0 = 1; l0 < nlev; l0 += 8) {
                unsigned cc[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) cc[q] = (col < NB && l0 + q < nlev) ? tab[(size_t)(l0 + q) * NB + col] : 0u;
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const unsigned nz = __ballot_sync(0xffffffffu, cc[q] != 0u);
                    if (nz && lane == 0) atomicAdd(&sEcursor[l0 + q], __popc(nz));
                }
            }
        }
        __syncthreads();
        if (tid == 0) {
            int run = 0;
            for (int l = 0; l < kLevels; ++l) { const int c = sCursor[l]; sPstart[l] = run; sCursor[l] = run; run += c; }
            sPstart[kLevels] = run;
        }
        if (tid == 32) {
            int run = 0;
            for (int l = 0; l < kLevels; ++l) { const int c = sEcursor[l]; sEstart[l] = run; sEcursor[l] = run; run += c; }
            sEstart[kLevels] = run;
        }
        __syncthreads();
        for (int i = tid; i < NP; i += nthr) {
            const unsigned long long p = pairs[i];
            ws.pairs2[atomicAdd(&sCursor[(int)(p >> 48)], 1)] = p;
        }
        for (int col = tid; col < ncols; col += nthr) {
            for (int l0 = 1; l0 < nlev; l0 += 8) {
                unsigned cc[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) cc[q] = (col < NB && l0 + q < nlev) ? tab[(size_t)(l0 + q) * NB + col] : 0u;
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const unsigned nz = __ballot_sync(0xffffffffu, cc[q] != 0u);
                    if (nz) {
                        int basepos = 0;
                        if (lane == 0) basepos = atomicAdd(&sEcursor[l0 + q], __popc(nz));
                        basepos = __shfl_sync(0xffffffffu, basepos, 0);
                        if (cc[q] != 0u) ws.elist[basepos + __popc(nz & ((1u << lane) - 1u))] = make_uint2((unsigned)col, cc[q]);
                    }
                }
            }
        }
        __syncthreads();
        PIPE_TICK(0)

        const double *__restrict__ powE = sd.powE;
        // register prefetch of the next level's inputs: one table entry and one candidate union per thread
        uint2 eq = make_uint2(0u, 0u);
        unsigned long long pq = 0ull;
        auto prefetch_level = [&](int l) {
            const int ie = sEstart[l] + tid;
            eq = (ie < sEstart[l + 1]) ? ws.elist[ie] : make_uint2(0u, 0u);
            const int i = sPstart[l] + tid;
            pq = (i < sPstart[l + 1]) ? ws.pairs2[i] : 0ull;
        };
        // increments of level l for the roots that were live at l (derivable from the hook log at any later time);
        // their pow(size, E) was fetched into bpw by the same thread during F3 of level l
        auto consume_level = [&](int l) {
            cp_async_wait_all();
            for (int bb = tid; bb < NB; bb += nthr) {
                const int cb = blev[bb];
                const int sg = cb >> 7;
                if ((cb & 0x7f) <= l && hooklev[bb] > l && l < (sg ? ns1 : ns0)) {
                    const float inc = __double2float_rn(__dmul_rn(bpw[bb], sHHd[sg][l]));
                    if (kMaxOnly) racc[bb] = __float_as_int(__fadd_rn(__int_as_float(racc[bb]), inc));
                    else ws.incseq[(size_t)l * nba + bb] = inc;
                }
            }
        };
        if (nlev > 1) prefetch_level(1);
        int nclass = 0;
        for (int lev = 1; lev < nlev; ++lev) {
            // ================= F1: unions of this level =================================================
            for (int i = sPstart[lev] + tid; i < sPstart[lev + 1]; i += nthr) {
                const unsigned long long p = (i == sPstart[lev] + tid && !(P.flags & 1024)) ? pq : ws.pairs2[i];
                int ru = pf_find(bparent, (int)((p >> 24) & 0xFFFFFFu));
                int ra = pf_find(bparent, (int)(p & 0xFFFFFFu));
                while (ru != ra) {
                    // total order on roots: (activation level, id); the later root goes under the earlier one
                    const int kru = ((blev[ru] & 0x7f) << 24) | ru, kra = ((blev[ra] & 0x7f) << 24) | ra;
                    const int hi = kru > kra ? ru : ra, lo = kru > kra ? ra : ru;
                    const int old = atomicCAS(bparent + hi, hi, lo);
                    if (old == hi) { // this thread hooked hi: log it
                        if (!kMaxOnly) hookpar[hi] = lo;
                        hooklev[hi] = (unsigned char)lev;
                        break;
                    }
                    const int nh = pf_find(bparent, old);
                    if (hi == ru) { ru = nh; ra = pf_find(bparent, ra); }
                    else          { ra = nh; ru = pf_find(bparent, ru); }
                }
            }
            __syncthreads();
            PIPE_TICK(3)
            // ================= F2a: the previous level's increments (its pow(size, E) fetches have landed) =
            if (lev > 1) consume_level(lev - 1);
            if (kMaxOnly) __syncthreads(); // leaders of level lev-1 final before the hand-over below
            // ================= F2b: sizes; one new class per component that gains vertices ================
            unsigned *__restrict__ row = tab + (size_t)lev * NB;
            {
                const int ebeg = sEstart[lev], eend = sEstart[lev + 1];
                for (int ew = ebeg + (tid & ~31); ew < eend; ew += nthr) { // warp-uniform trip counts (ballots below)
                    const int ie = ew + lane;
                    uint2 ent = make_uint2(0u, 0u);
                    if (ie < eend) ent = (ew == ebeg + (tid & ~31) && !(P.flags & 1024)) ? eq : ws.elist[ie];
                    const unsigned c = ent.y;
                    int r = -1;
                    if (c != 0u) r = pf_find(bparent, (int)ent.x);
                    // Late levels send most basins to a few giant roots: same-address shared-memory atomics would
                    // serialise.  The lanes that agree with the first active lane's root are summed by one redux.
                    const unsigned am = __ballot_sync(0xffffffffu, c != 0u);
                    if (am) {
                        const int lead = __ffs(am) - 1;
                        const int r0 = __shfl_sync(0xffffffffu, r, lead);
                        const bool same = (c != 0u) && r == r0;
                        const int sum = __reduce_add_sync(0xffffffffu, same ? (int)c : 0);
                        const bool own = (c != 0u) && !same; // a different root: on its own
                        if (lane == lead) atomicAdd(bsize + r0, sum);
                        if (own) atomicAdd(bsize + r, (int)c);
                        if (!kMaxOnly && (lane == lead || own)) {
                            const unsigned bit = 1u << (r & 31);
                            if (!(atomicOr(bitsC + (r >> 5), bit) & bit)) {
                                const int j = atomicAdd(&sC, 1);
                                ws.cls[j] = make_int2(r, lev);
                                bcur[r] = j;
                            }
                        }
                    }
                }
                // older roots hooked in this level hand their size (and their leader) over; a root of this very level
                // has neither yet
                for (int bb = tid; bb < NB; bb += nthr)
                    if (hooklev[bb] == lev && (blev[bb] & 0x7f) != lev) {
                        const int rr = pf_find(bparent, bb);
                        atomicAdd(bsize + rr, bsize[bb]);
                        if (kMaxOnly) atomicMax(racc + rr, racc[bb]); // sums are >= 0: integer order == float order
                    }
            }
            __syncthreads();
            PIPE_TICK(4)
            // ================= F3: fetch pow(size, E) of every live root (asynchronously) ================
            if (lev + 1 < nlev && !(P.flags & 1024)) prefetch_level(lev + 1);
            for (int bb = tid; bb < NB; bb += nthr) {
                const int cb = blev[bb];
                if (bparent[bb] == bb && (cb & 0x7f) <= lev && lev < ((cb >> 7) ? ns1 : ns0)) {
                    const int sz = bsize[bb];
                    if (sz < kPowSmem) bpw[bb] = sPow[sz];
                    else cp_async8(bpw + bb, powE + sz); // the few large components: asynchronous, consumed next level
                }
            }
            if (!kMaxOnly) {
                if (P.want_vertex_pass)
                    for (int ie = sEstart[lev] + tid; ie < sEstart[lev + 1]; ie += nthr) {
                        const int bb = (int)ws.elist[ie].x;
                        row[bb] = 0x80000000u | (unsigned)bcur[pf_find(bparent, bb)];
                    }
                for (int i = tid; i < bits_words; i += nthr) bitsC[i] = 0u;
            }
            __syncthreads();
            PIPE_TICK(5)
        }
        if (nlev > 1) consume_level(nlev - 1);
        __syncthreads();
        nclass = kMaxOnly ? 0 : sC;
        // ---- fold: the value of a class = the increments of its component in level order, following the merge
        // history.  Level-synchronous with the accumulators in REGISTERS: kFold classes per thread and pass; per
        // level the increments of all roots (one coalesced row) are staged in shared memory, so a (class, level)
        // step is two shared-memory loads and one fp32 add.  Class ids ascend with the creation level: the classes
        // of one pass start at about the same level and earlier levels are skipped.
        const float d0 = sDelta[0], d1 = sDelta[1];
        float m0 = 0.f, m1 = 0.f;
        constexpr int kFold = 16;
        if (kMaxOnly) {
            for (int bb = tid; bb < NB; bb += nthr)
                if (bparent[bb] == bb) { // final roots carry the leaders
                    const int sg = blev[bb] >> 7;
                    const float sc = __fmul_rn(__int_as_float(racc[bb]), sg ? d1 : d0);
                    if (sg) m1 = fmaxf(m1, sc); else m0 = fmaxf(m0, sc);
                }
        }
        for (int cbase = 0; !kMaxOnly && cbase < nclass; cbase += kFold * nthr) {
            int pk[kFold];    // root (16 bits) | creation level << 16 (7 bits) | last level << 23 (7 bits); -1 = no class
            float ac[kFold];
#pragma unroll
            for (int k = 0; k < kFold; ++k) {
                const int j = cbase + k * nthr + tid;
                pk[k] = -1;
                ac[k] = 0.f;
                if (j < nclass) {
                    const int2 rec = ws.cls[j];
                    const int sg = blev[rec.x] >> 7;
                    pk[k] = rec.x | (rec.y << 16) | (((sg ? ns1 : ns0) - 1) << 23);
                }
            }
            const int lmin = ws.cls[cbase].y; // creation level of the pass's first (= earliest) class
            __syncthreads();
            for (int i = tid; i < NB; i += nthr) rowbuf[(lmin & 1) * nba + i] = ws.incseq[(size_t)lmin * nba + i];
            __syncthreads();
            for (int l = lmin; l < nlev; ++l) {
                float nxt[8];
                const bool more = l + 1 < nlev;
                if (more) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) nxt[q] = (tid + q * nthr < NB) ? ws.incseq[(size_t)(l + 1) * nba + tid + q * nthr] : 0.f;
                }
                const float *__restrict__ cur = rowbuf + (l & 1) * nba;
#pragma unroll
                for (int k = 0; k < kFold; ++k) {
                    const int w = pk[k];
                    if (w >= 0 && l >= ((w >> 16) & 0x7f) && l <= ((w >> 23) & 0x7f)) {
                        int r = w & 0xffff;
                        while (hooklev[r] <= l) r = hookpar[r];
                        pk[k] = (w & ~0xffff) | r;
                        ac[k] = __fadd_rn(ac[k], cur[r]);
                    }
                }
                if (more) {
                    float *__restrict__ dstb = rowbuf + ((l + 1) & 1) * nba;
#pragma unroll
                    for (int q = 0; q < 8; ++q)
                        if (tid + q * nthr < NB) dstb[tid + q * nthr] = nxt[q];
                    for (int i = tid + 8 * nthr; i < NB; i += nthr) dstb[i] = ws.incseq[(size_t)(l + 1) * nba + i];
                }
                __syncthreads();
            }
#pragma unroll
            for (int k = 0; k < kFold; ++k) {
                const int j = cbase + k * nthr + tid;
                if (pk[k] >= 0) {
                    if (P.want_vertex_pass) ws.cls[j].y = __float_as_int(ac[k]);
                    const int sg = blev[pk[k] & 0xffff] >> 7;
                    const float sc = __fmul_rn(ac[k], sg ? d1 : d0);
                    if (sg) m1 = fmaxf(m1, sc); else m0 = fmaxf(m0, sc);
                }
            }
        }
        PIPE_TICK(6)
        // ---- outputs
        if (!kMaxOnly && P.want_vertex_pass) {
            // class ids -> values in the table; the scaled maxima are taken per vertex by K_G (vertex weights)
            __syncthreads();
            const int64_t n = (int64_t)nlev * NB;
            for (int64_t i = NB + tid; i < n; i += nthr) {
                const unsigned e = tab[i];
                if (e & 0x80000000u) tab[i] = (unsigned)ws.cls[e & 0x7fffffffu].y;
            }
            if (tid < 2 && P.max_out) P.max_out[e0 + tid] = 0.f;
        } else {
            m0 = pwarp_max(m0);
            m1 = pwarp_max(m1);
            if (lane == 0) { sRed[0][wid] = m0; sRed[1][wid] = m1; }
            __syncthreads();
            if (tid == 0 && P.max_out) {
                float a = 0.f, c = 0.f;
                for (int w = 0; w < nthr / 32; ++w) { a = fmaxf(a, sRed[0][w]); c = fmaxf(c, sRed[1][w]); }
                P.max_out[e0] = a;
                P.max_out[e0 + 1] = c;
            }
        }
        PIPE_TICK(2)
        if (P.timing && tid == 0) atomicAdd(P.timing + 7, (unsigned long long)nclass);
#undef PIPE_TICK
    }
}

// ------------------------------------------------------------------------------------------- K_S (max-only maps)
// No vertex weights and no maps requested: only the per-map maximum leaves the kernel, and that needs ONE
// accumulator per live root instead of one per class.  All increments are >= 0 and fp32 round-to-nearest
// addition is monotone, so among the classes of a component the one with the largest sum so far keeps the
// largest sum for ever (they all add the same increments from here on): the "leader".  A root's leader is its
// own first class (born with the root, at its peak); when roots merge the leader of the union is the larger of
// the two.  max_v TFCE(v) is therefore the maximum over the final roots of the leader sums -- bit-identical
// to the maximum of the full map, with no per-class state at all.
//
// Per level only LIVE roots are visited: a compact list (double-buffered, rebuilt in passing every level)
// instead of a dense loop over all basins; roots enter it at the level of their peak (basins bucketed by
// birth level) and leave it when they are hooked.  pow(size, E) comes from a shared-memory copy of the table
// for small components; the few large ones are fetched asynchronously (cp.async) and added one barrier later.
// Two geometries: one 1,024-thread CTA per SM with everything in shared memory, or (kSmall) two 512-thread CTAs per
// SM that hide each other's barrier intervals -- then the root lists live in global memory (L2) and the pow table and
// the pending slots are halved, so that 14 bytes per basin fit twice into an SM.  Maps too large for the small
// geometry are marked (meta[2] = 2) and taken by a second launch of the large one.
__host__ __device__ inline size_t pipe_sweep_max_smem_bytes(int NB, bool small, bool weighted) {
    const size_t nba = ((size_t)NB + 7) / 8 * 8;
    // weighted: + wnew (4), w0 (2), fcnt (1); its live-root lists always sit in the slot's global scratch
    return nba * (4 + 4 + 4 + 1 + 1 + ((small || weighted) ? 0 : 6) + (weighted ? 7 : 0)) + 16;
}

// Weighted maxima (kW).  The scaled maximum of pyfunc.py:116-117 / tm_func.py:173-174 is max_v fl32(fl32(tfce_v * delta) * w_v).
// All vertices of one class (component, activation level) share tfce_v, so only the class's largest weight matters; and
// among the classes of one component, class a can be dropped for ever once another class b has sum_b >= sum_a AND
// w_b >= w_a: both receive the same fp32 increments from now on (round-to-nearest addition is monotone) and the product
// is monotone in the sum and in the weight (weights >= 0).  Each live root therefore carries the Pareto front of its
// classes' (sum, weight rank) pairs -- a handful of entries: a late class of a big component sees thousands of new
// vertices, hence nearly the largest weight, and dominates everything younger.  Entry 0 lives in shared memory, further
// entries in the slot's global scratch.  The front of a root is only ever written by the thread that owns the root in
// F3 (merging the fronts of the roots hooked under it this level, then the new class, then the increment), so no locks.
static constexpr int kFront = 16;           // entries per root; a fuller front flags the map for the one-kernel sweep

template <bool kW>
__device__ __forceinline__ void ld_ent_if(unsigned &dst, const void *base, int idx, bool ok) {
    if (kW)
        asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\t@q ld.global.u32 %0, [%1];\n\t}" : "+r"(dst) : "l"(reinterpret_cast<const unsigned *>(base) + idx), "r"((int)ok) : "memory");
    else
        asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\t@q ld.global.u16 %0, [%1];\n\t}" : "+r"(dst) : "l"(reinterpret_cast<const unsigned short *>(base) + idx), "r"((int)ok) : "memory");
}
// + per-warp level histograms of the vertex sort: (threads / 32) * 128 ints, added by the launcher

template <int kThreads, int kMinBlocks, bool kSmall, bool kW>
__global__ void __launch_bounds__(kThreads, kMinBlocks) pipe_sweep_max_kernel(PipeParams P, int smem_bytes, int stage) {
    constexpr int kPowN = kSmall ? 1024 : 2048; // pow(n, E) entries kept in shared memory
    constexpr int kPend = kSmall ? 256 : 512;   // asynchronous pow fetches in flight per level
    extern __shared__ __align__(16) unsigned char sDyn[];
    __shared__ double sHHd[2][kLevels];
    __shared__ double sPow[kPowN];
    __shared__ double sPendPw[kPend];
    __shared__ int sPendBb[kPend];
    __shared__ int sPstart[kLevels + 1], sEstart[kLevels + 1], sBstart[kLevels + 1];
    __shared__ int sCurP[kLevels], sCurE[kLevels], sCurB[kLevels];
    __shared__ int sNs[2];
    __shared__ float sDelta[2];
    __shared__ int sItem, sNpend, sNalive[2], sOverflow;
    constexpr int kMergeCap = kW ? 256 : 1;   // weighted: (hooked root | new root << 16) pairs whose fronts the new root must
    __shared__ unsigned sMerge[2][kMergeCap]; // merge this level; double-buffered by level parity, rarely more than a few
    __shared__ int sMergeN[2];
    __shared__ float sRed[2][kThreads / 32];

    const int tid = threadIdx.x;
    constexpr int nthr = kThreads;
    const int lane = tid & 31, wid = tid >> 5;
    const SweepSlot ws = carve_slot(P.slot_ws + (size_t)blockIdx.x * P.slot_stride, P.Vmax, P.nbcap, P.paircap);
    const int total_items = P.B * P.S;

    for (;;) {
        __syncthreads();
        if (tid == 0) { sItem = atomicAdd(P.work_counter, 1); sOverflow = 0; }
        __syncthreads();
        const int item = sItem;
        if (item >= total_items) break;
        int s, b;
        item_coords(P, item, s, b);
        const SurfDesc sd = P.surfs[s];
        int *meta = P.meta + (size_t)item * 4;
        const int NB = meta[0], NP = meta[1];
        const size_t e0 = ((size_t)b * P.S + s) * 2;
        long long tk = P.timing ? clock64() : 0;
#define PIPE_TICK(i)                                                                 \
        if (P.timing && tid == 0) {                                                  \
            const long long now = clock64();                                         \
            atomicAdd(P.timing + (i), (unsigned long long)(now - tk));               \
            tk = now;                                                                \
        }
        // stage 0: the only sweep launch; 1: small geometry first (maps it cannot hold are marked 2); 2: the large
        // geometry takes the maps marked 2.  meta[2] == 1 always means "redo with tfce_basin_kernel".
        const int flag = meta[2];
        if (stage == 2 ? flag != 2 : flag != 0) continue;
        const bool hopeless = NB > P.nbcap || NB > 65534 || NP > P.paircap || (kW && NB > kSweepBasinCap);
        const bool too_big = pipe_sweep_max_smem_bytes(NB, kSmall, kW) > (size_t)smem_bytes;
        if (hopeless || too_big) {
            __syncthreads(); // everybody has read the flag
            if (tid == 0) {
                meta[2] = (!hopeless && stage == 1) ? 2 : 1;
                if (P.timing && (hopeless || stage != 1)) atomicAdd(P.timing + 10, 1ull);
            }
            continue;
        }
        if (stage == 2) {
            __syncthreads();
            if (tid == 0) meta[2] = 0;
        }
        if (P.timing && tid == 0) {
            atomicAdd(P.timing + 11, (unsigned long long)NB);
            atomicAdd(P.timing + 12, (unsigned long long)NP);
            atomicAdd(P.timing + 13, 1ull);
        }
        // ---- shared-memory layout of the per-basin state
        const int nba = (NB + 7) / 8 * 8;
        int *bparent = reinterpret_cast<int *>(sDyn);
        int *bsize = bparent + nba;
        int *racc = bsize + nba;                                                  // leader sum of a root (fp32 bits)
        unsigned short *alive[2];                                                 // live roots (double-buffered) ...
        unsigned short *birth;                                                    // ... and basins bucketed by the level of their peak
        unsigned char *hooklev;                                                   // level at which a root was hooked (255: never)
        constexpr bool kListsGlobal = kSmall || kW;
        if (kListsGlobal) { // lists in the slot's global scratch (the class path's increment log, unused here)
            alive[0] = reinterpret_cast<unsigned short *>(ws.incseq);
            alive[1] = alive[0] + nba;
            birth = alive[1] + nba;
            hooklev = reinterpret_cast<unsigned char *>(racc + nba);
        } else {
            alive[0] = reinterpret_cast<unsigned short *>(racc + nba);
            alive[1] = alive[0] + nba;
            birth = alive[1] + nba;
            hooklev = reinterpret_cast<unsigned char *>(birth + nba);
        }
        unsigned char *blev = hooklev + nba;                                      // level | sign << 7 of the peak
        // weighted: Pareto front of (sum, weight rank) per root -- entry 0 in racc / w0, entries 1.. in global scratch
        int *wnew = reinterpret_cast<int *>(blev + nba);                          // largest weight rank + 1 among the vertices a root gains this level
        unsigned short *w0 = reinterpret_cast<unsigned short *>(wnew + nba);
        unsigned char *fcnt = reinterpret_cast<unsigned char *>(w0 + nba);        // entries in the root's front
        uint2 *fr = reinterpret_cast<uint2 *>(reinterpret_cast<char *>(ws.incseq) + (kListsGlobal ? ((size_t)6 * nba + 15) / 16 * 16 : 0));
        auto fget = [&](int r, int k) -> uint2 {
            return k == 0 ? make_uint2((unsigned)racc[r], (unsigned)w0[r]) : fr[(size_t)r * kFront + k];
        };
        auto fset = [&](int r, int k, uint2 e) {
            if (k == 0) { racc[r] = (int)e.x; w0[r] = (unsigned short)e.y; }
            else fr[(size_t)r * kFront + k] = e;
        };
        // add class (sum bits s, weight rank w) to the front of root r (sums are >= 0: integer order == float order)
        auto finsert = [&](int r, unsigned s_, unsigned w_) {
            int c = fcnt[r];
            for (int k = 0; k < c; ++k) {
                const uint2 e = fget(r, k);
                if (e.x >= s_ && e.y >= w_) return;      // dominated by an existing class
            }
            for (int k = 0; k < c;) {                    // drop the classes it dominates
                const uint2 e = fget(r, k);
                if (e.x <= s_ && e.y <= w_) { --c; if (k < c) fset(r, k, fget(r, c)); }
                else ++k;
            }
            if (c >= P.front_cap) { sOverflow = 1; fcnt[r] = (unsigned char)c; return; }
            fset(r, c, make_uint2(s_, w_));
            fcnt[r] = (unsigned char)(c + 1);
        };
        auto fadd_all = [&](int r, float inc) {         // one level's increment to every class of the front
            racc[r] = __float_as_int(__fadd_rn(__int_as_float(racc[r]), inc));
            if (kW) {
                const int c = fcnt[r];
                for (int k = 1; k < c; ++k) {
                    uint2 e = fr[(size_t)r * kFront + k];
                    e.x = __float_as_uint(__fadd_rn(__uint_as_float(e.x), inc));
                    fr[(size_t)r * kFront + k] = e;
                }
            }
        };
        // basin of every active vertex, bucketed by level: built by K_C (see pipe_basin_kernel)
        const void *__restrict__ elist = kW ? static_cast<const void *>(reinterpret_cast<const unsigned *>(P.vlist) + (size_t)item * P.vstride)
                                            : static_cast<const void *>(P.vlist + (size_t)item * P.vstride);

        for (int i = tid; i < 2 * kLevels; i += nthr)
            sHHd[i / kLevels][i % kLevels] = (double)P.tab_HH[(e0 + i / kLevels) * kLevels + i % kLevels];
        for (int i = tid; i < kPowN; i += nthr) sPow[i] = (i <= sd.V) ? sd.powE[i] : 0.0;
        if (tid < 2) {
            const bool on = (tid == 0) || P.two_sided;
            sNs[tid] = on ? P.tab_ns[e0 + tid] : 0;
            sDelta[tid] = P.tab_scale ? P.tab_scale[e0 + tid] : P.tab_delta[e0 + tid]; // factor of the scaled maximum
            if (P.status) P.status[e0 + tid] = on ? P.tab_status[e0 + tid] : 0;
            sNalive[tid] = 0;
        }
        if (tid == 0) { sNpend = 0; sMergeN[0] = 0; sMergeN[1] = 0; }
        for (int i = tid; i < kLevels; i += nthr) { sCurP[i] = 0; sCurE[i] = 0; sCurB[i] = 0; }
        const unsigned char *__restrict__ gblev = P.blev + (size_t)item * P.nbcap;
        __syncthreads();
        for (int i = tid; i < NB; i += nthr) {
            bparent[i] = i;
            bsize[i] = 0;
            racc[i] = 0; // +0.0f
            hooklev[i] = 255;
            if (kW) { wnew[i] = 0; fcnt[i] = 0; }
            const int cb = gblev[i];
            blev[i] = (unsigned char)cb;
            atomicAdd(&sCurB[cb & 0x7f], 1);
        }
        // ---- candidate unions, basins (by the level of their peak) and VERTICES (by their own level) bucketed by
        //      level: counting sorts, order inside a level is irrelevant.  The vertex sort uses per-warp private
        //      histograms (no contention between warps); warp w owns the vertex blocks w, w + nwarps, ... in BOTH
        //      passes, 128 vertices per block (one 4-byte load of level codes per lane).
        const unsigned long long *__restrict__ pairs = P.pairs + (size_t)item * P.paircap;
        auto for_each_batched = [&](const unsigned long long *__restrict__ src, int n, auto &&fn) {
            for (int i0 = tid; i0 < n; i0 += 8 * nthr) {
                unsigned long long v[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) v[q] = ld_u64_if(src + i0 + q * nthr, i0 + q * nthr < n);
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    if (i0 + q * nthr < n) fn(v[q]);
            }
        };
        PIPE_TICK(20)
        for_each_batched(pairs, NP, [&](unsigned long long p) { atomicAdd(&sCurP[(int)(p >> 48)], 1); });
        __syncthreads();
        PIPE_TICK(21)
        // vertices per level: the map's histogram from K_A (the lists themselves were written by K_C)
        if (tid < kLevels) sCurE[tid] = P.lhist[(size_t)item * 256 + tid];
        __syncthreads();
        PIPE_TICK(22)
        if (tid < 3) {
            int *cur = tid == 0 ? sCurP : tid == 1 ? sCurE : sCurB;
            int *start = tid == 0 ? sPstart : tid == 1 ? sEstart : sBstart;
            int run = 0;
            for (int l = 0; l < kLevels; ++l) { const int c = cur[l]; start[l] = run; cur[l] = run; run += c; }
            start[kLevels] = run;
        }
        __syncthreads();
        PIPE_TICK(23)
        for_each_batched(pairs, NP, [&](unsigned long long p) { ws.pairs2[atomicAdd(&sCurP[(int)(p >> 48)], 1)] = p; });
        __syncthreads();
        PIPE_TICK(24)
        for (int i = tid; i < NB; i += nthr) birth[atomicAdd(&sCurB[blev[i] & 0x7f], 1)] = (unsigned short)i;
        __syncthreads();
        const int ns0 = sNs[0], ns1 = sNs[1];
        const int nlev = max(ns0, ns1);
        PIPE_TICK(0)

        const double *__restrict__ powE = sd.powE;
        // register pipeline of the coming levels' candidate unions, two levels deep (one per thread and level); the loads
        // go straight into their destination registers: first use two levels later
        unsigned long long pq0 = 0ull, pq1 = 0ull;
        auto prefetch_pairs = [&](int l) {
            if (l >= nlev) return; // block-uniform
            const int i = sPstart[l] + tid;
            const bool ok = i < sPstart[l + 1];
            if (l & 1) asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\t@q ld.global.u64 %0, [%1];\n\t}" : "+l"(pq1) : "l"(ws.pairs2 + i), "r"((int)ok) : "memory");
            else       asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\t@q ld.global.u64 %0, [%1];\n\t}" : "+l"(pq0) : "l"(ws.pairs2 + i), "r"((int)ok) : "memory");
        };
        // the same for the level's vertex list: four entries per thread (basins of vertices sEstart[l] + q * nthr + tid)
        unsigned eqA[4] = {0u, 0u, 0u, 0u}, eqB[4] = {0u, 0u, 0u, 0u};
        auto prefetch_entry = [&](int l) {
            if (l >= nlev) return; // block-uniform
            const int ie = sEstart[l] + tid, iend = sEstart[l + 1];
            if (l & 1) {
                ld_ent_if<kW>(eqB[0], elist, ie, ie < iend);
                ld_ent_if<kW>(eqB[1], elist, ie + nthr, ie + nthr < iend);
                ld_ent_if<kW>(eqB[2], elist, ie + 2 * nthr, ie + 2 * nthr < iend);
                ld_ent_if<kW>(eqB[3], elist, ie + 3 * nthr, ie + 3 * nthr < iend);
            } else {
                ld_ent_if<kW>(eqA[0], elist, ie, ie < iend);
                ld_ent_if<kW>(eqA[1], elist, ie + nthr, ie + nthr < iend);
                ld_ent_if<kW>(eqA[2], elist, ie + 2 * nthr, ie + 2 * nthr < iend);
                ld_ent_if<kW>(eqA[3], elist, ie + 3 * nthr, ie + 3 * nthr < iend);
            }
        };
        prefetch_pairs(1);
        prefetch_pairs(2);
        prefetch_entry(1);
        prefetch_entry(2);
        int cur = 0; // which alive list is current
        // ================= F1: unions of a level =========================================================
        auto do_unions = [&](int lev) {
            for (int i = sPstart[lev] + tid; i < sPstart[lev + 1]; i += nthr) {
                const unsigned long long p = (i == sPstart[lev] + tid) ? ((lev & 1) ? pq1 : pq0) : ws.pairs2[i];
                int ru = pf_find(bparent, (int)((p >> 24) & 0xFFFFFFu));
                int ra = pf_find(bparent, (int)(p & 0xFFFFFFu));
                while (ru != ra) {
                    // total order on roots: (activation level, id); the later root goes under the earlier one
                    const int kru = ((blev[ru] & 0x7f) << 24) | ru, kra = ((blev[ra] & 0x7f) << 24) | ra;
                    const int hi = kru > kra ? ru : ra, lo = kru > kra ? ra : ru;
                    const int old = atomicCAS(bparent + hi, hi, lo);
                    if (old == hi) { hooklev[hi] = (unsigned char)lev; break; } // this thread hooked hi
                    const int nh = pf_find(bparent, old);
                    if (hi == ru) { ru = nh; ra = pf_find(bparent, ra); }
                    else          { ra = nh; ru = pf_find(bparent, ru); }
                }
            }
        };
        // Two barrier intervals per level: [F3 of level l-1 + F1 of level l] | [F2 of level l].  F3 decides "live at
        // level l-1" from the hook log (hooklev > l-1), which the concurrent unions of level l cannot invalidate.
        if (nlev > 1) do_unions(1);
        __syncthreads();
        PIPE_TICK(3)
        for (int lev = 1; lev < nlev; ++lev) {
            // ================= F2a: previous level's increments of the large components ===================
            const int npend = sNpend;
            if (npend > 0) { // block-uniform
                for (int i = tid; i < npend; i += nthr) {
                    const int bb = sPendBb[i];
                    const float inc = __double2float_rn(__dmul_rn(sPendPw[i], sHHd[blev[bb] >> 7][lev - 1]));
                    fadd_all(bb, inc);
                }
                __syncthreads();
                if (tid == 0) sNpend = 0;
            }
            // ================= F2b: sizes =================================================================
            {
                // the small geometry keeps the live lists in global memory (L2): the first entry of the hooked-roots pass
                // below is requested now, so that its latency passes behind the vertex loop
                const int na2 = sNalive[cur];
                int bbpre = -1;
                if (kSmall && tid < na2) bbpre = (int)__ldcg(alive[cur] + tid);
                const int ebeg = sEstart[lev], eend = sEstart[lev + 1];
                // Passes of nthr entries, four at a time: the first four entries of a thread were prefetched two levels ago,
                // later groups are loaded together (four independent L2 loads in flight instead of one exposed load per
                // pass -- the dense late levels have ten and more passes).
                for (int eg = ebeg + (tid & ~31), grp = 0; eg < eend; eg += 4 * nthr, ++grp) { // warp-uniform trip counts
                    unsigned e4[4];
                    if (grp == 0) {
#pragma unroll
                        for (int q = 0; q < 4; ++q) e4[q] = (lev & 1) ? eqB[q] : eqA[q];
                    } else {
#pragma unroll
                        for (int q = 0; q < 4; ++q) { e4[q] = 0u; ld_ent_if<kW>(e4[q], elist, eg + q * nthr + lane, eg + q * nthr + lane < eend); }
                    }
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int ew = eg + q * nthr;
                        if (ew >= eend) break; // warp-uniform
                        const bool act = ew + lane < eend;
                        int r = -1;
                        if (act) r = pf_find(bparent, kW ? (int)(e4[q] & 0xffffu) : (int)e4[q]);
                        // Late levels send most vertices to a few giant roots: same-address shared-memory atomics would
                        // serialise.  The lanes that agree with the first active lane's root are counted by one ballot.
                        // (A second round for the other sign's giant root was measured slower: 8.07 vs 7.87 ms per 1,024 maps.)
                        const unsigned am = __ballot_sync(0xffffffffu, act);
                        const int lead = __ffs(am) - 1;
                        const int r0 = __shfl_sync(0xffffffffu, r, lead);
                        const unsigned same = __ballot_sync(0xffffffffu, act && r == r0);
                        if (lane == lead) atomicAdd(bsize + r0, __popc(same));
                        if (act && r != r0) atomicAdd(bsize + r, 1);
                        if (kW) { // largest weight among the vertices every root gains at this level
                            const int wr1 = act ? (int)(e4[q] >> 16) + 1 : 0;
                            const int mw = __reduce_max_sync(0xffffffffu, (act && r == r0) ? wr1 : 0);
                            if (lane == lead) atomicMax(wnew + r0, mw);
                            if (act && r != r0) atomicMax(wnew + r, wr1);
                        }
                    }
                }
                // older roots hooked in this level hand their size and their leader over (a root of this very level has
                // neither yet); they are still on the live list of the previous level
                const unsigned short *__restrict__ al = alive[cur];
                const int na = na2;
                for (int i0 = tid; i0 < na; i0 += 4 * nthr) { // four list entries in flight per thread
                    int b4[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int i = i0 + q * nthr;
                        b4[q] = -1;
                        if (i < na) b4[q] = kSmall ? ((q == 0 && i0 == tid) ? bbpre : (int)__ldcg(al + i)) : (kListsGlobal ? (int)__ldcg(al + i) : (int)al[i]);
                    }
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int bb = b4[q];
                        if (bb >= 0 && hooklev[bb] == lev) {
                            const int rr = pf_find(bparent, bb);
                            atomicAdd(bsize + rr, bsize[bb]);
                            if (!kW) {
                                atomicMax(racc + rr, racc[bb]); // sums are >= 0: integer order == float order
                            } else {
                                // does any class of bb survive against rr's front as it stands?  (read-only here; the owner
                                // of rr merges the survivors' roots in F3 and checks again)
                                const int cb2 = fcnt[bb], cr = fcnt[rr];
                                bool push = false;
                                for (int k = 0; k < cb2 && !push; ++k) {
                                    const uint2 e = fget(bb, k);
                                    bool dom = false;
                                    for (int j = 0; j < cr && !dom; ++j) {
                                        const uint2 f = fget(rr, j);
                                        dom = f.x >= e.x && f.y >= e.y;
                                    }
                                    push = !dom;
                                }
                                if (push) {
                                    const int slot = atomicAdd(&sMergeN[lev & 1], 1);
                                    if (slot < kMergeCap) sMerge[lev & 1][slot] = (unsigned)bb | ((unsigned)rr << 16);
                                    else sOverflow = 1;
                                }
                            }
                        }
                    }
                }
            }
            __syncthreads();
            if (P.timing && tid == 0) atomicAdd(P.timing + 16 + lev, (unsigned long long)(clock64() - tk));
            PIPE_TICK(4)
            // ================= F3: this level's increment of every live root; next live list ==============
            prefetch_pairs(lev + 2); // this level's registers are free again
            prefetch_entry(lev + 2);
            {
                const unsigned short *__restrict__ al = alive[cur];
                unsigned short *__restrict__ nx = alive[cur ^ 1];
                const int na = sNalive[cur];
                const int nb = sBstart[lev + 1] - sBstart[lev];   // roots born at this level
                const int tot = na + nb;
                const int nmerge = kW ? min(sMergeN[lev & 1], kMergeCap) : 0;
                if (kW && tid == 0) sMergeN[(lev + 1) & 1] = 0;   // the other buffer was last read one barrier ago
                // small geometry: request this thread's first list entry (L2), run the next level's unions while it is in
                // flight.  The order inside the interval is free: the unions of level lev+1 touch neither the sizes nor
                // the leader sums, and the hooks they log carry level lev+1 > lev.
                int bbfirst = -1;
                if (kSmall) {
                    if (tid < na) bbfirst = (int)__ldcg(al + tid);
                    else if (tid < tot) bbfirst = (int)__ldcg(birth + sBstart[lev] + tid - na);
                    if (lev + 1 < nlev) do_unions(lev + 1);
                }
                for (int ig = (tid & ~31); ig < tot; ig += 4 * nthr) { // warp-uniform trip counts; four entries in flight
                  int b4[4];
#pragma unroll
                  for (int q = 0; q < 4; ++q) {
                      const int i = ig + q * nthr + lane;
                      b4[q] = -1;
                      if (kSmall && q == 0 && ig == (tid & ~31)) b4[q] = bbfirst;
                      else if (i < na) b4[q] = kListsGlobal ? (int)__ldcg(al + i) : (int)al[i];
                      else if (i < tot) b4[q] = kListsGlobal ? (int)__ldcg(birth + sBstart[lev] + i - na) : (int)birth[sBstart[lev] + i - na];
                  }
#pragma unroll
                  for (int q = 0; q < 4; ++q) {
                    if (ig + q * nthr >= tot) break; // warp-uniform
                    const int bb = b4[q];
                    const bool live = bb >= 0 && hooklev[bb] > lev; // 255 = never hooked
                    if (live) {
                        const int cb = blev[bb];
                        const int sg = cb >> 7;
                        if (kW) {
                            // the fronts of the roots hooked under bb at this level, then the class of the vertices gained
                            // at this level (sum 0 so far); all sums are those after level lev - 1
                            for (int i = 0; i < nmerge; ++i) {
                                const unsigned e2 = sMerge[lev & 1][i];
                                if ((int)(e2 >> 16) != bb) continue;
                                const int hb = (int)(e2 & 0xffffu), c2 = fcnt[hb];
                                for (int k = 0; k < c2; ++k) { const uint2 e = fget(hb, k); finsert(bb, e.x, e.y); }
                            }
                            const int wn = wnew[bb];
                            if (wn > 0) { wnew[bb] = 0; finsert(bb, 0u, (unsigned)(wn - 1)); }
                        }
                        if (lev < (sg ? ns1 : ns0)) {
                            const int sz = bsize[bb];
                            if (sz < kPowN) {
                                const float inc = __double2float_rn(__dmul_rn(sPow[sz], sHHd[sg][lev]));
                                fadd_all(bb, inc);
                            } else {
                                const int slot = atomicAdd(&sNpend, 1);
                                if (slot < kPend) {
                                    sPendBb[slot] = bb;
                                    cp_async8(sPendPw + slot, powE + sz); // added at the next level's F2a
                                } else { // more large components than slots: fetch synchronously
                                    atomicSub(&sNpend, 1);
                                    const float inc = __double2float_rn(__dmul_rn(powE[sz], sHHd[sg][lev]));
                                    fadd_all(bb, inc);
                                }
                            }
                        }
                    }
                    const unsigned lm = __ballot_sync(0xffffffffu, live);
                    if (lm) {
                        int basepos = 0;
                        if (lane == 0) basepos = atomicAdd(&sNalive[cur ^ 1], __popc(lm));
                        basepos = __shfl_sync(0xffffffffu, basepos, 0);
                        if (live) nx[basepos + __popc(lm & ((1u << lane) - 1u))] = (unsigned short)bb;
                    }
                  }
                }
            }
            if (!kSmall && lev + 1 < nlev) do_unions(lev + 1);
            cp_async_wait_all(); // this thread's pow(size, E) fetches
            __syncthreads();
            if (tid == 0) sNalive[cur] = 0; // becomes the next "next" list (first touched again after the next barrier)
            cur ^= 1;
            if (P.timing && tid == 0) atomicAdd(P.timing + 144 + lev, (unsigned long long)(clock64() - tk));
            PIPE_TICK(5)
        }
        // the last level's pending increments
        cp_async_wait_all();
        __syncthreads();
        {
            const int npend = sNpend;
            for (int i = tid; i < npend; i += nthr) {
                const int bb = sPendBb[i];
                const float inc = __double2float_rn(__dmul_rn(sPendPw[i], sHHd[blev[bb] >> 7][nlev - 1]));
                fadd_all(bb, inc);
            }
        }
        __syncthreads();
        if (tid == 0) sNpend = 0;
        // ---- outputs: the final roots carry the leaders
        const float d0 = sDelta[0], d1 = sDelta[1];
        float m0 = 0.f, m1 = 0.f;
        for (int bb = tid; bb < NB; bb += nthr)
            if (bparent[bb] == bb) {
                const int sg = blev[bb] >> 7;
                if (!kW) {
                    const float sc = __fmul_rn(__int_as_float(racc[bb]), sg ? d1 : d0);
                    if (sg) m1 = fmaxf(m1, sc); else m0 = fmaxf(m0, sc);
                } else {
                    const int c = fcnt[bb];
                    for (int k = 0; k < c; ++k) { // fl32(fl32(tfce * delta) * w), the reference's order (pyfunc.py:116-117)
                        const uint2 e = fget(bb, k);
                        float sc = __fmul_rn(__uint_as_float(e.x), sg ? d1 : d0);
                        if (sd.wtab64) sc = __double2float_rn(__dmul_rn((double)sc, sd.wtab64[e.y]));
                        else if (sd.wtab) sc = __fmul_rn(sc, sd.wtab[e.y]);
                        if (sg) m1 = fmaxf(m1, sc); else m0 = fmaxf(m0, sc);
                    }
                }
            }
        m0 = pwarp_max(m0);
        m1 = pwarp_max(m1);
        if (lane == 0) { sRed[0][wid] = m0; sRed[1][wid] = m1; }
        __syncthreads();
        if (tid == 0 && P.max_out) {
            float a = 0.f, c = 0.f;
            for (int w = 0; w < nthr / 32; ++w) { a = fmaxf(a, sRed[0][w]); c = fmaxf(c, sRed[1][w]); }
            P.max_out[e0] = a;
            P.max_out[e0 + 1] = c;
        }
        if (kW && tid == 0 && sOverflow) meta[2] = 1; // a front outgrew kFront entries: redone by tfce_basin_kernel
        PIPE_TICK(2)
#undef PIPE_TICK
    }
}

// ------------------------------------------------------------------------------------------- K_G
__global__ void __launch_bounds__(256) pipe_output_kernel(PipeParams P, int chunks) {
    int item, chunk, s, b;
    grid_coords(P, item, chunk, s, b);
    const SurfDesc sd = P.surfs[s];
    const int *meta = P.meta + (size_t)item * 4;
    if (chunk * 256 >= sd.V || meta[2]) return; // CTA-uniform
    const int v = chunk * 256 + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const size_t base = (size_t)item * P.vstride;
    const size_t e0 = ((size_t)b * P.S + s) * 2;
    float sc0 = 0.f, sc1 = 0.f;
    if (v < sd.V) {
        const int bu = P.basin[base + v];
        float val = 0.f;
        int neg = 0;
        if (bu >= 0) {
            const int cv = P.lev8[base + v];
            neg = cv >> 7;
            val = __uint_as_float(P.table[(size_t)item * P.tabcap + (size_t)(cv & 0x7f) * meta[0] + bu]);
            const float sc = scaled_vertex_value(val, P.tab_scale ? P.tab_scale[e0 + neg] : P.tab_delta[e0 + neg], sd, v);
            if (neg) sc1 = sc; else sc0 = sc;
        }
        const int32_t *__restrict__ vmap = (P.flags & 4) ? nullptr : sd.vmap;
        const size_t o = (size_t)b * P.ld + sd.col_off + (vmap ? vmap[v] : v);
        if (P.tfce_pos) P.tfce_pos[o] = neg ? 0.f : val;
        if (P.tfce_neg) P.tfce_neg[o] = neg ? val : 0.f;
    }
    sc0 = pwarp_max(sc0);
    sc1 = pwarp_max(sc1);
    if (lane == 0 && P.max_out) { // values are >= 0: integer order == float order
        if (sc0 > 0.f) atomicMax(reinterpret_cast<int *>(P.max_out + e0), __float_as_int(sc0));
        if (sc1 > 0.f) atomicMax(reinterpret_cast<int *>(P.max_out + e0 + 1), __float_as_int(sc1));
    }
}

// ------------------------------------------------------------------------------------------- host
// Sweep geometry by surface size.  A map's ~100 levels are a chain of short barrier intervals bound by memory latency, so
// small maps -- whose basin state needs little shared memory -- run with more, smaller CTAs per SM, i.e. more maps in
// flight: BASELINE config 1 (10,242 vertices, 4,096 maps per step): 6.23 ms with two 512-thread CTAs per SM, 4.41 ms with
// four of 256, 3.37 ms with eight of 128.
static constexpr int kTinyVmax = 40000;  // up to this many vertices: four 256-thread CTAs per SM (36 KB of basin state each)
static constexpr int kMicroVmax = 16000; // up to this many: eight 128-thread CTAs per SM (11 KB each)
int pipe_slots_per_sm(int Vmax) { return Vmax <= kMicroVmax ? 8 : Vmax <= kTinyVmax ? 4 : 2; }

int pipe_sweep_max_smem() {
    int v = 200 * 1024;
    if (const char *g = getenv("TMB_PIPE_SMEM_KB")) { const int kb = atoi(g); if (kb >= 16 && kb <= 224) v = kb * 1024; }
    return v;
}

int launch_tfce_tables(const SurfDesc *surfs, int S, int count, const float *maxima, int two_sided, int32_t *ns,
                       float *delta, float *T, float *HH, int32_t *st, cudaStream_t stream) {
    if (count <= 0) return 0;
    pipe_tables_kernel<<<(count + 63) / 64, 64, 0, stream>>>(surfs, S, count, maxima, two_sided, ns, delta, T, HH, st);
    count_launch();
    TMB_CUDA(cudaGetLastError());
    return 0;
}

int launch_tfce_pipeline(const PipeParams &p_in, int num_slots, int sm_count, cudaStream_t stream) {
    PipeParams p = p_in;
    if (const char *d = getenv("TMB_PIPE_DEBUG")) p.flags |= atoi(d) & ~7; // timing experiments only (results invalid)
    const int items = p.B * p.S;
    if (items <= 0) return 0;
    p.front_cap = kFront;
    if (const char *fc = getenv("TMB_PIPE_FRONT")) { const int v = atoi(fc); if (v >= 1 && v < kFront) p.front_cap = v; }
    const int chunksA = (p.Vmax + kChunkA - 1) / kChunkA;
    const int chunks = (p.Vmax + 255) / 256;
    TMB_REQUIRE(p.B <= 65535 && p.S <= 65535, "tfce pipeline: at most 65535 rows and surfaces per launch (got %d, %d)", p.B, p.S);
    TMB_CUDA(cudaMemsetAsync(p.lhist, 0, sizeof(int) * 256 * (size_t)items, stream));
    pipe_levels_kernel<<<dim3(chunksA, p.B, p.S), 256, 0, stream>>>(p, chunksA);
    // Mixed plans: the leading narrow_slots surface slots (triangle meshes, self-padded width-8 rows) take the fixed-width
    // ascent / count kernels -- 14 % less stage time than sliced rows on such graphs (config 2: 13.2 against 15.3 ms) --
    // and the remaining slots the sliced-row kernels; each group is its own launch over its slot range.
    const int Sn = p.sell_words ? std::min(p.narrow_slots, p.S) : p.S;      // slots on fixed-width rows
    const int Sw = p.S - Sn;                                                // slots on sliced rows
    PipeParams pw = p;
    pw.z0 = Sn;
    if (Sn > 0) {
        const dim3 gridB(chunks, p.B, Sn);
        const int maxdeg = p.sell_words ? p.narrow_max_degree : p.max_degree;
        const int self = p.sell_words ? 1 : p.ell_self;
        if (maxdeg > 0 && maxdeg <= 6 && self) pipe_ascent_kernel<true, true><<<gridB, 256, 0, stream>>>(p, chunks);
        else if (maxdeg > 0 && maxdeg <= 6) pipe_ascent_kernel<true, false><<<gridB, 256, 0, stream>>>(p, chunks);
        else if (self) pipe_ascent_kernel<false, true><<<gridB, 256, 0, stream>>>(p, chunks);
        else pipe_ascent_kernel<false, false><<<gridB, 256, 0, stream>>>(p, chunks);
    }
    if (Sw > 0) {
        const dim3 gridB(chunks, p.B, Sw);
        switch (p.sell_words) {
        case 1: pipe_ascent_wide_kernel<1><<<gridB, 256, 0, stream>>>(pw, chunks); break;
        case 2: pipe_ascent_wide_kernel<2><<<gridB, 256, 0, stream>>>(pw, chunks); break;
        case 4: pipe_ascent_wide_kernel<4><<<gridB, 256, 0, stream>>>(pw, chunks); break;
        case 8: pipe_ascent_wide_kernel<8><<<gridB, 256, 0, stream>>>(pw, chunks); break;
        default: set_error("tfce pipeline: sell_words must be 0, 1, 2, 4 or 8 (got %d)", p.sell_words); return 1;
        }
    }
    const int chunksC = (p.Vmax + kBasinChunk - 1) / kBasinChunk;
    if (p.weighted) pipe_basin_kernel<true><<<dim3(chunksC, p.B, p.S), 256, 0, stream>>>(p, chunksC);
    else pipe_basin_kernel<false><<<dim3(chunksC, p.B, p.S), 256, 0, stream>>>(p, chunksC);
    const int chunksD = (p.Vmax + kCountChunk - 1) / kCountChunk;
    if (Sn > 0) {
        const dim3 gridD(chunksD, p.B, Sn);
        if (p.want_vertex_pass) pipe_count_kernel<true><<<gridD, 256, 0, stream>>>(p, chunksD);
        else pipe_count_kernel<false><<<gridD, 256, 0, stream>>>(p, chunksD);
    }
    if (Sw > 0) {
        const dim3 gridD(chunksD, p.B, Sw);
#define TMB_COUNT_WIDE(W)                                                                                  \
        if (p.want_vertex_pass) pipe_count_wide_kernel<true, W><<<gridD, 256, 0, stream>>>(pw, chunksD);   \
        else pipe_count_wide_kernel<false, W><<<gridD, 256, 0, stream>>>(pw, chunksD)
        switch (p.sell_words) {
        case 1: TMB_COUNT_WIDE(1); break;
        case 2: TMB_COUNT_WIDE(2); break;
        case 4: TMB_COUNT_WIDE(4); break;
        default: TMB_COUNT_WIDE(8); break;
        }
#undef TMB_COUNT_WIDE
    }
    if (Sn > 0 && Sw > 0) count_launch(2);
    TMB_CUDA(cudaMemsetAsync(p.work_counter, 0, sizeof(int), stream));
    if (p.want_vertex_pass) {
        // class path (values per vertex): one 1024-thread CTA per SM
        const int smem = pipe_sweep_max_smem();
        TMB_CUDA(cudaFuncSetAttribute(pipe_sweep_kernel<1024, 1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        const int grid = items < sm_count ? items : sm_count;
        pipe_sweep_kernel<1024, 1, false><<<grid, 1024, smem, stream>>>(p, smem);
    } else {
        // max-only path.  Default for surfaces above kTinyVmax vertices (TMB_PIPE_GEOM=2): two 512-thread CTAs per SM (small geometry) that hide each other's
        // barrier intervals, then one launch of the large geometry (one 1,024-thread CTA per SM) for the maps the small
        // one could not hold.  TMB_PIPE_GEOM=1: large geometry only.  Measured on config 2 (B200, whole TFCE stage):
        // 1,024 maps 9.15 -> 9.04 ms, 2,048 maps 18.05 -> 17.20 ms (the small geometry's longer maps cost more at the
        // tail of a launch, so it pays with more maps per launch).
        int geom = 0;                                               // 0: choose by surface size
        if (const char *g = getenv("TMB_PIPE_GEOM")) geom = atoi(g);
        const int smem_large = 196 * 1024, smem_small = 94 * 1024, smem_tiny = 36 * 1024;
        const int grid_large = items < sm_count ? items : sm_count;
        const int grid_small = items < 2 * sm_count ? items : 2 * sm_count;
        const int grid_tiny = items < 4 * sm_count ? items : 4 * sm_count;
        // Small surfaces: geometry 3 (four 256-thread CTAs per SM) or 4 (eight of 128), see kTinyVmax / kMicroVmax.
        if (geom < 1 || geom > 4) geom = p.Vmax <= kMicroVmax ? 4 : p.Vmax <= kTinyVmax ? 3 : 2;
        if (geom == 4 && num_slots < 8 * sm_count) geom = 3;
        if (geom == 3 && num_slots < 4 * sm_count) geom = 2;
#define TMB_SWEEP_MAX(W)                                                                                                   \
        TMB_CUDA(cudaFuncSetAttribute(pipe_sweep_max_kernel<1024, 1, false, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_large)); \
        if (geom == 1) {                                                                                                   \
            pipe_sweep_max_kernel<1024, 1, false, W><<<grid_large, 1024, smem_large, stream>>>(p, smem_large, 0);          \
        } else if (geom == 4) {                                                                                            \
            TMB_CUDA(cudaFuncSetAttribute(pipe_sweep_max_kernel<128, 8, true, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, 11 * 1024)); \
            pipe_sweep_max_kernel<128, 8, true, W><<<(items < 8 * sm_count ? items : 8 * sm_count), 128, 11 * 1024, stream>>>(p, 11 * 1024, 1); \
            TMB_CUDA(cudaMemsetAsync(p.work_counter, 0, sizeof(int), stream));                                             \
            pipe_sweep_max_kernel<1024, 1, false, W><<<grid_large, 1024, smem_large, stream>>>(p, smem_large, 2);          \
            count_launch();                                                                                                \
        } else if (geom == 3) {                                                                                            \
            TMB_CUDA(cudaFuncSetAttribute(pipe_sweep_max_kernel<256, 4, true, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_tiny)); \
            pipe_sweep_max_kernel<256, 4, true, W><<<grid_tiny, 256, smem_tiny, stream>>>(p, smem_tiny, 1);                \
            TMB_CUDA(cudaMemsetAsync(p.work_counter, 0, sizeof(int), stream));                                             \
            pipe_sweep_max_kernel<1024, 1, false, W><<<grid_large, 1024, smem_large, stream>>>(p, smem_large, 2);          \
            count_launch();                                                                                                \
        } else {                                                                                                           \
            TMB_CUDA(cudaFuncSetAttribute(pipe_sweep_max_kernel<512, 2, true, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_small)); \
            pipe_sweep_max_kernel<512, 2, true, W><<<grid_small, 512, smem_small, stream>>>(p, smem_small, 1);             \
            TMB_CUDA(cudaMemsetAsync(p.work_counter, 0, sizeof(int), stream));                                             \
            pipe_sweep_max_kernel<1024, 1, false, W><<<grid_large, 1024, smem_large, stream>>>(p, smem_large, 2);          \
            count_launch();                                                                                                \
        }
        if (p.weighted) { TMB_SWEEP_MAX(true) } else { TMB_SWEEP_MAX(false) }
#undef TMB_SWEEP_MAX
    }
    count_launch(5);
    if (p.want_vertex_pass) {
        pipe_output_kernel<<<dim3(chunks, p.B, p.S), 256, 0, stream>>>(p, chunks);
        count_launch();
    }
    TMB_CUDA(cudaGetLastError());
    return 0;
}

} // namespace tmb
