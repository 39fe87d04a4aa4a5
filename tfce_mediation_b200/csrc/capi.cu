// C ABI of libtfce_b200.so: handles (graph, plan) and the TFCE entry points.
// The GLM entry points live in glm_kernels.cu, the voxel adjacency builder in adjacency_kernels.cu.
#include "common.cuh"
#include "../../include/tfce_b200.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <string>
#include <thread>
#include <vector>

namespace tmb {

static thread_local std::string g_error;
std::atomic<int64_t> g_launch_count{0};

void set_error(const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_error = buf;
}

} // namespace tmb

using namespace tmb;

struct tmb_graph {
    int device = 0;
    int32_t V = 0;
    int64_t nnz = 0;
    float H = 2.f, E = 1.f;
    int64_t *d_indptr = nullptr;
    int32_t *d_indices = nullptr;
    double *d_powE = nullptr;
    int32_t *d_vmap = nullptr;          // internal -> caller index (nullptr: identity)
    int32_t *d_ell = nullptr;           // fixed-width adjacency rows (low-degree graphs)
    int32_t *d_ell_self = nullptr;      // the same rows for width 8 with pad = the vertex itself (pipeline ascent kernel)
    int32_t ell_width = 0;
    int32_t max_degree = 0;
    // sliced rows (SELL-32-4, see SurfDesc): built on the host at creation for every symmetric graph of degree <= 256,
    // uploaded by the first plan that runs the wide pipeline kernels
    std::vector<int32_t> sell_host, sell_off_host;
    int32_t sell_words = 0;             // ceil(max_degree / 32); 0: no sliced rows
    int4 *d_sell = nullptr;
    int32_t *d_sell_off = nullptr;
    std::vector<int32_t> vmap;          // host copy
    bool symmetric = true;
    tmb_plan *self_plan = nullptr; // lazily created single-surface plan for tmb_tfce_run
    float *d_image = nullptr, *d_enhn = nullptr;
    int32_t *d_labels = nullptr, *d_extents = nullptr, *d_status = nullptr;
    float *d_thr = nullptr;
    char *d_tabs = nullptr; // [ns(2) | status(2) | delta(2) | T(2x128) | HH(2x128)] for the single-map entry points
};

struct tmb_plan {
    int device = 0;
    int S = 0;
    int32_t Vmax = 0;
    int num_slots = 0;
    std::vector<tmb_graph *> graphs;
    std::vector<int64_t> col_offset;
    SurfDesc *d_surfs = nullptr;
    int32_t *d_order = nullptr;
    std::vector<float *> d_weights;
    std::vector<unsigned short *> d_wrank;  // weight ranks per surface (leader sweep with weights), or nullptr
    std::vector<float *> d_wtab;            // distinct weight values per surface, ascending
    std::vector<double *> d_weights64;      // float64 weights per surface (non-low-RAM mmr), or nullptr
    std::vector<double *> d_wtab64;
    int pipe_wfast = 0;                     // every weighted surface has ranks: max-only maps need no per-vertex pass
    char *d_workspace = nullptr;
    size_t slot_stride = 0;
    int *d_counter = nullptr;
    int use_basin = 0;          // every surface symmetric -> V3 basin sweep
    int internal_order = 0;     // statistic rows already use the graphs' internal vertex order (vmap ignored)
    unsigned long long *d_timing = nullptr; // TMB_PHASE_TIMING=1: per-phase cycle totals
    // streaming pipeline (tfce_pipeline.cu): per-item buffers for `pipe_items` work items, grown on demand
    int pipe_ok = 0;            // basin path and every surface has fixed-width rows of at most 32 slots
    int pipe_weights = 0;       // some surface carries vertex weights (scaled maxima then need the per-vertex pass)
    int pipe_words = 0;         // 0: fixed-width rows; W > 0: sliced rows, W mask words per vertex (wide pipeline kernels)
    int pipe_narrow = 0;        // mixed plans (pipe_words > 0): leading surface slots that stay on fixed-width rows
    int pipe_narrow_degree = 0; // largest degree among them
    int pipe_items = 0;
    int pipe_has_table = 0;     // the per-item buffers include the class path's [level][basin] table (32 B per vertex)
    int64_t pipe_vstride = 0, pipe_tabcap = 0;
    int pipe_nbcap = 0, pipe_paircap = 0;
    char *d_pipe = nullptr;     // one allocation: lev8 | up | emask | basin | meta | blev | pairs | table
    char *d_pipe_slots = nullptr;
    size_t pipe_slot_stride = 0;
    int pipe_slots = 0;
    int pipe_sms = 0;           // SMs of the device (the sweep grids are multiples of it)
    // device-built threshold tables (exact_pow = False) for `tab_items` work items
    int tab_items = 0;
    char *d_tabs = nullptr;     // maxima | ns | status | delta | T | HH
};

extern "C" const char *tmb_last_error(void) { return g_error.c_str(); }
extern "C" int tmb_abi_version(void) { return TMB_ABI_VERSION; }
extern "C" int64_t tmb_launch_count(void) { return g_launch_count.load(); }

extern "C" int tmb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}


// ---- host-side graph analysis (one-time, at CreateAdjSet construction) ---------------------------
// Symmetric adjacency lets the sweep compare activation levels only; otherwise the reference's
// directional rule (fast_tfce.hpp:47-65) is honoured with value comparisons.
static bool csr_is_symmetric(int32_t V, const int64_t *indptr, const int32_t *indices) {
    std::vector<int32_t> sorted(indices, indices + indptr[V]);
    for (int32_t v = 0; v < V; ++v) std::sort(sorted.begin() + indptr[v], sorted.begin() + indptr[v + 1]);
    for (int32_t u = 0; u < V; ++u)
        for (int64_t e = indptr[u]; e < indptr[u + 1]; ++e) {
            const int32_t a = indices[e];
            if (a == u) continue;
            if (!std::binary_search(sorted.begin() + indptr[a], sorted.begin() + indptr[a + 1], u)) return false;
        }
    return true;
}

// Reverse Cuthill-McKee ordering: neighbours end up close in index, so the random parent/level
// lookups of one warp fall into a few cache lines.  Pure relabelling: TFCE values do not depend on
// vertex numbering for symmetric graphs.  Returns internal -> caller index.
static std::vector<int32_t> rcm_order(int32_t V, const int64_t *indptr, const int32_t *indices) {
    std::vector<int32_t> order;
    order.reserve(V);
    std::vector<char> seen(V, 0);
    std::vector<int32_t> by_degree(V);
    std::iota(by_degree.begin(), by_degree.end(), 0);
    auto deg = [&](int32_t v) { return indptr[v + 1] - indptr[v]; };
    std::stable_sort(by_degree.begin(), by_degree.end(), [&](int32_t a, int32_t b) { return deg(a) < deg(b); });
    std::vector<int32_t> nb;
    for (int32_t start : by_degree) {
        if (seen[start]) continue;
        seen[start] = 1;
        size_t head = order.size();
        order.push_back(start);
        while (head < order.size()) {
            const int32_t v = order[head++];
            nb.clear();
            for (int64_t e = indptr[v]; e < indptr[v + 1]; ++e) {
                const int32_t a = indices[e];
                if (!seen[a]) { seen[a] = 1; nb.push_back(a); }
            }
            std::sort(nb.begin(), nb.end(), [&](int32_t a, int32_t b) { return deg(a) != deg(b) ? deg(a) < deg(b) : a < b; });
            order.insert(order.end(), nb.begin(), nb.end());
        }
    }
    std::reverse(order.begin(), order.end());
    return order;
}

// Patch ordering: walk the RCM order and grow compact BFS blobs of up to `patch` still-unassigned
// vertices; blobs get consecutive ids.  Most neighbours of a vertex then sit within the same few
// hundred bytes, and consecutive blobs stay adjacent on the surface.
static std::vector<int32_t> patch_order(int32_t V, const int64_t *indptr, const int32_t *indices, int patch) {
    const std::vector<int32_t> base = rcm_order(V, indptr, indices);
    std::vector<char> done(V, 0);
    std::vector<int32_t> order;
    order.reserve(V);
    for (int32_t seed : base) {
        if (done[seed]) continue;
        const size_t start = order.size();
        done[seed] = 1;
        order.push_back(seed);
        size_t head = start;
        while (head < order.size() && order.size() - start < (size_t)patch) {
            const int32_t v = order[head++];
            for (int64_t e = indptr[v]; e < indptr[v + 1] && order.size() - start < (size_t)patch; ++e) {
                const int32_t a = indices[e];
                if (!done[a]) { done[a] = 1; order.push_back(a); }
            }
        }
    }
    return order;
}

extern "C" int tmb_graph_create(int device, int32_t V, const int64_t *indptr, const int32_t *indices, float H,
                                float E, tmb_graph **out) {
    TMB_REQUIRE(out && indptr && V > 0, "tmb_graph_create: bad arguments (V=%d)", V);
    TMB_REQUIRE(V < 0x00FFFFFF, "tmb_graph_create: at most %d vertices per graph are supported", 0x00FFFFFF - 1);
    TMB_REQUIRE(indptr[0] == 0, "tmb_graph_create: indptr[0] must be 0");
    for (int32_t v = 0; v < V; ++v)
        TMB_REQUIRE(indptr[v + 1] >= indptr[v], "tmb_graph_create: indptr not monotone at %d", v);
    const int64_t nnz = indptr[V];
    TMB_REQUIRE(nnz == 0 || indices, "tmb_graph_create: indices is null");
    for (int64_t e = 0; e < nnz; ++e)
        TMB_REQUIRE(indices[e] >= 0 && indices[e] < V, "tmb_graph_create: neighbour index %d out of range [0,%d)",
                    indices[e], V);
    TMB_ON_DEVICE(device);
    tmb_graph *g = new tmb_graph();
    g->device = device; g->V = V; g->nnz = nnz; g->H = H; g->E = E;
    g->symmetric = csr_is_symmetric(V, indptr, indices);
    // locality reordering (symmetric graphs only; TMB_NO_REORDER=1 disables it for A/B measurements)
    const char *no_reorder = getenv("TMB_NO_REORDER");
    const char *reorder_mode = getenv("TMB_REORDER"); // "rcm" | "patch<N>" (default patch64)
    std::vector<int64_t> r_indptr;
    std::vector<int32_t> r_indices;
    const int64_t *use_indptr = indptr;
    const int32_t *use_indices = indices;
    if (g->symmetric && !(no_reorder && no_reorder[0] == '1') && nnz > 0) {
        int patch = 64;
        if (reorder_mode && strncmp(reorder_mode, "patch", 5) == 0 && atoi(reorder_mode + 5) > 0) patch = atoi(reorder_mode + 5);
        if (reorder_mode && strcmp(reorder_mode, "rcm") == 0) g->vmap = rcm_order(V, indptr, indices);
        else g->vmap = patch_order(V, indptr, indices, patch);
        std::vector<int32_t> inv(V);
        for (int32_t i = 0; i < V; ++i) inv[g->vmap[i]] = i;
        r_indptr.assign((size_t)V + 1, 0);
        r_indices.resize((size_t)nnz);
        for (int32_t i = 0; i < V; ++i) r_indptr[i + 1] = r_indptr[i] + (indptr[g->vmap[i] + 1] - indptr[g->vmap[i]]);
        for (int32_t i = 0; i < V; ++i) {
            const int32_t o = g->vmap[i];
            int64_t w = r_indptr[i];
            for (int64_t e = indptr[o]; e < indptr[o + 1]; ++e) r_indices[w++] = inv[indices[e]];
            std::sort(r_indices.begin() + r_indptr[i], r_indices.begin() + r_indptr[i + 1]);
        }
        use_indptr = r_indptr.data();
        use_indices = r_indices.data();
    }
    // pow(n, E) table with the host C library: identical to the reference's
    // pow(c->size(), E) double overload (fast_tfce.hpp:77)
    std::vector<double> powE((size_t)V + 1);
    for (int64_t n = 0; n <= V; ++n) powE[(size_t)n] = std::pow((double)n, (double)E);
    auto fail = [&](const char *what, cudaError_t e) {
        set_error("tmb_graph_create: %s: %s", what, cudaGetErrorString(e));
        tmb_graph_destroy(g);
        return 1;
    };
    cudaError_t e;
    if ((e = cudaMalloc(&g->d_indptr, sizeof(int64_t) * ((size_t)V + 1))) != cudaSuccess) return fail("malloc", e);
    if ((e = cudaMalloc(&g->d_indices, sizeof(int32_t) * (size_t)std::max<int64_t>(nnz, 1))) != cudaSuccess)
        return fail("malloc", e);
    if ((e = cudaMalloc(&g->d_powE, sizeof(double) * ((size_t)V + 1))) != cudaSuccess) return fail("malloc", e);
    if ((e = cudaMemcpy(g->d_indptr, use_indptr, sizeof(int64_t) * ((size_t)V + 1), cudaMemcpyHostToDevice)) != cudaSuccess)
        return fail("memcpy", e);
    if (nnz && (e = cudaMemcpy(g->d_indices, use_indices, sizeof(int32_t) * (size_t)nnz, cudaMemcpyHostToDevice)) != cudaSuccess)
        return fail("memcpy", e);
    if ((e = cudaMemcpy(g->d_powE, powE.data(), sizeof(double) * ((size_t)V + 1), cudaMemcpyHostToDevice)) != cudaSuccess)
        return fail("memcpy", e);
    {
        // fixed-width (ELL) copy of the rows for low-degree graphs: one or two 128-bit loads fetch a whole
        // row from a single 32-byte sector instead of indptr + one scalar load per neighbour
        int64_t maxdeg = 0;
        for (int32_t v = 0; v < V; ++v) maxdeg = std::max<int64_t>(maxdeg, use_indptr[v + 1] - use_indptr[v]);
        g->max_degree = (int32_t)std::min<int64_t>(maxdeg, INT32_MAX);
        int width = maxdeg <= 8 ? 8 : maxdeg <= 16 ? 16 : maxdeg <= 32 ? 32 : 0;
        if (width && nnz > 0 && (double)nnz / ((double)V * width) >= 0.4) {
            std::vector<int32_t> ell((size_t)V * width, -1);
            for (int32_t v = 0; v < V; ++v)
                std::copy(use_indices + use_indptr[v], use_indices + use_indptr[v + 1], ell.begin() + (size_t)v * width);
            if ((e = cudaMalloc(&g->d_ell, sizeof(int32_t) * ell.size())) != cudaSuccess) return fail("malloc", e);
            if ((e = cudaMemcpy(g->d_ell, ell.data(), sizeof(int32_t) * ell.size(), cudaMemcpyHostToDevice)) != cudaSuccess)
                return fail("memcpy", e);
            g->ell_width = width;
            if (width == 8) {
                for (int32_t v = 0; v < V; ++v)
                    for (int j = (int)(use_indptr[v + 1] - use_indptr[v]); j < 8; ++j) ell[(size_t)v * 8 + j] = v;
                if ((e = cudaMalloc(&g->d_ell_self, sizeof(int32_t) * ell.size())) != cudaSuccess) return fail("malloc", e);
                if ((e = cudaMemcpy(g->d_ell_self, ell.data(), sizeof(int32_t) * ell.size(), cudaMemcpyHostToDevice)) != cudaSuccess)
                    return fail("memcpy", e);
            }
        }
        // sliced rows (SELL-32-4) for everything else -- the reference's default geodesic adjacency sets have ~60
        // neighbours per vertex (STEP_1_vertex_tfce_multiple_regression.py:71-76,155-158) -- kept on the host until
        // a plan needs them.  Per slice of 32 vertices: width = largest degree rounded up to 4.
        if (g->symmetric && nnz > 0 && maxdeg <= 256) {
            const int32_t nsl = (V + 31) / 32;
            g->sell_off_host.assign((size_t)nsl + 1, 0);
            int64_t run = 0;
            for (int32_t sl = 0; sl < nsl; ++sl) {
                int64_t w = 0;
                for (int32_t v = sl * 32; v < std::min(V, sl * 32 + 32); ++v) w = std::max<int64_t>(w, use_indptr[v + 1] - use_indptr[v]);
                run += (w + 3) / 4 * 32;
                g->sell_off_host[(size_t)sl + 1] = (int32_t)run;
            }
            if (run < (int64_t)INT32_MAX / 2) {
                // pad slots hold the vertex ITSELF: a self-reference is never "earlier" and never an ascent target, so
                // the kernels need no pad test in front of their gathers
                g->sell_host.assign((size_t)run * 4, 0);
                for (int32_t v = 0; v < V; ++v) {
                    const int64_t o0 = g->sell_off_host[v >> 5];
                    const int64_t w4 = (g->sell_off_host[(v >> 5) + 1] - o0) / 32;
                    int j = 0;
                    for (int64_t e2 = use_indptr[v]; e2 < use_indptr[v + 1]; ++e2, ++j)
                        g->sell_host[(size_t)(o0 + (j >> 2) * 32 + (v & 31)) * 4 + (j & 3)] = use_indices[e2];
                    for (; j < w4 * 4; ++j) g->sell_host[(size_t)(o0 + (j >> 2) * 32 + (v & 31)) * 4 + (j & 3)] = v;
                }
                g->sell_words = (int32_t)std::max<int64_t>(1, (maxdeg + 31) / 32);
            } else {
                g->sell_off_host.clear();
            }
        }
    }
    if (!g->vmap.empty()) {
        if ((e = cudaMalloc(&g->d_vmap, sizeof(int32_t) * (size_t)V)) != cudaSuccess) return fail("malloc", e);
        if ((e = cudaMemcpy(g->d_vmap, g->vmap.data(), sizeof(int32_t) * (size_t)V, cudaMemcpyHostToDevice)) != cudaSuccess)
            return fail("memcpy", e);
    }
    *out = g;
    return 0;
}

extern "C" int tmb_graph_destroy(tmb_graph *g) {
    if (!g) return 0;
    DeviceGuard guard(g->device);
    if (g->self_plan) tmb_plan_destroy(g->self_plan);
    cudaFree(g->d_indptr); cudaFree(g->d_indices); cudaFree(g->d_powE); cudaFree(g->d_vmap); cudaFree(g->d_ell); cudaFree(g->d_ell_self); cudaFree(g->d_sell); cudaFree(g->d_sell_off);
    cudaFree(g->d_image); cudaFree(g->d_enhn); cudaFree(g->d_labels); cudaFree(g->d_extents);
    cudaFree(g->d_status); cudaFree(g->d_thr); cudaFree(g->d_tabs);
    delete g;
    return 0;
}

extern "C" int tmb_graph_vmap(const tmb_graph *g, int32_t *vmap_host) {
    TMB_REQUIRE(g && vmap_host, "tmb_graph_vmap: null pointer");
    for (int32_t i = 0; i < g->V; ++i) vmap_host[i] = g->vmap.empty() ? i : g->vmap[i];
    return 0;
}

extern "C" int tmb_plan_set_internal_order(tmb_plan *p, int on) {
    TMB_REQUIRE(p, "tmb_plan_set_internal_order: null plan");
    p->internal_order = on ? 1 : 0;
    return 0;
}

extern "C" int tmb_graph_num_vertices(const tmb_graph *g, int32_t *V, int64_t *nnz) {
    TMB_REQUIRE(g, "tmb_graph_num_vertices: null graph");
    if (V) *V = g->V;
    if (nnz) *nnz = g->nnz;
    return 0;
}

extern "C" int tmb_plan_create(int device, int S, tmb_graph *const *graphs, const int64_t *col_offset,
                               const float *const *weight_host, const double *const *weight64_host, int max_slots,
                               tmb_plan **out) {
    TMB_REQUIRE(out && graphs && col_offset && S > 0, "tmb_plan_create: bad arguments");
    for (int s = 0; s < S; ++s) {
        TMB_REQUIRE(graphs[s], "tmb_plan_create: graph %d is null", s);
        TMB_REQUIRE(graphs[s]->device == device, "tmb_plan_create: graph %d lives on device %d, plan on %d", s,
                    graphs[s]->device, device);
        TMB_REQUIRE(col_offset[s] >= 0, "tmb_plan_create: negative column offset");
    }
    TMB_ON_DEVICE(device);
    tmb_plan *p = new tmb_plan();
    p->device = device; p->S = S;
    p->graphs.assign(graphs, graphs + S);
    p->col_offset.assign(col_offset, col_offset + S);
    p->d_weights.assign(S, nullptr);
    p->d_wrank.assign(S, nullptr);
    p->d_wtab.assign(S, nullptr);
    p->d_weights64.assign(S, nullptr);
    p->d_wtab64.assign(S, nullptr);
    std::vector<SurfDesc> descs(S);
    std::vector<int32_t> order(S);
    std::iota(order.begin(), order.end(), 0);
    for (int s = 0; s < S; ++s) p->Vmax = std::max(p->Vmax, graphs[s]->V);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return graphs[a]->V > graphs[b]->V; });
    cudaDeviceProp prop;
    cudaError_t e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) { set_error("tmb_plan_create: %s", cudaGetErrorString(e)); delete p; return 1; }
    {
        const char *sw = getenv("TMB_SWEEP"); // "v2" forces the vertex-level sweep (A/B measurements)
        p->use_basin = !(sw && strcmp(sw, "v2") == 0);
        for (int s = 0; s < S; ++s)
            if (!graphs[s]->symmetric) p->use_basin = 0; // directed adjacency: reference's directional rule (V2)
    }
    p->slot_stride = tfce_slot_bytes_for(p->Vmax, p->use_basin);
    {
        int threads, per_sm;
        size_t dyn;
        tfce_sweep_geometry(p->Vmax, p->use_basin, &threads, &per_sm, &dyn);
        p->num_slots = max_slots > 0 ? max_slots : prop.multiProcessorCount * per_sm;
    }
    {   // keep the workspace within a quarter of the free HBM (huge merged graphs get fewer slots)
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) {
            const size_t budget = free_b / 4;
            const size_t fit = std::max<size_t>(1, budget / std::max<size_t>(p->slot_stride, 1));
            if ((size_t)p->num_slots > fit) p->num_slots = (int)fit;
        }
    }
    auto fail = [&](const char *what, cudaError_t err) {
        set_error("tmb_plan_create: %s: %s", what, cudaGetErrorString(err));
        tmb_plan_destroy(p);
        return 1;
    };
    {
        // which row format the streaming pipeline uses: fixed-width rows when every surface has them (degree <= 32,
        // round-1 kernels), else sliced rows for all surfaces (wide kernels); neither: the one-kernel sweep
        const char *tf = getenv("TMB_TFCE");        // "basin" forces the one-kernel sweep (A/B measurements)
        const char *fs = getenv("TMB_PIPE_ROWS");   // "sell" forces sliced rows (tests, A/B measurements)
        p->pipe_ok = p->use_basin && !(tf && strcmp(tf, "basin") == 0);
        bool all_ell = true, all_sell = true;
        int words = 1;
        for (int s = 0; s < S; ++s) {
            // fixed-width rows only for width 8 (triangle meshes): wider rows -- 26-connectivity voxel graphs sit in 32 slots
            // half empty -- run 25% faster on sliced rows (BASELINE config 3: 76.0 k against 60.9 k shuffles/s)
            const char *keep = getenv("TMB_PIPE_ROWS");
            const bool ell_ok = graphs[s]->d_ell && (graphs[s]->ell_width == 8 || (keep && strcmp(keep, "ell") == 0 && graphs[s]->ell_width <= 32));
            if (!ell_ok) all_ell = false;
            if (graphs[s]->sell_words == 0) all_sell = false;
            words = std::max(words, (int)graphs[s]->sell_words);
        }
        const bool force_sell = fs && strcmp(fs, "sell") == 0;
        if (force_sell && all_sell) all_ell = false;
        if (all_ell) p->pipe_words = 0;
        else if (all_sell) {
            // Mixed plans (mmr: triangle meshes next to voxel graphs): the meshes keep their fixed-width rows.  The
            // surface slots are issued meshes first (each group still by descending size), so that every group is one
            // contiguous slot range of the streaming kernels' grids.  TMB_PIPE_MIXED=0: sliced rows for all surfaces.
            const char *mx = getenv("TMB_PIPE_MIXED");
            std::vector<char> narrow((size_t)S, 0);
            int nn = 0;
            if (!force_sell && !(mx && mx[0] == '0'))
                for (int s = 0; s < S; ++s)
                    if (graphs[s]->d_ell && graphs[s]->ell_width == 8 && graphs[s]->d_ell_self) { narrow[s] = 1; ++nn; }
            if (nn > 0 && nn < S) {
                std::stable_partition(order.begin(), order.end(), [&](int s) { return narrow[s] != 0; });
                p->pipe_narrow = nn;
                words = 1;
                for (int s = 0; s < S; ++s) {
                    if (narrow[s]) p->pipe_narrow_degree = std::max(p->pipe_narrow_degree, (int)graphs[s]->max_degree);
                    else words = std::max(words, (int)graphs[s]->sell_words);
                }
            }
            p->pipe_words = words <= 1 ? 1 : words <= 2 ? 2 : words <= 4 ? 4 : 8;
        } else p->pipe_ok = 0;
        if (p->pipe_ok && p->pipe_words)
            for (int s = 0; s < S; ++s) {
                tmb_graph *g = graphs[s];
                if (g->d_sell) continue;
                if ((e = cudaMalloc(&g->d_sell, sizeof(int32_t) * std::max<size_t>(g->sell_host.size(), 4))) != cudaSuccess) return fail("malloc", e);
                if ((e = cudaMalloc(&g->d_sell_off, sizeof(int32_t) * g->sell_off_host.size())) != cudaSuccess) return fail("malloc", e);
                if ((e = cudaMemcpy(g->d_sell, g->sell_host.data(), sizeof(int32_t) * g->sell_host.size(), cudaMemcpyHostToDevice)) != cudaSuccess)
                    return fail("memcpy", e);
                if ((e = cudaMemcpy(g->d_sell_off, g->sell_off_host.data(), sizeof(int32_t) * g->sell_off_host.size(), cudaMemcpyHostToDevice)) != cudaSuccess)
                    return fail("memcpy", e);
            }
        // sweep CTAs per SM: two of 512 threads; four of 256 / eight of 128 for plans of small surfaces (launch_tfce_pipeline)
        p->pipe_sms = prop.multiProcessorCount;
        p->pipe_slots = pipe_slots_per_sm(p->Vmax) * prop.multiProcessorCount;
    }
    bool wfast_ok = true;
    for (int s = 0; s < S; ++s) {
        const tmb_graph *g = graphs[s];
        if (weight64_host && weight64_host[s]) {
            // float64 weights: same ranks, the product is formed in double (scaled_vertex_value)
            std::vector<double> w(weight64_host[s], weight64_host[s] + g->V);
            if (!g->vmap.empty())
                for (int32_t i = 0; i < g->V; ++i) w[i] = weight64_host[s][g->vmap[i]];
            if ((e = cudaMalloc(&p->d_weights64[s], sizeof(double) * (size_t)g->V)) != cudaSuccess) return fail("malloc", e);
            if ((e = cudaMemcpy(p->d_weights64[s], w.data(), sizeof(double) * (size_t)g->V, cudaMemcpyHostToDevice)) != cudaSuccess)
                return fail("memcpy", e);
            p->pipe_weights = 1;
            bool ok = true;
            for (double x : w) ok = ok && std::isfinite(x) && x >= 0.0;
            std::vector<double> vals(w);
            std::sort(vals.begin(), vals.end());
            vals.resize((size_t)(std::unique(vals.begin(), vals.end()) - vals.begin()));
            if (ok && vals.size() <= 65535) {
                std::vector<unsigned short> rk((size_t)g->V);
                for (int32_t i = 0; i < g->V; ++i)
                    rk[i] = (unsigned short)(std::lower_bound(vals.begin(), vals.end(), w[i]) - vals.begin());
                if ((e = cudaMalloc(&p->d_wrank[s], sizeof(unsigned short) * (size_t)g->V)) != cudaSuccess) return fail("malloc", e);
                if ((e = cudaMalloc(&p->d_wtab64[s], sizeof(double) * vals.size())) != cudaSuccess) return fail("malloc", e);
                if ((e = cudaMemcpy(p->d_wrank[s], rk.data(), sizeof(unsigned short) * (size_t)g->V, cudaMemcpyHostToDevice)) != cudaSuccess)
                    return fail("memcpy", e);
                if ((e = cudaMemcpy(p->d_wtab64[s], vals.data(), sizeof(double) * vals.size(), cudaMemcpyHostToDevice)) != cudaSuccess)
                    return fail("memcpy", e);
            } else {
                wfast_ok = false;
            }
        } else if (weight_host && weight_host[s]) {
            std::vector<float> w(weight_host[s], weight_host[s] + g->V);
            if (!g->vmap.empty())
                for (int32_t i = 0; i < g->V; ++i) w[i] = weight_host[s][g->vmap[i]];
            if ((e = cudaMalloc(&p->d_weights[s], sizeof(float) * (size_t)g->V)) != cudaSuccess) return fail("malloc", e);
            if ((e = cudaMemcpy(p->d_weights[s], w.data(), sizeof(float) * (size_t)g->V, cudaMemcpyHostToDevice)) != cudaSuccess)
                return fail("memcpy", e);
            p->pipe_weights = 1;
            // weight ranks for the leader sweep: finite, non-negative weights with at most 65,535 distinct values
            // (density weights are a function of the neighbour count: a few dozen values)
            bool ok = true;
            for (float x : w) ok = ok && std::isfinite(x) && x >= 0.0f;
            std::vector<float> vals(w);
            std::sort(vals.begin(), vals.end());
            vals.resize((size_t)(std::unique(vals.begin(), vals.end()) - vals.begin()));
            if (ok && vals.size() <= 65535) {
                std::vector<unsigned short> rk((size_t)g->V);
                for (int32_t i = 0; i < g->V; ++i)
                    rk[i] = (unsigned short)(std::lower_bound(vals.begin(), vals.end(), w[i]) - vals.begin());
                if ((e = cudaMalloc(&p->d_wrank[s], sizeof(unsigned short) * (size_t)g->V)) != cudaSuccess) return fail("malloc", e);
                if ((e = cudaMalloc(&p->d_wtab[s], sizeof(float) * vals.size())) != cudaSuccess) return fail("malloc", e);
                if ((e = cudaMemcpy(p->d_wrank[s], rk.data(), sizeof(unsigned short) * (size_t)g->V, cudaMemcpyHostToDevice)) != cudaSuccess)
                    return fail("memcpy", e);
                if ((e = cudaMemcpy(p->d_wtab[s], vals.data(), sizeof(float) * vals.size(), cudaMemcpyHostToDevice)) != cudaSuccess)
                    return fail("memcpy", e);
            } else {
                wfast_ok = false;
            }
        }
        SurfDesc d{};
        d.indptr = g->d_indptr; d.indices = g->d_indices; d.ell = g->d_ell; d.ell_width = g->ell_width; d.ell_self = g->d_ell_self;
        d.sell = p->pipe_words ? g->d_sell : nullptr; d.sell_off = p->pipe_words ? g->d_sell_off : nullptr;
        d.wrank = p->d_wrank[s]; d.wtab = p->d_wtab[s]; d.weight64 = p->d_weights64[s]; d.wtab64 = p->d_wtab64[s];
        d.powE = g->d_powE; d.weight = p->d_weights[s]; d.vmap = g->d_vmap; d.col_off = col_offset[s]; d.V = g->V; d.H = g->H;
        d.directed = g->symmetric ? 0 : 1;
        descs[s] = d;
    }
    {
        const char *wm = getenv("TMB_PIPE_WEIGHTS"); // "class" forces the per-vertex class path (A/B measurements, tests)
        p->pipe_wfast = p->pipe_weights && wfast_ok && !(wm && strcmp(wm, "class") == 0);
    }
    if ((e = cudaMalloc(&p->d_surfs, sizeof(SurfDesc) * S)) != cudaSuccess) return fail("malloc", e);
    if ((e = cudaMalloc(&p->d_order, sizeof(int32_t) * S)) != cudaSuccess) return fail("malloc", e);
    if ((e = cudaMalloc(&p->d_counter, sizeof(int))) != cudaSuccess) return fail("malloc", e);
    {
        const char *pt = getenv("TMB_PHASE_TIMING");
        if (pt && pt[0] == '1') {
            if ((e = cudaMalloc(&p->d_timing, sizeof(unsigned long long) * 272)) != cudaSuccess) return fail("malloc", e);
            cudaMemset(p->d_timing, 0, sizeof(unsigned long long) * 272);
        }
    }
    if ((e = cudaMalloc(&p->d_workspace, p->slot_stride * (size_t)p->num_slots)) != cudaSuccess) return fail("malloc workspace", e);
    if ((e = cudaMemcpy(p->d_surfs, descs.data(), sizeof(SurfDesc) * S, cudaMemcpyHostToDevice)) != cudaSuccess) return fail("memcpy", e);
    if ((e = cudaMemcpy(p->d_order, order.data(), sizeof(int32_t) * S, cudaMemcpyHostToDevice)) != cudaSuccess) return fail("memcpy", e);
    *out = p;
    return 0;
}

extern "C" int tmb_plan_destroy(tmb_plan *p) {
    if (!p) return 0;
    DeviceGuard guard(p->device);
    cudaFree(p->d_pipe); cudaFree(p->d_pipe_slots); cudaFree(p->d_tabs);
    p->d_pipe = p->d_pipe_slots = p->d_tabs = nullptr;
    if (p->d_timing) {
        unsigned long long t[272];
        if (cudaMemcpy(t, p->d_timing, sizeof(t), cudaMemcpyDeviceToHost) == cudaSuccess) {
            static const char *names_v2[7] = {"tables", "levels+sort", "X: P2bc+P1", "Y: P2a", "tail", "node walk", "finalize"};
            static const char *names_v3[7] = {"tables", "levels+sort", "ascent+init", "I1: unions", "I2: sizes", "node walk", "finalize"};
            static const char *names_pipe[7] = {"S: init+sort", "S: (loop rest)", "S: outputs", "S: F1 unions", "S: F2 sizes", "S: F3 incr", "S: fold"};
            const char **names = p->pipe_ok ? names_pipe : p->use_basin ? names_v3 : names_v2;
            if (p->pipe_ok)
                fprintf(stderr, "[tmb pipeline] maps swept %llu (mean basins %.0f, candidate unions %.0f, classes %.0f), flagged for the sweep kernel %llu\n",
                        t[13], t[13] ? (double)t[11] / t[13] : 0.0, t[13] ? (double)t[12] / t[13] : 0.0,
                        t[13] ? (double)t[7] / t[13] : 0.0, t[10]);
            double tot = 0;
            for (int i = 0; i < 7; ++i) tot += (double)t[i];
            tot += (double)t[8] + (double)t[9];
            fprintf(stderr, "[tmb phase timing] total %.3e cycles, nodes %llu\n", tot, t[7]);
            for (int i = 0; i < 7; ++i) fprintf(stderr, "  %-12s %6.2f%%\n", names[i], tot > 0 ? 100.0 * t[i] / tot : 0.0);
            if (p->pipe_ok)
                fprintf(stderr, "  init detail (%% of total): setup %.2f, pair hist %.2f, vertex hist %.2f, scans %.2f, pair scatter %.2f (rest of init: births + vertex scatter)\n",
                        100.0 * t[20] / tot, 100.0 * t[21] / tot, 100.0 * t[22] / tot, 100.0 * t[23] / tot, 100.0 * t[24] / tot);

        }
        if (p->use_basin || p->pipe_ok) {
            double tot = 0;
            for (int i = 0; i < 7; ++i) tot += (double)t[i];
            fprintf(stderr, "  per-level share of I1 / I2 (pipeline: F2 / F3+F1) (levels 1..127, %% of total):\n");
            for (int l = 1; l < 128; ++l)
                if (t[16 + l] || t[144 + l])
                    fprintf(stderr, "   L%-3d %5.2f %5.2f\n", l, 100.0 * t[16 + l] / tot, 100.0 * t[144 + l] / tot);
        }
        cudaFree(p->d_timing);
    }
    for (float *w : p->d_weights) cudaFree(w);
    for (unsigned short *w : p->d_wrank) cudaFree(w);
    for (float *w : p->d_wtab) cudaFree(w);
    for (double *w : p->d_weights64) cudaFree(w);
    for (double *w : p->d_wtab64) cudaFree(w);
    cudaFree(p->d_surfs); cudaFree(p->d_order); cudaFree(p->d_counter); cudaFree(p->d_workspace);
    delete p;
    return 0;
}

struct TableSet {
    const int32_t *ns = nullptr;
    const float *delta = nullptr;
    const float *T = nullptr;
    const float *HH = nullptr;
    const int32_t *status = nullptr;
    const float *scale = nullptr;   // optional factor of the scaled maxima (default: delta)
};

static size_t al256(size_t x) { return (x + 255) / 256 * 256; }

// per-item bytes of the streaming pipeline's buffers
static size_t pipe_item_bytes(const tmb_plan *p, bool with_table) {
    const size_t words = (size_t)std::max(1, p->pipe_words);
    return al256((size_t)p->pipe_vstride) + (2 + words) * al256(sizeof(int) * (size_t)p->pipe_vstride) + 16 + 1024 + al256((p->pipe_wfast ? 4 : 2) * (size_t)p->pipe_vstride) +
           al256((size_t)p->pipe_nbcap) +
           al256(sizeof(unsigned long long) * (size_t)p->pipe_paircap) +
           (with_table ? al256(sizeof(unsigned) * (size_t)p->pipe_tabcap) : 0);
}

// (re)allocate the pipeline buffers for at least `items` work items; returns the number of items that fit
// the memory budget (a third of the free HBM), 0 when not even one does.
static int pipe_ensure(tmb_plan *p, int items, bool need_table) {
    if (p->pipe_vstride == 0) {
        p->pipe_vstride = ((int64_t)p->Vmax + 127) / 128 * 128;
        p->pipe_nbcap = (int)std::min<int64_t>(p->pipe_vstride, std::max<int64_t>(4096, p->pipe_vstride / 8));
        p->pipe_paircap = (int)p->pipe_vstride;
        p->pipe_tabcap = 8 * p->pipe_vstride;
        // capacity overrides (tests: force the over-capacity path, whose maps are redone by tfce_basin_kernel)
        if (const char *e = getenv("TMB_PIPE_NBCAP")) { const int v = atoi(e); if (v >= 1) p->pipe_nbcap = std::min<int>(v, p->pipe_nbcap); }
        if (const char *e = getenv("TMB_PIPE_PAIRCAP")) { const int v = atoi(e); if (v >= 1) p->pipe_paircap = std::min<int>(v, p->pipe_paircap); }
    }
    if (!p->d_pipe_slots) {
        p->pipe_slot_stride = pipe_slot_bytes(p->Vmax, p->pipe_nbcap, p->pipe_paircap);
        if (cudaMalloc(&p->d_pipe_slots, p->pipe_slot_stride * (size_t)p->pipe_slots) != cudaSuccess) {
            cudaGetLastError();
            p->d_pipe_slots = nullptr;
            return 0;
        }
    }
    const bool table = need_table || p->pipe_has_table;
    if (items <= p->pipe_items && (!need_table || p->pipe_has_table)) return items;
    const size_t per = pipe_item_bytes(p, table);
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) { cudaGetLastError(); return p->pipe_has_table || !need_table ? p->pipe_items : 0; }
    const size_t have = (size_t)p->pipe_items * pipe_item_bytes(p, p->pipe_has_table != 0);
    const size_t fit = (free_b + have) / 3 / per;
    int want = (int)std::min<size_t>((size_t)std::max(items, p->pipe_items), fit);
    if (want <= p->pipe_items && (!need_table || p->pipe_has_table)) return p->pipe_items;
    cudaFree(p->d_pipe);
    p->d_pipe = nullptr;
    p->pipe_items = 0;
    p->pipe_has_table = 0;
    if (want < 1) return 0;
    if (cudaMalloc(&p->d_pipe, per * (size_t)want + 8192) != cudaSuccess) { cudaGetLastError(); return 0; }
    p->pipe_items = want;
    p->pipe_has_table = table ? 1 : 0;
    return std::min(want, items);
}

static int tabs_ensure(tmb_plan *p, int items) {
    if (items <= p->tab_items) return 0;
    cudaFree(p->d_tabs);
    p->d_tabs = nullptr;
    p->tab_items = 0;
    const size_t cnt = (size_t)items * 2;
    TMB_CUDA(cudaMalloc(&p->d_tabs, cnt * (sizeof(float) * 2 + sizeof(int32_t) * 2 + sizeof(float) * 256)));
    p->tab_items = items;
    return 0;
}

static int plan_launch_sweep(tmb_plan *p, const float *stat, int64_t ld, int B, int two_sided, int accumulate,
                             float *max_dev, float *tfce_pos, float *tfce_neg, int32_t *status, int stop_level,
                             int32_t *labels, int32_t *extents, float *thr, cudaStream_t stream, const TableSet *tabs,
                             const int *only_flagged) {
    SweepParams sp{};
    if (tabs) { sp.tab_ns = tabs->ns; sp.tab_delta = tabs->delta; sp.tab_T = tabs->T; sp.tab_HH = tabs->HH; sp.tab_status = tabs->status; sp.tab_scale = tabs->scale; }
    sp.surfs = p->d_surfs; sp.surf_order = p->d_order; sp.S = p->S; sp.B = B; sp.two_sided = two_sided;
    sp.accumulate = accumulate; sp.stat = stat; sp.ld = ld; sp.max_out = max_dev; sp.tfce_pos = tfce_pos;
    sp.tfce_neg = tfce_neg; sp.status = status; sp.stop_level = stop_level; sp.labels = labels;
    sp.extents = extents; sp.threshold_out = thr; sp.workspace = p->d_workspace; sp.slot_stride = p->slot_stride;
    sp.Vmax = p->Vmax; sp.work_counter = p->d_counter; sp.timing = only_flagged ? nullptr : p->d_timing;
    sp.only_flagged = only_flagged;
    {
        const char *pc = getenv("TMB_PARENT_UNCACHED");
        sp.flags = ((pc && pc[0] == '1') ? 0 : 1) | (p->use_basin ? 2 : 0) | (p->internal_order ? 4 : 0); // see SweepParams::flags
    }
    return launch_tfce_sweep(sp, p->num_slots, stream);
}

static int plan_launch(tmb_plan *p, const float *stat, int64_t ld, int B, int two_sided, int accumulate,
                       float *max_dev, float *tfce_pos, float *tfce_neg, int32_t *status, int stop_level,
                       int32_t *labels, int32_t *extents, float *thr, cudaStream_t stream,
                       const TableSet *tabs = nullptr) {
    const bool pipeline = p->pipe_ok && !accumulate && stop_level < 0 && B > 0;
    int fit = 0;
    const bool class_path = tfce_pos || tfce_neg || (p->pipe_weights && !p->pipe_wfast);
    if (pipeline) fit = pipe_ensure(p, B * p->S, class_path) / p->S; // statistic rows per pipeline pass
    if (pipeline && fit >= 1)
        if (const char *wv = getenv("TMB_PIPE_WAVE")) { const int v = atoi(wv); if (v >= 1) fit = std::min(fit, v); } // experiments
    if (!pipeline || fit < 1)
        return plan_launch_sweep(p, stat, ld, B, two_sided, accumulate, max_dev, tfce_pos, tfce_neg, status, stop_level,
                                 labels, extents, thr, stream, tabs, nullptr);
    // ---- streaming pipeline, `fit` rows at a time; maps it flags as over capacity are redone by the sweep kernel
    TableSet dev_tabs;
    if (!tabs) {
        // exact_pow = False: correctly rounded tables built on the device (no host round trip)
        if (tabs_ensure(p, B * p->S)) return 1;
        const size_t cnt = (size_t)B * p->S * 2;
        float *mx = reinterpret_cast<float *>(p->d_tabs);
        float *delta = mx + cnt;
        int32_t *ns = reinterpret_cast<int32_t *>(delta + cnt);
        int32_t *st = ns + cnt;
        float *T = reinterpret_cast<float *>(st + cnt);
        float *HH = T + cnt * 128;
        if (launch_tfce_maxima(p->d_surfs, p->S, stat, ld, B, mx, stream)) return 1;
        if (launch_tfce_tables(p->d_surfs, p->S, (int)cnt, mx, two_sided, ns, delta, T, HH, st, stream)) return 1;
        dev_tabs.ns = ns; dev_tabs.delta = delta; dev_tabs.T = T; dev_tabs.HH = HH; dev_tabs.status = st;
        tabs = &dev_tabs;
    }
    const size_t vs = (size_t)p->pipe_vstride;
    for (int b0 = 0; b0 < B; b0 += fit) {
        const int Bc = std::min(fit, B - b0);
        const size_t items = (size_t)Bc * p->S;
        const size_t eoff = (size_t)b0 * p->S * 2;
        PipeParams pp{};
        pp.surfs = p->d_surfs; pp.surf_order = p->d_order; pp.S = p->S; pp.B = Bc; pp.two_sided = two_sided;
        pp.stat = stat + (size_t)b0 * ld; pp.ld = ld;
        pp.tab_ns = tabs->ns + eoff; pp.tab_delta = tabs->delta + eoff; pp.tab_T = tabs->T + eoff * 128;
        pp.tab_HH = tabs->HH + eoff * 128; pp.tab_status = tabs->status + eoff;
        pp.tab_scale = tabs->scale ? tabs->scale + eoff : nullptr;
        pp.flags = p->internal_order ? 4 : 0;
        pp.Vmax = p->Vmax; pp.vstride = p->pipe_vstride;
        char *base = p->d_pipe;
        pp.lev8 = reinterpret_cast<unsigned char *>(base); base += al256(items * vs);
        pp.up = reinterpret_cast<int *>(base); base += al256(items * vs * sizeof(int));
        pp.emask = reinterpret_cast<unsigned *>(base); base += al256(items * vs * sizeof(int) * (size_t)std::max(1, p->pipe_words));
        pp.basin = reinterpret_cast<int *>(base); base += al256(items * vs * sizeof(int));
        pp.meta = reinterpret_cast<int *>(base); base += al256(items * 16);
        pp.lhist = reinterpret_cast<int *>(base); base += al256(items * 1024);
        pp.vlist = reinterpret_cast<unsigned short *>(base); base += al256(items * vs * (p->pipe_wfast ? 4 : 2));
        pp.blev = reinterpret_cast<unsigned char *>(base); base += al256(items * (size_t)p->pipe_nbcap);
        pp.pairs = reinterpret_cast<unsigned long long *>(base); base += al256(items * (size_t)p->pipe_paircap * sizeof(unsigned long long));
        pp.table = reinterpret_cast<unsigned *>(base);
        pp.nbcap = p->pipe_nbcap; pp.paircap = p->pipe_paircap; pp.tabcap = p->pipe_tabcap;
        pp.max_out = max_dev ? max_dev + eoff : nullptr;
        pp.tfce_pos = tfce_pos ? tfce_pos + (size_t)b0 * ld : nullptr;
        pp.tfce_neg = tfce_neg ? tfce_neg + (size_t)b0 * ld : nullptr;
        pp.status = status ? status + eoff : nullptr;
        pp.want_vertex_pass = (tfce_pos || tfce_neg || (p->pipe_weights && !p->pipe_wfast)) ? 1 : 0;
        pp.weighted = (p->pipe_wfast && !pp.want_vertex_pass) ? 1 : 0;
        pp.slot_ws = p->d_pipe_slots; pp.slot_stride = p->pipe_slot_stride; pp.work_counter = p->d_counter;
        pp.timing = p->d_timing;
        pp.max_degree = 0;
        pp.sell_words = p->pipe_words;
        pp.narrow_slots = p->pipe_narrow; pp.narrow_max_degree = p->pipe_narrow_degree; pp.z0 = 0;
        pp.ell_self = 1;
        for (int s = 0; s < p->S; ++s) if (!p->graphs[s]->d_ell_self) pp.ell_self = 0;
        for (int s = 0; s < p->S; ++s) pp.max_degree = std::max(pp.max_degree, (int)p->graphs[s]->max_degree);
        if (launch_tfce_pipeline(pp, p->pipe_slots, p->pipe_sms, stream)) return 1;
        TableSet sub;
        sub.ns = pp.tab_ns; sub.delta = pp.tab_delta; sub.T = pp.tab_T; sub.HH = pp.tab_HH; sub.status = pp.tab_status;
        sub.scale = pp.tab_scale;
        if (plan_launch_sweep(p, pp.stat, ld, Bc, two_sided, 0, pp.max_out, pp.tfce_pos, pp.tfce_neg, pp.status, -1, nullptr,
                              nullptr, nullptr, stream, &sub, pp.meta))
            return 1;
    }
    return 0;
}

extern "C" int tmb_plan_run(tmb_plan *p, const float *stat_dev, int64_t ld, int B, int two_sided, float *max_dev,
                            float *tfce_pos_dev, float *tfce_neg_dev, int32_t *status_dev, void *stream) {
    TMB_REQUIRE(p && stat_dev && B >= 0, "tmb_plan_run: bad arguments");
    TMB_REQUIRE(max_dev || tfce_pos_dev || tfce_neg_dev, "tmb_plan_run: no output requested");
    for (int s = 0; s < p->S; ++s)
        TMB_REQUIRE(p->col_offset[s] + p->graphs[s]->V <= ld, "tmb_plan_run: surface %d exceeds row length %lld", s,
                    (long long)ld);
    TMB_ON_DEVICE(p->device);
    return plan_launch(p, stat_dev, ld, B, two_sided, 0, max_dev, tfce_pos_dev, tfce_neg_dev, status_dev, -1, nullptr,
                       nullptr, nullptr, (cudaStream_t)stream);
}


// Threshold sequence and height terms of fast_tfce.hpp:32-39,70 computed on the HOST with the very
// call the reference compiles to (std::pow(float, float) == libm powf, which is not correctly
// rounded: about 6 in 10^4 thresholds differ from the exact square by one ulp).
static void host_tables(float mx, float H, int32_t *ns_out, float *delta_out, float *T, float *HH,
                        int32_t *status_out) {
    int ns = 0, st = 0;
    volatile float d = 0.f;
    if (mx >= 0.f) {
        d = mx / 100;
        if (d == 0.f) {
            st = TMB_MAP_MAX_IS_ZERO;
        } else {
            for (volatile float t = mx; t >= 0.f; t -= d) {
                if (ns == 128) { st = TMB_MAP_STEP_OVERFLOW; ns = 0; break; }
                const float tc = t;
                T[ns] = tc;
                HH[ns] = std::pow(tc, H);
                ++ns;
            }
        }
    }
    *ns_out = ns; *delta_out = d; *status_out = st;
}

extern "C" int tmb_threshold_tables(const float *maxima_host, const float *H_host, int count, int32_t *ns_host,
                                    float *delta_host, float *T_host, float *HH_host, int32_t *status_host) {
    TMB_REQUIRE(maxima_host && H_host && ns_host && delta_host && T_host && HH_host && status_host && count >= 0,
                "tmb_threshold_tables: bad arguments");
    // ~100 powf calls per entry: a block of 512 two-sided shuffles on two surfaces has 2,048 entries (~3 ms on one core,
    // on the critical path of the host that feeds the GPU), so the entries are split over a few threads
    auto work = [&](int lo, int hi) {
        for (int i = lo; i < hi; ++i)
            host_tables(maxima_host[i], H_host[i], ns_host + i, delta_host + i, T_host + (size_t)i * 128,
                        HH_host + (size_t)i * 128, status_host + i);
    };
    int nt = (int)std::min<unsigned>(8u, std::max(1u, std::thread::hardware_concurrency() / 2));
    if (const char *lw = getenv("LOCAL_WORLD_SIZE")) {   // one process per GPU (torchrun): share the host cores between the ranks
        const int ranks = std::max(1, atoi(lw));
        nt = std::max(1, std::min(nt, (int)std::thread::hardware_concurrency() / (2 * ranks)));
    }
    if (const char *e = getenv("TMB_TABLE_THREADS")) nt = std::max(1, atoi(e));
    if (count < 256 || nt == 1) { work(0, count); return 0; }
    std::vector<std::thread> pool;
    const int per = (count + nt - 1) / nt;
    for (int t = 1; t < nt; ++t)
        if (t * per < count) pool.emplace_back(work, t * per, std::min(count, (t + 1) * per));
    work(0, std::min(count, per));
    for (auto &th : pool) th.join();
    return 0;
}

extern "C" int tmb_plan_maxima(tmb_plan *p, const float *stat_dev, int64_t ld, int B, float *max_dev, void *stream) {
    TMB_REQUIRE(p && stat_dev && max_dev && B >= 0, "tmb_plan_maxima: bad arguments");
    TMB_ON_DEVICE(p->device);
    return launch_tfce_maxima(p->d_surfs, p->S, stat_dev, ld, B, max_dev, (cudaStream_t)stream);
}

extern "C" int tmb_plan_run_tables(tmb_plan *p, const float *stat_dev, int64_t ld, int B, int two_sided,
                                   const int32_t *ns_dev, const float *delta_dev, const float *T_dev,
                                   const float *HH_dev, const int32_t *tstatus_dev, const float *scale_dev, float *max_dev,
                                   float *tfce_pos_dev, float *tfce_neg_dev, int32_t *status_dev, void *stream) {
    TMB_REQUIRE(p && stat_dev && B >= 0 && ns_dev && delta_dev && T_dev && HH_dev && tstatus_dev,
                "tmb_plan_run_tables: bad arguments");
    TMB_REQUIRE(max_dev || tfce_pos_dev || tfce_neg_dev, "tmb_plan_run_tables: no output requested");
    for (int s = 0; s < p->S; ++s)
        TMB_REQUIRE(p->col_offset[s] + p->graphs[s]->V <= ld, "tmb_plan_run_tables: surface %d exceeds row length %lld",
                    s, (long long)ld);
    TMB_ON_DEVICE(p->device);
    TableSet t; t.ns = ns_dev; t.delta = delta_dev; t.T = T_dev; t.HH = HH_dev; t.status = tstatus_dev; t.scale = scale_dev;
    return plan_launch(p, stat_dev, ld, B, two_sided, 0, max_dev, tfce_pos_dev, tfce_neg_dev, status_dev, -1, nullptr,
                       nullptr, nullptr, (cudaStream_t)stream, &t);
}

static int ensure_self_plan(tmb_graph *g) {
    if (g->self_plan) return 0;
    const int64_t off = 0;
    tmb_graph *gs[1] = {g};
    if (tmb_plan_create(g->device, 1, gs, &off, nullptr, nullptr, 1, &g->self_plan)) return 1;
    TMB_CUDA(cudaMalloc(&g->d_image, sizeof(float) * (size_t)g->V));
    TMB_CUDA(cudaMalloc(&g->d_enhn, sizeof(float) * (size_t)g->V));
    TMB_CUDA(cudaMalloc(&g->d_labels, sizeof(int32_t) * (size_t)g->V));
    TMB_CUDA(cudaMalloc(&g->d_extents, sizeof(int32_t) * (size_t)g->V));
    TMB_CUDA(cudaMalloc(&g->d_status, sizeof(int32_t) * 2));
    TMB_CUDA(cudaMalloc(&g->d_thr, sizeof(float)));
    TMB_CUDA(cudaMalloc(&g->d_tabs, sizeof(int32_t) * 4 + sizeof(float) * (2 + 4 * 128)));
    return 0;
}

// tables for ONE host image (positive side only): exact libm arithmetic, uploaded to g->d_tabs
static int upload_single_tables(tmb_graph *g, const float *image_host, TableSet *t) {
    float mx = -INFINITY;
    for (int32_t v = 0; v < g->V; ++v) mx = std::fmax(mx, image_host[v]); // fmax ignores NaN
    struct { int32_t ns[2]; int32_t st[2]; float delta[2]; float T[2][128]; float HH[2][128]; } h;
    memset(&h, 0, sizeof(h));
    host_tables(mx, g->H, &h.ns[0], &h.delta[0], h.T[0], h.HH[0], &h.st[0]);
    TMB_CUDA(cudaMemcpy(g->d_tabs, &h, sizeof(h), cudaMemcpyHostToDevice));
    char *base = g->d_tabs;
    t->ns = reinterpret_cast<const int32_t *>(base);
    t->status = reinterpret_cast<const int32_t *>(base + sizeof(int32_t) * 2);
    t->delta = reinterpret_cast<const float *>(base + sizeof(int32_t) * 4);
    t->T = reinterpret_cast<const float *>(base + sizeof(int32_t) * 4 + sizeof(float) * 2);
    t->HH = t->T + 2 * 128;
    return 0;
}

extern "C" int tmb_tfce_run(tmb_graph *g, const float *image_host, float *enhn_host, int *map_status) {
    TMB_REQUIRE(g && image_host && enhn_host, "tmb_tfce_run: null pointer");
    TMB_ON_DEVICE(g->device);
    if (ensure_self_plan(g)) return 1;
    const size_t bytes = sizeof(float) * (size_t)g->V;
    TMB_CUDA(cudaMemcpy(g->d_image, image_host, bytes, cudaMemcpyHostToDevice));
    TMB_CUDA(cudaMemcpy(g->d_enhn, enhn_host, bytes, cudaMemcpyHostToDevice));
    // one-sided; accumulate into enhn like the reference's `enhn[v] += increment`.  Callers always pass
    // zeros (pyfunc.py:110-111,123): then the per-node walk is exact and cheaper; a non-zero enhn takes
    // the per-vertex walk that starts every sum from enhn[v].
    bool all_zero = true;
    for (int32_t v = 0; v < g->V && all_zero; ++v) all_zero = (enhn_host[v] == 0.0f) && !std::signbit(enhn_host[v]);
    TableSet tabs;
    if (upload_single_tables(g, image_host, &tabs)) return 1;
    if (plan_launch(g->self_plan, g->d_image, g->V, 1, 0, all_zero ? 0 : 1, nullptr, g->d_enhn, nullptr, g->d_status,
                    -1, nullptr, nullptr, nullptr, nullptr, &tabs))
        return 1;
    TMB_CUDA(cudaMemcpy(enhn_host, g->d_enhn, bytes, cudaMemcpyDeviceToHost));
    int32_t st[2] = {0, 0};
    TMB_CUDA(cudaMemcpy(st, g->d_status, sizeof(st), cudaMemcpyDeviceToHost));
    if (map_status) *map_status = st[0];
    return 0;
}

extern "C" int tmb_tfce_components(tmb_graph *g, const float *image_host, int level, int32_t *labels_host,
                                   int32_t *extents_host, float *threshold_out) {
    TMB_REQUIRE(g && image_host && labels_host && extents_host && level >= 0, "tmb_tfce_components: bad arguments");
    TMB_ON_DEVICE(g->device);
    if (ensure_self_plan(g)) return 1;
    const size_t bytes = sizeof(float) * (size_t)g->V;
    TMB_CUDA(cudaMemcpy(g->d_image, image_host, bytes, cudaMemcpyHostToDevice));
    TableSet tabs;
    if (upload_single_tables(g, image_host, &tabs)) return 1;
    if (plan_launch(g->self_plan, g->d_image, g->V, 1, 0, 0, nullptr, nullptr, nullptr, nullptr, level, g->d_labels,
                    g->d_extents, g->d_thr, nullptr, &tabs))
        return 1;
    TMB_CUDA(cudaMemcpy(labels_host, g->d_labels, sizeof(int32_t) * (size_t)g->V, cudaMemcpyDeviceToHost));
    TMB_CUDA(cudaMemcpy(extents_host, g->d_extents, sizeof(int32_t) * (size_t)g->V, cudaMemcpyDeviceToHost));
    {
        // the kernel labels a component by the INTERNAL index of its union-find root; the canonical label
        // is the smallest caller index of the component.  labels_host[] currently holds internal ids.
        std::vector<int32_t> canon((size_t)g->V, INT32_MAX);
        for (int32_t o = 0; o < g->V; ++o)
            if (labels_host[o] >= 0) canon[labels_host[o]] = std::min(canon[labels_host[o]], o);
        for (int32_t o = 0; o < g->V; ++o)
            if (labels_host[o] >= 0) labels_host[o] = canon[labels_host[o]];
    }
    if (threshold_out) TMB_CUDA(cudaMemcpy(threshold_out, g->d_thr, sizeof(float), cudaMemcpyDeviceToHost));
    return 0;
}
