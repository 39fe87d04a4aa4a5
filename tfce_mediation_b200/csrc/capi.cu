// C ABI of libtfce_b200.so: handles (graph, plan) and the TFCE entry points.
// The GLM entry points live in glm_kernels.cu, the voxel adjacency builder in adjacency_kernels.cu.
#include "common.cuh"
#include "../../include/tfce_b200.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <numeric>
#include <string>
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
    tmb_plan *self_plan = nullptr; // lazily created single-surface plan for tmb_tfce_run
    float *d_image = nullptr, *d_enhn = nullptr;
    int32_t *d_labels = nullptr, *d_extents = nullptr, *d_status = nullptr;
    float *d_thr = nullptr;
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
    char *d_workspace = nullptr;
    size_t slot_stride = 0;
    int *d_counter = nullptr;
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

extern "C" int tmb_graph_create(int device, int32_t V, const int64_t *indptr, const int32_t *indices, float H,
                                float E, tmb_graph **out) {
    TMB_REQUIRE(out && indptr && V > 0, "tmb_graph_create: bad arguments (V=%d)", V);
    TMB_REQUIRE(indptr[0] == 0, "tmb_graph_create: indptr[0] must be 0");
    for (int32_t v = 0; v < V; ++v)
        TMB_REQUIRE(indptr[v + 1] >= indptr[v], "tmb_graph_create: indptr not monotone at %d", v);
    const int64_t nnz = indptr[V];
    TMB_REQUIRE(nnz == 0 || indices, "tmb_graph_create: indices is null");
    for (int64_t e = 0; e < nnz; ++e)
        TMB_REQUIRE(indices[e] >= 0 && indices[e] < V, "tmb_graph_create: neighbour index %d out of range [0,%d)",
                    indices[e], V);
    TMB_CUDA(cudaSetDevice(device));
    tmb_graph *g = new tmb_graph();
    g->device = device; g->V = V; g->nnz = nnz; g->H = H; g->E = E;
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
    if ((e = cudaMemcpy(g->d_indptr, indptr, sizeof(int64_t) * ((size_t)V + 1), cudaMemcpyHostToDevice)) != cudaSuccess)
        return fail("memcpy", e);
    if (nnz && (e = cudaMemcpy(g->d_indices, indices, sizeof(int32_t) * (size_t)nnz, cudaMemcpyHostToDevice)) != cudaSuccess)
        return fail("memcpy", e);
    if ((e = cudaMemcpy(g->d_powE, powE.data(), sizeof(double) * ((size_t)V + 1), cudaMemcpyHostToDevice)) != cudaSuccess)
        return fail("memcpy", e);
    *out = g;
    return 0;
}

extern "C" int tmb_graph_destroy(tmb_graph *g) {
    if (!g) return 0;
    cudaSetDevice(g->device);
    if (g->self_plan) tmb_plan_destroy(g->self_plan);
    cudaFree(g->d_indptr); cudaFree(g->d_indices); cudaFree(g->d_powE);
    cudaFree(g->d_image); cudaFree(g->d_enhn); cudaFree(g->d_labels); cudaFree(g->d_extents);
    cudaFree(g->d_status); cudaFree(g->d_thr);
    delete g;
    return 0;
}

extern "C" int tmb_graph_num_vertices(const tmb_graph *g, int32_t *V, int64_t *nnz) {
    TMB_REQUIRE(g, "tmb_graph_num_vertices: null graph");
    if (V) *V = g->V;
    if (nnz) *nnz = g->nnz;
    return 0;
}

extern "C" int tmb_plan_create(int device, int S, tmb_graph *const *graphs, const int64_t *col_offset,
                               const float *const *weight_host, int max_slots, tmb_plan **out) {
    TMB_REQUIRE(out && graphs && col_offset && S > 0, "tmb_plan_create: bad arguments");
    for (int s = 0; s < S; ++s) {
        TMB_REQUIRE(graphs[s], "tmb_plan_create: graph %d is null", s);
        TMB_REQUIRE(graphs[s]->device == device, "tmb_plan_create: graph %d lives on device %d, plan on %d", s,
                    graphs[s]->device, device);
        TMB_REQUIRE(col_offset[s] >= 0, "tmb_plan_create: negative column offset");
    }
    TMB_CUDA(cudaSetDevice(device));
    tmb_plan *p = new tmb_plan();
    p->device = device; p->S = S;
    p->graphs.assign(graphs, graphs + S);
    p->col_offset.assign(col_offset, col_offset + S);
    p->d_weights.assign(S, nullptr);
    std::vector<SurfDesc> descs(S);
    std::vector<int32_t> order(S);
    std::iota(order.begin(), order.end(), 0);
    for (int s = 0; s < S; ++s) p->Vmax = std::max(p->Vmax, graphs[s]->V);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return graphs[a]->V > graphs[b]->V; });
    cudaDeviceProp prop;
    cudaError_t e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) { set_error("tmb_plan_create: %s", cudaGetErrorString(e)); delete p; return 1; }
    p->num_slots = max_slots > 0 ? max_slots : prop.multiProcessorCount * 2;
    p->slot_stride = tfce_slot_bytes(p->Vmax);
    auto fail = [&](const char *what, cudaError_t err) {
        set_error("tmb_plan_create: %s: %s", what, cudaGetErrorString(err));
        tmb_plan_destroy(p);
        return 1;
    };
    for (int s = 0; s < S; ++s) {
        const tmb_graph *g = graphs[s];
        if (weight_host && weight_host[s]) {
            if ((e = cudaMalloc(&p->d_weights[s], sizeof(float) * (size_t)g->V)) != cudaSuccess) return fail("malloc", e);
            if ((e = cudaMemcpy(p->d_weights[s], weight_host[s], sizeof(float) * (size_t)g->V, cudaMemcpyHostToDevice)) != cudaSuccess)
                return fail("memcpy", e);
        }
        descs[s] = SurfDesc{g->d_indptr, g->d_indices, g->d_powE, p->d_weights[s], col_offset[s], g->V, g->H};
    }
    if ((e = cudaMalloc(&p->d_surfs, sizeof(SurfDesc) * S)) != cudaSuccess) return fail("malloc", e);
    if ((e = cudaMalloc(&p->d_order, sizeof(int32_t) * S)) != cudaSuccess) return fail("malloc", e);
    if ((e = cudaMalloc(&p->d_counter, sizeof(int))) != cudaSuccess) return fail("malloc", e);
    if ((e = cudaMalloc(&p->d_workspace, p->slot_stride * (size_t)p->num_slots)) != cudaSuccess) return fail("malloc workspace", e);
    if ((e = cudaMemcpy(p->d_surfs, descs.data(), sizeof(SurfDesc) * S, cudaMemcpyHostToDevice)) != cudaSuccess) return fail("memcpy", e);
    if ((e = cudaMemcpy(p->d_order, order.data(), sizeof(int32_t) * S, cudaMemcpyHostToDevice)) != cudaSuccess) return fail("memcpy", e);
    *out = p;
    return 0;
}

extern "C" int tmb_plan_destroy(tmb_plan *p) {
    if (!p) return 0;
    cudaSetDevice(p->device);
    for (float *w : p->d_weights) cudaFree(w);
    cudaFree(p->d_surfs); cudaFree(p->d_order); cudaFree(p->d_counter); cudaFree(p->d_workspace);
    delete p;
    return 0;
}

static int plan_launch(tmb_plan *p, const float *stat, int64_t ld, int B, int two_sided, int accumulate,
                       float *max_dev, float *tfce_pos, float *tfce_neg, int32_t *status, int stop_level,
                       int32_t *labels, int32_t *extents, float *thr, cudaStream_t stream) {
    SweepParams sp{};
    sp.surfs = p->d_surfs; sp.surf_order = p->d_order; sp.S = p->S; sp.B = B; sp.two_sided = two_sided;
    sp.accumulate = accumulate; sp.stat = stat; sp.ld = ld; sp.max_out = max_dev; sp.tfce_pos = tfce_pos;
    sp.tfce_neg = tfce_neg; sp.status = status; sp.stop_level = stop_level; sp.labels = labels;
    sp.extents = extents; sp.threshold_out = thr; sp.workspace = p->d_workspace; sp.slot_stride = p->slot_stride;
    sp.Vmax = p->Vmax; sp.work_counter = p->d_counter;
    return launch_tfce_sweep(sp, p->num_slots, stream);
}

extern "C" int tmb_plan_run(tmb_plan *p, const float *stat_dev, int64_t ld, int B, int two_sided, float *max_dev,
                            float *tfce_pos_dev, float *tfce_neg_dev, int32_t *status_dev, void *stream) {
    TMB_REQUIRE(p && stat_dev && B >= 0, "tmb_plan_run: bad arguments");
    TMB_REQUIRE(max_dev || tfce_pos_dev || tfce_neg_dev, "tmb_plan_run: no output requested");
    for (int s = 0; s < p->S; ++s)
        TMB_REQUIRE(p->col_offset[s] + p->graphs[s]->V <= ld, "tmb_plan_run: surface %d exceeds row length %lld", s,
                    (long long)ld);
    TMB_CUDA(cudaSetDevice(p->device));
    return plan_launch(p, stat_dev, ld, B, two_sided, 0, max_dev, tfce_pos_dev, tfce_neg_dev, status_dev, -1, nullptr,
                       nullptr, nullptr, (cudaStream_t)stream);
}

static int ensure_self_plan(tmb_graph *g) {
    if (g->self_plan) return 0;
    TMB_CUDA(cudaSetDevice(g->device));
    const int64_t off = 0;
    tmb_graph *gs[1] = {g};
    if (tmb_plan_create(g->device, 1, gs, &off, nullptr, 1, &g->self_plan)) return 1;
    TMB_CUDA(cudaMalloc(&g->d_image, sizeof(float) * (size_t)g->V));
    TMB_CUDA(cudaMalloc(&g->d_enhn, sizeof(float) * (size_t)g->V));
    TMB_CUDA(cudaMalloc(&g->d_labels, sizeof(int32_t) * (size_t)g->V));
    TMB_CUDA(cudaMalloc(&g->d_extents, sizeof(int32_t) * (size_t)g->V));
    TMB_CUDA(cudaMalloc(&g->d_status, sizeof(int32_t) * 2));
    TMB_CUDA(cudaMalloc(&g->d_thr, sizeof(float)));
    return 0;
}

extern "C" int tmb_tfce_run(tmb_graph *g, const float *image_host, float *enhn_host, int *map_status) {
    TMB_REQUIRE(g && image_host && enhn_host, "tmb_tfce_run: null pointer");
    if (ensure_self_plan(g)) return 1;
    const size_t bytes = sizeof(float) * (size_t)g->V;
    TMB_CUDA(cudaMemcpy(g->d_image, image_host, bytes, cudaMemcpyHostToDevice));
    TMB_CUDA(cudaMemcpy(g->d_enhn, enhn_host, bytes, cudaMemcpyHostToDevice));
    // one-sided, accumulate into enhn like the reference's `enhn[v] += increment`
    if (plan_launch(g->self_plan, g->d_image, g->V, 1, 0, 1, nullptr, g->d_enhn, nullptr, g->d_status, -1, nullptr,
                    nullptr, nullptr, nullptr))
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
    if (ensure_self_plan(g)) return 1;
    const size_t bytes = sizeof(float) * (size_t)g->V;
    TMB_CUDA(cudaMemcpy(g->d_image, image_host, bytes, cudaMemcpyHostToDevice));
    if (plan_launch(g->self_plan, g->d_image, g->V, 1, 0, 0, nullptr, nullptr, nullptr, nullptr, level, g->d_labels,
                    g->d_extents, g->d_thr, nullptr))
        return 1;
    TMB_CUDA(cudaMemcpy(labels_host, g->d_labels, sizeof(int32_t) * (size_t)g->V, cudaMemcpyDeviceToHost));
    TMB_CUDA(cudaMemcpy(extents_host, g->d_extents, sizeof(int32_t) * (size_t)g->V, cudaMemcpyDeviceToHost));
    if (threshold_out) TMB_CUDA(cudaMemcpy(threshold_out, g->d_thr, sizeof(float), cudaMemcpyDeviceToHost));
    return 0;
}
