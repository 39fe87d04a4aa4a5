// Shared helpers for libtfce_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>

namespace tmb {

void set_error(const char *fmt, ...);
extern std::atomic<int64_t> g_launch_count;

inline void count_launch(int n = 1) { g_launch_count.fetch_add(n, std::memory_order_relaxed); }

#define TMB_CUDA(call)                                                                              \
    do {                                                                                            \
        cudaError_t e__ = (call);                                                                   \
        if (e__ != cudaSuccess) {                                                                   \
            ::tmb::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__,     \
                             __LINE__);                                                             \
            return 1;                                                                               \
        }                                                                                           \
    } while (0)

#define TMB_REQUIRE(cond, ...)                                                                      \
    do {                                                                                            \
        if (!(cond)) {                                                                              \
            ::tmb::set_error(__VA_ARGS__);                                                          \
            return 1;                                                                               \
        }                                                                                           \
    } while (0)

// Scoped current-device switch: every entry point runs on the device that owns its handle / buffers and leaves the
// calling thread's current device as it found it (the caller is a PyTorch process with its own notion of it).
struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) { cudaGetLastError(); prev = -1; }
        if (prev != dev) ok = cudaSetDevice(dev) == cudaSuccess;
        if (!ok) cudaGetLastError();
    }
    ~DeviceGuard() {
        int cur = -1;
        if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
    }
    DeviceGuard(const DeviceGuard &) = delete;
    DeviceGuard &operator=(const DeviceGuard &) = delete;
};
#define TMB_ON_DEVICE(dev)                                                                          \
    ::tmb::DeviceGuard guard__(dev);                                                                \
    TMB_REQUIRE(guard__.ok, "cannot switch to CUDA device %d", (int)(dev))

// device that owns a device pointer (-1: not a device pointer)
inline int device_of(const void *ptr) {
    cudaPointerAttributes a;
    if (!ptr || cudaPointerGetAttributes(&a, ptr) != cudaSuccess) { cudaGetLastError(); return -1; }
    return (a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged) ? a.device : -1;
}

#define TMB_DEVICE_OF(ptr, who)                                                                     \
    const int dev__ = ::tmb::device_of(ptr);                                                        \
    TMB_REQUIRE(dev__ >= 0, "%s: %s is not a device pointer", who, #ptr);                           \
    TMB_ON_DEVICE(dev__)

// ---- per-surface descriptor consumed by the TFCE kernels (device-resident array) -------------
struct SurfDesc {
    const int64_t *indptr;  // [V+1]
    const int32_t *indices; // [nnz]
    const int32_t *ell;     // [V * ell_width] fixed-width rows padded with -1 (32-byte aligned), or nullptr
    int32_t ell_width;      // 8, 16 or 32 when ell != nullptr
    const int32_t *ell_self; // width-8 rows padded with the vertex ITSELF instead of -1 (pipeline ascent: no pad test), or nullptr
    // sliced rows (SELL-32-4) for graphs wider than 32 neighbours or with very uneven degrees: the 32 vertices of a
    // slice share a width (their largest degree rounded up to 4); slot group j4 of the slice is 32 consecutive int4
    // (one per vertex), so a warp that owns the slice reads 512 contiguous bytes per group.  Pad = the vertex itself.
    const int4 *sell;       // [sell_off[nslices]] or nullptr (uploaded only for plans that use the wide kernels)
    const int32_t *sell_off; // [nslices + 1] first int4 of every slice
    const double *powE;     // [V+1]  pow((double)n, (double)E) tabulated with the host libm
    const float *weight;    // [V] or nullptr (internal vertex order)
    const double *weight64; // [V] or nullptr: float64 weights (non-low-RAM mmr builds its density weights with np.hstack
                            // from [], i.e. float64: the product is then fl32(double(fl32(tfce * delta)) * w), tm_func.py:83-91)
    // weighted maxima without the per-vertex pass (weights finite and >= 0, at most 65,535 distinct values): the
    // weight of a vertex as the RANK of its value among the surface's distinct weights, and the values by rank
    const unsigned short *wrank; // [V] or nullptr
    const float *wtab;           // [distinct weights] or nullptr
    const double *wtab64;        // the same for float64 weights
    const int32_t *vmap;    // [V] internal (locality-reordered) index -> caller's index, or nullptr
    int64_t col_off;        // first column of this surface in a statistic row
    int32_t V;
    float H;
    int32_t directed;       // adjacency is not symmetric: honour the reference's directional join rule
};

// fl32(fl32(tfce * scale) * w): the reference's order of operations (pyfunc.py:116-117, tm_func.py:173-174), with the
// float64 variant of the non-low-RAM mmr path (tm_func.py:83-91)
__device__ __forceinline__ float scaled_vertex_value(float val, float scale, const SurfDesc &sd, int v) {
    float sc = __fmul_rn(val, scale);
    if (sd.weight64) sc = __double2float_rn(__dmul_rn((double)sc, sd.weight64[v]));
    else if (sd.weight) sc = __fmul_rn(sc, sd.weight[v]);
    return sc;
}

// per-launch parameters of the sweep kernel
struct SweepParams {
    const SurfDesc *surfs;
    const int32_t *surf_order; // surfaces sorted by descending V (work items issue big ones first)
    int S;
    int B;
    int two_sided;
    int accumulate;   // acc starts from tfce_pos/tfce_neg contents (CreateAdjSet.run `+=` semantics)
    const float *stat;
    int64_t ld;
    float *max_out;   // [B, S, 2] or nullptr
    float *tfce_pos;  // [B, ld] or nullptr
    float *tfce_neg;  // [B, ld] or nullptr
    int32_t *status;  // [B, S, 2] or nullptr
    // component inspection (tmb_tfce_components): stop after this level and write labels/extents
    int stop_level;
    int32_t *labels;
    int32_t *extents;
    float *threshold_out;
    // optional host-computed threshold tables (exact libm powf, see tmb_threshold_tables):
    // entry e = (b*S + s)*2 + sign;  T/HH rows of 128 floats
    const int32_t *tab_ns;
    const float *tab_delta;
    const float *tab_T;
    const float *tab_HH;
    const int32_t *tab_status;
    const float *tab_scale;     // optional: factor of the scaled maximum per entry when it is NOT the threshold step
                                // (non-low-RAM mmr: thresholds from the maximum over ALL surfaces, scale from the surface's own)
    int flags;                  // bit 0: union-find pointer loads go through L1 (ld.ca) instead of ld.cg;
                                // bit 1: basin sweep (symmetric adjacency only)
                                // bit 2: statistic rows are in the graphs' internal vertex order (ignore vmap)
    unsigned long long *timing; // optional [8] per-phase cycle totals (development aid), or nullptr
    // workspace
    char *workspace;
    size_t slot_stride;
    int32_t Vmax;
    int *work_counter;
    const int *only_flagged;    // optional [B*S][4]: process only work items with only_flagged[item*4 + 2] != 0
                                // (maps the streaming pipeline could not take, see tfce_pipeline.cu)
};

// per-launch parameters of the streaming TFCE pipeline (tfce_pipeline.cu); work item = (row b, surface s)
struct PipeParams {
    const SurfDesc *surfs;
    const int32_t *surf_order;
    int S;
    int B;
    int two_sided;
    const float *stat;
    int64_t ld;
    // threshold tables, entry e = (b*S + s)*2 + sign (host-built exact-libm tables or pipe_tables_kernel)
    const int32_t *tab_ns;
    const float *tab_delta;
    const float *tab_T;
    const float *tab_HH;
    const int32_t *tab_status;
    const float *tab_scale; // optional, see SweepParams
    int flags;              // bit 2: statistic rows are in the graphs' internal vertex order
    int max_degree;         // largest vertex degree over the plan's graphs (triangle meshes: 6 -> the ascent kernel skips the two pad slots)
    int ell_self;           // every surface has self-padded width-8 rows (SurfDesc::ell_self)
    int sell_words;         // 0: fixed-width rows (ell); W > 0: sliced rows (sell) with W 32-bit words of earlier-neighbour mask per vertex
    int narrow_slots;       // mixed plans (sell_words > 0): the first narrow_slots surface slots of surf_order are triangle meshes
                            // with self-padded width-8 rows and run the fixed-width ascent / count kernels, the rest the sliced ones
    int narrow_max_degree;  // largest degree among those
    int z0;                 // surface-slot offset of this launch's grid (blockIdx.z + z0 indexes surf_order)
    int32_t Vmax;
    // per-item buffers
    int64_t vstride;        // elements per work item in the per-vertex arrays
    unsigned char *lev8;    // activation level | sign << 7; 0 = never active
    int *up;                // ascent target; -1 - basin for a peak; INT_MIN when inactive
    unsigned *emask;        // row slots holding an earlier-activated same-sign neighbour (ascent target excluded);
                            // sliced rows: sell_words words per vertex, word w of item i at ((i * sell_words + w) * vstride + v)
    int *basin;             // compact basin id, -1 when inactive
    int *meta;              // [items][4]: basins, candidate unions, over-capacity flag, unused
    int *lhist;             // [items][256]: [0,128) vertices per activation level (K_A), [128,256) the cursors with which
                            // K_C reserves each CTA's range of a level's vertex list (max-only maps)
    unsigned short *vlist;  // [items][vstride] basin of every active vertex, bucketed by activation level (K_C -> sweep);
                            // weighted: 32-bit entries basin | weight rank << 16
    unsigned char *blev;    // [items][nbcap] level | sign of each peak
    int nbcap;
    unsigned long long *pairs; // [items][paircap] (level << 48) | (basin << 24) | basin
    int paircap;
    unsigned *table;        // [items][tabcap]: [level][basin] vertex counts -> class ids -> TFCE values
    int64_t tabcap;
    // outputs
    float *max_out;
    float *tfce_pos;
    float *tfce_neg;
    int32_t *status;
    int want_vertex_pass;   // maps requested (or weights the leader sweep cannot take): values go through the table and K_G
    int front_cap;          // weighted sweep: entries a root's front may hold (<= 16; TMB_PIPE_FRONT lowers it in tests)
    int weighted;           // max-only maps with vertex weights: vertex lists carry {basin, weight rank} (4 bytes), the sweep
                            // keeps a Pareto front of (sum, weight) leaders per live root (pipe_sweep_max_kernel<.., true>)
    // sweep slots
    char *slot_ws;
    size_t slot_stride;
    int *work_counter;
    unsigned long long *timing;
};

int launch_tfce_sweep(const SweepParams &p, int num_slots, cudaStream_t stream);
int launch_tfce_maxima(const SurfDesc *surfs, int S, const float *stat, int64_t ld, int B, float *max_out,
                       cudaStream_t stream);
size_t tfce_slot_bytes_for(int32_t Vmax, int use_basin);
int launch_tfce_pipeline(const PipeParams &p, int num_slots, int sm_count, cudaStream_t stream);
int pipe_slots_per_sm(int Vmax); // sweep CTAs per SM for a plan whose largest surface has Vmax vertices (2, 4 or 8)
int launch_tfce_tables(const SurfDesc *surfs, int S, int count, const float *maxima, int two_sided, int32_t *ns,
                       float *delta, float *T, float *HH, int32_t *st, cudaStream_t stream);
size_t pipe_slot_bytes(int32_t Vmax, int nbcap, int paircap);
void tfce_sweep_geometry(int32_t Vmax, int use_basin, int *threads, int *ctas_per_sm, size_t *dyn_smem);

} // namespace tmb
