// Voxel adjacency builder on the GPU (sm_100a).
//
// Replaces pyfunc.py:48-76 create_adjac_voxel (variant 0) and
// tools/tm_mulitmodality_adjacency.py:40-66 (variant 1): voxels of a 3-D mask are labelled in
// C (np.where) order; each voxel lists the labels > 0 found in its clipped 3x3x3 box
// (26-connectivity) or among its 6 face neighbours (only when the whole box is inside the
// volume -- at walls the reference degenerates to the voxel itself, pyfunc.py:65-67).
//   variant 0: self removed, lists sorted, adjacency[0] = [] and label 0 never listed
//   variant 1: self kept when its label > 0 (label 0 keeps its neighbours, nobody lists 0)
// Neighbour offsets are visited in C order, so every list comes out sorted without a sort.
#include "common.cuh"
#include "../../include/tfce_b200.h"

namespace tmb {

static constexpr int kScanThreads = 256;
static constexpr int kScanItems = 16;
static constexpr int kScanTile = kScanThreads * kScanItems;

template <typename In>
__global__ void scan_partial_kernel(const In *__restrict__ in, int64_t n, int64_t *__restrict__ partial) {
    __shared__ int64_t red[kScanThreads / 32];
    const int64_t base = (int64_t)blockIdx.x * kScanTile;
    int64_t s = 0;
    for (int i = threadIdx.x; i < kScanTile; i += kScanThreads) {
        const int64_t idx = base + i;
        if (idx < n) s += (int64_t)in[idx];
    }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        int64_t t = 0;
        for (int w = 0; w < kScanThreads / 32; ++w) t += red[w];
        partial[blockIdx.x] = t;
    }
}

__global__ void scan_spine_kernel(int64_t *partial, int nblocks, int64_t *total) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        int64_t run = 0;
        for (int b = 0; b < nblocks; ++b) { int64_t t = partial[b]; partial[b] = run; run += t; }
        *total = run;
    }
}

// exclusive scan; each thread owns kScanItems consecutive elements of its block's tile
template <typename In>
__global__ void scan_apply_kernel(const In *__restrict__ in, int64_t n, const int64_t *__restrict__ partial,
                                  int64_t *__restrict__ out) {
    __shared__ int64_t warp_tot[kScanThreads / 32];
    const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
    int64_t local[kScanItems];
    int64_t s = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        const int64_t idx = base + i;
        const int64_t v = (idx < n) ? (int64_t)in[idx] : 0;
        local[i] = s;
        s += v;
    }
    // exclusive scan of per-thread totals across the block
    int64_t incl = s;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int o = 1; o < 32; o <<= 1) {
        int64_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) warp_tot[wid] = incl;
    __syncthreads();
    int64_t woff = 0;
    for (int w = 0; w < wid; ++w) woff += warp_tot[w];
    const int64_t thread_off = partial[blockIdx.x] + woff + incl - s;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        const int64_t idx = base + i;
        if (idx < n) out[idx] = thread_off + local[i];
    }
}

template <typename In>
static int exclusive_scan(const In *in, int64_t n, int64_t *out, int64_t *partial, int64_t *total_dev,
                          cudaStream_t st) {
    const int nblocks = (int)((n + kScanTile - 1) / kScanTile);
    scan_partial_kernel<In><<<nblocks, kScanThreads, 0, st>>>(in, n, partial);
    scan_spine_kernel<<<1, 32, 0, st>>>(partial, nblocks, total_dev);
    scan_apply_kernel<In><<<nblocks, kScanThreads, 0, st>>>(in, n, partial, out);
    count_launch(3);
    TMB_CUDA(cudaGetLastError());
    return 0;
}

__device__ __forceinline__ bool face_offset(int dx, int dy, int dz) {
    return (abs(dx) + abs(dy) + abs(dz)) <= 1; // centre and the 6 faces, as the reference's 3x3x3 stencil
}

// pass 0: count -> degree[label];  pass 1: fill indices[indptr[label] ...]
template <int PASS>
__global__ void voxel_neighbours_kernel(const uint8_t *__restrict__ mask, const int64_t *__restrict__ label, int nx,
                                        int ny, int nz, int conn, int variant, int32_t *__restrict__ degree,
                                        const int64_t *__restrict__ indptr, int32_t *__restrict__ indices) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nvol = (int64_t)nx * ny * nz;
    if (idx >= nvol || !mask[idx]) return;
    const int z = (int)(idx % nz);
    const int y = (int)((idx / nz) % ny);
    const int x = (int)(idx / ((int64_t)nz * ny));
    const int64_t me = label[idx];
    int cnt = 0;
    int64_t wr = (PASS == 1) ? indptr[me] : 0;
    const bool interior = x > 0 && y > 0 && z > 0 && x < nx - 1 && y < ny - 1 && z < nz - 1;
    const bool skip_all = (variant == 0 && me == 0); // pyfunc.py:75 adjacency[0] = []
    if (!skip_all) {
        if (conn == 6 && !interior) {
            // walls: the box collapses to the voxel itself (kept only by variant 1 when label > 0)
            if (variant == 1 && me > 0) { if (PASS == 1) indices[wr] = (int32_t)me; ++cnt; }
        } else {
            for (int dx = -1; dx <= 1; ++dx) {
                const int xx = x + dx;
                if (xx < 0 || xx >= nx) continue;
                for (int dy = -1; dy <= 1; ++dy) {
                    const int yy = y + dy;
                    if (yy < 0 || yy >= ny) continue;
                    for (int dz = -1; dz <= 1; ++dz) {
                        const int zz = z + dz;
                        if (zz < 0 || zz >= nz) continue;
                        if (conn == 6 && !face_offset(dx, dy, dz)) continue;
                        const int64_t j = ((int64_t)xx * ny + yy) * nz + zz;
                        if (!mask[j]) continue;
                        const int64_t lj = label[j];
                        if (lj <= 0) continue;                       // label 0 is never listed
                        if (variant == 0 && lj == me) continue;      // self removed
                        if (PASS == 1) indices[wr + cnt] = (int32_t)lj;
                        ++cnt;
                    }
                }
            }
        }
    }
    if (PASS == 0) degree[me] = cnt;
}

} // namespace tmb

using namespace tmb;

extern "C" int tmb_voxel_adjacency(int device, const uint8_t *mask_host, int nx, int ny, int nz, int conn,
                                   int variant, int32_t *num_voxel, int64_t *nnz, int64_t *indptr_host,
                                   int32_t *indices_host) {
    TMB_REQUIRE(mask_host && num_voxel && nnz, "tmb_voxel_adjacency: null pointer");
    TMB_REQUIRE(nx > 0 && ny > 0 && nz > 0, "tmb_voxel_adjacency: bad volume shape");
    TMB_REQUIRE(conn == 26 || conn == 6, "tmb_voxel_adjacency: conn must be 26 or 6");
    TMB_REQUIRE(variant == 0 || variant == 1, "tmb_voxel_adjacency: variant must be 0 (pyfunc) or 1 (tools)");
    TMB_ON_DEVICE(device);
    const int64_t nvol = (int64_t)nx * ny * nz;
    const int nblocks = (int)((nvol + kScanTile - 1) / kScanTile);
    uint8_t *d_mask = nullptr;
    int64_t *d_label = nullptr, *d_partial = nullptr, *d_total = nullptr, *d_indptr = nullptr;
    int32_t *d_degree = nullptr, *d_indices = nullptr;
    int rc = 1;
    int64_t V = 0, total = 0;
    cudaStream_t st = nullptr;
#define ADJ_CUDA(call)                                                                               \
    do {                                                                                             \
        cudaError_t e__ = (call);                                                                    \
        if (e__ != cudaSuccess) {                                                                    \
            set_error("%s failed: %s", #call, cudaGetErrorString(e__));                              \
            goto cleanup;                                                                            \
        }                                                                                            \
    } while (0)
    ADJ_CUDA(cudaMalloc(&d_mask, nvol));
    ADJ_CUDA(cudaMalloc(&d_label, sizeof(int64_t) * nvol));
    ADJ_CUDA(cudaMalloc(&d_partial, sizeof(int64_t) * (nblocks + 1)));
    ADJ_CUDA(cudaMalloc(&d_total, sizeof(int64_t)));
    ADJ_CUDA(cudaMemcpy(d_mask, mask_host, nvol, cudaMemcpyHostToDevice));
    if (exclusive_scan<uint8_t>(d_mask, nvol, d_label, d_partial, d_total, st)) goto cleanup;
    ADJ_CUDA(cudaMemcpy(&V, d_total, sizeof(int64_t), cudaMemcpyDeviceToHost));
    if (V > INT32_MAX - 1) { set_error("tmb_voxel_adjacency: too many voxels"); goto cleanup; }
    *num_voxel = (int32_t)V;
    if (V == 0) { *nnz = 0; if (indptr_host) indptr_host[0] = 0; rc = 0; goto cleanup; }
    ADJ_CUDA(cudaMalloc(&d_degree, sizeof(int32_t) * V));
    ADJ_CUDA(cudaMalloc(&d_indptr, sizeof(int64_t) * (V + 1)));
    {
        const int threads = 256;
        const unsigned grid = (unsigned)((nvol + threads - 1) / threads);
        voxel_neighbours_kernel<0><<<grid, threads, 0, st>>>(d_mask, d_label, nx, ny, nz, conn, variant, d_degree,
                                                             nullptr, nullptr);
        count_launch();
        ADJ_CUDA(cudaGetLastError());
        if (exclusive_scan<int32_t>(d_degree, V, d_indptr, d_partial, d_total, st)) goto cleanup;
        ADJ_CUDA(cudaMemcpy(&total, d_total, sizeof(int64_t), cudaMemcpyDeviceToHost));
        *nnz = total;
        if (indices_host == nullptr || indptr_host == nullptr) { rc = 0; goto cleanup; } // size query
        ADJ_CUDA(cudaMemcpy(d_indptr + V, d_total, sizeof(int64_t), cudaMemcpyDeviceToDevice));
        if (total > 0) {
            ADJ_CUDA(cudaMalloc(&d_indices, sizeof(int32_t) * total));
            voxel_neighbours_kernel<1><<<grid, threads, 0, st>>>(d_mask, d_label, nx, ny, nz, conn, variant, nullptr,
                                                                 d_indptr, d_indices);
            count_launch();
            ADJ_CUDA(cudaGetLastError());
            ADJ_CUDA(cudaMemcpy(indices_host, d_indices, sizeof(int32_t) * total, cudaMemcpyDeviceToHost));
        }
        ADJ_CUDA(cudaMemcpy(indptr_host, d_indptr, sizeof(int64_t) * (V + 1), cudaMemcpyDeviceToHost));
    }
    rc = 0;
cleanup:
    cudaFree(d_mask); cudaFree(d_label); cudaFree(d_partial); cudaFree(d_total);
    cudaFree(d_degree); cudaFree(d_indptr); cudaFree(d_indices);
#undef ADJ_CUDA
    return rc;
}
