// FWER-corrected p lookup (sm_100a).  Replaces the per-vertex Python loop of
// tmanalysis/calculate_fweP_vertex.py:37-42,61-69 and calculate_fweP_voxel.py:23-50:
//   sorted = sort(perm_max);  p_array[j] = j / n;
//   corrp[v] = p_array[max(searchsorted(sorted, tfce[v], side="left") - 1, 0)]
// One thread per statistic value, binary search in the sorted null maxima (float64, as genfromtxt reads them).
#include "common.cuh"
#include "../../include/tfce_b200.h"

namespace tmb {

__global__ void fwe_lookup_kernel(const double *__restrict__ sorted_max, int n, const float *__restrict__ values,
                                  int64_t m, double *__restrict__ corrp) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const double v = (double)values[i];
    int lo = 0, hi = n; // first index with sorted_max[idx] >= v  (searchsorted side="left")
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (sorted_max[mid] < v) lo = mid + 1; else hi = mid;
    }
    const int idx = lo > 0 ? lo - 1 : 0;
    corrp[i] = __ddiv_rn((double)idx, (double)n); // np.true_divide(j, num_perm)
}

} // namespace tmb

using namespace tmb;

extern "C" int tmb_fwe_lookup(const double *sorted_max_dev, int n, const float *values_dev, int64_t m,
                              double *corrp_dev, void *stream) {
    TMB_REQUIRE(sorted_max_dev && values_dev && corrp_dev && n > 0 && m >= 0, "tmb_fwe_lookup: bad arguments");
    if (m == 0) return 0;
    TMB_DEVICE_OF(sorted_max_dev, "tmb_fwe_lookup");
    const int threads = 256;
    fwe_lookup_kernel<<<(unsigned)((m + threads - 1) / threads), threads, 0, (cudaStream_t)stream>>>(
        sorted_max_dev, n, values_dev, m, corrp_dev);
    count_launch();
    TMB_CUDA(cudaGetLastError());
    return 0;
}
