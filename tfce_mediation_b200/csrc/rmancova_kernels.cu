// Repeated-measures ANCOVA statistics (sm_100a) for the permutation loop of tmanalysis/tm_models_randomise.py:522-677,
// i.e. pyfunc.py:1712-2280 reg_rm_ancova_{one,two}_bs_factor.  The reference shuffles the ROWS of the long-format data
// (np.random.shuffle, pyfunc.py:1826 / :2148) and then runs a chain of long-format regressions whose residual sums of
// squares it combines into Type I F statistics.  Here the data stay where they are in HBM and every design is a
// whole-row permutation of a fixed base design, so per shuffle
//   * ONE contraction c = Z' Y (tmb_glm_beta; Z: the centred union of all the designs' columns, rows permuted) gives the
//     cross-products of every design at once; the explained sum of squares of a design with column set S is
//     c_S' inv(G_SS) c_S with G = Z'Z invariant under the permutation (rm_ancova_stats_kernel),
//   * the subject term (residual of the regression on the subject dummies) is the within-group sum of squares of the
//     shuffled rows, and SS_Total is accumulated in the DATA's precision in the shuffled row order, exactly as numpy
//     reduces a [rows, V] array along axis 0 (rm_totals_kernel),
//   * a short host-written program over those terms follows the reference's own sequence of subtractions and
//     divisions (the two functions differ only in their designs and program).
#include "common.cuh"
#include "../../include/tfce_b200.h"
#include <cstdlib>
#include <cstring>

namespace tmb {

// One thread per (shuffle, vertex).  order: shuffled row i holds original row order[i]; grp_rows: the original rows of
// the shuffled data listed subject by subject (grp_size rows each).
template <typename YT>
__global__ void __launch_bounds__(128) rm_totals_kernel(const YT *__restrict__ Y, int N, int64_t V, int64_t ldy,
                                                        const int32_t *__restrict__ order,
                                                        const int32_t *__restrict__ grp_rows,
                                                        const int32_t *__restrict__ grp_size, int ngroups,
                                                        double *__restrict__ sstot, double *__restrict__ ssw, int64_t ldo) {
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int perm = blockIdx.y;
    if (v >= ldo) return;
    if (v >= V) {
        sstot[(size_t)perm * ldo + v] = 0.0;
        if (ssw) ssw[(size_t)perm * ldo + v] = 0.0;
        return;
    }
    const YT *col = Y + v;
    // SS_Total = np.sum((endog - np.mean(endog, 0))**2, 0) (pyfunc.py:1858 / :2171): numpy reduces axis 0 of a C-ordered
    // array row after row with an accumulator of the array's own type -- float32 data give a float32 mean and sum.
    const int32_t *ord = order ? order + (size_t)perm * N : nullptr;   // null: the rows as they are stored
    YT acc = (YT)0;
    for (int i = 0; i < N; ++i) acc = acc + col[(size_t)(ord ? __ldg(ord + i) : i) * ldy];
    const YT mean = acc / (YT)N;
    YT ss = (YT)0;
    for (int i = 0; i < N; ++i) {
        const YT d = col[(size_t)(ord ? __ldg(ord + i) : i) * ldy] - mean;
        ss = ss + d * d;
    }
    sstot[(size_t)perm * ldo + v] = (double)ss;
    if (!ssw) return;
    // residual of the regression on [1, subject dummies]: deviations from the subject means (float64)
    const int32_t *gr = grp_rows + (size_t)perm * N;
    double w = 0.0;
    int at = 0;
    for (int g = 0; g < ngroups; ++g) {
        const int sz = __ldg(grp_size + g);
        double sg = 0.0;
        for (int j = 0; j < sz; ++j) sg += (double)col[(size_t)__ldg(gr + at + j) * ldy];
        const double mg = sg / (double)sz;
        for (int j = 0; j < sz; ++j) {
            const double d = (double)col[(size_t)__ldg(gr + at + j) * ldy] - mg;
            w += d * d;
        }
        at += sz;
    }
    ssw[(size_t)perm * ldo + v] = w;
}

// The same, with the CTA's data columns staged in shared memory once and reused by all P shuffles: every shuffle reads
// every row four times (mean and squares of the total, mean and squares per subject) in its own order, so the global
// kernel above moves 4*P*N*V elements through L2/HBM while this one moves N*V.  TV vertices per CTA (1,024 threads =
// 1024/TV shuffles in flight x TV vertices: one CTA per SM, and the serial, latency-bound walks need the warps); dynamic
// shared memory N*TV elements.
template <typename YT, int TV>
__global__ void __launch_bounds__(1024) rm_totals_tile_kernel(const YT *__restrict__ Y, int N, int64_t V, int64_t ldy,
                                                             const int32_t *__restrict__ order,
                                                             const int32_t *__restrict__ grp_rows,
                                                             const int32_t *__restrict__ grp_size, int ngroups, int P,
                                                             double *__restrict__ sstot, double *__restrict__ ssw,
                                                             int64_t ldo) {
    extern __shared__ __align__(16) unsigned char rm_smem[];
    YT *tile = reinterpret_cast<YT *>(rm_smem);
    constexpr int kLanes = 1024 / TV;                      // shuffles in flight
    const int c = threadIdx.x % TV, sp = threadIdx.x / TV;
    const int64_t v = (int64_t)blockIdx.x * TV + c;
    for (int row = sp; row < N; row += kLanes) tile[(size_t)row * TV + c] = (v < V) ? Y[(size_t)row * ldy + v] : (YT)0;
    __syncthreads();
    if (v >= ldo) return;
    const YT *col = tile + c;
    for (int perm = blockIdx.y * kLanes + sp; perm < P; perm += gridDim.y * kLanes) {
        if (v >= V) {
            sstot[(size_t)perm * ldo + v] = 0.0;
            if (ssw) ssw[(size_t)perm * ldo + v] = 0.0;
            continue;
        }
        const int32_t *ord = order ? order + (size_t)perm * N : nullptr;
        YT acc = (YT)0;
        for (int i = 0; i < N; ++i) acc = acc + col[(size_t)(ord ? __ldg(ord + i) : i) * TV];
        const YT mean = acc / (YT)N;
        YT ss = (YT)0;
        for (int i = 0; i < N; ++i) {
            const YT d = col[(size_t)(ord ? __ldg(ord + i) : i) * TV] - mean;
            ss = ss + d * d;
        }
        sstot[(size_t)perm * ldo + v] = (double)ss;
        if (!ssw) continue;
        const int32_t *gr = grp_rows + (size_t)perm * N;
        double w = 0.0;
        int at = 0;
        for (int g = 0; g < ngroups; ++g) {
            const int sz = __ldg(grp_size + g);
            double sg = 0.0;
            for (int j = 0; j < sz; ++j) sg += (double)col[(size_t)__ldg(gr + at + j) * TV];
            const double mg = sg / (double)sz;
            for (int j = 0; j < sz; ++j) {
                const double d = (double)col[(size_t)__ldg(gr + at + j) * TV] - mg;
                w += d * d;
            }
            at += sz;
        }
        ssw[(size_t)perm * ldo + v] = w;
    }
}

template <typename YT, int TV>
static int launch_rm_totals_tile(const YT *Y, int N, int64_t V, int64_t ldy, const int32_t *order, const int32_t *grp_rows,
                                 const int32_t *grp_size, int ngroups, int P, double *sstot, double *ssw, int64_t ldo,
                                 cudaStream_t stream) {
    const size_t smem = (size_t)N * TV * sizeof(YT);
    TMB_CUDA(cudaFuncSetAttribute(rm_totals_tile_kernel<YT, TV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t xtiles = (ldo + TV - 1) / TV;
    // enough CTAs for a few waves: split the shuffles when there are few column tiles
    int ysplit = 1;
    constexpr int kLanes = 1024 / TV;
    while (xtiles * ysplit < 4 * 148 && ysplit * kLanes * 2 <= P) ysplit *= 2;
    rm_totals_tile_kernel<YT, TV><<<dim3((unsigned)xtiles, (unsigned)ysplit), 1024, smem, stream>>>(
        Y, N, V, ldy, order, grp_rows, grp_size, ngroups, P, sstot, ssw, ldo);
    return 0;
}

template <typename YT>
static int launch_rm_totals(const YT *Y, int N, int64_t V, int64_t ldy, const int32_t *order, const int32_t *grp_rows,
                            const int32_t *grp_size, int ngroups, int P, double *sstot, double *ssw, int64_t ldo,
                            cudaStream_t stream) {
    const size_t budget = 200 * 1024, row = (size_t)N * sizeof(YT);
    const char *force = getenv("TMB_RM_TOTALS");            // "global": the unstaged kernel (A/B measurements, tests)
    const bool global_only = force && strcmp(force, "global") == 0;
    if (!global_only && P >= 4) {
        if (row * 128 <= budget) return launch_rm_totals_tile<YT, 128>(Y, N, V, ldy, order, grp_rows, grp_size, ngroups, P, sstot, ssw, ldo, stream);
        if (row * 64 <= budget) return launch_rm_totals_tile<YT, 64>(Y, N, V, ldy, order, grp_rows, grp_size, ngroups, P, sstot, ssw, ldo, stream);
        if (row * 32 <= budget) return launch_rm_totals_tile<YT, 32>(Y, N, V, ldy, order, grp_rows, grp_size, ngroups, P, sstot, ssw, ldo, stream);
    }
    const dim3 grid((unsigned)((ldo + 127) / 128), (unsigned)P);
    rm_totals_kernel<YT><<<grid, 128, 0, stream>>>(Y, N, V, ldy, order, grp_rows, grp_size, ngroups, sstot, ssw, ldo);
    return 0;
}

static constexpr int kRmMaxCols = 64;      // columns of the union design
static constexpr int kRmMaxRegs = 96;      // registers of the program (terms + temporaries)
static constexpr int kRmDesignStride = 2 + kRmMaxCols;

// meta (int32): [0] designs D, [1] program length, [2] output rows, [3] columns rU of the union design;
//   [8 + d*66 ...]: k_d, offset of inv(G_SS) in `mats`, the k_d column indices;
//   then 4 ints per operation (op, dst, a, b), then the output registers.
// Registers: 0 SS_Total, 1 within-subject residual, 2 + d the residual sum of squares of design d.
// Operations: 0 dst = a - b; 1 dst = a + b; 2 dst = a / consts[b]; 3 dst = a / b; 4 dst = 0.
__global__ void __launch_bounds__(128) rm_ancova_stats_kernel(const double *__restrict__ cross, int64_t ldb, int64_t V,
                                                              const int32_t *__restrict__ meta,
                                                              const double *__restrict__ mats,
                                                              const double *__restrict__ consts,
                                                              const double *__restrict__ yy,
                                                              const double *__restrict__ sstot,
                                                              const double *__restrict__ ssw, int64_t ldo,
                                                              float *__restrict__ out32, double *__restrict__ out64,
                                                              int64_t ldt, int nan_to_zero) {
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int perm = blockIdx.y;
    if (v >= ldt) return;
    const bool inside = v < V;
    const int D = __ldg(meta), nops = __ldg(meta + 1), nout = __ldg(meta + 2), rU = __ldg(meta + 3);
    double c[kRmMaxCols];
    for (int j = 0; j < rU; ++j) c[j] = inside ? cross[((size_t)perm * rU + j) * ldb + v] : 0.0;
    double reg[kRmMaxRegs];
    const double yyv = inside ? yy[v] : 0.0;
    reg[0] = inside ? sstot[(size_t)perm * ldo + v] : 0.0;
    reg[1] = inside ? ssw[(size_t)perm * ldo + v] : 0.0;
    for (int d = 0; d < D; ++d) {
        const int32_t *des = meta + 8 + d * kRmDesignStride;
        const int k = __ldg(des);
        const double *M = mats + __ldg(des + 1);
        double q = 0.0;
        for (int a = 0; a < k; ++a) {
            double inner = 0.0;
            for (int b = 0; b < k; ++b) inner = __fma_rn(__ldg(M + a * k + b), c[__ldg(des + 2 + b)], inner);
            q = __fma_rn(c[__ldg(des + 2 + a)], inner, q);
        }
        reg[2 + d] = yyv - q;
    }
    const int32_t *prog = meta + 8 + D * kRmDesignStride;
    for (int o = 0; o < nops; ++o) {
        const int op = __ldg(prog + 4 * o), dst = __ldg(prog + 4 * o + 1), a = __ldg(prog + 4 * o + 2),
                  b = __ldg(prog + 4 * o + 3);
        double x;
        switch (op) {
        case 0: x = __dsub_rn(reg[a], reg[b]); break;
        case 1: x = __dadd_rn(reg[a], reg[b]); break;
        case 2: x = __ddiv_rn(reg[a], __ldg(consts + b)); break;
        case 3: x = __ddiv_rn(reg[a], reg[b]); break;
        default: x = 0.0; break;
        }
        reg[dst] = x;
    }
    const int32_t *outs = prog + 4 * nops;
    for (int r = 0; r < nout; ++r) {
        double x = reg[__ldg(outs + r)];
        if (nan_to_zero && x != x) x = 0.0;
        if (!inside) x = 0.0;
        const size_t off = ((size_t)perm * nout + r) * ldt + v;
        if (out32) out32[off] = __double2float_rn(x);
        if (out64) out64[off] = x;
    }
}

} // namespace tmb

using namespace tmb;

extern "C" int tmb_rm_totals(const void *Y_dev, int ydtype, int N, int64_t V, int64_t ldy, const int32_t *order_dev,
                             const int32_t *grp_rows_dev, const int32_t *grp_size_dev, int ngroups, int P,
                             double *sstotal_dev, double *sswithin_dev, int64_t ldo, void *stream) {
    TMB_REQUIRE(Y_dev && sstotal_dev, "tmb_rm_totals: null pointer");
    TMB_REQUIRE((grp_rows_dev && grp_size_dev && sswithin_dev && ngroups > 0) || (!grp_rows_dev && !sswithin_dev),
                "tmb_rm_totals: the subject term needs grp_rows_dev, grp_size_dev and sswithin_dev together");
    if (!sswithin_dev) ngroups = 0;
    TMB_REQUIRE(ydtype == TMB_F32 || ydtype == TMB_F64, "ydtype must be TMB_F32 (0) or TMB_F64 (1)");
    TMB_REQUIRE(N > 0 && V > 0 && ldy >= V && ldo >= V && P >= 1 && P <= 65535,
                "tmb_rm_totals: bad shape (N=%d V=%lld groups=%d P=%d)", N, (long long)V, ngroups, P);
    TMB_DEVICE_OF(Y_dev, "tmb_rm_totals");
    const int rc = ydtype == TMB_F64
                       ? launch_rm_totals<double>((const double *)Y_dev, N, V, ldy, order_dev, grp_rows_dev, grp_size_dev, ngroups, P,
                                                  sstotal_dev, sswithin_dev, ldo, (cudaStream_t)stream)
                       : launch_rm_totals<float>((const float *)Y_dev, N, V, ldy, order_dev, grp_rows_dev, grp_size_dev, ngroups, P,
                                                 sstotal_dev, sswithin_dev, ldo, (cudaStream_t)stream);
    if (rc) return rc;
    count_launch();
    TMB_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int tmb_rm_ancova_stats(const double *cross_dev, int64_t ldb, int64_t V, int P, const int32_t *meta_dev,
                                   const int32_t *meta_host, const double *mats_dev, const double *consts_dev,
                                   const double *yy_dev, const double *sstotal_dev, const double *sswithin_dev,
                                   int64_t ldo, float *f32_dev, double *f64_dev, int64_t ldt, int nan_to_zero,
                                   void *stream) {
    TMB_REQUIRE(cross_dev && meta_dev && meta_host && mats_dev && consts_dev && yy_dev && sstotal_dev && sswithin_dev &&
                    (f32_dev || f64_dev), "tmb_rm_ancova_stats: null pointer");
    const int D = meta_host[0], nops = meta_host[1], nout = meta_host[2], rU = meta_host[3];
    TMB_REQUIRE(V > 0 && ldb >= V && ldo >= V && ldt >= V && P >= 1 && P <= 65535, "tmb_rm_ancova_stats: bad shape");
    TMB_REQUIRE(D >= 1 && 2 + D <= kRmMaxRegs && nops >= 1 && nout >= 1 && rU >= 1 && rU <= kRmMaxCols,
                "tmb_rm_ancova_stats: %d designs, %d operations, %d outputs, %d columns (at most %d columns)", D, nops,
                nout, rU, kRmMaxCols);
    for (int d = 0; d < D; ++d) {
        const int32_t *des = meta_host + 8 + d * kRmDesignStride;
        TMB_REQUIRE(des[0] >= 0 && des[0] <= rU && des[1] >= 0, "tmb_rm_ancova_stats: design %d has %d columns", d, des[0]);
        for (int j = 0; j < des[0]; ++j)
            TMB_REQUIRE(des[2 + j] >= 0 && des[2 + j] < rU, "tmb_rm_ancova_stats: design %d, column %d out of range", d, des[2 + j]);
    }
    const int32_t *prog = meta_host + 8 + D * kRmDesignStride;
    for (int o = 0; o < nops; ++o) {
        const int op = prog[4 * o], dst = prog[4 * o + 1], a = prog[4 * o + 2], b = prog[4 * o + 3];
        TMB_REQUIRE(op >= 0 && op <= 4 && dst >= 2 + D && dst < kRmMaxRegs && a >= 0 && a < kRmMaxRegs && b >= 0 &&
                        (op == 2 || b < kRmMaxRegs), "tmb_rm_ancova_stats: operation %d is malformed", o);
    }
    for (int r = 0; r < nout; ++r)
        TMB_REQUIRE(prog[4 * nops + r] >= 0 && prog[4 * nops + r] < kRmMaxRegs, "tmb_rm_ancova_stats: output %d out of range", r);
    TMB_DEVICE_OF(cross_dev, "tmb_rm_ancova_stats");
    const dim3 grid((unsigned)((ldt + 127) / 128), (unsigned)P);
    rm_ancova_stats_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(cross_dev, ldb, V, meta_dev, mats_dev, consts_dev, yy_dev,
                                                                   sstotal_dev, sswithin_dev, ldo, f32_dev, f64_dev, ldt,
                                                                   nan_to_zero);
    count_launch();
    TMB_CUDA(cudaGetLastError());
    return 0;
}
