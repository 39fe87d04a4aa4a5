/*
 * tfce_b200.h -- C ABI of libtfce_b200.so: the B200 (sm_100a) implementation of the
 * TFCE_mediation permutation hot path (permuted OLS fit + t  ->  TFCE  ->  scaled max).
 *
 * The reference has no C/FFI plugin interface: its boundary is the Python API of two compiled
 * extension modules.  Each entry point below names the reference
 * interface it replaces (paths relative to /root/reference/tfce_mediation/); the ctypes
 * binding a maintainer adds on the reference side is in INTEGRATION.md and shipped as
 * tfce_mediation_b200/{tfce,cynumstats,pyfunc,tm_func}.py.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; tmb_last_error() then returns
 *     a thread-local human-readable message.  No exceptions cross the ABI.
 *   - *_host pointers are host memory owned by the caller; *_dev pointers are device memory on
 *     the handle's device owned by the caller (the Python host layer allocates them with
 *     PyTorch and passes tensor.data_ptr()).  `stream` is a cudaStream_t passed as void*
 *     (NULL = the legacy default stream).  Calls taking a stream are asynchronous on it.
 *   - handles are opaque, own their device state, and are not re-entrant.
 *   - there is NO CPU fallback: without a CUDA device every compute call fails.
 */
#ifndef TFCE_B200_H
#define TFCE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TMB_ABI_VERSION 3
#define TMB_F32 0
#define TMB_F64 1

typedef struct tmb_graph tmb_graph; /* one adjacency graph + (H, E): == a CreateAdjSet object  */
typedef struct tmb_plan tmb_plan;   /* a set of graphs laid out along one statistic row          */

/* status bits reported per TFCE map (tmb_plan_run `status_dev`, tmb_tfce_run return detail) */
#define TMB_MAP_OK 0
#define TMB_MAP_MAX_IS_ZERO 1   /* max == 0: the reference never returns (fast_tfce.hpp:34-39); we return zeros */
#define TMB_MAP_STEP_OVERFLOW 2 /* threshold sequence longer than 127 steps (cannot happen for finite maxima)   */

const char *tmb_last_error(void);
int tmb_abi_version(void);
/* number of visible CUDA devices (0 when there is none; never fails) */
int tmb_device_count(void);
/* number of kernel launches issued by this library since load (bench.py's gpu_launches) */
int64_t tmb_launch_count(void);

/* ---------------------------------------------------------------------------------------------
 * Graph  ==  tfce.pyx:24-42  `CreateAdjSet.__init__(H, E, pyAdjacency)`
 * The Python layer flattens pyAdjacency (lists / sets / arrays, item order kept) into CSR.
 * H and E are stored as C float exactly like tfce.pyx:27-33.  pow(n, E) for n = 0..V is
 * tabulated on the host with the C library's double pow so that the increment
 * (float)(pow((double)n,(double)E) * (double)powf(T,H)) of fast_tfce.hpp:70-77 is bit-identical.
 * Out-of-range neighbour indices are rejected.  Directed entries are honoured with the
 * reference's rule (entry a of adjacency[u] joins only if a activated before u).
 * ------------------------------------------------------------------------------------------- */
int tmb_graph_create(int device, int32_t V, const int64_t *indptr_host, const int32_t *indices_host,
                     float H, float E, tmb_graph **out);
int tmb_graph_destroy(tmb_graph *g);
int tmb_graph_num_vertices(const tmb_graph *g, int32_t *V, int64_t *nnz);
/* The library relabels the vertices of symmetric graphs for memory locality (reverse Cuthill-McKee +
 * 64-vertex patches); callers never see it unless they opt in: vmap_host[i] (int32 [V]) is the caller's
 * index of internal vertex i (identity when the graph was not relabelled). */
int tmb_graph_vmap(const tmb_graph *g, int32_t *vmap_host);

/* == tfce.pyx:44-45  `CreateAdjSet.run(image, enhn)`:  enhn[v] += TFCE(image)[v], fp32, host buffers.
 * map_status (may be NULL) receives TMB_MAP_* bits. */
int tmb_tfce_run(tmb_graph *g, const float *image_host, float *enhn_host, int *map_status);

/* Test/inspection entry (parity of labels and extents): connected components
 * of {v : image[v] > T_level} with T from the reference threshold sequence.  labels[v] = smallest
 * vertex index in v's component or -1; extents[v] = component size or 0.  Host buffers. */
int tmb_tfce_components(tmb_graph *g, const float *image_host, int level, int32_t *labels_host,
                        int32_t *extents_host, float *threshold_out);

/* ---------------------------------------------------------------------------------------------
 * Plan: S graphs ("surfaces": lh/rh hemispheres, a voxel skeleton, mmr surfaces) laid along one
 * statistic row; surface s covers columns [col_offset[s], col_offset[s] + V_s).
 * weight_host[s] is NULL (weight 1) or V_s floats (vertex-density correction, vdensity_?h of
 * STEP_1_vertex_tfce_multiple_regression.py:161-173; float32 there and in mmr-lr, tm_mmr_rand_low_ram.py:165).
 * weight64_host (may be NULL) / weight64_host[s]: V_s DOUBLES instead -- the non-low-RAM mmr path builds its density
 * weights in float64 (tm_multimodality_multisurface_regression.py:451-459) and the reference then multiplies in
 * double: fl32(double(fl32(tfce * scale)) * w), tm_func.py:83-91.  max_slots bounds how many maps are in
 * flight at once (workspace is max_slots * O(V_max)); 0 = library default.
 *
 * tmb_plan_run == the body of pyfunc.py:107-126 write_perm_maxTFCE_vertex / _voxel and of
 * tm_func.py:160-182 for B statistic rows at once:
 *   for each row b and surface s:  TFCE of +stat (and of -stat when two_sided), then
 *   max_dev[(b*S + s)*2 + sign] = max_v fl32( fl32(tfce[v] * fl32(max(stat)/100)) * weight[v] )
 * (0 when nothing is positive).  tfce_pos_dev / tfce_neg_dev (may be NULL) receive the unscaled
 * TFCE maps, rows of leading dimension ld like stat_dev.  status_dev (may be NULL): int32 [B*S*2].
 * ------------------------------------------------------------------------------------------- */
int tmb_plan_create(int device, int S, tmb_graph *const *graphs, const int64_t *col_offset,
                    const float *const *weight_host, const double *const *weight64_host, int max_slots,
                    tmb_plan **out);
int tmb_plan_destroy(tmb_plan *p);
/* Opt-in fast path of the batched engine: declare that statistic rows (and requested TFCE maps) use the
 * graphs' INTERNAL vertex order (tmb_graph_vmap), e.g. because the data columns were permuted once at
 * upload; the per-map gather through vmap is then skipped. */
int tmb_plan_set_internal_order(tmb_plan *p, int on);
int tmb_plan_run(tmb_plan *p, const float *stat_dev, int64_t ld, int B, int two_sided, float *max_dev,
                 float *tfce_pos_dev, float *tfce_neg_dev, int32_t *status_dev, void *stream);

/* Exact-libm mode.  The reference evaluates the height term with std::pow(float,float) (fast_tfce.hpp:70),
 * i.e. the host C library's powf, which is NOT correctly rounded: about 6 in 10^4 thresholds differ from
 * the exact square by one ulp, so bit-identical TFCE values require that very function.
 *   tmb_plan_maxima:       max_dev[(b*S+s)*2 + sign] = max of +stat / -stat over surface s of row b (NaN ignored)
 *   tmb_threshold_tables:  HOST function; for each maximum builds the reference's threshold sequence
 *                          (T_0 = max, T_{i+1} = T_i - max/100 while T_i >= 0, fast_tfce.hpp:32-39) and
 *                          HH_i = powf(T_i, H) into rows of 128 floats, plus step count, delta and TMB_MAP_* status
 *   tmb_plan_run_tables:   tmb_plan_run consuming those tables (device copies) instead of computing
 *                          correctly rounded ones on the device.  scale_dev (may be NULL): float32 [B*S*2], the factor of
 *                          the scaled maximum per entry when it is not the threshold step delta -- the non-low-RAM mmr
 *                          path thresholds ONE merged graph, i.e. delta = (maximum over all surfaces) / 100
 *                          (fast_tfce.hpp:32-36 on tm_func.py:77-78's merged image), but rescales every surface with
 *                          its own max/100 (tm_func.py:83-91): tables from the group maximum, scale from the surface's.
 * tmb_tfce_run and tmb_tfce_components always use host tables (they have the host image anyway). */
int tmb_plan_maxima(tmb_plan *p, const float *stat_dev, int64_t ld, int B, float *max_dev, void *stream);
int tmb_threshold_tables(const float *maxima_host, const float *H_host, int count, int32_t *ns_host,
                         float *delta_host, float *T_host, float *HH_host, int32_t *status_host);
int tmb_plan_run_tables(tmb_plan *p, const float *stat_dev, int64_t ld, int B, int two_sided, const int32_t *ns_dev,
                        const float *delta_dev, const float *T_dev, const float *HH_dev, const int32_t *tstatus_dev,
                        const float *scale_dev, float *max_dev, float *tfce_pos_dev, float *tfce_neg_dev,
                        int32_t *status_dev, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Permuted-design fit + t.  Replaces cynumstats.pyx:28-29 cy_lin_lstsqr_mat, :47-52 se_of_slope,
 * :59-64 tval_int (and :66-74 calc_beta_se) for P designs at once.
 *
 * Y_dev: float32 (ydtype TMB_F32) or float64 (TMB_F64) [n, ldy] row-major subject-by-vertex data
 *        (== merge_y; the reference feeds float64 in step 1 and float32 in the randomise step); ldy is a multiple of 128
 *        covering V rounded up to 128 (pad columns zero); 16-byte aligned.
 * At_dev: float64 [n, ldA]: row i of the j-th design's pseudo-inverse (X_j'X_j)^-1 X_j' as one column
 *         (k-major so one subject's coefficients are contiguous); rp is r padded to 1, 2, 4 or 8 with
 *         zero columns.  Column order `layout` (ask tmb_glm_layout(ydtype, rp), which knows the kernel that will run):
 *           0: column j*rp + i                               (fp64 vector kernel), ldA >= P*rp
 *           1: "tile8", column (j/8)*8*rp + i*8 + j%8        (fp64 tensor-core kernels: one 8-row DMMA tile
 *              holds one regressor of 8 designs),            ldA >= ceil(P/8)*8*rp
 *         ldA is a multiple of 128 covering those columns; columns beyond them are zero.
 * G_dev:  float64 [P, r, r] = X_j'X_j ;  d_dev: float64 [P, r] = diag((X_j'X_j)^-1).
 * yy_dev: float64 [V] = sum_i Y[i,v]^2 (tmb_glm_sumsq; colsum_dev optionally receives sum_i Y[i,v]).
 * SSE = yy - b'Gb ; sigma2 = SSE/dof ; se = (float)sqrt(sigma2 * d)  (the fp32 rounding of
 * cynumstats.pyx:49-51 is kept) ; t = b / (double)se.
 * Rows [row0, row0 + nrows) of every design are written:
 *   t32_dev  float32 [P, nrows, ldt]  and/or  t64_dev float64 [P, nrows, ldt]  (either may be NULL)
 * nan_to_zero != 0 applies voxel_tfce_multiple_regression_randomise.py:109 (t[isnan] = 0).
 * When the design has an intercept the host passes mean-centred designs (Frisch-Waugh), see
 * DESIGN.md; yy is then the centred sum of squares (center = 1 in tmb_glm_sumsq).
 * ------------------------------------------------------------------------------------------- */
int tmb_glm_sumsq(const void *Y_dev, int ydtype, int n, int64_t V, int64_t ldy, int center, double *yy_dev,
                  double *colsum_dev, void *stream);
int tmb_glm_tstat(const void *Y_dev, int ydtype, int n, int64_t V, int64_t ldy, const double *At_dev, int64_t ldA,
                  const double *G_dev, const double *d_dev, int P, int r, int rp, int row0, int nrows,
                  double dof, const double *yy_dev, float *t32_dev, double *t64_dev, int64_t ldt,
                  int nan_to_zero, int layout, void *stream);
/* What the fit kernels expect for this data type: the column order of At_dev (see above), the padded regressor count rp
 * for r regressors (0: r is outside 1..8 -- use the *_beta entry points below), and the columns of At_dev they read
 * for P designs (a multiple of 128; allocate At_dev with ldA >= this, zero-filled beyond the packed columns). */
int tmb_glm_layout(int ydtype, int rp);
int tmb_glm_rp(int ydtype, int r);
int64_t tmb_glm_packed_columns(int ydtype, int P, int rp);
/* F statistics of pyfunc.py:2282-2401 glm_typeI (the tm-models GLM branch, tmanalysis/tm_models_randomise.py:197-272)
 * for P designs at once, from ONE fit per design: operands as in tmb_glm_tstat (centred designs, r = k-1 regressors).
 * Tested variable i covers regressor rows [var_lo[i], var_lo[i] + var_k[i]) (host arrays, nvar <= 8); M_dev float64
 * [P, sum_i var_k[i]^2] holds, per design, the matrices inv(C[S_i, S_i]) one after the other, C = (X'X)^-1, because
 * RSS_without_i - RSS = b_S' inv(C_SS) b_S.  Output rows per design (ldt-strided, float32 and/or float64):
 *   [model F = ((TSS-RSS)/r) / (RSS/dof)  -- only when want_model != 0],  then per variable
 *   F_i = (RSS_without_i - RSS) / ((RSS/dof) * var_k[i])                  (pyfunc.py:2336-2354).
 * sstotal_dev float64 [V] or NULL: the TSS of the model F as the reference accumulates it (pyfunc.py:2331: in float32
 * for float32 data; tmb_rm_totals with order_dev = NULL); NULL uses the float64 explained sum of squares. */
int tmb_glm_fstat(const void *Y_dev, int ydtype, int n, int64_t V, int64_t ldy, const double *At_dev, int64_t ldA,
                  const double *G_dev, const double *M_dev, int P, int r, int rp, int nvar, const int32_t *var_lo,
                  const int32_t *var_k, int want_model, double dof, const double *yy_dev, const double *sstotal_dev,
                  float *f32_dev, double *f64_dev, int64_t ldt, int nan_to_zero, int layout, void *stream);

/* Stacked pseudo-inverses of P row-permuted copies of ONE design (the permutation loop of
 * vertex_tfce_multiple_regression_randomise.py:104-106, `nx = X[np.random.permutation(...)]`): permuting whole rows
 * permutes the columns of pinv(X), so At[k, col(p, i)] = pinv[i, perm_idx[p, k]] is a gather done on the device
 * (col as in tmb_glm_tstat's `layout`).
 * pinv_dev float64 [r, n] (centred regressors' pseudo-inverse), perm_idx_dev int32 [P, n], At_dev float64 [n, ldA]
 * (unused columns and rows i >= r are zero-filled). */
int tmb_glm_pack_rowperm(const double *pinv_dev, int r, int n, const int32_t *perm_idx_dev, int P, int rp,
                         double *At_dev, int64_t ldA, int layout, void *stream);

/* Sobel z for medtype 'M' / 'I' (pyfunc.py:130-162 under the drivers' rule that only pred_x is permuted,
 * vertex_tfce_mediation_randomise.py:82-90) from ONE contraction row per shuffle instead of three.  At_dev float64
 * [n, ldA]: column p = the centred, permuted predictor of shuffle p (tmb_glm_pack_rowperm with r = 1 on the centred
 * predictor; ldA = tmb_glm_packed_columns(TMB_F32, P, 1)).  cd_dev float64 [V]: the centred cross-product dep'y of the
 * un-permuted second variable (one tmb_glm_beta row, once).  xx = x'x (centred).  CB_dev float64 [P, 8]: per shuffle the
 * inverse C of the centred Gram matrix of path B's two regressors in design order (C00, C01, C10, C11), then
 * C[rowB][rowB] / dofB, then padding; xpos = position of the predictor among the two (1 for 'M': [dep, x]; 0 for 'I':
 * [x, dep]); rowB = the tested one (0 for both).  float32 data only. */
int tmb_sobelz_cross(const void *Y_dev, int ydtype, int n, int64_t V, int64_t ldy, const double *At_dev, int64_t ldA,
                     const double *cd_dev, double xx, double dofA, const double *CB_dev, int xpos, int rowB, double dofB,
                     const double *yy_dev, int P, int alg, float *z32_dev, double *z64_dev, int64_t ldt, void *stream);

/* Sobel z for two paths whose designs share a set of PERMUTED columns and a set of FIXED columns (tm-models mediation,
 * tm_models_randomise.py:430-520: left variable permuted; right variable and covariates fixed -- or permuted too for
 * medtype 'Y'), from centred cross-products c = Z'y: cperm_dev float64 [P*m, ldb] = the m permuted columns' rows of every
 * shuffle (tmb_glm_beta with the centred, permuted columns as At), cfix_dev float64 [f, ldf] = the fixed columns' rows
 * (once).  CA_dev [P, rA, rA] / CB_dev [P, rB, rB]: INVERSE centred Gram matrices of the shuffle's designs (rA, rB <= 16);
 * colmap (the same int32 array on the device and on the host): for path A's rA, then path B's rB regressors in design
 * order the source row (< m: permuted row, else m + fixed row).  ta_scalar_dev replaces path A as in tmb_sobelz. */
int tmb_sobelz_cross_rows(const double *cperm_dev, int64_t ldb, int m, const double *cfix_dev, int64_t ldf, int f, int64_t V,
                          const double *CA_dev, int rA, int rowA, double dofA, const double *CB_dev, int rB, int rowB,
                          double dofB, const int32_t *colmap_dev, const int32_t *colmap_host, const double *yy_dev,
                          const double *ta_scalar_dev, int P, int alg, float *z32_dev, double *z64_dev, int64_t ldt,
                          void *stream);

/* t statistics of designs of which only SOME columns change between shuffles (the randomise drivers' `-v first last`:
 * vertex_tfce_multiple_regression_randomise.py:84-97 permutes the chosen regressors and leaves the covariates alone), from
 * centred cross-products: cperm_dev float64 [P*m, ldb] = the rows of the m changing columns of every shuffle (tmb_glm_beta),
 * cfix_dev [f, ldf] = the rows of the fixed columns (once), C_dev [P, r, r] = the shuffle's INVERSE centred Gram matrix
 * (r <= 16), colmap: source row of each regressor in design order (< m: changing, else m + fixed row).  Output rows
 * row0 .. row0+nrows-1 per shuffle, as tmb_glm_tstat. */
int tmb_glm_tstat_cross_rows(const double *cperm_dev, int64_t ldb, int m, const double *cfix_dev, int64_t ldf, int f, int64_t V,
                             const double *C_dev, int r, const int32_t *colmap_dev, const int32_t *colmap_host, int row0,
                             int nrows, double dof, const double *yy_dev, int P, float *t32_dev, double *t64_dev, int64_t ldt,
                             int nan_to_zero, void *stream);

/* Designs with MORE than 8 non-intercept regressors (the reference accepts any k: cynumstats.pyx:28-29,59-64 -- e.g.
 * dummy-coded sites plus covariates): the betas are formed first by tmb_glm_beta (every pseudo-inverse row of every
 * design is one column of At_dev and one output row, design-major: row p*r + i), then these evaluate the same statistics
 * as tmb_glm_tstat / tmb_glm_fstat / tmb_sobelz per (design, vertex) from beta_dev float64 [P*r, ldb].  r <= 64. */
int tmb_glm_tstat_beta(const double *beta_dev, int64_t ldb, int64_t V, const double *G_dev, const double *d_dev, int P,
                       int r, int row0, int nrows, double dof, const double *yy_dev, float *t32_dev, double *t64_dev,
                       int64_t ldt, int nan_to_zero, void *stream);
int tmb_glm_fstat_beta(const double *beta_dev, int64_t ldb, int64_t V, const double *G_dev, const double *M_dev, int P,
                       int r, int nvar, const int32_t *var_lo, const int32_t *var_k, int want_model, double dof,
                       const double *yy_dev, const double *sstotal_dev, float *f32_dev, double *f64_dev, int64_t ldt,
                       int nan_to_zero, void *stream);
int tmb_sobelz_beta(const double *beta_dev, int64_t ldb, int64_t V, const double *GA_dev, const double *dA_dev, int rA,
                    int rowA, double dofA, const double *GB_dev, const double *dB_dev, int rB, int rowB, double dofB,
                    const double *yy_dev, const double *ta_scalar_dev, int P, int alg, float *z32_dev, double *z64_dev,
                    int64_t ldt, void *stream);

/* Cosinor statistics (pyfunc.py:2406-2563 glm_cosinor, the permutation branch used by tm_models_randomise.py:274-381)
 * from stored betas of the centred design [cos_0, sin_0, ..., cos_{nper-1}, sin_{nper-1}, nexog tested columns,
 * covariates] (r regressors, r <= 64).  G_dev, C_dev float64 [P, r, r]: the centred Gram matrix and its inverse.
 * sstotal_dev float64 [V] or NULL: SS_Total as the reference accumulates it (tmb_rm_totals; float32 arithmetic for
 * float32 data) for the model F's numerator SS_Total - SS_Residuals; NULL uses the float64 explained sum of squares.
 * mediation == 0: rows per design [model F, (|t amplitude_i|, |t acrophase_i|) per period, t of each tested column]
 * (1 + 2*nper + nexog rows).  mediation != 0 (tm_models_randomise.py:383-412): the single row
 * calc_indirect(ta, t of tested column 0) with `ta` the un-permuted path-A amplitude t (alg as in tmb_sobelz). */
int tmb_glm_cosinor_beta(const double *beta_dev, int64_t ldb, int64_t V, const double *G_dev, const double *C_dev, int P,
                         int r, int nper, int nexog, double dof, const double *yy_dev, const double *sstotal_dev,
                         int mediation, double ta, int alg, float *s32_dev, double *s64_dev, int64_t ldt, int nan_to_zero,
                         void *stream);

/* Repeated-measures ANCOVA (pyfunc.py:1712-2280 reg_rm_ancova_{one,two}_bs_factor in the permutation loop of
 * tm_models_randomise.py:522-677).  The reference shuffles the rows of the long-format data [N = intervals*subjects, V]
 * and regresses them on a chain of designs; here the data stay in place and every design is a whole-row permutation
 * of a fixed base design.  Per shuffle: (1) tmb_glm_beta with the centred, row-permuted union of all design columns
 * gives the cross-products c = Z'Y [rU, V]; (2) tmb_rm_totals gives SS_Total accumulated in the data's own precision in
 * the shuffled row order (as numpy reduces axis 0; order_dev int32 [P, N]: shuffled row i = original row order[i]) and
 * the residual of the subject-dummy regression (grp_rows_dev int32 [P, N]: the original rows subject by subject,
 * grp_size_dev int32 [ngroups] rows each; order_dev NULL = the rows as stored, grp_rows_dev = sswithin_dev = NULL = SS_Total
 * only -- that form also gives glm_cosinor's SS_Total, pyfunc.py:2492); (3) tmb_rm_ancova_stats forms the residual SS of every design,
 * yy - c_S' inv(G_SS) c_S, and runs the host-written program that combines them as the reference does.
 * meta (int32, the same array on the host and on the device): [0] designs D, [1] operations, [2] output rows, [3] rU,
 * [8 + 66*d ..] = k_d, offset of inv(G_SS) (k_d x k_d, row-major) in mats_dev, k_d column indices; then 4 ints per
 * operation (op, dst, a, b) and the output registers.  Registers: 0 SS_Total, 1 subject residual, 2+d residual SS of
 * design d, the rest temporaries (< 96).  op 0: dst = a - b; 1: a + b; 2: a / consts_dev[b]; 3: a / b; 4: 0. */
int tmb_rm_totals(const void *Y_dev, int ydtype, int N, int64_t V, int64_t ldy, const int32_t *order_dev,
                  const int32_t *grp_rows_dev, const int32_t *grp_size_dev, int ngroups, int P, double *sstotal_dev,
                  double *sswithin_dev, int64_t ldo, void *stream);
int tmb_rm_ancova_stats(const double *cross_dev, int64_t ldb, int64_t V, int P, const int32_t *meta_dev,
                        const int32_t *meta_host, const double *mats_dev, const double *consts_dev, const double *yy_dev,
                        const double *sstotal_dev, const double *sswithin_dev, int64_t ldo, float *f32_dev,
                        double *f64_dev, int64_t ldt, int nan_to_zero, void *stream);

/* betas only == cynumstats.pyx:28-29 cy_lin_lstsqr_mat: beta64_dev float64 [nrows, ldt] for the nrows
 * pseudo-inverse rows stored as the first nrows columns of At_dev (ldA a multiple of 128). */
int tmb_glm_beta(const void *Y_dev, int ydtype, int n, int64_t V, int64_t ldy, const double *At_dev, int64_t ldA,
                 int nrows, double *beta64_dev, int64_t ldt, void *stream);

/* Direct single-design fit with the reference's EXPLICIT residual pass (cynumstats.pyx:61), one
 * vertex per thread; serves the API-parity entry points on one design:
 *   cynumstats.pyx:59-64 tval_int, :66-74 calc_beta_se, :54-57 resid_covars, :31-36 calcF,
 *   :109-112 cy_lin_lstsqr_mat_residual.
 * X_dev float64 [n,k] row-major, pinv_dev float64 [k,n] row-major = (X'X)^-1 X', k <= 16.
 * d_dev float64 [k] = diag of the caller's invXX (may be NULL when no t/se is requested).
 * Outputs (each may be NULL): beta64/t64 float64 [k, ldt]; se32 float32 [k, ldt];
 * resid64/resid32 [n, ldr]; sse float64 [V]; tss float64 [V] = sum_i (y_i - grand_mean)^2. */
int tmb_glm_direct(const void *Y_dev, int ydtype, int n, int64_t V, int64_t ldy, const double *X_dev,
                   const double *pinv_dev, int k, const double *d_dev, double dof, double grand_mean,
                   double *beta64_dev, double *t64_dev, float *se32_dev, int64_t ldt, double *resid64_dev,
                   float *resid32_dev, int64_t ldr, double *sse_dev, double *tss_dev, void *stream);

/* == cynumstats.pyx:47-52 se_of_slope: se32[j, v] = (float)sqrt(sigma2[v] * d[j]). */
int tmb_se_of_slope(const double *sigma2_dev, int64_t V, const double *d_dev, int k, float *se32_dev, int64_t ld,
                    void *stream);

/* ---------------------------------------------------------------------------------------------
 * Sobel / Aroian / Goodman mediation z == pyfunc.py:130-162 calc_sobelz for P permutations.
 * Path A and path B are two fits of the same data; their pseudo-inverse rows are stacked in one
 * operand laid out like tmb_glm_tstat's At_dev with group width rp (1,2,4,8) and the same `layout`: regressors
 * [0,rA) of permutation j are path A's rows, regressors rA + [0,rB) path B's.
 * GA/dA [P,rA,rA]/[P,rA] and GB/dB [P,rB,rB]/[P,rB] as in tmb_glm_tstat; rowA/rowB select the
 * coefficient whose t enters (calc_beta_se's a[1] / se[1], cynumstats.pyx:66-74).
 * ta_scalar_dev (may be NULL): float64 [P] path-A t as a per-permutation scalar (medtype 'Y',
 * where path A is scipy.stats.linregress(x, dep)); then rA may be 0.
 * alg: 0 = aroian, 1 = sobel, 2 = goodman.   z32_dev float32 [P, ldt] and/or z64_dev float64.
 * ------------------------------------------------------------------------------------------- */
int tmb_sobelz(const void *Y_dev, int ydtype, int n, int64_t V, int64_t ldy, const double *At_dev, int64_t ldA, int rp,
               const double *GA_dev, const double *dA_dev, int rA, int rowA, double dofA, const double *GB_dev,
               const double *dB_dev, int rB, int rowB, double dofB, const double *yy_dev,
               const double *ta_scalar_dev, int P, int alg, float *z32_dev, double *z64_dev, int64_t ldt,
               int layout, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Voxel adjacency == pyfunc.py:48-76 create_adjac_voxel (variant 0) and
 * tools/tm_mulitmodality_adjacency.py:40-66 (variant 1).  mask_host: uint8 [nx,ny,nz] C order.
 * Two-call protocol: first with indices_host == NULL to obtain *num_voxel and *nnz, then with
 * indptr_host int64 [num_voxel+1] and indices_host int32 [nnz].  conn is 26 or 6.
 * ------------------------------------------------------------------------------------------- */
int tmb_voxel_adjacency(int device, const uint8_t *mask_host, int nx, int ny, int nz, int conn, int variant,
                        int32_t *num_voxel, int64_t *nnz, int64_t *indptr_host, int32_t *indices_host);

/* ---------------------------------------------------------------------------------------------
 * FWER-corrected p lookup == tmanalysis/calculate_fweP_vertex.py:37-42,61-69 (and _voxel.py:23-50):
 * corrp[i] = max(searchsorted(sorted_max, values[i], "left") - 1, 0) / n for the ASCENDING sorted null
 * maxima sorted_max_dev (float64 [n], as np.genfromtxt reads the CSV) and TFCE values float32 [m].
 * ------------------------------------------------------------------------------------------- */
int tmb_fwe_lookup(const double *sorted_max_dev, int n, const float *values_dev, int64_t m, double *corrp_dev,
                   void *stream);

/* ---------------------------------------------------------------------------------------------
 * Multi-GPU collection of the per-shuffle maxima (replaces the reference's `echo ... >> perm_*.csv` from N worker
 * processes, STEP_2_tfce_randomise_parallel.py:139-157 + pyfunc.py:119).  One process per GPU; every rank owns a
 * contiguous slice of the permutation range; ONE all-gather per job over NCCL / NVLink.
 *   tmb_comm_unique_id: rank 0 obtains the 128-byte NCCL id and ships it to the other ranks by any host channel
 *   tmb_comm_create:    collective over all ranks
 *   tmb_allgather_max:  global_dev[r*count + i] = rank r's local_dev[i]  (count floats per rank; short slices are
 *                       zero-padded by the caller, whose shard rule tells it every rank's real count)
 * NCCL is bound at run time (dlopen libnccl.so.2); processes that never call these never load it.
 * ------------------------------------------------------------------------------------------- */
typedef struct tmb_comm tmb_comm;
int tmb_comm_unique_id(void *id_out128);
int tmb_comm_create(const void *id128, int rank, int world, int device, tmb_comm **out);
int tmb_comm_destroy(tmb_comm *c);
int tmb_allgather_max(tmb_comm *c, const float *local_dev, int64_t count, float *global_dev, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* TFCE_B200_H */
