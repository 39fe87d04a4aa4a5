// Permuted-design OLS fit with fused SSE -> se -> t epilogue (sm_100a).
//
// Replaces cynumstats.pyx:28-29 (cy_lin_lstsqr_mat), :47-52 (se_of_slope), :59-64 (tval_int),
// :66-74 (calc_beta_se) and the fit half of pyfunc.py:130-162 (calc_sobelz) for P designs at once.
//
// One dense fp64 contraction  B[P*rp, V] = At^T[P*rp, n] . Y[n, V]  with the stacked
// pseudo-inverses of P permuted designs as the left operand, tiled 128 (design rows) x 128
// (vertices) x 32 (subjects) per CTA.  One elected thread streams both operands into a 3-stage
// shared-memory ring with bulk async copies (cp.async.bulk -> UBLKCP, completion on mbarriers); 8 consumer warps run an 8x8 register tile of DFMAs per thread (64 DFMA per 6 LDS.128).  The betas never go to
// HBM: the epilogue turns them into t (or Sobel z) in registers:
//     SSE = yy - b'Gb,  sigma2 = SSE/dof,  se = fl32(sqrt(sigma2 * d)),  t = b / (double)se
// keeping the reference's fp32 rounding of se (cynumstats.pyx:49-51).
#include "common.cuh"

#include <cuda.h>   // CUtensorMap and the cuTensorMapEncodeTiled prototype only: the entry point is fetched at run time
#include <cstdlib>
#include <cstring>

namespace tmb {

static constexpr int BM = 128;      // design rows per tile
static constexpr int TN = 8;        // vertices per thread: columns tn*4..+3 and 64+tn*4..+3 of the tile
static constexpr int BN = 128;      // vertices per tile
static constexpr int BK = 32;       // subjects per stage
static constexpr int STAGES = 3;
static constexpr int kConsumers = 256;
static constexpr int kGlmThreads = kConsumers; // 8 warps; lane 0 of each warp doubles as a copy issuer (a 9th warp would
                                               // round the CTA up to 12 warps of registers and cap the tile at 168 regs)

// ---------------------------------------------------------------- mbarrier / bulk-copy PTX
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// 2-D tiled TMA load (cp.async.bulk.tensor -> UTMALDG): one instruction brings a [rows x box] tile of a row-major matrix
// into shared memory and completes `bytes` on the mbarrier.  The box is four (A, fp64) / eight (Y, fp32) elements WIDER
// than the tile the CTA uses: TMA packs box rows densely, so the box width IS the shared-memory row pitch, and the
// extra columns give exactly the padded pitch whose fragment loads take the minimum number of wavefronts.
__device__ __forceinline__ void tma_load_2d(void *dst_smem, const CUtensorMap *tmap, int x, int y, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            smem_u32(dst_smem)),
        "l"(tmap), "r"(x), "r"(y), "r"(smem_u32(bar))
        : "memory");
}

struct GlmParams {
    const void *Y; int y_is_f64; int n; int64_t V; int64_t ldy;
    const double *At; int64_t ldA;
    const double *G; const double *d;   // [P, r, r], [P, r]
    int P, r, rp, row0, nrows;
    double dof;
    const double *yy;
    float *t32; double *t64; int64_t ldt;
    int nan_to_zero;
    int mode;                           // 0 t-stat, 1 betas, 2 sobel, 3 F statistics, 4 cosinor (stored betas only), 5 sobel from cross-products (DMMA, one row per design), 6 the same for any designs (stored rows), 7 t from cross-products (stored rows)
    // sobel: rows [0, rA) of each design group are path A, rows [rA, rA+rB) path B
    const double *GB; const double *dB; int rA, rB, rowA, rowB; double dofB;
    const double *ta_scalar; int alg;
    const double *sstot;                                         // the reference's own SS_Total (model F numerator) or null
    const double *cfix; double xx; int xpos;
    const int32_t *colmap; int cross_m, cross_f; int64_t ldf;    // mode 6: column sources of the two paths, permuted / fixed cross-product rows                     // mode 5 (Sobel from cross-products): dep'Y per vertex, x'x, position of x in path B
    int cos_nexog, cos_mediation; double cos_ta;                 // mode 4 (cosinor): tested columns, mediation row, path-A t
    // F statistics (mode 3): per design the inverse blocks M_i = inv((X'X)^-1[S_i, S_i]) of every tested variable,
    // stored one after the other (k_i x k_i each, msz doubles per design); variable i covers rows [var_lo[i], +var_k[i])
    const double *M; int msz, nvar; int var_lo[8], var_k[8];
    int exact_epilogue;                 // DMMA kernel: always take the fp64 square root / division (TMB_GLM_EPILOGUE=exact)
    int use_tma;                        // DMMA kernels: operands arrive by 2-D tensor-map loads (else one bulk copy per row)
    int layout;                         // column order of At: 0 = p * rp + i (DFMA tile kernel); 1 = "tile8": 8 designs x rp
                                        // regressors per block, column (p / 8) * 8 * rp + i * 8 + p % 8 (DMMA kernels)
};

// sum_{a,b in [lo,lo+r)} acc[g*RP+a][c] * G[(a-lo)*r + (b-lo)] * acc[g*RP+b][c]; every loop is fully
// unrolled with predicates so the accumulator tile stays in registers.
template <int RP>
__device__ __forceinline__ double quad_form(const double (&acc)[8][TN], int g, int c, const double *__restrict__ G,
                                            int lo, int r) {
    double q = 0.0;
#pragma unroll
    for (int a = 0; a < RP; ++a) {
        if (a >= lo && a < lo + r) {
            double inner = 0.0;
#pragma unroll
            for (int b2 = 0; b2 < RP; ++b2)
                if (b2 >= lo && b2 < lo + r) inner = __fma_rn(G[(a - lo) * r + (b2 - lo)], acc[g * RP + b2][c], inner);
            q = __fma_rn(acc[g * RP + a][c], inner, q);
        }
    }
    return q;
}

template <int RP>
__device__ __forceinline__ double pick_row(const double (&acc)[8][TN], int g, int c, int idx) {
    double v = 0.0;
#pragma unroll
    for (int a = 0; a < RP; ++a)
        if (a == idx) v = acc[g * RP + a][c];
    return v;
}

// se = fl32(sqrt(sigma2 * d)) (cynumstats.pyx:49-51), t = beta / (double)se (cynumstats.pyx:63)
__device__ __forceinline__ double t_from(double beta, double sse, double dof, double d) {
    if (sse < 0.0) sse = 0.0;
    const float se = __double2float_rn(__dsqrt_rn(__dmul_rn(__ddiv_rn(sse, dof), d)));
    return __ddiv_rn(beta, (double)se);
}

// the same with d / dof folded into one factor
__device__ __forceinline__ double t_from_scaled(double beta, double sse, double d_over_dof) {
    if (sse < 0.0) sse = 0.0;
    const float se = __double2float_rn(__dsqrt_rn(__dmul_rn(sse, d_over_dof)));
    return __ddiv_rn(beta, (double)se);
}

// exact widening of a positive normal float (integer pipe instead of the conversion unit)
__device__ __forceinline__ double widen_pos_normal(float f) {
    const unsigned u = __float_as_uint(f);
    return __hiloint2double((int)((u >> 3) + 0x38000000u), (int)(u << 29));
}

// fl32(t) of t_from_scaled with ~8 fp64-pipe instructions instead of ~35.  The epilogue's dependent fp64 chain (square
// root, division) queues behind the other CTA's DMMAs on the one fp64 pipe and held the warps for most of their
// lifetime (ncu: 40% of the stall samples on these two lines).  Both results are only needed to fp32 accuracy:
//   se = fl32(RN64(sqrt(s))),   t32 = fl32(RN64(beta / se))
// so a 2^-43-accurate square root and a 2^-45-accurate quotient (fp32 MUFU seed + one fp64 Newton step each) round
// to the same fp32 value as the correctly rounded fp64 results UNLESS they fall within 2^-41 (relative) of an fp32
// rounding boundary -- then, and for zero / out-of-range / non-finite operands, `slow` is set and the caller takes the
// exact path (probability ~3e-5 per value).  Bit-identical to the exact path by construction.
__device__ __forceinline__ float t32_fast(double beta, double sse, double d_over_dof, bool &slow) {
    if (sse < 0.0) sse = 0.0;
    const double s = __dmul_rn(sse, d_over_dof);
    const float sf = __double2float_rn(s);
    slow = !(sf > 1e-30f && sf < 1e30f);                    // also catches 0, NaN, inf
    float y32;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y32) : "f"(slow ? 1.0f : sf));   // rel. error < 2^-22
    const double y = widen_pos_normal(y32);
    const double g = __dmul_rn(s, y);                        // ~ sqrt(s) (1 + e0)
    const double h = __dmul_rn(0.5, y);
    const double e = __fma_rn(-h, g, 0.5);                   // ~ -e0
    const double g1 = __fma_rn(g, e, g);                     // sqrt(s) (1 - 1.5 e0^2): rel. error < 2^-43
    const unsigned lowg = (unsigned)__double2loint(g1) & 0x1FFFFFFFu;   // the 29 bits below fp32 precision
    slow |= (lowg - (0x10000000u - 4096u)) < 8192u;          // within 2^12 ulp64 of a rounding boundary
    const float se = __double2float_rn(g1);
    float r32;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r32) : "f"(slow ? 1.0f : se));     // rel. error < 2^-22.9
    const double sed = widen_pos_normal(slow ? 1.0f : se);
    const double r = widen_pos_normal(r32);
    const double e2 = __fma_rn(-sed, r, 1.0);
    const double r1 = __fma_rn(r, e2, r);                    // 1/se: rel. error < 2^-45
    const double q = __dmul_rn(beta, r1);
    const unsigned lowq = (unsigned)__double2loint(q) & 0x1FFFFFFFu;
    slow |= (lowq - (0x10000000u - 4096u)) < 8192u;
    const unsigned ex = ((unsigned)__double2hiint(q) >> 20) & 0x7ffu;  // fp32-normal quotient (or exactly zero) only
    slow |= !((ex - (1023u - 120u)) <= 240u || q == 0.0);
    return __double2float_rn(q);
}

// beta / fl32(sqrt(sse * dscale)) in float64 with relative error < 2^-44 (the float32 rounding of se is exact: `slow` is set
// when the square root lies within 2^12 ulp64 of a float32 rounding boundary or outside the float32 normal range).  Same
// seeds and Newton steps as t32_fast.
__device__ __forceinline__ double t64_fast(double beta, double sse, double dscale, bool &slow) {
    if (sse < 0.0) sse = 0.0;
    const double s = __dmul_rn(sse, dscale);
    const float sf = __double2float_rn(s);
    const bool bad = !(sf > 1e-30f && sf < 1e30f);
    slow |= bad;
    float y32;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y32) : "f"(bad ? 1.0f : sf));
    const double y = widen_pos_normal(y32);
    const double g = __dmul_rn(s, y);
    const double h = __dmul_rn(0.5, y);
    const double e = __fma_rn(-h, g, 0.5);
    const double g1 = __fma_rn(g, e, g);                     // sqrt(s): rel. error < 2^-43
    const unsigned lowg = (unsigned)__double2loint(g1) & 0x1FFFFFFFu;
    slow |= (lowg - (0x10000000u - 4096u)) < 8192u;
    const float se = __double2float_rn(g1);
    const bool bad2 = !(se > 1e-30f && se < 1e30f);
    slow |= bad2;
    float r32;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r32) : "f"(bad2 ? 1.0f : se));
    const double sed = widen_pos_normal(bad2 ? 1.0f : se);
    const double r = widen_pos_normal(r32);
    const double e2 = __fma_rn(-sed, r, 1.0);
    const double r1 = __fma_rn(r, e2, r);                    // 1 / se: rel. error < 2^-45
    return __dmul_rn(beta, r1);
}

// fl32(sqrt(num / den)) for positive operands known to ~2^-42: float32 seeds + one Newton step each, accepted unless the
// result lies within 2^14 ulp64 of a float32 rounding boundary (accumulated relative error of the caller < 2^-41.5) or
// outside the float32 normal range.
__device__ __forceinline__ float sqrt_ratio32_fast(double num, double den, bool &slow) {
    const float df = __double2float_rn(den);
    const bool bad = !(df > 1e-30f && df < 1e30f);
    slow |= bad;
    float r32;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r32) : "f"(bad ? 1.0f : df));
    const double r = widen_pos_normal(r32);
    const double e = __fma_rn(-den, r, 1.0);
    const double r1 = __fma_rn(r, e, r);                     // 1 / den: rel. error < 2^-44
    const double w = __dmul_rn(num, r1);
    const float wf = __double2float_rn(w);
    const bool bad2 = !(wf > 1e-30f && wf < 1e30f);
    slow |= bad2;
    float y32;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y32) : "f"(bad2 ? 1.0f : wf));
    const double y = widen_pos_normal(y32);
    const double g = __dmul_rn(w, y);
    const double h = __dmul_rn(0.5, y);
    const double e2 = __fma_rn(-h, g, 0.5);
    const double g1 = __fma_rn(g, e2, g);                    // sqrt(w): rel. error < 2^-43
    const unsigned low = (unsigned)__double2loint(g1) & 0x1FFFFFFFu;
    slow |= (low - (0x10000000u - 16384u)) < 32768u;
    return __double2float_rn(g1);
}

// one output row of a thread's tile: two 128-bit stores (float) / eight scalar stores (double)
__device__ __forceinline__ void store_tile_row(float *t32, double *t64, size_t off, int64_t v0, int tn,
                                               const float (&o32)[8], const double (&o64)[8]) {
    const size_t c0 = off + v0 + tn * 4;
    if (t32) {
        *reinterpret_cast<float4 *>(t32 + c0) = make_float4(o32[0], o32[1], o32[2], o32[3]);
        *reinterpret_cast<float4 *>(t32 + c0 + 64) = make_float4(o32[4], o32[5], o32[6], o32[7]);
    }
    if (t64) {
#pragma unroll
        for (int c = 0; c < 4; ++c) { t64[c0 + c] = o64[c]; t64[c0 + 64 + c] = o64[4 + c]; }
    }
}

// column c (0..7) of a thread's tile: two groups of four, 64 vertices apart (conflict-free LDS.128)
__device__ __forceinline__ int64_t tile_col(int64_t v0, int tn, int c) { return v0 + (c >> 2) * 64 + tn * 4 + (c & 3); }

template <int RP>
__device__ __forceinline__ void epilogue(const GlmParams &p, const double (&acc)[8][TN], int m_base, int64_t v0, int tn) {
    // thread owns design rows m_base .. m_base+7 and the vertices tile_col(v0, tn, 0..7)
    if (p.mode == 1) { // betas
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int m = m_base + i;
            if (m < p.P * RP) {
                double *dst = p.t64 + (size_t)m * p.ldt;
#pragma unroll
                for (int c = 0; c < TN; ++c)
                    if (tile_col(v0, tn, c) < p.V) dst[tile_col(v0, tn, c)] = acc[i][c];
            }
        }
        return;
    }
    double yy[TN];
#pragma unroll
    for (int c = 0; c < TN; ++c) yy[c] = (tile_col(v0, tn, c) < p.V) ? p.yy[tile_col(v0, tn, c)] : 0.0;

#pragma unroll
    for (int g = 0; g < 8 / RP; ++g) {
        const int perm = m_base / RP + g;
        if (perm >= p.P) continue;
        if (p.mode == 0) {
            const int r = p.r;
            const double *G = p.G + (size_t)perm * r * r;
            const double *dg = p.d + (size_t)perm * r;
            double sse[TN];
#pragma unroll
            for (int c = 0; c < TN; ++c) sse[c] = yy[c] - quad_form<RP>(acc, g, c, G, 0, r);
#pragma unroll
            for (int a = 0; a < RP; ++a) {
                if (a < p.row0 || a >= p.row0 + p.nrows || a >= r) continue;
                const double da = dg[a];
                float o32[TN];
                double o64[TN];
#pragma unroll
                for (int c = 0; c < TN; ++c) {
                    double t = t_from(acc[g * RP + a][c], sse[c], p.dof, da);
                    if (p.nan_to_zero && t != t) t = 0.0;
                    if (tile_col(v0, tn, c) >= p.V) t = 0.0;
                    o64[c] = t;
                    o32[c] = __double2float_rn(t);
                }
                const size_t off = ((size_t)perm * p.nrows + (a - p.row0)) * p.ldt;
                store_tile_row(p.t32, p.t64, off, v0, tn, o32, o64);
            }
        } else if (p.mode == 3) {
            // F statistics of pyfunc.py:2282-2401 glm_typeI: model F = ((TSS - RSS)/(k-1)) / (RSS/(n-k)) and per tested
            // variable i the partial F = ((RSS_without_i - RSS)/k_i) / (RSS/(n-k)).  The extra sum of squares of a set S
            // of regressors needs no second fit: RSS_without_S - RSS = b_S' inv(C_SS) b_S with C = (X'X)^-1.
            // Output rows: [model F when row0 == 0], then one row per variable.
            const int r = p.r;
            const double *G = p.G + (size_t)perm * r * r;
            double ssb[TN], ms[TN];
#pragma unroll
            for (int c = 0; c < TN; ++c) {
                ssb[c] = quad_form<RP>(acc, g, c, G, 0, r);
                ms[c] = __ddiv_rn(yy[c] - ssb[c], p.dof);
            }
            const int first = p.row0 == 0 ? 1 : 0;
            float o32[TN];
            double o64[TN];
            if (first) {
#pragma unroll
                for (int c = 0; c < TN; ++c) {
                    // model F = ((SS_Total - SS_Residuals) / (k-1)) / MS_Residuals with the reference's own SS_Total
                    const int64_t vc = tile_col(v0, tn, c);
                    const double between = (p.sstot && vc < p.V) ? __dsub_rn(p.sstot[vc], yy[c] - ssb[c]) : ssb[c];
                    double f = __ddiv_rn(__ddiv_rn(between, (double)r), ms[c]);
                    if (p.nan_to_zero && f != f) f = 0.0;
                    if (vc >= p.V) f = 0.0;
                    o64[c] = f;
                    o32[c] = __double2float_rn(f);
                }
                store_tile_row(p.t32, p.t64, (size_t)perm * p.nrows * p.ldt, v0, tn, o32, o64);
            }
            const double *M = p.M + (size_t)perm * p.msz;
#pragma unroll 1
            for (int i = 0; i < p.nvar; ++i) {
                const int lo = p.var_lo[i], ki = p.var_k[i];
#pragma unroll
                for (int c = 0; c < TN; ++c) {
                    const double num = quad_form<RP>(acc, g, c, M, lo, ki);
                    double f = __ddiv_rn(num, __dmul_rn(ms[c], (double)ki));
                    if (p.nan_to_zero && f != f) f = 0.0;
                    if (tile_col(v0, tn, c) >= p.V) f = 0.0;
                    o64[c] = f;
                    o32[c] = __double2float_rn(f);
                }
                store_tile_row(p.t32, p.t64, ((size_t)perm * p.nrows + first + i) * p.ldt, v0, tn, o32, o64);
                M += ki * ki;
            }
        } else { // sobel (pyfunc.py:130-162)
            const int rA = p.rA, rB = p.rB;
            const double *GB = p.GB + (size_t)perm * rB * rB;
            const double dBr = p.dB[(size_t)perm * rB + p.rowB];
            float o32[TN];
            double o64[TN];
#pragma unroll
            for (int c = 0; c < TN; ++c) {
                double ta;
                if (p.ta_scalar) {
                    ta = p.ta_scalar[perm];
                } else {
                    const double sseA = yy[c] - quad_form<RP>(acc, g, c, p.G + (size_t)perm * rA * rA, 0, rA);
                    ta = t_from(pick_row<RP>(acc, g, c, p.rowA), sseA, p.dof, p.d[(size_t)perm * rA + p.rowA]);
                }
                const double sseB = yy[c] - quad_form<RP>(acc, g, c, GB, rA, rB);
                const double tb = t_from(pick_row<RP>(acc, g, c, rA + p.rowB), sseB, p.dofB, dBr);
                // 1/sqrt(1/tb^2 + 1/ta^2 (+|-) 1/(ta^2 tb^2)), evaluated in the reference's order
                const double ta2 = __dmul_rn(ta, ta), tb2 = __dmul_rn(tb, tb);
                double s = __dadd_rn(__ddiv_rn(1.0, tb2), __ddiv_rn(1.0, ta2));
                const double cross = __ddiv_rn(1.0, __dmul_rn(ta2, tb2));
                if (p.alg == 0) s = __dadd_rn(s, cross);
                else if (p.alg == 2) s = __dsub_rn(s, cross);
                double z = __ddiv_rn(1.0, __dsqrt_rn(s));
                if (tile_col(v0, tn, c) >= p.V) z = 0.0;
                o64[c] = z;
                o32[c] = __double2float_rn(z);
            }
            const size_t off = (size_t)perm * p.ldt;
            store_tile_row(p.t32, p.t64, off, v0, tn, o32, o64);
        }
    }
}

template <typename YT>
__device__ __forceinline__ void load_y4(const YT *p, double *y);
template <>
__device__ __forceinline__ void load_y4<float>(const float *p, double *y) {
    const float4 f = *reinterpret_cast<const float4 *>(p);
    y[0] = (double)f.x; y[1] = (double)f.y; y[2] = (double)f.z; y[3] = (double)f.w;
}
template <>
__device__ __forceinline__ void load_y4<double>(const double *p, double *y) {
    const double2 a = *reinterpret_cast<const double2 *>(p);
    const double2 b = *reinterpret_cast<const double2 *>(p + 2);
    y[0] = a.x; y[1] = a.y; y[2] = b.x; y[3] = b.y;
}

template <int RP, typename YT>
__global__ void __launch_bounds__(kGlmThreads, 1) glm_tile_kernel(GlmParams p, int mtiles) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *sA = reinterpret_cast<double *>(smem_raw);                             // [STAGES][BK][BM]
    YT *sY = reinterpret_cast<YT *>(smem_raw + sizeof(double) * STAGES * BK * BM); // [STAGES][BK][BN]
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + sizeof(double) * STAGES * BK * BM +
                                                   sizeof(YT) * STAGES * BK * BN);
    uint64_t *empty = full + STAGES;

    const int tid = threadIdx.x;
    const int tile = blockIdx.x;
    const int mt = tile % mtiles;
    const int64_t vt = tile / mtiles;
    const int m0 = mt * BM;
    const int64_t v0 = vt * BN;
    const int nchunks = (p.n + BK - 1) / BK;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, kConsumers / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    // issue the bulk copies of chunk kc into its ring slot.  Lane 0 of EVERY warp issues the rows warp, warp + 8, ... of
    // the chunk (8 of the 64 row copies): one thread issuing all 64 took about as long as a warp's share of the chunk's
    // arithmetic (ncu: half of the stall samples sat on the consumers' wait for the ring), and it skewed warp 0 against
    // the others.  Warp 0 also posts the expected byte count; the slot's phase cannot complete before it has (the
    // barrier's one pending arrival), whatever the order in which the copies land.
    const int lane_i = tid & 31, warp_i = tid >> 5;
    auto issue_chunk = [&](int kc) {
        const int s = kc % STAGES;
        const int round = kc / STAGES;
        mbar_wait(empty + s, (round & 1) ^ 1); // every consumer warp has released the slot
        const int k0 = kc * BK;
        const int rows = min(BK, p.n - k0);
        if (warp_i == 0) mbar_expect_tx(full + s, (uint32_t)rows * (BM * 8 + BN * (uint32_t)sizeof(YT)));
        for (int kk = warp_i; kk < rows; kk += kConsumers / 32) {
            bulk_g2s(sA + ((size_t)s * BK + kk) * BM, p.At + (size_t)(k0 + kk) * p.ldA + m0, BM * 8, full + s);
            bulk_g2s(sY + ((size_t)s * BK + kk) * BN, reinterpret_cast<const YT *>(p.Y) + (size_t)(k0 + kk) * p.ldy + v0,
                     BN * (uint32_t)sizeof(YT), full + s);
        }
    };
    if (lane_i == 0)
        for (int kc = 0; kc < STAGES - 1 && kc < nchunks; ++kc) issue_chunk(kc); // prologue: fill the ring

    // ===== consumers =====
    const int tm = tid >> 4;  // 0..15 -> design rows tm*8 .. +7   (two values per warp: A reads broadcast per half-warp)
    const int tn = tid & 15;  // 0..15 -> vertices tn*4 .. +3 and 64+tn*4 .. +3
    double acc[8][TN];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int c = 0; c < TN; ++c) acc[i][c] = 0.0;

    for (int kc = 0; kc < nchunks; ++kc) {
        const int s = kc % STAGES;
        const int round = kc / STAGES;
        // keep STAGES-1 chunks in flight: the slot released by iteration kc-1 is refilled now
        if (lane_i == 0 && kc + STAGES - 1 < nchunks) issue_chunk(kc + STAGES - 1);
        mbar_wait(full + s, round & 1);
        const int rows = min(BK, p.n - kc * BK);
        const double *a_base = sA + (size_t)s * BK * BM + tm * 8;
        const YT *y_base = sY + (size_t)s * BK * BN + tn * 4;
#pragma unroll 4
        for (int kk = 0; kk < rows; ++kk) {
            const double2 a01 = *reinterpret_cast<const double2 *>(a_base + kk * BM);
            const double2 a23 = *reinterpret_cast<const double2 *>(a_base + kk * BM + 2);
            const double2 a45 = *reinterpret_cast<const double2 *>(a_base + kk * BM + 4);
            const double2 a67 = *reinterpret_cast<const double2 *>(a_base + kk * BM + 6);
            const double a[8] = {a01.x, a01.y, a23.x, a23.y, a45.x, a45.y, a67.x, a67.y};
            double y[TN];
            load_y4<YT>(y_base + kk * BN, y);
            load_y4<YT>(y_base + kk * BN + 64, y + 4);
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int c = 0; c < TN; ++c) acc[i][c] = __fma_rn(a[i], y[c], acc[i][c]);
        }
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(empty + s); // one arrival per consumer warp
    }
    epilogue<RP>(p, acc, m0 + tm * 8, v0, tn);
}

// ---------------------------------------------------------------- fp64 tensor-core (DMMA) variant
// The same contraction on the fp64 tensor cores: mma.sync.m8n8k4.f64 (SASS DMMA).  On B200 the DFMA
// version above saturates the fp64 vector pipe (ncu: math-pipe throttle); the tensor path has twice the
// fp64 throughput.  Used for the headline case r == 1 (one slope per design, k = 2), whose epilogue needs
// no exchange between threads: every accumulator element is one (design, vertex) pair.
// CTA tile 64 (designs) x 128 (vertices) x 32 (subjects); 8 warps as 2 (m) x 4 (n), warp tile 32 x 32 =
// 4 x 4 DMMA tiles (4 + 4 fragment loads and 4 conversions per 16 DMMAs).  Shared-memory rows are pitched (+4 doubles / +8 floats) so that the fragment loads
// -- A[m = lane/4][k = lane%4], B[k = lane%4][n = lane/4] -- are bank-conflict free.
static constexpr int DM = 64, DN = 128, DK = 32, DSTAGES = 3;
static constexpr int DPA = DM + 4;   // pitch of an A row (doubles)
static constexpr int DPY = DN + 8;   // pitch of a Y row (elements)

__device__ __forceinline__ void dmma_884(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

template <typename YT>
__global__ void __launch_bounds__(256, 2) glm_dmma_kernel(GlmParams p, int mtiles, const __grid_constant__ CUtensorMap tmA,
                                                          const __grid_constant__ CUtensorMap tmY) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *sA = reinterpret_cast<double *>(smem_raw);                                  // [DSTAGES][DK][DPA]
    YT *sY = reinterpret_cast<YT *>(smem_raw + sizeof(double) * DSTAGES * DK * DPA);    // [DSTAGES][DK][DPY]
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + sizeof(double) * DSTAGES * DK * DPA +
                                                   sizeof(YT) * DSTAGES * DK * DPY);
    uint64_t *empty = full + DSTAGES;
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int tile = blockIdx.x;
    const int mt = tile % mtiles;
    const int64_t vt = tile / mtiles;
    const int m0 = mt * DM;
    const int64_t v0 = vt * DN;
    const int nchunks = (p.n + DK - 1) / DK;
    if (tid == 0) {
        for (int s = 0; s < DSTAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    auto issue_chunk = [&](int kc) {
        const int s = kc % DSTAGES;
        const int round = kc / DSTAGES;
        mbar_wait(empty + s, (round & 1) ^ 1);
        const int k0 = kc * DK;
        const int rows = min(DK, p.n - k0);
        if (p.use_tma) {
            // two tensor-map loads per chunk, issued by warp 0 alone; rows beyond n arrive as zeros (out-of-bounds fill)
            if (warp == 0) {
                mbar_expect_tx(full + s, (uint32_t)DK * (DPA * 8 + DPY * (uint32_t)sizeof(YT)));
                tma_load_2d(sA + (size_t)s * DK * DPA, &tmA, m0, k0, full + s);
                tma_load_2d(sY + (size_t)s * DK * DPY, &tmY, (int)v0, k0, full + s);
            }
            return;
        }
        if (warp == 0) mbar_expect_tx(full + s, (uint32_t)rows * (DM * 8 + DN * (uint32_t)sizeof(YT)));
        for (int kk = warp; kk < rows; kk += 8) { // lane 0 of every warp issues an eighth of the row copies (see glm_tile_kernel)
            bulk_g2s(sA + ((size_t)s * DK + kk) * DPA, p.At + (size_t)(k0 + kk) * p.ldA + m0, DM * 8, full + s);
            bulk_g2s(sY + ((size_t)s * DK + kk) * DPY, reinterpret_cast<const YT *>(p.Y) + (size_t)(k0 + kk) * p.ldy + v0,
                     DN * (uint32_t)sizeof(YT), full + s);
        }
    };
    const bool issuer = lane == 0 && (warp == 0 || !p.use_tma);   // tensor-map loads: warp 0 alone feeds the ring
    if (issuer)
        for (int kc = 0; kc < DSTAGES - 1 && kc < nchunks; ++kc) issue_chunk(kc);

    const int wm = warp >> 2, wn = warp & 3;      // warp tile: rows wm*32.., cols wn*32..
    const int g = lane >> 2, t4 = lane & 3;       // fragment coordinates
    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    for (int kc = 0; kc < nchunks; ++kc) {
        const int s = kc % DSTAGES;
        const int round = kc / DSTAGES;
        if (issuer && kc + DSTAGES - 1 < nchunks) issue_chunk(kc + DSTAGES - 1);
        mbar_wait(full + s, round & 1);
        const int rows = min(DK, p.n - kc * DK);
        const double *a_st = sA + (size_t)s * DK * DPA + wm * 32 + g;
        const YT *y_st = sY + (size_t)s * DK * DPY + wn * 32 + g;
        for (int k4 = 0; k4 < rows; k4 += 4) {
            const int kr = k4 + t4;
            const bool ok = kr < rows;   // the last chunk of n may be ragged: pad the k-slice with zeros
            double a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = ok ? a_st[(size_t)kr * DPA + i * 8] : 0.0;
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = ok ? (double)y_st[(size_t)kr * DPY + j * 8] : 0.0;
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma_884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + s);
    }
    // epilogue (r == 1): element (i, j, e) is design m0 + wm*32 + i*8 + g, vertex v0 + wn*32 + j*8 + t4*2 + e
    if (p.mode == 1) { // the plain contraction (tmb_glm_beta): row m of the output is column m of At
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int row = m0 + wm * 32 + i * 8 + g;
            if (row >= p.P) continue;
            double *dst = p.t64 + (size_t)row * p.ldt;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int64_t v = v0 + wn * 32 + j * 8 + t4 * 2;
                if (v < p.V) dst[v] = acc[i][j][0];
                if (v + 1 < p.V) dst[v + 1] = acc[i][j][1];
            }
        }
        return;
    }
    const bool fast = p.t64 == nullptr && !p.exact_epilogue; // fp32 output only
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int perm = m0 + wm * 32 + i * 8 + g;
        if (perm >= p.P) continue;
        // sigma2 * d = sse * (d / dof): the quotient is formed once per design, which removes one of the epilogue's two
        // fp64 divisions per statistic (ncu: the epilogue's divisions and square root keep the fp64 pipe throttled for
        // ~45% of the kernel's samples).  fp64 rounding differs from (sse / dof) * d by <= 1 ulp before the reference's
        // own rounding of se to fp32.
        const double G = p.G[perm], dscale = __ddiv_rn(p.d[perm], p.dof);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int64_t v = v0 + wn * 32 + j * 8 + t4 * 2;
            float o32[2];
            double o64[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const double beta = acc[i][j][e];
                const double yyv = (v + e < p.V) ? p.yy[v + e] : 0.0;
                const double sse = yyv - __dmul_rn(beta, __dmul_rn(G, beta));
                bool slow = true;
                float tf = 0.f;
                if (fast) tf = t32_fast(beta, sse, dscale, slow);
                double t = 0.0;
                if (slow) { // exact path: fp64 outputs requested, or the fast value is too close to an fp32 rounding boundary
                    t = p.exact_epilogue == 2 ? t_from(beta, sse, p.dof, p.d[perm]) : t_from_scaled(beta, sse, dscale);
                    if (p.nan_to_zero && t != t) t = 0.0;
                    tf = __double2float_rn(t);
                }
                if (v + e >= p.V) { t = 0.0; tf = 0.f; }
                o64[e] = t;
                o32[e] = tf;
            }
            const size_t off = (size_t)perm * p.ldt + v;
            if (p.t32) *reinterpret_cast<float2 *>(p.t32 + off) = make_float2(o32[0], o32[1]);
            if (p.t64) { p.t64[off] = o64[0]; p.t64[off + 1] = o64[1]; }
        }
    }
}


// ---------------------------------------------------------------- tensor maps for the DMMA kernels' operands
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        const char *off = getenv("TMB_GLM_TMA");     // "0": one bulk copy per operand row instead (A/B measurements)
        if (!(off && off[0] == '0')) {
            void *sym = nullptr;
            cudaDriverEntryPointQueryResult q;
            if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess &&
                q == cudaDriverEntryPointSuccess)
                fn = reinterpret_cast<EncodeTiledFn>(sym);
            else
                cudaGetLastError();
        }
    }
    return fn;
}

// row-major [rows, ld] matrix of fp64 / fp32 elements, box = box_cols x box_rows (inner dimension first)
static bool make_tmap(CUtensorMap *m, const void *base, bool f64, int64_t ld, int64_t rows, int box_cols, int box_rows) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)ld, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * (f64 ? 8u : 4u)};
    const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    return fn(m, f64 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void *>(base), dims,
              strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <typename YT>
static int launch_dmma(const GlmParams &p, cudaStream_t stream) {
    const int mtiles = (p.P + DM - 1) / DM;
    const int64_t vtiles = (p.V + DN - 1) / DN;
    const int64_t tiles = (int64_t)mtiles * vtiles;
    TMB_REQUIRE(tiles < (int64_t)INT32_MAX, "glm: too many tiles");
    const size_t smem = sizeof(double) * DSTAGES * DK * DPA + sizeof(YT) * DSTAGES * DK * DPY + sizeof(uint64_t) * 2 * DSTAGES;
    TMB_CUDA(cudaFuncSetAttribute(glm_dmma_kernel<YT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GlmParams q = p;
    CUtensorMap tmA, tmY;
    memset(&tmA, 0, sizeof(tmA)); memset(&tmY, 0, sizeof(tmY));
    q.use_tma = make_tmap(&tmA, p.At, true, p.ldA, p.n, DPA, DK) && make_tmap(&tmY, p.Y, sizeof(YT) == 8, p.ldy, p.n, DPY, DK);
    glm_dmma_kernel<YT><<<(unsigned)tiles, 256, smem, stream>>>(q, mtiles, tmA, tmY);
    count_launch();
    TMB_CUDA(cudaGetLastError());
    return 0;
}


// ---------------------------------------------------------------- DMMA, several regressors per design (rp = 2, 4, 8)
// Same contraction and operand ring as glm_dmma_kernel.  The columns of At come in the "tile8" order -- blocks of 8
// designs, inside a block first regressor 0 of the 8 designs, then regressor 1, ... -- so that one 8-row DMMA tile holds
// ONE regressor of 8 designs and a thread's accumulators of RP consecutive row tiles hold ALL the betas of one
// (design, vertex) pair: the fused epilogue (t, partial F, Sobel z) needs no exchange between threads.
//   RP <= 4: 8 warps as 2 (m) x 4 (n), warp tile 32 x 32 (4 x 4 DMMA tiles, 4 / RP design groups per warp)
//   RP == 8: 8 warps as 1 (m) x 8 (n), warp tile 64 x 16 (8 x 2 DMMA tiles, one design group)
//   RP = 3     : CTA tile 48 rows, 2 x 4 warps, warp tile 24 x 32 (no padded regressor row: k = 4 regression, Sobel 'M'/'I')
//   RP = 5,6,7 : CTA tile 8 * RP rows, 1 x 8 warps, warp tile (8 * RP) x 16
template <int RP>
struct DmmaShape {
    static constexpr int WM = RP <= 4 ? 2 : 1;                        // warps along m
    static constexpr int MT = RP <= 2 ? 4 : RP;                       // DMMA row tiles per warp (a multiple of RP)
    static constexpr int NT = RP <= 4 ? 4 : 2;                        // DMMA column tiles per warp
    static constexpr int ROWS = WM * MT * 8;                          // CTA tile rows: 64, 64, 48, 64, 40, 48, 56, 64
    static constexpr int PITCH = ROWS + 4;                            // doubles per staged A row (pitch mod 16 = 4 or 12:
                                                                      // fragment loads take the minimum two wavefronts)
};

// b'Gb over regressors [lo, lo + r) of one (design, vertex); G row-major r x r in global memory (L1-resident: 8 designs per warp)
template <int RP>
__device__ __forceinline__ double dq_form(const double (&b)[RP], const double *__restrict__ G, int lo, int r) {
    double q = 0.0;
#pragma unroll
    for (int a = 0; a < RP; ++a) {
        if (a >= lo && a < lo + r) {
            double inner = 0.0;
#pragma unroll
            for (int c = 0; c < RP; ++c)
                if (c >= lo && c < lo + r) inner = __fma_rn(__ldg(G + (a - lo) * r + (c - lo)), b[c], inner);
            q = __fma_rn(b[a], inner, q);
        }
    }
    return q;
}

template <int RP>
__device__ __forceinline__ double dpick(const double (&b)[RP], int idx) {
    double v = 0.0;
#pragma unroll
    for (int a = 0; a < RP; ++a)
        if (a == idx) v = b[a];
    return v;
}

// Sobel z of pyfunc.py:130-162 for medtype 'M' / 'I' from ONE contraction row per shuffle (mode 5, tmb_sobelz_cross).  Only
// pred_x is permuted (vertex_tfce_mediation_randomise.py:82-90), so of the centred cross-products c_x = x_p'y and
// c_d = dep'y only c_x changes with the shuffle; c_d comes from one extra row fitted once (p.cfix).  Path A, y ~ [1, x_p]:
// beta = c_x / x'x, SSE = yy - c_x beta.  Path B, y ~ [1, dep, x_p] ('M', xpos 1) or [1, x_p, dep] ('I', xpos 0):
// beta = C c with C = inverse of the shuffle's centred 2 x 2 Gram matrix, SSE = yy - c'beta.  The per-design constants
// (C00, C01, C10, C11, C[rowB][rowB] / dofB) are loaded once per design by the caller.
// The epilogue shares the fp64 pipe with the DMMAs, so it is kept to three divisions and three square roots per value:
// 1 / x'x and d / dof are formed on the host (the k = 2 kernel does the same: <= 1 ulp in float64 before the reference's
// own rounding of se to float32), and z = 1 / sqrt(1/tb^2 + 1/ta^2 + 1/(ta^2 tb^2)) is evaluated as
// sqrt(ta^2 tb^2 / (ta^2 + tb^2 + 1)) unless a square is zero, subnormal or overflowing -- then the reference's own sequence
// runs, for its inf / NaN results.  float32 output only: the two t values and the square root of the ratio come from
// float32 seeds plus one Newton step each (t64_fast, sqrt_ratio32_fast), accepted unless an intermediate float32 rounding
// (se of either path, z itself) is too close to call -- then the exact sequence runs.  Bit-identical to it by construction
// (tests compare the two on whole blocks).
__device__ __forceinline__ void sobel_cross_value(const GlmParams &p, int perm, double cx, double cd, double yyv, int64_t v,
                                                  double C00, double C01, double C10, double C11, double dsB) {
    const bool inside = v < p.V;
    const double bA = __dmul_rn(cx, p.xx);                               // p.xx = 1 / x'x
    const double sseA = yyv - __dmul_rn(cx, bA);
    const double c0 = p.xpos == 0 ? cx : cd, c1 = p.xpos == 0 ? cd : cx;
    const double b0 = __fma_rn(C01, c1, __dmul_rn(C00, c0));
    const double b1 = __fma_rn(C11, c1, __dmul_rn(C10, c0));
    const double sseB = yyv - __fma_rn(c1, b1, __dmul_rn(c0, b0));
    const double bB = p.rowB ? b1 : b0;
    if (p.t64 == nullptr && !p.exact_epilogue) {
        bool slow = false;
        const double fa = t64_fast(bA, sseA, p.dof, slow), fb = t64_fast(bB, sseB, dsB, slow);
        const double fa2 = __dmul_rn(fa, fa), fb2 = __dmul_rn(fb, fb);
        const double fsum = __dadd_rn(fa2, fb2);
        const double fden = p.alg == 0 ? __dadd_rn(fsum, 1.0) : p.alg == 2 ? __dsub_rn(fsum, 1.0) : fsum;
        if (p.alg == 2 && !(fsum > 2.0)) slow = true;       // Goodman: ta^2 + tb^2 - 1 cancels, the error bound does not hold
        const float zf = sqrt_ratio32_fast(__dmul_rn(fa2, fb2), fden, slow);
        if (!slow) {
            p.t32[(size_t)perm * p.ldt + v] = inside ? zf : 0.f;
            return;
        }
    }
    const double ta = t_from_scaled(bA, sseA, p.dof);                    // p.dof = (1 / x'x) / dofA
    const double tb = t_from_scaled(bB, sseB, dsB);
    const double ta2 = __dmul_rn(ta, ta), tb2 = __dmul_rn(tb, tb);
    const double prod = __dmul_rn(ta2, tb2);
    double z;
    if (prod > 1e-280 && prod < 1e280) {
        const double sum = __dadd_rn(ta2, tb2);
        const double den = p.alg == 0 ? __dadd_rn(sum, 1.0) : p.alg == 2 ? __dsub_rn(sum, 1.0) : sum;
        z = __dsqrt_rn(__ddiv_rn(prod, den));
    } else {
        double s = __dadd_rn(__ddiv_rn(1.0, tb2), __ddiv_rn(1.0, ta2));
        const double cross = __ddiv_rn(1.0, prod);
        if (p.alg == 0) s = __dadd_rn(s, cross);
        else if (p.alg == 2) s = __dsub_rn(s, cross);
        z = __ddiv_rn(1.0, __dsqrt_rn(s));
    }
    if (!inside) z = 0.0;
    const size_t off = (size_t)perm * p.ldt + v;
    if (p.t32) p.t32[off] = __double2float_rn(z);
    if (p.t64) p.t64[off] = z;
}

// all statistics of one (design, vertex) from its RP betas; same arithmetic as epilogue<RP> of the DFMA kernel
template <int RP>
__device__ __forceinline__ void dmma_vertex_stats(const GlmParams &p, int perm, const double (&b)[RP], double yyv, int64_t v) {
    const bool inside = v < p.V;
    if (p.mode == 0) {
        const int r = p.r;
        const double sse = yyv - dq_form<RP>(b, p.G + (size_t)perm * r * r, 0, r);
        const double *dg = p.d + (size_t)perm * r;
#pragma unroll
        for (int a = 0; a < RP; ++a) {
            if (a < p.row0 || a >= p.row0 + p.nrows || a >= r) continue;
            double t = t_from(b[a], sse, p.dof, __ldg(dg + a));
            if (p.nan_to_zero && t != t) t = 0.0;
            if (!inside) t = 0.0;
            const size_t off = ((size_t)perm * p.nrows + (a - p.row0)) * p.ldt + v;
            if (p.t32) p.t32[off] = __double2float_rn(t);
            if (p.t64) p.t64[off] = t;
        }
    } else if (p.mode == 3) {
        const int r = p.r;
        const double ssb = dq_form<RP>(b, p.G + (size_t)perm * r * r, 0, r);
        const double ms = __ddiv_rn(yyv - ssb, p.dof);
        const int first = p.row0 == 0 ? 1 : 0;
        if (first) {
            const double between = (p.sstot && inside) ? __dsub_rn(p.sstot[v], yyv - ssb) : ssb;
            double f = __ddiv_rn(__ddiv_rn(between, (double)r), ms);
            if (p.nan_to_zero && f != f) f = 0.0;
            if (!inside) f = 0.0;
            const size_t off = (size_t)perm * p.nrows * p.ldt + v;
            if (p.t32) p.t32[off] = __double2float_rn(f);
            if (p.t64) p.t64[off] = f;
        }
        const double *M = p.M + (size_t)perm * p.msz;
#pragma unroll 1
        for (int i = 0; i < p.nvar; ++i) {
            const int lo = p.var_lo[i], ki = p.var_k[i];
            double f = __ddiv_rn(dq_form<RP>(b, M, lo, ki), __dmul_rn(ms, (double)ki));
            if (p.nan_to_zero && f != f) f = 0.0;
            if (!inside) f = 0.0;
            const size_t off = ((size_t)perm * p.nrows + first + i) * p.ldt + v;
            if (p.t32) p.t32[off] = __double2float_rn(f);
            if (p.t64) p.t64[off] = f;
            M += ki * ki;
        }
    } else { // sobel (pyfunc.py:130-162)
        const int rA = p.rA, rB = p.rB;
        double ta;
        if (p.ta_scalar) {
            ta = p.ta_scalar[perm];
        } else {
            const double sseA = yyv - dq_form<RP>(b, p.G + (size_t)perm * rA * rA, 0, rA);
            ta = t_from(dpick<RP>(b, p.rowA), sseA, p.dof, p.d[(size_t)perm * rA + p.rowA]);
        }
        const double sseB = yyv - dq_form<RP>(b, p.GB + (size_t)perm * rB * rB, rA, rB);
        const double tb = t_from(dpick<RP>(b, rA + p.rowB), sseB, p.dofB, p.dB[(size_t)perm * rB + p.rowB]);
        const double ta2 = __dmul_rn(ta, ta), tb2 = __dmul_rn(tb, tb);
        double s = __dadd_rn(__ddiv_rn(1.0, tb2), __ddiv_rn(1.0, ta2));
        const double cross = __ddiv_rn(1.0, __dmul_rn(ta2, tb2));
        if (p.alg == 0) s = __dadd_rn(s, cross);
        else if (p.alg == 2) s = __dsub_rn(s, cross);
        double z = __ddiv_rn(1.0, __dsqrt_rn(s));
        if (!inside) z = 0.0;
        const size_t off = (size_t)perm * p.ldt + v;
        if (p.t32) p.t32[off] = __double2float_rn(z);
        if (p.t64) p.t64[off] = z;
    }
}

template <int RP, typename YT>
__global__ void __launch_bounds__(256, 2) glm_dmma_multi_kernel(GlmParams p, int mtiles, const __grid_constant__ CUtensorMap tmA,
                                                                const __grid_constant__ CUtensorMap tmY) {
    constexpr int MT = DmmaShape<RP>::MT, NT = DmmaShape<RP>::NT, WM = DmmaShape<RP>::WM;
    constexpr int DM = DmmaShape<RP>::ROWS, DPA = DmmaShape<RP>::PITCH;   // shadow the r == 1 kernel's constants
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *sA = reinterpret_cast<double *>(smem_raw);                                  // [DSTAGES][DK][DPA]
    YT *sY = reinterpret_cast<YT *>(smem_raw + sizeof(double) * DSTAGES * DK * DPA);    // [DSTAGES][DK][DPY]
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + sizeof(double) * DSTAGES * DK * DPA +
                                                   sizeof(YT) * DSTAGES * DK * DPY);
    uint64_t *empty = full + DSTAGES;
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int tile = blockIdx.x;
    const int mt = tile % mtiles;
    const int64_t vt = tile / mtiles;
    const int m0 = mt * DM;
    const int64_t v0 = vt * DN;
    const int nchunks = (p.n + DK - 1) / DK;
    if (tid == 0) {
        for (int s = 0; s < DSTAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    auto issue_chunk = [&](int kc) {
        const int s = kc % DSTAGES;
        const int round = kc / DSTAGES;
        mbar_wait(empty + s, (round & 1) ^ 1);
        const int k0 = kc * DK;
        const int rows = min(DK, p.n - k0);
        if (p.use_tma) {
            if (warp == 0) {
                mbar_expect_tx(full + s, (uint32_t)DK * (DPA * 8 + DPY * (uint32_t)sizeof(YT)));
                tma_load_2d(sA + (size_t)s * DK * DPA, &tmA, m0, k0, full + s);
                tma_load_2d(sY + (size_t)s * DK * DPY, &tmY, (int)v0, k0, full + s);
            }
            return;
        }
        if (warp == 0) mbar_expect_tx(full + s, (uint32_t)rows * (DM * 8 + DN * (uint32_t)sizeof(YT)));
        for (int kk = warp; kk < rows; kk += 8) {
            bulk_g2s(sA + ((size_t)s * DK + kk) * DPA, p.At + (size_t)(k0 + kk) * p.ldA + m0, DM * 8, full + s);
            bulk_g2s(sY + ((size_t)s * DK + kk) * DPY, reinterpret_cast<const YT *>(p.Y) + (size_t)(k0 + kk) * p.ldy + v0,
                     DN * (uint32_t)sizeof(YT), full + s);
        }
    };
    const bool issuer = lane == 0 && (warp == 0 || !p.use_tma);   // tensor-map loads: warp 0 alone feeds the ring
    if (issuer)
        for (int kc = 0; kc < DSTAGES - 1 && kc < nchunks; ++kc) issue_chunk(kc);

    const int wm = WM == 1 ? 0 : (warp >> 2), wn = WM == 1 ? warp : (warp & 3);
    const int g = lane >> 2, t4 = lane & 3;
    double acc[MT][NT][2];
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    for (int kc = 0; kc < nchunks; ++kc) {
        const int s = kc % DSTAGES;
        const int round = kc / DSTAGES;
        if (issuer && kc + DSTAGES - 1 < nchunks) issue_chunk(kc + DSTAGES - 1);
        mbar_wait(full + s, round & 1);
        const int rows = min(DK, p.n - kc * DK);
        const double *a_st = sA + (size_t)s * DK * DPA + wm * (MT * 8) + g;
        const YT *y_st = sY + (size_t)s * DK * DPY + wn * (NT * 8) + g;
        for (int k4 = 0; k4 < rows; k4 += 4) {
            const int kr = k4 + t4;
            const bool ok = kr < rows;
            double a[MT], b[NT];
#pragma unroll
            for (int i = 0; i < MT; ++i) a[i] = ok ? a_st[(size_t)kr * DPA + i * 8] : 0.0;
#pragma unroll
            for (int j = 0; j < NT; ++j) b[j] = ok ? (double)y_st[(size_t)kr * DPY + j * 8] : 0.0;
#pragma unroll
            for (int i = 0; i < MT; ++i)
#pragma unroll
                for (int j = 0; j < NT; ++j) dmma_884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + s);
    }
    // epilogue: row tile i of the warp is regressor i % RP of the designs ((m0 / 8 + wm * MT + i) / RP) * 8 + 0..7
#pragma unroll
    for (int gi = 0; gi < MT / RP; ++gi) {
        const int perm = ((m0 / 8 + wm * MT) / RP + gi) * 8 + g;
        if (perm >= p.P) continue;
        if (RP == 1 && p.mode == 5) { // Sobel from cross-products: the design's five constants once, then its 8 values
            const double *C = p.GB + (size_t)perm * 8;
            const double C00 = __ldg(C), C01 = __ldg(C + 1), C10 = __ldg(C + 2), C11 = __ldg(C + 3), dsB = __ldg(C + 4);
#pragma unroll
            for (int j = 0; j < NT; ++j) {
                const int64_t v = v0 + wn * (NT * 8) + j * 8 + t4 * 2;
                if (v >= p.ldt) continue;
                // v is even and ldt a multiple of 4: the pair (v, v + 1) is inside the padded row
                const double2 cd = (v + 1 < p.V) ? __ldg(reinterpret_cast<const double2 *>(p.cfix + v))
                                                 : make_double2(v < p.V ? __ldg(p.cfix + v) : 0.0, 0.0);
                const double2 yv = (v + 1 < p.V) ? __ldg(reinterpret_cast<const double2 *>(p.yy + v))
                                                 : make_double2(v < p.V ? __ldg(p.yy + v) : 0.0, 0.0);
                sobel_cross_value(p, perm, acc[gi][j][0], cd.x, yv.x, v, C00, C01, C10, C11, dsB);
                sobel_cross_value(p, perm, acc[gi][j][1], cd.y, yv.y, v + 1, C00, C01, C10, C11, dsB);
            }
            continue;
        }
#pragma unroll
        for (int j = 0; j < NT; ++j) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int64_t v = v0 + wn * (NT * 8) + j * 8 + t4 * 2 + e;
                double b[RP];
#pragma unroll
                for (int i = 0; i < RP; ++i) b[i] = acc[gi * RP + i][j][e];
                const double yyv = (v < p.V) ? p.yy[v] : 0.0;
                if (v < p.ldt) dmma_vertex_stats<RP>(p, perm, b, yyv, v);
            }
        }
    }
}

template <int RP, typename YT>
static int launch_dmma_multi(const GlmParams &p, cudaStream_t stream) {
    constexpr int DM = DmmaShape<RP>::ROWS, DPA = DmmaShape<RP>::PITCH;
    const int64_t rows = ((int64_t)p.P + 7) / 8 * 8 * RP;
    const int mtiles = (int)((rows + DM - 1) / DM);
    const int64_t vtiles = (p.V + DN - 1) / DN;
    const int64_t tiles = (int64_t)mtiles * vtiles;
    TMB_REQUIRE(tiles < (int64_t)INT32_MAX, "glm: too many tiles");
    TMB_REQUIRE(p.ldA >= (int64_t)mtiles * DM, "glm: ldA must cover %lld columns", (long long)mtiles * DM);
    const size_t smem = sizeof(double) * DSTAGES * DK * DPA + sizeof(YT) * DSTAGES * DK * DPY + sizeof(uint64_t) * 2 * DSTAGES;
    TMB_CUDA(cudaFuncSetAttribute(glm_dmma_multi_kernel<RP, YT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GlmParams q = p;
    CUtensorMap tmA, tmY;
    memset(&tmA, 0, sizeof(tmA)); memset(&tmY, 0, sizeof(tmY));
    q.use_tma = make_tmap(&tmA, p.At, true, p.ldA, p.n, DPA, DK) && make_tmap(&tmY, p.Y, sizeof(YT) == 8, p.ldy, p.n, DPY, DK);
    glm_dmma_multi_kernel<RP, YT><<<(unsigned)tiles, 256, smem, stream>>>(q, mtiles, tmA, tmY);
    count_launch();
    TMB_CUDA(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------- stacked pseudo-inverses of row-permuted designs
// Permuting whole rows of a design permutes the columns of its pseudo-inverse (X'X is invariant), so the left operand
// of the batched fit is a gather: At[k, p*rp + i] = pinv[i, perm_idx[p, k]].  Done here instead of on the host: the
// host then ships only the index rows (n * 4 bytes per shuffle) and keeps ~3 ms per 512 shuffles off its critical path.
__global__ void glm_pack_rowperm_kernel(const double *__restrict__ pinv, int r, int n, const int32_t *__restrict__ idx,
                                        int P, int rp, double *__restrict__ At, int64_t ldA, int layout) {
    const int64_t col = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int k = blockIdx.y;
    if (col >= ldA) return;
    int p, i;
    if (layout == 0) {
        p = (int)(col / rp); i = (int)(col - (int64_t)p * rp);
    } else { // tile8: blocks of 8 designs x rp regressors, regressor-major inside a block
        const int64_t blk = col / (8 * rp);
        const int rem = (int)(col - blk * 8 * rp);
        i = rem >> 3; p = (int)(blk * 8 + (rem & 7));
    }
    double v = 0.0;
    if (p < P && i < r) v = pinv[(size_t)i * n + idx[(size_t)p * n + k]];
    At[(size_t)k * ldA + col] = v;
}

// ---------------------------------------------------------------- per-vertex sum of squares
template <typename YT>
__global__ void glm_sumsq_kernel(const YT *__restrict__ Y, int n, int64_t V, int64_t ldy, int center,
                                 double *__restrict__ yy, double *__restrict__ colsum) {
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V) return;
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += (double)Y[(size_t)i * ldy + v];
    const double mean = center ? s / n : 0.0;
    double q = 0.0;
    for (int i = 0; i < n; ++i) {
        const double c = (double)Y[(size_t)i * ldy + v] - mean;
        q = __fma_rn(c, c, q);
    }
    if (yy) yy[v] = q;
    if (colsum) colsum[v] = s;
}

// ---------------------------------------------------------------- direct per-vertex fit
// One thread per vertex, subjects streamed coalesced across the warp: beta = pinv . y in registers,
// then the explicit residual pass r_i = y_i - x_i . beta exactly as cynumstats.pyx:61 does it.
// Serves the API-parity entry points (tval_int / calc_beta_se / resid_covars / calcF /
// cy_lin_lstsqr_mat_residual on single designs); the permutation loop uses glm_tile_kernel.
static constexpr int kMaxK = 16;
struct DirectParams {
    const void *Y; int y_is_f64; int n; int64_t V; int64_t ldy;
    const double *X; const double *pinv; int k;
    const double *d; double dof; double grand_mean;
    double *beta64; double *t64; float *se32; int64_t ldt;
    double *r64; float *r32; int64_t ldr;
    double *sse; double *tss;
};

template <typename YT>
__global__ void glm_direct_kernel(DirectParams p) {
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= p.V) return;
    const YT *__restrict__ Y = reinterpret_cast<const YT *>(p.Y);
    const int n = p.n, k = p.k;
    double beta[kMaxK];
#pragma unroll
    for (int j = 0; j < kMaxK; ++j) beta[j] = 0.0;
    for (int i = 0; i < n; ++i) {
        const double y = (double)Y[(size_t)i * p.ldy + v];
#pragma unroll
        for (int j = 0; j < kMaxK; ++j)
            if (j < k) beta[j] = __fma_rn(p.pinv[(size_t)j * n + i], y, beta[j]);
    }
    // Sums follow the reference's evaluation order: Python/numpy add the rows one after another
    // (cynumstats.pyx:34-35,61).  For float32 data the reference's TSS is float32 arithmetic
    // ((y - np.mean(y))**2 stays float32), reproduced here with explicit fp32 ops.
    double sse = 0.0, tss = 0.0;
    float tss32 = 0.f;
    const float gm32 = (float)p.grand_mean;
    for (int i = 0; i < n; ++i) {
        double fit = 0.0;
#pragma unroll
        for (int j = 0; j < kMaxK; ++j)
            if (j < k) fit = __fma_rn(p.X[(size_t)i * k + j], beta[j], fit);
        const double y = (double)Y[(size_t)i * p.ldy + v];
        const double res = y - fit;
        if (p.r64) p.r64[(size_t)i * p.ldr + v] = res;
        if (p.r32) p.r32[(size_t)i * p.ldr + v] = (float)res;
        sse = __dadd_rn(sse, __dmul_rn(res, res));
        if (sizeof(YT) == 4) {
            const float c = __fsub_rn((float)y, gm32);
            tss32 = __fadd_rn(tss32, __fmul_rn(c, c));
        } else {
            const double c = y - p.grand_mean;
            tss = __dadd_rn(tss, __dmul_rn(c, c));
        }
    }
    if (p.sse) p.sse[v] = sse;
    if (p.tss) p.tss[v] = (sizeof(YT) == 4) ? (double)tss32 : tss;
    const double sigma2 = __ddiv_rn(sse, p.dof);
#pragma unroll
    for (int j = 0; j < kMaxK; ++j) {
        if (j < k) {
            if (p.beta64) p.beta64[(size_t)j * p.ldt + v] = beta[j];
            if (p.d) {
                const float se = __double2float_rn(__dsqrt_rn(__dmul_rn(sigma2, p.d[j])));
                if (p.se32) p.se32[(size_t)j * p.ldt + v] = se;
                if (p.t64) p.t64[(size_t)j * p.ldt + v] = __ddiv_rn(beta[j], (double)se);
            }
        }
    }
}

// se[j, v] = fl32(sqrt(sigma2[v] * d[j]))   (cynumstats.pyx:47-52 se_of_slope, without the Python loop)
__global__ void se_of_slope_kernel(const double *__restrict__ sigma2, int64_t V, const double *__restrict__ d, int k,
                                   float *__restrict__ se, int64_t ld) {
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V) return;
    const double s = sigma2[v];
    for (int j = 0; j < k; ++j) se[(size_t)j * ld + v] = __double2float_rn(__dsqrt_rn(__dmul_rn(s, d[j])));
}

template <typename YT>
static size_t glm_smem_bytes() {
    return sizeof(double) * STAGES * BK * BM + sizeof(YT) * STAGES * BK * BN + sizeof(uint64_t) * 2 * STAGES;
}

template <int RP, typename YT>
static int launch_tile(const GlmParams &p, cudaStream_t stream) {
    const int mtiles = (p.P * RP + BM - 1) / BM;
    const int64_t vtiles = (p.V + BN - 1) / BN;
    const int64_t tiles = (int64_t)mtiles * vtiles;
    TMB_REQUIRE(tiles < (int64_t)INT32_MAX, "glm: too many tiles");
    TMB_CUDA(cudaFuncSetAttribute(glm_tile_kernel<RP, YT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)glm_smem_bytes<YT>()));
    glm_tile_kernel<RP, YT><<<(unsigned)tiles, kGlmThreads, glm_smem_bytes<YT>(), stream>>>(p, mtiles);
    count_launch();
    TMB_CUDA(cudaGetLastError());
    return 0;
}

template <typename YT>
static int launch_glm_t(const GlmParams &p, cudaStream_t stream) {
    switch (p.rp) {
    case 1: return launch_tile<1, YT>(p, stream);
    case 2: return launch_tile<2, YT>(p, stream);
    case 4: return launch_tile<4, YT>(p, stream);
    case 8: return launch_tile<8, YT>(p, stream);
    default: set_error("glm: rp must be 1, 2, 4 or 8 (got %d)", p.rp); return 1;
    }
}


// ---------------------------------------------------------------- any number of regressors: statistics from stored betas
// Designs with more than 8 non-intercept regressors (dummy-coded sites plus covariates; the reference accepts any k,
// cynumstats.pyx:28-29,59-64) do not fit the register-resident epilogues above.  Their betas are written to HBM by the
// plain contraction (tmb_glm_beta: every pseudo-inverse row is its own output row) and this kernel evaluates the same
// statistics per (design, vertex) with run-time loops.  Slower (the betas make one round trip, the quadratic form is
// O(r^2) per value) but exact and without a limit below kMaxGenericR.
static constexpr int kMaxGenericR = 64;

__device__ __forceinline__ double rq_form(const double *b, const double *__restrict__ G, int lo, int r) {
    double q = 0.0;
    for (int a = 0; a < r; ++a) {
        double inner = 0.0;
        for (int c = 0; c < r; ++c) inner = __fma_rn(__ldg(G + a * r + c), b[lo + c], inner);
        q = __fma_rn(b[lo + a], inner, q);
    }
    return q;
}

__global__ void __launch_bounds__(128) glm_stats_from_beta_kernel(GlmParams p, const double *__restrict__ beta, int64_t ldb) {
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int perm = blockIdx.y;
    if (v >= p.ldt) return;
    const bool inside = v < p.V;
    const int rt = p.mode == 2 ? p.rA + p.rB : (p.mode == 6 || p.mode == 7) ? p.cross_m : p.r;
    double b[kMaxGenericR];
    for (int i = 0; i < rt; ++i) b[i] = inside ? beta[((size_t)perm * rt + i) * ldb + v] : 0.0;
    const double yyv = inside ? p.yy[v] : 0.0;
    if (p.mode == 0) {
        const int r = p.r;
        const double sse = yyv - rq_form(b, p.G + (size_t)perm * r * r, 0, r);
        for (int a = p.row0; a < p.row0 + p.nrows; ++a) {
            double t = t_from(b[a], sse, p.dof, p.d[(size_t)perm * r + a]);
            if (p.nan_to_zero && t != t) t = 0.0;
            if (!inside) t = 0.0;
            const size_t off = ((size_t)perm * p.nrows + (a - p.row0)) * p.ldt + v;
            if (p.t32) p.t32[off] = __double2float_rn(t);
            if (p.t64) p.t64[off] = t;
        }
    } else if (p.mode == 3) {
        const int r = p.r;
        const double ssb = rq_form(b, p.G + (size_t)perm * r * r, 0, r);
        const double ms = __ddiv_rn(yyv - ssb, p.dof);
        const int first = p.row0 == 0 ? 1 : 0;
        if (first) {
            const double between = (p.sstot && inside) ? __dsub_rn(p.sstot[v], yyv - ssb) : ssb;
            double f = __ddiv_rn(__ddiv_rn(between, (double)r), ms);
            if (p.nan_to_zero && f != f) f = 0.0;
            if (!inside) f = 0.0;
            const size_t off = (size_t)perm * p.nrows * p.ldt + v;
            if (p.t32) p.t32[off] = __double2float_rn(f);
            if (p.t64) p.t64[off] = f;
        }
        const double *M = p.M + (size_t)perm * p.msz;
        for (int i = 0; i < p.nvar; ++i) {
            const int lo = p.var_lo[i], ki = p.var_k[i];
            double f = __ddiv_rn(rq_form(b, M, lo, ki), __dmul_rn(ms, (double)ki));
            if (p.nan_to_zero && f != f) f = 0.0;
            if (!inside) f = 0.0;
            const size_t off = ((size_t)perm * p.nrows + first + i) * p.ldt + v;
            if (p.t32) p.t32[off] = __double2float_rn(f);
            if (p.t64) p.t64[off] = f;
            M += ki * ki;
        }
    } else if (p.mode == 7) {
        // t statistics of a design of which only some columns are permuted (the drivers' -v mode,
        // vertex_tfce_multiple_regression_randomise.py:84-97): centred cross-products of the permuted columns from the
        // contraction (b[0 .. m)), of the fixed columns from p.cfix (fitted once); colmap: source row of each of the r
        // regressors in design order; p.G: the shuffle's INVERSE centred Gram matrix.  beta = C c, SSE = yy - c'beta.
        const int m = p.cross_m, r = p.r;
        const double *C = p.G + (size_t)perm * r * r;
        double c[16], bt[16], q = 0.0;
        for (int i = 0; i < r; ++i) {
            const int src = __ldg(p.colmap + i);
            c[i] = src < m ? b[src] : (inside ? __ldg(p.cfix + (size_t)(src - m) * p.ldf + v) : 0.0);
        }
        for (int a = 0; a < r; ++a) {
            double bi = 0.0;
            for (int j = 0; j < r; ++j) bi = __fma_rn(__ldg(C + a * r + j), c[j], bi);
            bt[a] = bi;
            q = __fma_rn(c[a], bi, q);
        }
        const double sse = yyv - q;
        for (int a = p.row0; a < p.row0 + p.nrows; ++a) {
            double t = t_from(bt[a], sse, p.dof, __ldg(C + a * r + a));
            if (p.nan_to_zero && t != t) t = 0.0;
            if (!inside) t = 0.0;
            const size_t off = ((size_t)perm * p.nrows + (a - p.row0)) * p.ldt + v;
            if (p.t32) p.t32[off] = __double2float_rn(t);
            if (p.t64) p.t64[off] = t;
        }
    } else if (p.mode == 6) {
        // Sobel z from centred CROSS-PRODUCTS c = Z'y instead of betas, for designs of which only some columns are
        // permuted (tm-models mediation, tm_models_randomise.py:430-520: the left variable; covariates and the right
        // variable stay): the permuted columns' cross-products come from the contraction (b[0 .. m)), the fixed columns'
        // from p.cfix (f rows, fitted once).  colmap lists, for path A's rA and then path B's rB regressors in design
        // order, the source row (< m: permuted, else fixed row - m).  p.G / p.GB hold the INVERSE centred Gram matrices of
        // the shuffle's two designs: beta = C c, SSE = yy - c'beta.
        const int m = p.cross_m, rA = p.rA, rB = p.rB;
        auto source = [&](int idx) { return idx < m ? b[idx] : (inside ? __ldg(p.cfix + (size_t)(idx - m) * p.ldf + v) : 0.0); };
        auto path_t = [&](const double *C, const int32_t *map, int r, int row, double dof) {
            double c[16], brow = 0.0, q = 0.0;
            for (int i = 0; i < r; ++i) c[i] = source(__ldg(map + i));
            for (int a = 0; a < r; ++a) {
                double bi = 0.0;
                for (int j = 0; j < r; ++j) bi = __fma_rn(__ldg(C + a * r + j), c[j], bi);
                q = __fma_rn(c[a], bi, q);
                if (a == row) brow = bi;
            }
            return t_from(brow, yyv - q, dof, __ldg(C + row * r + row));
        };
        const double ta = p.ta_scalar ? p.ta_scalar[perm] : path_t(p.G + (size_t)perm * rA * rA, p.colmap, rA, p.rowA, p.dof);
        const double tb = path_t(p.GB + (size_t)perm * rB * rB, p.colmap + rA, rB, p.rowB, p.dofB);
        const double ta2 = __dmul_rn(ta, ta), tb2 = __dmul_rn(tb, tb);
        double s = __dadd_rn(__ddiv_rn(1.0, tb2), __ddiv_rn(1.0, ta2));
        const double cross = __ddiv_rn(1.0, __dmul_rn(ta2, tb2));
        if (p.alg == 0) s = __dadd_rn(s, cross);
        else if (p.alg == 2) s = __dsub_rn(s, cross);
        double z = __ddiv_rn(1.0, __dsqrt_rn(s));
        if (!inside) z = 0.0;
        const size_t off = (size_t)perm * p.ldt + v;
        if (p.t32) p.t32[off] = __double2float_rn(z);
        if (p.t64) p.t64[off] = z;
    } else if (p.mode == 4) {
        // Cosinor statistics of pyfunc.py:2406-2563 glm_cosinor (permutation branch): regressors 2i, 2i+1 are the
        // cos / sin pair of period i, then nexog tested columns, then covariates.  C = inv(G) (= the slope block of
        // inv(X'X), intercept included).  Output rows: [model F, (|t amplitude|, |t acrophase|) per period, t of every
        // tested column] -- or, for the cosinor mediation (tm_models_randomise.py:383-412), the single row
        // calc_indirect(ta, t of tested column 0) with the un-permuted path-A amplitude t handed in by the host.
        const int r = p.r, nper = p.nvar, nexog = p.cos_nexog;
        const double *C = p.M + (size_t)perm * r * r;
        const double ssb = rq_form(b, p.G + (size_t)perm * r * r, 0, r);
        double sse = yyv - ssb;
        if (sse < 0.0) sse = 0.0;
        const double ms = __ddiv_rn(sse, p.dof);
        const double sigma = __dsqrt_rn(ms);
        const size_t base = (size_t)perm * p.nrows * p.ldt + v;
        auto put = [&](int row, double x) {
            if (p.nan_to_zero && x != x) x = 0.0;
            if (!inside) x = 0.0;
            if (p.t32) p.t32[base + (size_t)row * p.ldt] = __double2float_rn(x);
            if (p.t64) p.t64[base + (size_t)row * p.ldt] = x;
        };
        const double s2 = __dmul_rn(sigma, sigma);                          // the reference re-squares sigma (:2507)
        auto t_exog = [&](int j) {
            const int col = 2 * nper + j;
            const float se = __double2float_rn(__dsqrt_rn(__dmul_rn(s2, __ldg(C + col * r + col))));
            return __ddiv_rn(b[col], (double)se);
        };
        if (p.cos_mediation) {                                               // mediation: one row
            const double ta = p.cos_ta, tb = t_exog(0);
            const double ta2 = __dmul_rn(ta, ta), tb2 = __dmul_rn(tb, tb);
            double s = __dadd_rn(__ddiv_rn(1.0, tb2), __ddiv_rn(1.0, ta2));
            const double cross = __ddiv_rn(1.0, __dmul_rn(ta2, tb2));
            if (p.alg == 0) s = __dadd_rn(s, cross);
            else if (p.alg == 2) s = __dsub_rn(s, cross);
            put(0, __ddiv_rn(1.0, __dsqrt_rn(s)));
        } else {
            // Fmodel = ((SS_Total - SS_Residuals) / DF_Between) / MS_Residuals (:2492-2495)
            const double between = p.sstot ? __dsub_rn(inside ? p.sstot[v] : 0.0, sse) : ssb;
            put(0, __ddiv_rn(__ddiv_rn(between, (double)r), ms));
            for (int i = 0; i < nper; ++i) {
                const double be = b[2 * i], ga = b[2 * i + 1];
                const double amp = __dsqrt_rn(__dadd_rn(__dmul_rn(be, be), __dmul_rn(ga, ga)));
                const double acr = atan(fabs(__ddiv_rn(-ga, be)));
                double sn, cs;
                sincos(acr, &sn, &cs);
                const double c11 = __ldg(C + (2 * i) * r + 2 * i), c12 = __ldg(C + (2 * i) * r + 2 * i + 1),
                             c22 = __ldg(C + (2 * i + 1) * r + 2 * i + 1);
                const double sn2 = __dmul_rn(sn, sn), cs2 = __dmul_rn(cs, cs);
                const double mix = __dmul_rn(__dmul_rn(__dmul_rn(2.0, c12), sn), cs);
                const double va = __dadd_rn(__dadd_rn(__dmul_rn(c11, sn2), mix), __dmul_rn(c22, cs2));   // acrophase
                const double vm = __dadd_rn(__dsub_rn(__dmul_rn(c11, cs2), mix), __dmul_rn(c22, sn2));   // amplitude
                const double se_acr = __ddiv_rn(__dmul_rn(sigma, __dsqrt_rn(va)), amp);
                const double se_amp = __dmul_rn(sigma, __dsqrt_rn(vm));
                put(1 + 2 * i, fabs(__ddiv_rn(amp, se_amp)));
                put(2 + 2 * i, fabs(__ddiv_rn(1.0, se_acr)));
            }
            for (int j = 0; j < nexog; ++j) put(1 + 2 * nper + j, t_exog(j));
        }
    } else {
        const int rA = p.rA, rB = p.rB;
        double ta;
        if (p.ta_scalar) {
            ta = p.ta_scalar[perm];
        } else {
            const double sseA = yyv - rq_form(b, p.G + (size_t)perm * rA * rA, 0, rA);
            ta = t_from(b[p.rowA], sseA, p.dof, p.d[(size_t)perm * rA + p.rowA]);
        }
        const double sseB = yyv - rq_form(b, p.GB + (size_t)perm * rB * rB, rA, rB);
        const double tb = t_from(b[rA + p.rowB], sseB, p.dofB, p.dB[(size_t)perm * rB + p.rowB]);
        const double ta2 = __dmul_rn(ta, ta), tb2 = __dmul_rn(tb, tb);
        double s = __dadd_rn(__ddiv_rn(1.0, tb2), __ddiv_rn(1.0, ta2));
        const double cross = __ddiv_rn(1.0, __dmul_rn(ta2, tb2));
        if (p.alg == 0) s = __dadd_rn(s, cross);
        else if (p.alg == 2) s = __dsub_rn(s, cross);
        double z = __ddiv_rn(1.0, __dsqrt_rn(s));
        if (!inside) z = 0.0;
        const size_t off = (size_t)perm * p.ldt + v;
        if (p.t32) p.t32[off] = __double2float_rn(z);
        if (p.t64) p.t64[off] = z;
    }
}

static int launch_stats_from_beta(const GlmParams &p, const double *beta, int64_t ldb, cudaStream_t stream) {
    const int rt = p.mode == 2 ? p.rA + p.rB : (p.mode == 6 || p.mode == 7) ? p.cross_m : p.r;
    TMB_REQUIRE(rt >= 1 && rt <= kMaxGenericR, "glm: at most %d non-intercept regressors per design (got %d)", kMaxGenericR, rt);
    TMB_REQUIRE(p.P >= 1 && p.P <= 65535, "glm (stored betas): 1..65535 designs per call (got %d)", p.P);
    const dim3 grid((unsigned)((p.ldt + 127) / 128), (unsigned)p.P);
    glm_stats_from_beta_kernel<<<grid, 128, 0, stream>>>(p, beta, ldb);
    count_launch();
    TMB_CUDA(cudaGetLastError());
    return 0;
}

// which column order of At the kernels behind launch_glm expect for (data type, padded regressors per design)
static bool glm_uses_dmma(int y_is_f64, int rp) {
    const char *force = getenv("TMB_GLM");
    const bool dfma = force && strcmp(force, "dfma") == 0;
    return !dfma && !y_is_f64 && rp >= 1 && rp <= 8;
}

// regressors per design after padding: the tensor-core kernels take any 1..8, the vector kernel 1, 2, 4 or 8
static int glm_padded_regressors(int y_is_f64, int r) {
    if (r < 1 || r > 8) return 0;
    if (glm_uses_dmma(y_is_f64, r)) return r;
    return r <= 1 ? 1 : r <= 2 ? 2 : r <= 4 ? 4 : 8;
}

// columns of At the kernels will read for P designs (a multiple of 128)
static int64_t glm_packed_columns(int y_is_f64, int P, int rp) {
    int64_t cols;
    if (glm_uses_dmma(y_is_f64, rp)) {
        const int dm = rp <= 2 ? 64 : rp <= 4 ? (rp == 3 ? 48 : 64) : 8 * rp;
        const int64_t rows = ((int64_t)P + 7) / 8 * 8 * rp;
        cols = (rows + dm - 1) / dm * dm;
    } else {
        cols = (int64_t)P * rp;
    }
    return (cols + BM - 1) / BM * BM;
}

int launch_glm(const GlmParams &p, cudaStream_t stream) {
    TMB_REQUIRE(p.ldy % BN == 0 && p.ldy >= (p.V + BN - 1) / BN * BN,
                "glm: ldy must be a multiple of %d covering V rounded up", BN);
    TMB_REQUIRE(p.ldA % BM == 0, "glm: ldA must be a multiple of %d", BM);
    TMB_REQUIRE(p.mode == 1 || p.ldA >= glm_packed_columns(p.y_is_f64, p.P, p.rp),
                "glm: ldA must be at least tmb_glm_packed_columns()");
    TMB_REQUIRE(p.ldt % 4 == 0 && p.ldt >= (p.V + 3) / 4 * 4, "glm: ldt must be a multiple of 4 covering V");
    TMB_REQUIRE((reinterpret_cast<uintptr_t>(p.Y) & 15) == 0 && (reinterpret_cast<uintptr_t>(p.At) & 15) == 0,
                "glm: Y and At must be 16-byte aligned");
    TMB_REQUIRE(!p.t32 || (reinterpret_cast<uintptr_t>(p.t32) & 15) == 0, "glm: t32 must be 16-byte aligned");
    {
        // float32 data: fp64 tensor cores (TMB_GLM=dfma forces the vector kernel).  One slope per design and the
        // t-statistic (the headline case) has its own kernel with the cheap fp32-seeded epilogue; everything else --
        // several regressors, partial F, Sobel z -- runs on glm_dmma_multi_kernel.
        // (float64 data: the doubled Y stage halves the resident CTAs and DFMA wins, measured in round 1.)
        if (p.mode == 1 && glm_uses_dmma(p.y_is_f64, 1) && p.ldA >= (p.P + DM - 1) / DM * DM) {
            // stored betas (designs with more than 8 regressors, cosinor, repeated-measures cross-products): the same
            // tensor-core contraction with a store-only epilogue
            return launch_dmma<float>(p, stream);
        }
        if (glm_uses_dmma(p.y_is_f64, p.rp) && p.mode != 1) {
            TMB_REQUIRE(p.layout == 1 || p.rp == 1, "glm: the DMMA kernels need the tile8 column order of At (tmb_glm_layout)");
            TMB_REQUIRE(p.ldA >= glm_packed_columns(p.y_is_f64, p.P, p.rp), "glm: ldA must be at least tmb_glm_packed_columns() = %lld",
                        (long long)glm_packed_columns(p.y_is_f64, p.P, p.rp));
            if (p.mode == 5) {
                const char *ep = getenv("TMB_GLM_EPILOGUE");
                GlmParams q = p;
                q.exact_epilogue = ep && strcmp(ep, "exact") == 0 ? 1 : 0;
                return launch_dmma_multi<1, float>(q, stream);
            }
            if (p.mode == 0 && p.r == 1 && p.rp == 1 && p.row0 == 0 && p.nrows == 1) {
                const char *ep = getenv("TMB_GLM_EPILOGUE");
                GlmParams q = p;
                q.exact_epilogue = !ep ? 0 : strcmp(ep, "exact") == 0 ? 1 : strcmp(ep, "reforder") == 0 ? 2 : 0; // reforder: (sse / dof) * d
                return launch_dmma<float>(q, stream);
            }
            switch (p.rp) {
            case 1: return launch_dmma_multi<1, float>(p, stream);
            case 2: return launch_dmma_multi<2, float>(p, stream);
            case 3: return launch_dmma_multi<3, float>(p, stream);
            case 4: return launch_dmma_multi<4, float>(p, stream);
            case 5: return launch_dmma_multi<5, float>(p, stream);
            case 6: return launch_dmma_multi<6, float>(p, stream);
            case 7: return launch_dmma_multi<7, float>(p, stream);
            case 8: return launch_dmma_multi<8, float>(p, stream);
            default: break;
            }
        }
        TMB_REQUIRE(p.layout == 0 || p.rp == 1, "glm: the DFMA tile kernel needs the plain column order of At (tmb_glm_layout)");
    }
    return p.y_is_f64 ? launch_glm_t<double>(p, stream) : launch_glm_t<float>(p, stream);
}

} // namespace tmb

// --------------------------------------------------------------------------------- C ABI
#include "../../include/tfce_b200.h"
using namespace tmb;

static int dtype_is_f64(int ydtype, int *out) {
    TMB_REQUIRE(ydtype == TMB_F32 || ydtype == TMB_F64, "ydtype must be TMB_F32 (0) or TMB_F64 (1)");
    *out = ydtype == TMB_F64;
    return 0;
}

extern "C" int tmb_glm_sumsq(const void *Y_dev, int ydtype, int n, int64_t V, int64_t ldy, int center,
                             double *yy_dev, double *colsum_dev, void *stream) {
    TMB_REQUIRE(Y_dev && (yy_dev || colsum_dev) && n > 0 && V > 0, "tmb_glm_sumsq: bad arguments");
    TMB_DEVICE_OF(Y_dev, "tmb_glm_sumsq");
    int f64;
    if (dtype_is_f64(ydtype, &f64)) return 1;
    const int threads = 256;
    const unsigned grid = (unsigned)((V + threads - 1) / threads);
    if (f64)
        glm_sumsq_kernel<double><<<grid, threads, 0, (cudaStream_t)stream>>>((const double *)Y_dev, n, V, ldy, center,
                                                                          yy_dev, colsum_dev);
    else
        glm_sumsq_kernel<float><<<grid, threads, 0, (cudaStream_t)stream>>>((const float *)Y_dev, n, V, ldy, center,
                                                                         yy_dev, colsum_dev);
    count_launch();
    TMB_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int tmb_glm_tstat(const void *Y_dev, int ydtype, int n, int64_t V, int64_t ldy, const double *At_dev,
                             int64_t ldA, const double *G_dev, const double *d_dev, int P, int r, int rp, int row0,
                             int nrows, double dof, const double *yy_dev, float *t32_dev, double *t64_dev,
                             int64_t ldt, int nan_to_zero, int layout, void *stream) {
    TMB_REQUIRE(Y_dev && At_dev && G_dev && d_dev && yy_dev && (t32_dev || t64_dev), "tmb_glm_tstat: null pointer");
    TMB_REQUIRE(n > 0 && V > 0 && P > 0 && r >= 1 && r <= rp && row0 >= 0 && nrows >= 1 && row0 + nrows <= r,
                "tmb_glm_tstat: bad shape (n=%d V=%lld P=%d r=%d rp=%d row0=%d nrows=%d)", n, (long long)V, P, r, rp,
                row0, nrows);
    TMB_DEVICE_OF(Y_dev, "tmb_glm_tstat");
    GlmParams p{};
    if (dtype_is_f64(ydtype, &p.y_is_f64)) return 1;
    p.Y = Y_dev; p.n = n; p.V = V; p.ldy = ldy; p.At = At_dev; p.ldA = ldA; p.G = G_dev; p.d = d_dev;
    p.P = P; p.r = r; p.rp = rp; p.row0 = row0; p.nrows = nrows; p.dof = dof; p.yy = yy_dev;
    p.t32 = t32_dev; p.t64 = t64_dev; p.ldt = ldt; p.nan_to_zero = nan_to_zero; p.mode = 0; p.layout = layout;
    return launch_glm(p, (cudaStream_t)stream);
}

extern "C" int tmb_glm_fstat(const void *Y_dev, int ydtype, int n, int64_t V, int64_t ldy, const double *At_dev,
                             int64_t ldA, const double *G_dev, const double *M_dev, int P, int r, int rp, int nvar,
                             const int32_t *var_lo, const int32_t *var_k, int want_model, double dof,
                             const double *yy_dev, const double *sstotal_dev, float *f32_dev, double *f64_dev,
                             int64_t ldt, int nan_to_zero, int layout, void *stream) {
    TMB_REQUIRE(Y_dev && At_dev && G_dev && yy_dev && (f32_dev || f64_dev), "tmb_glm_fstat: null pointer");
    TMB_REQUIRE(n > 0 && V > 0 && P > 0 && r >= 1 && r <= rp && nvar >= 0 && nvar <= 8 && (nvar == 0 || (M_dev && var_lo && var_k)) &&
                    (nvar > 0 || want_model),
                "tmb_glm_fstat: bad shape (n=%d V=%lld P=%d r=%d rp=%d nvar=%d)", n, (long long)V, P, r, rp, nvar);
    TMB_DEVICE_OF(Y_dev, "tmb_glm_fstat");
    GlmParams p{};
    if (dtype_is_f64(ydtype, &p.y_is_f64)) return 1;
    p.Y = Y_dev; p.n = n; p.V = V; p.ldy = ldy; p.At = At_dev; p.ldA = ldA; p.G = G_dev; p.M = M_dev;
    p.P = P; p.r = r; p.rp = rp; p.row0 = want_model ? 0 : 1; p.nrows = nvar + (want_model ? 1 : 0); p.dof = dof; p.yy = yy_dev;
    p.t32 = f32_dev; p.t64 = f64_dev; p.ldt = ldt; p.nan_to_zero = nan_to_zero; p.mode = 3; p.nvar = nvar; p.layout = layout;
    p.sstot = sstotal_dev;
    for (int i = 0; i < nvar; ++i) {
        TMB_REQUIRE(var_lo[i] >= 0 && var_k[i] >= 1 && var_lo[i] + var_k[i] <= r,
                    "tmb_glm_fstat: variable %d covers rows [%d, %d) of %d", i, var_lo[i], var_lo[i] + var_k[i], r);
        p.var_lo[i] = var_lo[i]; p.var_k[i] = var_k[i]; p.msz += var_k[i] * var_k[i];
    }
    return launch_glm(p, (cudaStream_t)stream);
}

extern "C" int tmb_glm_layout(int ydtype, int rp) {
    return glm_uses_dmma(ydtype == TMB_F64, rp) ? 1 : 0;
}

extern "C" int tmb_glm_rp(int ydtype, int r) { return glm_padded_regressors(ydtype == TMB_F64, r); }

extern "C" int64_t tmb_glm_packed_columns(int ydtype, int P, int rp) {
    return glm_packed_columns(ydtype == TMB_F64, P, rp);
}

extern "C" int tmb_glm_pack_rowperm(const double *pinv_dev, int r, int n, const int32_t *perm_idx_dev, int P, int rp,
                                    double *At_dev, int64_t ldA, int layout, void *stream) {
    TMB_REQUIRE(pinv_dev && perm_idx_dev && At_dev, "tmb_glm_pack_rowperm: null pointer");
    TMB_REQUIRE(r >= 1 && r <= rp && rp <= 8 && n > 0 && P > 0 && n <= 65535 && (layout == 0 || layout == 1) &&
                    ldA >= (layout ? ((int64_t)P + 7) / 8 * 8 * rp : (int64_t)P * rp),
                "tmb_glm_pack_rowperm: bad shape (r=%d rp=%d n=%d P=%d ldA=%lld)", r, rp, n, P, (long long)ldA);
    TMB_DEVICE_OF(pinv_dev, "tmb_glm_pack_rowperm");
    const dim3 grid((unsigned)((ldA + 255) / 256), (unsigned)n);
    glm_pack_rowperm_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(pinv_dev, r, n, perm_idx_dev, P, rp, At_dev, ldA, layout);
    count_launch();
    TMB_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int tmb_glm_beta(const void *Y_dev, int ydtype, int n, int64_t V, int64_t ldy, const double *At_dev,
                            int64_t ldA, int nrows, double *beta64_dev, int64_t ldt, void *stream) {
    TMB_REQUIRE(Y_dev && At_dev && beta64_dev && n > 0 && V > 0 && nrows > 0, "tmb_glm_beta: bad arguments");
    TMB_DEVICE_OF(Y_dev, "tmb_glm_beta");
    GlmParams p{};
    if (dtype_is_f64(ydtype, &p.y_is_f64)) return 1;
    p.Y = Y_dev; p.n = n; p.V = V; p.ldy = ldy; p.At = At_dev; p.ldA = ldA; p.P = nrows; p.r = 1; p.rp = 1;
    p.t64 = beta64_dev; p.ldt = ldt; p.mode = 1;
    return launch_glm(p, (cudaStream_t)stream);
}

extern "C" int tmb_glm_direct(const void *Y_dev, int ydtype, int n, int64_t V, int64_t ldy, const double *X_dev,
                              const double *pinv_dev, int k, const double *d_dev, double dof, double grand_mean,
                              double *beta64_dev, double *t64_dev, float *se32_dev, int64_t ldt, double *resid64_dev,
                              float *resid32_dev, int64_t ldr, double *sse_dev, double *tss_dev, void *stream) {
    TMB_REQUIRE(Y_dev && X_dev && pinv_dev && n > 0 && V > 0, "tmb_glm_direct: bad arguments");
    TMB_REQUIRE(k >= 1 && k <= kMaxK, "tmb_glm_direct: k must be in 1..%d (got %d)", kMaxK, k);
    TMB_REQUIRE(!(t64_dev || se32_dev) || d_dev, "tmb_glm_direct: t/se requested without diag(inv(X'X))");
    TMB_DEVICE_OF(Y_dev, "tmb_glm_direct");
    DirectParams p{};
    if (dtype_is_f64(ydtype, &p.y_is_f64)) return 1;
    p.Y = Y_dev; p.n = n; p.V = V; p.ldy = ldy; p.X = X_dev; p.pinv = pinv_dev; p.k = k; p.d = d_dev; p.dof = dof;
    p.grand_mean = grand_mean; p.beta64 = beta64_dev; p.t64 = t64_dev; p.se32 = se32_dev; p.ldt = ldt;
    p.r64 = resid64_dev; p.r32 = resid32_dev; p.ldr = ldr; p.sse = sse_dev; p.tss = tss_dev;
    const int threads = 128;
    const unsigned grid = (unsigned)((V + threads - 1) / threads);
    if (p.y_is_f64) glm_direct_kernel<double><<<grid, threads, 0, (cudaStream_t)stream>>>(p);
    else glm_direct_kernel<float><<<grid, threads, 0, (cudaStream_t)stream>>>(p);
    count_launch();
    TMB_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int tmb_se_of_slope(const double *sigma2_dev, int64_t V, const double *d_dev, int k, float *se32_dev,
                               int64_t ld, void *stream) {
    TMB_REQUIRE(sigma2_dev && d_dev && se32_dev && V > 0 && k > 0, "tmb_se_of_slope: bad arguments");
    TMB_DEVICE_OF(sigma2_dev, "tmb_se_of_slope");
    const int threads = 256;
    se_of_slope_kernel<<<(unsigned)((V + threads - 1) / threads), threads, 0, (cudaStream_t)stream>>>(
        sigma2_dev, V, d_dev, k, se32_dev, ld);
    count_launch();
    TMB_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int tmb_sobelz(const void *Y_dev, int ydtype, int n, int64_t V, int64_t ldy, const double *At_dev,
                          int64_t ldA, int rp, const double *GA_dev, const double *dA_dev, int rA, int rowA,
                          double dofA, const double *GB_dev, const double *dB_dev, int rB, int rowB, double dofB,
                          const double *yy_dev, const double *ta_scalar_dev, int P, int alg, float *z32_dev,
                          double *z64_dev, int64_t ldt, int layout, void *stream) {
    TMB_REQUIRE(Y_dev && At_dev && GB_dev && dB_dev && yy_dev && (z32_dev || z64_dev), "tmb_sobelz: null pointer");
    TMB_REQUIRE(rA >= 0 && rB >= 1 && rA + rB <= rp && rowB >= 0 && rowB < rB &&
                    (ta_scalar_dev || (rA >= 1 && rowA >= 0 && rowA < rA && GA_dev && dA_dev)),
                "tmb_sobelz: bad shape");
    TMB_REQUIRE(alg >= 0 && alg <= 2, "tmb_sobelz: alg must be 0 (aroian), 1 (sobel) or 2 (goodman)");
    TMB_DEVICE_OF(Y_dev, "tmb_sobelz");
    GlmParams p{};
    if (dtype_is_f64(ydtype, &p.y_is_f64)) return 1;
    p.Y = Y_dev; p.n = n; p.V = V; p.ldy = ldy; p.At = At_dev; p.ldA = ldA; p.G = GA_dev; p.d = dA_dev;
    p.P = P; p.r = rA + rB; p.rp = rp; p.dof = dofA; p.yy = yy_dev; p.t32 = z32_dev; p.t64 = z64_dev; p.ldt = ldt;
    p.mode = 2; p.GB = GB_dev; p.dB = dB_dev; p.rA = rA; p.rB = rB; p.rowA = rowA; p.rowB = rowB; p.dofB = dofB;
    p.ta_scalar = ta_scalar_dev; p.alg = alg; p.layout = layout;
    return launch_glm(p, (cudaStream_t)stream);
}

extern "C" int tmb_sobelz_cross(const void *Y_dev, int ydtype, int n, int64_t V, int64_t ldy, const double *At_dev,
                                int64_t ldA, const double *cd_dev, double xx, double dofA, const double *CB_dev, int xpos,
                                int rowB, double dofB, const double *yy_dev, int P, int alg, float *z32_dev,
                                double *z64_dev, int64_t ldt, void *stream) {
    TMB_REQUIRE(Y_dev && At_dev && cd_dev && CB_dev && yy_dev && (z32_dev || z64_dev), "tmb_sobelz_cross: null pointer");
    TMB_REQUIRE(n > 0 && V > 0 && P > 0 && xx > 0.0 && (xpos == 0 || xpos == 1) && (rowB == 0 || rowB == 1),
                "tmb_sobelz_cross: bad shape (n=%d V=%lld P=%d x'x=%g)", n, (long long)V, P, xx);
    TMB_REQUIRE(alg >= 0 && alg <= 2, "tmb_sobelz_cross: alg must be 0 (aroian), 1 (sobel) or 2 (goodman)");
    TMB_REQUIRE(ydtype == TMB_F32, "tmb_sobelz_cross: float32 data only (tmb_sobelz serves float64 data)");
    TMB_REQUIRE(((reinterpret_cast<uintptr_t>(cd_dev) | reinterpret_cast<uintptr_t>(yy_dev)) & 15) == 0 && ldt % 4 == 0,
                "tmb_sobelz_cross: cd_dev and yy_dev must be 16-byte aligned, ldt a multiple of 4");
    TMB_DEVICE_OF(Y_dev, "tmb_sobelz_cross");
    GlmParams p{};
    if (dtype_is_f64(ydtype, &p.y_is_f64)) return 1;
    TMB_REQUIRE(glm_uses_dmma(p.y_is_f64, 1), "tmb_sobelz_cross: needs the tensor-core fit (TMB_GLM=dfma is set)");
    p.Y = Y_dev; p.n = n; p.V = V; p.ldy = ldy; p.At = At_dev; p.ldA = ldA; p.P = P; p.r = 1; p.rp = 1;
    p.yy = yy_dev; p.t32 = z32_dev; p.t64 = z64_dev; p.ldt = ldt; p.mode = 5; p.GB = CB_dev; p.rowB = rowB; p.dofB = dofB;
    p.cfix = cd_dev; p.xx = 1.0 / xx; p.dof = (1.0 / xx) / dofA; p.xpos = xpos; p.alg = alg; p.layout = 1;
    return launch_glm(p, (cudaStream_t)stream);
}

extern "C" int tmb_glm_tstat_cross_rows(const double *cperm_dev, int64_t ldb, int m, const double *cfix_dev, int64_t ldf, int f,
                                        int64_t V, const double *C_dev, int r, const int32_t *colmap_dev,
                                        const int32_t *colmap_host, int row0, int nrows, double dof, const double *yy_dev,
                                        int P, float *t32_dev, double *t64_dev, int64_t ldt, int nan_to_zero, void *stream) {
    TMB_REQUIRE(cperm_dev && C_dev && colmap_dev && colmap_host && yy_dev && (t32_dev || t64_dev), "tmb_glm_tstat_cross_rows: null pointer");
    TMB_REQUIRE(m >= 1 && m <= kMaxGenericR && f >= 0 && (f == 0 || cfix_dev) && r >= 1 && r <= 16 && row0 >= 0 && nrows >= 1 &&
                    row0 + nrows <= r && ldb >= V && ldt >= V && (f == 0 || ldf >= V),
                "tmb_glm_tstat_cross_rows: bad shape (m=%d f=%d r=%d rows %d..%d; at most 16 regressors)", m, f, r, row0, row0 + nrows);
    for (int i = 0; i < r; ++i)
        TMB_REQUIRE(colmap_host[i] >= 0 && colmap_host[i] < m + f, "tmb_glm_tstat_cross_rows: column source %d out of range", colmap_host[i]);
    TMB_DEVICE_OF(cperm_dev, "tmb_glm_tstat_cross_rows");
    GlmParams p{};
    p.V = V; p.G = C_dev; p.r = r; p.row0 = row0; p.nrows = nrows; p.dof = dof; p.cfix = cfix_dev; p.ldf = ldf; p.cross_m = m;
    p.cross_f = f; p.colmap = colmap_dev; p.yy = yy_dev; p.P = P; p.t32 = t32_dev; p.t64 = t64_dev; p.ldt = ldt;
    p.nan_to_zero = nan_to_zero; p.mode = 7;
    return launch_stats_from_beta(p, cperm_dev, ldb, (cudaStream_t)stream);
}

extern "C" int tmb_sobelz_cross_rows(const double *cperm_dev, int64_t ldb, int m, const double *cfix_dev, int64_t ldf, int f,
                                     int64_t V, const double *CA_dev, int rA, int rowA, double dofA, const double *CB_dev,
                                     int rB, int rowB, double dofB, const int32_t *colmap_dev, const int32_t *colmap_host,
                                     const double *yy_dev, const double *ta_scalar_dev, int P, int alg, float *z32_dev,
                                     double *z64_dev, int64_t ldt, void *stream) {
    TMB_REQUIRE(cperm_dev && CB_dev && colmap_dev && colmap_host && yy_dev && (z32_dev || z64_dev), "tmb_sobelz_cross_rows: null pointer");
    TMB_REQUIRE(m >= 1 && m <= kMaxGenericR && f >= 0 && (f == 0 || cfix_dev) && rA >= 0 && rA <= 16 && rB >= 1 && rB <= 16 &&
                    rowB >= 0 && rowB < rB && (ta_scalar_dev || (rA >= 1 && rowA >= 0 && rowA < rA && CA_dev)) && ldb >= V &&
                    ldt >= V && (f == 0 || ldf >= V),
                "tmb_sobelz_cross_rows: bad shape (m=%d f=%d rA=%d rB=%d; at most 16 regressors per path)", m, f, rA, rB);
    TMB_REQUIRE(alg >= 0 && alg <= 2, "tmb_sobelz_cross_rows: alg must be 0 (aroian), 1 (sobel) or 2 (goodman)");
    for (int i = 0; i < rA + rB; ++i)
        TMB_REQUIRE(colmap_host[i] >= 0 && colmap_host[i] < m + f, "tmb_sobelz_cross_rows: column source %d out of range", colmap_host[i]);
    TMB_DEVICE_OF(cperm_dev, "tmb_sobelz_cross_rows");
    GlmParams p{};
    p.V = V; p.G = CA_dev; p.GB = CB_dev; p.rA = rA; p.rB = rB; p.rowA = rowA; p.rowB = rowB; p.dof = dofA; p.dofB = dofB;
    p.cfix = cfix_dev; p.ldf = ldf; p.cross_m = m; p.cross_f = f; p.colmap = colmap_dev; p.yy = yy_dev; p.ta_scalar = ta_scalar_dev;
    p.P = P; p.alg = alg; p.t32 = z32_dev; p.t64 = z64_dev; p.ldt = ldt; p.mode = 6;
    return launch_stats_from_beta(p, cperm_dev, ldb, (cudaStream_t)stream);
}

// ---- statistics from stored betas (designs with more than 8 regressors; see glm_stats_from_beta_kernel) ----------
extern "C" int tmb_glm_tstat_beta(const double *beta_dev, int64_t ldb, int64_t V, const double *G_dev, const double *d_dev,
                                  int P, int r, int row0, int nrows, double dof, const double *yy_dev, float *t32_dev,
                                  double *t64_dev, int64_t ldt, int nan_to_zero, void *stream) {
    TMB_REQUIRE(beta_dev && G_dev && d_dev && yy_dev && (t32_dev || t64_dev), "tmb_glm_tstat_beta: null pointer");
    TMB_REQUIRE(V > 0 && P > 0 && r >= 1 && row0 >= 0 && nrows >= 1 && row0 + nrows <= r && ldb >= V && ldt >= V,
                "tmb_glm_tstat_beta: bad shape");
    TMB_DEVICE_OF(beta_dev, "tmb_glm_tstat_beta");
    GlmParams p{};
    p.V = V; p.G = G_dev; p.d = d_dev; p.P = P; p.r = r; p.row0 = row0; p.nrows = nrows; p.dof = dof; p.yy = yy_dev;
    p.t32 = t32_dev; p.t64 = t64_dev; p.ldt = ldt; p.nan_to_zero = nan_to_zero; p.mode = 0;
    return launch_stats_from_beta(p, beta_dev, ldb, (cudaStream_t)stream);
}

extern "C" int tmb_glm_fstat_beta(const double *beta_dev, int64_t ldb, int64_t V, const double *G_dev, const double *M_dev,
                                  int P, int r, int nvar, const int32_t *var_lo, const int32_t *var_k, int want_model,
                                  double dof, const double *yy_dev, const double *sstotal_dev, float *f32_dev,
                                  double *f64_dev, int64_t ldt, int nan_to_zero, void *stream) {
    TMB_REQUIRE(beta_dev && G_dev && yy_dev && (f32_dev || f64_dev), "tmb_glm_fstat_beta: null pointer");
    TMB_REQUIRE(V > 0 && P > 0 && r >= 1 && nvar >= 0 && nvar <= 8 && (nvar == 0 || (M_dev && var_lo && var_k)) &&
                    (nvar > 0 || want_model) && ldb >= V && ldt >= V, "tmb_glm_fstat_beta: bad shape");
    TMB_DEVICE_OF(beta_dev, "tmb_glm_fstat_beta");
    GlmParams p{};
    p.V = V; p.G = G_dev; p.M = M_dev; p.P = P; p.r = r; p.row0 = want_model ? 0 : 1; p.nrows = nvar + (want_model ? 1 : 0);
    p.dof = dof; p.yy = yy_dev; p.t32 = f32_dev; p.t64 = f64_dev; p.ldt = ldt; p.nan_to_zero = nan_to_zero; p.mode = 3; p.nvar = nvar;
    p.sstot = sstotal_dev;
    for (int i = 0; i < nvar; ++i) {
        TMB_REQUIRE(var_lo[i] >= 0 && var_k[i] >= 1 && var_lo[i] + var_k[i] <= r,
                    "tmb_glm_fstat_beta: variable %d covers rows [%d, %d) of %d", i, var_lo[i], var_lo[i] + var_k[i], r);
        p.var_lo[i] = var_lo[i]; p.var_k[i] = var_k[i]; p.msz += var_k[i] * var_k[i];
    }
    return launch_stats_from_beta(p, beta_dev, ldb, (cudaStream_t)stream);
}

extern "C" int tmb_glm_cosinor_beta(const double *beta_dev, int64_t ldb, int64_t V, const double *G_dev, const double *C_dev,
                                    int P, int r, int nper, int nexog, double dof, const double *yy_dev,
                                    const double *sstotal_dev, int mediation, double ta, int alg, float *s32_dev,
                                    double *s64_dev, int64_t ldt, int nan_to_zero, void *stream) {
    TMB_REQUIRE(beta_dev && G_dev && C_dev && yy_dev && (s32_dev || s64_dev), "tmb_glm_cosinor_beta: null pointer");
    TMB_REQUIRE(V > 0 && P > 0 && nper >= 1 && nexog >= 0 && 2 * nper + nexog <= r && ldb >= V && ldt >= V &&
                    (!mediation || nexog >= 1) && alg >= 0 && alg <= 2,
                "tmb_glm_cosinor_beta: bad shape (P=%d r=%d periods=%d tested columns=%d)", P, r, nper, nexog);
    TMB_DEVICE_OF(beta_dev, "tmb_glm_cosinor_beta");
    GlmParams p{};
    p.V = V; p.G = G_dev; p.M = C_dev; p.P = P; p.r = r; p.nvar = nper; p.cos_nexog = nexog; p.dof = dof; p.yy = yy_dev;
    p.cos_mediation = mediation ? 1 : 0; p.cos_ta = ta; p.alg = alg; p.sstot = sstotal_dev; p.nrows = mediation ? 1 : 1 + 2 * nper + nexog;
    p.t32 = s32_dev; p.t64 = s64_dev; p.ldt = ldt; p.nan_to_zero = nan_to_zero; p.mode = 4;
    return launch_stats_from_beta(p, beta_dev, ldb, (cudaStream_t)stream);
}

extern "C" int tmb_sobelz_beta(const double *beta_dev, int64_t ldb, int64_t V, const double *GA_dev, const double *dA_dev,
                               int rA, int rowA, double dofA, const double *GB_dev, const double *dB_dev, int rB, int rowB,
                               double dofB, const double *yy_dev, const double *ta_scalar_dev, int P, int alg,
                               float *z32_dev, double *z64_dev, int64_t ldt, void *stream) {
    TMB_REQUIRE(beta_dev && GB_dev && dB_dev && yy_dev && (z32_dev || z64_dev), "tmb_sobelz_beta: null pointer");
    TMB_REQUIRE(rA >= 0 && rB >= 1 && rowB >= 0 && rowB < rB && ldb >= V && ldt >= V &&
                    (ta_scalar_dev || (rA >= 1 && rowA >= 0 && rowA < rA && GA_dev && dA_dev)), "tmb_sobelz_beta: bad shape");
    TMB_REQUIRE(alg >= 0 && alg <= 2, "tmb_sobelz_beta: alg must be 0 (aroian), 1 (sobel) or 2 (goodman)");
    TMB_DEVICE_OF(beta_dev, "tmb_sobelz_beta");
    GlmParams p{};
    p.V = V; p.G = GA_dev; p.d = dA_dev; p.P = P; p.r = rA + rB; p.dof = dofA; p.yy = yy_dev; p.t32 = z32_dev; p.t64 = z64_dev;
    p.ldt = ldt; p.mode = 2; p.GB = GB_dev; p.dB = dB_dev; p.rA = rA; p.rB = rB; p.rowA = rowA; p.rowB = rowB; p.dofB = dofB;
    p.ta_scalar = ta_scalar_dev; p.alg = alg;
    return launch_stats_from_beta(p, beta_dev, ldb, (cudaStream_t)stream);
}
