"""Drop-in for the hot-path subset of the reference's ``tfce_mediation.pyfunc``.

    create_adjac_vertex(vertices, faces)                         pyfunc.py:37-46
    create_adjac_voxel(data_index, data_mask, num_voxel, dirtype) pyfunc.py:48-76
    write_perm_maxTFCE_vertex(...)                                pyfunc.py:107-119
    write_perm_maxTFCE_voxel(...)                                 pyfunc.py:121-126
    calc_sobelz(medtype, pred_x, depend_y, merge_y, n, num_vertex, alg) pyfunc.py:130-162

Same names, arguments and side effects (one formatted line appended to a CSV in the current
directory).  These are the single-map forms; the randomise drivers use the batched engine
(engine.PermutationEngine) which produces identical rows for whole blocks of shuffles.
"""
import ctypes

import numpy as np

from . import _lib
from .cynumstats import calc_beta_se  # noqa: F401  (re-exported like the reference does, pyfunc.py:34)


def _append_line(path, text):
    # the reference shells out to `echo ... >> file` (pyfunc.py:119,126); same bytes, no fork
    with open(path, "a") as f:
        f.write(text + "\n")


def create_adjac_vertex(vertices, faces):
    """1-ring neighbour sets from triangle faces (pyfunc.py:37-46).  Host-side integer set building."""
    adjacency = [set([]) for _ in range(vertices.shape[0])]
    f = np.asarray(faces)
    for a, b in ((0, 1), (0, 2), (1, 0), (1, 2), (2, 0), (2, 1)):
        for u, v in zip(f[:, a].tolist(), f[:, b].tolist()):
            adjacency[u].add(v)
    return adjacency


def _voxel_csr(mask, dirtype, variant):
    _lib.require_device()
    import torch
    m = np.ascontiguousarray(np.asarray(mask) != 0, dtype=np.uint8)
    if m.ndim != 3:
        raise ValueError("mask must be 3-D")
    nx, ny, nz = m.shape
    nv, nnz = ctypes.c_int32(0), ctypes.c_int64(0)
    dev = torch.cuda.current_device()
    L = _lib.lib()
    _lib.check(L.tmb_voxel_adjacency(dev, _lib.ptr(m), nx, ny, nz, int(dirtype), variant, ctypes.byref(nv),
                                     ctypes.byref(nnz), None, None))
    indptr = np.zeros(nv.value + 1, dtype=np.int64)
    indices = np.zeros(max(nnz.value, 1), dtype=np.int32)
    _lib.check(L.tmb_voxel_adjacency(dev, _lib.ptr(m), nx, ny, nz, int(dirtype), variant, ctypes.byref(nv),
                                     ctypes.byref(nnz), _lib.ptr(indptr), _lib.ptr(indices)))
    return indptr, indices[:nnz.value]


def create_adjac_voxel(data_index, data_mask, num_voxel, dirtype=26):
    """26- or 6-connectivity voxel adjacency (pyfunc.py:48-76): object array of sorted lists,
    self excluded, ``adjacency[0] = []`` and label 0 never listed.  Built by the GPU kernel.
    (data_mask only supplies the volume shape in the reference; it does the same here.)"""
    data_index = np.asarray(data_index)
    if data_index.shape != np.asarray(data_mask).shape:
        raise ValueError("data_index and data_mask must have the same shape")
    indptr, indices = _voxel_csr(data_index, dirtype, 0)
    if indptr.shape[0] - 1 != int(num_voxel):
        raise ValueError("num_voxel=%d but the mask holds %d voxels" % (int(num_voxel), indptr.shape[0] - 1))
    out = np.empty(indptr.shape[0] - 1, dtype=object)
    for i in range(out.shape[0]):
        out[i] = indices[indptr[i]:indptr[i + 1]].tolist()
    return out


def create_adjac_voxel_tools(data_index, dirtype=26):
    """tools/tm_mulitmodality_adjacency.py:40-66 variant: list of sets, self kept, voxel 0 asymmetric."""
    indptr, indices = _voxel_csr(data_index, dirtype, 1)
    return [set(indices[indptr[i]:indptr[i + 1]].tolist()) for i in range(indptr.shape[0] - 1)]


_PLANS = []        # [(weakref to every CreateAdjSet, key, plan)]: the single-map writers reuse one TFCE plan per graph set


def _plan_for(adjsets, masks, weights):
    """A TfcePlan over the kept vertices of the given CreateAdjSet objects (cached while the same objects are passed):
    masked-out vertices carry statistic 0 in the reference (pyfunc.py:108-113) and can never activate, so the plan runs
    on the induced sub-graphs, one surface per hemisphere, laid side by side along one statistic row."""
    import weakref
    from ._graph import induced_subgraph
    from .engine import Surface, TfcePlan
    from .tfce import CreateAdjSet
    key = tuple(id(w) if np.ndim(w) else float(w) for w in weights)
    for refs, k, plan in _PLANS:
        if k == key and len(refs) == len(adjsets) and all(r() is a for r, a in zip(refs, adjsets)):
            return plan
    surfs, off = [], 0
    for adj, mask, w in zip(adjsets, masks, weights):
        keep = None if mask is None else np.asarray(mask, dtype=bool)
        graph = (adj.indptr, adj.indices) if keep is None else induced_subgraph(adj.indptr, adj.indices, keep)
        if np.ndim(w):
            w = np.asarray(w)[keep] if keep is not None and np.shape(w)[0] == keep.shape[0] else np.asarray(w)
        surfs.append(Surface(CreateAdjSet(adj.H, adj.E, graph), off, w))
        off += graph[0].shape[0] - 1
    plan = TfcePlan(surfs)
    _PLANS.append(([weakref.ref(a) for a in adjsets], key, plan))
    del _PLANS[:-4]
    return plan


def _single_map_max(plan, stat):
    """Scaled TFCE maxima [S] of ONE one-sided statistic row (a block of one map through the batched pipeline)."""
    import torch
    row = torch.zeros((1, (plan.row_len + 3) // 4 * 4), dtype=torch.float32, device=plan.device)
    row[0, :stat.shape[0]] = torch.from_numpy(np.ascontiguousarray(stat, dtype=np.float32)).to(plan.device)
    mx, _, _ = plan.run(row, two_sided=False)
    return mx[0, :, 0].cpu().numpy()


def write_perm_maxTFCE_vertex(statname, vertStat, num_vertex, bin_mask_lh, bin_mask_rh, calcTFCE_lh, calcTFCE_rh,
                              density_corr_lh=1, density_corr_rh=1):
    """pyfunc.py:107-119: TFCE of the statistic on both hemispheres, max over both of tfce * (max(stat)/100) * density,
    one '%.4f' row appended to perm_<statname>_TFCE_maxVertex.csv.  vertStat holds the kept vertices of lh (the first
    num_vertex entries) and rh; the two hemispheres run as the two surfaces of one cached plan."""
    plan = _plan_for((calcTFCE_lh, calcTFCE_rh), (bin_mask_lh, bin_mask_rh), (density_corr_lh, density_corr_rh))
    stat = np.asarray(vertStat, dtype=np.float32)
    stat = np.where(np.isfinite(stat), stat, np.float32(0))          # the reference drops non-finite entries (:116-117)
    _append_line("perm_%s_TFCE_maxVertex.csv" % statname, "%.4f" % _single_map_max(plan, stat).max())


def write_perm_maxTFCE_voxel(statname, voxelStat, TFCEfunc):
    """pyfunc.py:121-126: TFCE of the voxel statistic, tfce.max() * (stat.max()/100), one '%1.4f' row."""
    plan = _plan_for((TFCEfunc,), (None,), (1,))
    _append_line("perm_%s_TFCE_maxVoxel.csv" % statname,
                 "%1.4f" % _single_map_max(plan, np.asarray(voxelStat, dtype=np.float32))[0])


def calc_sobelz(medtype, pred_x, depend_y, merge_y, n, num_vertex, alg="aroian"):
    """pyfunc.py:130-162: Sobel-family z from two fits; float64 [V].  Runs the fused GPU kernel
    (tmb_sobelz) through a one-shuffle engine call with the identity permutation."""
    from .engine import sobelz_single
    return sobelz_single(medtype, pred_x, depend_y, merge_y, alg)


def typeI_design(exog, dmy_covariates, n):
    """The design of pyfunc.py:2304-2315: [1, exog variables ..., covariates] and the column count of each variable."""
    blocks = [np.asarray(v, dtype=np.float64).reshape(n, -1) for v in exog]
    cols = [np.ones((n, 1))] + blocks
    if dmy_covariates is not None:
        cols.append(np.asarray(dmy_covariates, dtype=np.float64).reshape(n, -1))
    return np.hstack(cols), [blk.shape[1] for blk in blocks]


def glm_typeI(endog, exog, dmy_covariates=None, output_fvalues=True, output_tvalues=False, output_pvalues=False,
              verbose=True, rand_array=None, use_reduced_residuals=False, output_reduced_residuals=False,
              exog_names=None):
    """pyfunc.py:2282-2401 glm_typeI: model F, per-variable F (Type I extra sum of squares) and t of the design
    [1, exog..., dmy_covariates] with rows optionally permuted by rand_array; same return tuples as the reference.
    The F statistics come from ONE fused GPU fit (tmb_glm_fstat: RSS_without_i - RSS = b_S' inv(C_SS) b_S), the t
    values from the tval_int drop-in; p-values (scipy) stay on the host like in the reference."""
    from . import cynumstats
    from .engine import PermutationEngine, design_stack, to_host
    endog = np.asarray(endog)
    if endog.ndim == 1:
        endog = endog[:, None]
    n = endog.shape[0]
    exog_vars, kvars = typeI_design(exog, dmy_covariates, n)
    if rand_array is not None:
        if use_reduced_residuals:
            endog = np.asarray(cynumstats.resid_covars(exog_vars, endog.T))
        exog_vars = exog_vars[rand_array]
    reduced_data = None
    if output_reduced_residuals:
        reduced_data = np.asarray(cynumstats.resid_covars(exog_vars, endog.T))
    k = exog_vars.shape[1]
    DF_Between, DF_Within, DF_Total = k - 1, n - k, n - 1
    V = endog.shape[1]
    Fvalues = Fvar = Tvalues = None
    if output_fvalues:
        eng = PermutationEngine(endog, None)
        var_lo = np.concatenate([[0], np.cumsum(kvars)[:-1]]).astype(np.int32)
        _, f64 = eng.fstat(design_stack(exog_vars[None], center=True), var_lo, kvars, want_model=True, want_f64=True)
        f64 = to_host(f64[0, :, :V])
        Fvalues, Fvar = f64[0], f64[1:]
        if verbose:
            print("Source\t\tDF\tF(Max)")
            print("Model\t\t(%d,%d)\t%.2f" % (DF_Between, DF_Within, Fvalues.max()))
            for i, col in enumerate(kvars):
                print("%s\t\t(%d,%d)\t%.2f" % (exog_names[i] if exog_names is not None else "Exog%d" % (i + 1), col,
                                                DF_Within, Fvar[i].max()))
    if output_tvalues:
        invXX = np.linalg.inv(np.dot(exog_vars.T, exog_vars))
        Tvalues = cynumstats.tval_int(exog_vars, invXX, endog, n, k, V)
    if output_pvalues:
        from scipy.stats import f as f_dist, t as t_dist
    if output_tvalues and output_fvalues:
        if output_pvalues:
            Pvar = np.array([f_dist.sf(Fvar[i], col, DF_Within) for i, col in enumerate(kvars)])
            return (Fvalues, Fvar, Tvalues, f_dist.sf(Fvalues, DF_Between, DF_Within), Pvar,
                    t_dist.sf(np.abs(Tvalues), DF_Total) * 2)
        return (Fvalues, Fvar, Tvalues, reduced_data) if output_reduced_residuals else (Fvalues, Fvar, Tvalues)
    if output_tvalues:
        if output_pvalues:
            return (Tvalues, t_dist.sf(np.abs(Tvalues), DF_Total) * 2)
        return (Tvalues, reduced_data) if output_reduced_residuals else Tvalues
    if output_fvalues:
        if output_pvalues:
            Pvar = np.array([f_dist.sf(Fvar[i], col, DF_Within) for i, col in enumerate(kvars)])
            return (Fvalues, Fvar, f_dist.sf(Fvalues, DF_Between, DF_Within), Pvar)
        return (Fvalues, Fvar, reduced_data) if output_reduced_residuals else (Fvalues, Fvar)
    print("No output has been selected")


def _rm_ancova(data, factors, dmy_subjects, dmy_covariates, data_format, output_sig, verbose, rand_array,
               use_reduced_residuals, output_reduced_residuals, labels):
    """Body of the two repeated-measures ANCOVA drop-ins: long-format view, the reference's side effects on `data`, the GPU
    statistics (engine.rm_ancova_stats), degrees of freedom, optional p-values / residuals."""
    from . import cynumstats
    from .engine import PermutationEngine, to_host
    from .rmancova import RmAncovaModel
    n = len(factors[0])
    if data_format == "short":
        if data.ndim == 2:
            data = data[:, :, np.newaxis]
        s = data.shape[0]
        endog = data.reshape(s * n, data.shape[2])          # a view of the caller's array when it is contiguous
    elif data_format == "long":
        s = int(len(data) // n)                             # the reference's len(data)/n is a float under Python 3
        endog = data if data.ndim == 2 else data[:, np.newaxis]
    else:
        raise ValueError("data format must be short or long")
    def between_design(order):
        """[1, factors(, interaction), covariates] in long format, subjects taken in `order` (pyfunc.py:1810-1823, 1866-1870,
        2142-2146, 2176-2178)."""
        long_ = lambda x: np.concatenate([np.asarray(x, dtype=np.float64).reshape(n, -1)[order]] * s, 0)   # noqa: E731
        cols = [long_(f) for f in factors]
        if len(factors) == 2:
            f1, f2 = (np.asarray(f, dtype=np.float64).reshape(n, -1) for f in factors)
            cols.append(long_(np.concatenate([f1[:, i:i + 1] * f2 for i in range(f1.shape[1])], axis=1)))
        if dmy_covariates is not None:
            cols.append(long_(dmy_covariates))
        return np.column_stack([np.ones(s * n)] + cols)

    if rand_array is not None:
        if use_reduced_residuals:
            endog = np.asarray(cynumstats.resid_covars(between_design(np.arange(n)), np.asarray(endog).T))
        order = np.arange(s * n)
        np.random.shuffle(order)                            # the draws of np.random.shuffle(endog_arr), pyfunc.py:1826 / 2148
        endog[:] = endog[order]                             # in place, like the reference: the caller's data stay shuffled
    model = RmAncovaModel(n, s, factors, dmy_subjects, dmy_covariates)
    eng = PermutationEngine(np.ascontiguousarray(endog), None)
    _, f64 = eng.rm_ancova_stats(model, None, [np.arange(n) if rand_array is None else np.asarray(rand_array)], want_f64=True)
    F = to_host(f64[0, :, :endog.shape[1]])
    df_w, df_s = model.df["within_factors"], model.df["s"]
    dfs = [(model.df[k], df_w) if "s" not in k else (df_s if k == "s" else model.df[k[1:]] * df_s, df_w * df_s)
           for k in model.names]
    if verbose:
        print("Source\t\tDF\tF(Max)")
        for lab, (d1, d2), f in zip(labels, dfs, F):
            print("%s\t(%d,%d)\t%.2f" % (lab, d1, d2, f.max()))
    out = tuple(F)
    if output_sig:
        from scipy.stats import f as f_dist
        return out + tuple(1 - f_dist.cdf(f, d1, d2) for f, (d1, d2) in zip(F, dfs))
    if output_reduced_residuals:
        order = np.arange(n) if rand_array is None else np.asarray(rand_array)
        return out + (np.asarray(cynumstats.resid_covars(between_design(order), np.asarray(endog).T)),)
    return out


def reg_rm_ancova_one_bs_factor(data, dmy_factor1, dmy_subjects, data_format="short", dmy_covariates=None, output_sig=False,
                                verbose=True, rand_array=None, use_reduced_residuals=False, output_reduced_residuals=False):
    """pyfunc.py:2068-2280: repeated-measures ANCOVA with one between-subject factor -> (F_a, F_s, F_sa[, P... | residuals]).
    Same arguments and side effects (with rand_array the rows of the long-format `data` are shuffled in place with the global
    numpy stream); the statistics run on the GPU (csrc/rmancova_kernels.cu)."""
    return _rm_ancova(data, [dmy_factor1], dmy_subjects, dmy_covariates, data_format, output_sig, verbose, rand_array,
                      use_reduced_residuals, output_reduced_residuals, ["Factor\t", "Time\t", "Factor*Time"])


def reg_rm_ancova_two_bs_factor(data, dmy_factor1, dmy_factor2, dmy_subjects, dmy_covariates=None, data_format="short",
                                output_sig=False, verbose=True, rand_array=None, use_reduced_residuals=False,
                                output_reduced_residuals=False):
    """pyfunc.py:1712-2052: two between-subject factors -> (F_a, F_b, F_ab, F_s, F_sa, F_sb, F_sab[, P... | residuals])."""
    return _rm_ancova(data, [dmy_factor1, dmy_factor2], dmy_subjects, dmy_covariates, data_format, output_sig, verbose,
                      rand_array, use_reduced_residuals, output_reduced_residuals,
                      ["Factor1\t", "Factor2\t", "F1*F2\t", "Time\t", "F1*Time\t", "F2*Time\t", "F1*F2*Time"])


def glm_cosinor(endog, time_var, exog=None, dmy_covariates=None, rand_array=None, interaction_var=None, period=[24.0],
                calc_MESOR=True, output_fit_only=False):
    """pyfunc.py:2406-2563: cosinor model -> (R2, MESOR, SE_MESOR, AMPLITUDE, SE_AMPLITUDE, ACROPHASE, SE_ACROPHASE, Fmodel,
    tMESOR, |tAMPLITUDE|, |tACROPHASE|, tEXOG), or (MESOR, AMPLITUDE, ACROPHASE) with output_fit_only.  The fit, the
    residual sums of squares and the standard errors come from the cynumstats drop-ins (GPU); the closed-form amplitude /
    acrophase algebra on [periods, V] arrays stays on the host.  (interaction_var: the reference multiplies a ROW of the
    design with it, pyfunc.py:2443-2445, which only works by accident of shapes; not supported.)"""
    from . import cynumstats
    from .engine import cosinor_design
    if interaction_var is not None:
        raise NotImplementedError("glm_cosinor: interaction_var is not supported (see the docstring)")
    endog = np.asarray(endog)
    one_d = endog.ndim == 1
    n = endog.shape[0]
    X, nper, _ = cosinor_design(time_var, period, exog, dmy_covariates)
    if rand_array is not None:
        X = X[rand_array]
    k = X.shape[1]
    a, ss_res = cynumstats.cy_lin_lstsqr_mat_residual(X, endog)
    a2 = a[:, None] if one_d else a
    cosr, sinr = a2[1:1 + 2 * nper:2], a2[2:2 + 2 * nper:2]             # [periods, V]
    amplitude = np.sqrt(cosr ** 2 + sinr ** 2)
    acro = np.arctan(np.abs(np.divide(-sinr, cosr)))

    def quadrant(phi):
        """Acrophase in (-2 pi, 0] from the signs of the two coefficients (pyfunc.py:2477-2481)."""
        out = phi.copy()
        out[(sinr > 0) & (cosr >= 0)] = -phi[(sinr > 0) & (cosr >= 0)]
        out[(sinr > 0) & (cosr < 0)] = -np.pi + phi[(sinr > 0) & (cosr < 0)]
        out[(sinr < 0) & (cosr <= 0)] = -np.pi - phi[(sinr < 0) & (cosr <= 0)]
        out[(sinr <= 0) & (cosr > 0)] = -2 * np.pi + phi[(sinr <= 0) & (cosr > 0)]
        return out

    if output_fit_only:
        return a[0], amplitude, quadrant(acro)
    ss_total = np.sum((endog - np.mean(endog, 0)) ** 2, 0)
    ms_res = ss_res / (n - k)
    fmodel = ((ss_total - ss_res) / (k - 1)) / ms_res
    sigma = np.sqrt(ss_res / (n - k))
    invXX = np.linalg.inv(np.dot(X.T, X))
    mesor = t_mesor = se_mesor = t_exog = None
    if calc_MESOR or exog is not None:
        if one_d:
            se = np.sqrt(np.diag(sigma * sigma * invXX))
        else:
            se = cynumstats.se_of_slope(endog.shape[1], invXX, sigma ** 2, k)
        tvalues = a / se
        mesor, t_mesor, se_mesor = a[0], tvalues[0], se[0]
        if exog is not None:
            t_exog = (tvalues[:, None] if one_d else tvalues)[1 + 2 * nper:]
    ci = 1 + 2 * np.arange(nper)
    c11, c12, c22 = invXX[ci, ci][:, None], invXX[ci, ci + 1][:, None], invXX[ci + 1, ci + 1][:, None]
    sn, cs = np.sin(acro), np.cos(acro)
    se_acro = sigma * np.sqrt((c11 * sn ** 2) + (2 * c12 * sn * cs) + (c22 * cs ** 2)) / amplitude
    se_amp = sigma * np.sqrt((c11 * cs ** 2) - (2 * c12 * sn * cs) + (c22 * sn ** 2))
    r2 = (1 - ss_res / ss_total) if rand_array is None else None
    return (r2, mesor, se_mesor, amplitude, se_amp, quadrant(acro) if rand_array is None else acro, se_acro, fmodel, t_mesor,
            np.abs(amplitude / se_amp), np.abs(1.0 / se_acro), np.array(t_exog))


# ---- small host helpers of the tm-models scripts (design coding; pyfunc.py:2565-2709), k x n work only -------------------
def calc_indirect(ta, tb, alg="aroian"):
    """Sobel-family z of two t maps (pyfunc.py:2678-2709): aroian 1/sqrt(1/tb^2 + 1/ta^2 + 1/(ta^2 tb^2)), sobel without the
    product term, goodman with it subtracted."""
    ta2, tb2 = np.square(ta), np.square(tb)
    terms = (1 / tb2) + (1 / ta2)
    if alg == "aroian":
        return 1 / np.sqrt(terms + (1 / (ta2 * tb2)))
    if alg == "sobel":
        return 1 / np.sqrt(terms)
    if alg == "goodman":
        return 1 / np.sqrt(terms - (1 / (ta2 * tb2)))
    raise ValueError("Unknown indirect test algorithm")


def dummy_code(variable, iscontinous=False, demean=True):
    """One indicator column per level except the first (int, squeezed; pyfunc.py:2565-2595), or the variable itself when
    continuous; optionally centred."""
    variable = np.asarray(variable)
    if iscontinous:
        return variable - np.mean(variable, 0) if demean else variable
    levels = np.unique(variable)[1:]
    coded = np.squeeze(np.array([(variable == lv) for lv in levels]).astype(int)).T
    return coded - np.mean(coded, 0) if demean else coded


def dummy_code_cosine(time, period=24.0):
    """[cos(2 pi t / T), sin(2 pi t / T)] (pyfunc.py:2597-2620)."""
    angle = np.divide(2.0 * np.pi * np.asarray(time), period)
    return np.column_stack((np.cos(angle), np.sin(angle)))


def column_product(arr1, arr2):
    """Every column of arr1 times every column of arr2, arr1-major (pyfunc.py:2622-2660)."""
    arr1, arr2 = np.array(arr1), np.array(arr2)
    if len(arr1) != len(arr2):
        raise ValueError("arrays must be of same length")
    if arr1.ndim == 1 or arr2.ndim == 1:
        a, b = (arr1, arr2) if arr1.ndim == 1 else (arr2, arr1)
        return (a * b.T).T + 0
    return np.concatenate([arr1[:, i:i + 1] * arr2 for i in range(arr1.shape[1])], axis=1) + 0


def stack_ones(arr):
    """The array with a leading column of ones (pyfunc.py:2662-2676)."""
    return np.column_stack([np.ones(len(arr)), arr])


def check_blocks(block_list):
    """pyfunc.py:2711-2731."""
    unique_blocks = np.unique(block_list)
    block_sizes = [len(block_list[block_list == block]) for block in unique_blocks]
    is_equal_sizes = all(x == block_sizes[0] for x in block_sizes)
    if not is_equal_sizes:
        print("Warning: blocks are not equal. Swaping with only occur within blocks, but not among blocks.")
    return is_equal_sizes


def rand_blocks(block_list, is_equal_sizes):
    """pyfunc.py:2733-2757: permutation index from exchangeability blocks -- the same numpy RNG calls in the same order:
    with equal block sizes the block ORDER is drawn first, then (always) one permutation inside every block."""
    block_list = np.asarray(block_list)
    where = np.arange(len(block_list))
    labels = np.unique(block_list)
    order = np.random.permutation(list(labels)) if is_equal_sizes is True else labels
    return np.concatenate([np.random.permutation(where[block_list == lab]) for lab in order])
