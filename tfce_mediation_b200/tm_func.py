"""Drop-in for the hot-path subset of the reference's ``tfce_mediation.tm_func`` (mmr / mmr-lr glue).

    create_full_mask(masking_array)                      tm_func.py:502-510
    merge_adjacency_array(adjacent_range, adjacency)     tm_func.py:521-540
    create_position_array(masking_array)                 tm_func.py:556-562
    calculate_tfce(...)                                  tm_func.py:54-123
    low_ram_calculate_tfce(...)                          tm_func.py:144-185
    calculate_mediation_tfce(...)                        tm_func.py:207-249
    low_ram_calculate_mediation_tfce(...)                tm_func.py:269-305
    calc_mixed_tfce(...)                                 tm_func.py:327-378

Same names, arguments, return values and CSV side effects ('%f' rows, positive then negative) as the
reference.  The per-vertex arithmetic runs on the GPU (cynumstats.tval_int -> tmb_glm_direct,
CreateAdjSet.run -> tmb_tfce_run, calc_sobelz -> tmb_sobelz); these single-shuffle forms exist for
API parity.  Whole permutation runs go through tm_multisurface.mmr_lr_randomise, which batches
shuffles x surfaces through engine.PermutationEngine and writes the same rows.
"""
import os
from time import time

import numpy as np

from .cynumstats import tval_int
from .pyfunc import calc_sobelz


def _append(path, value):
    with open(path, "a") as f:
        f.write("%f\n" % value)


def _time_seed(perm_number):
    # tm_func.py:57,150,209,275: perm_number + the last six CHARACTERS of str(time()) times 100
    return perm_number + int(float(str(time())[-6:]) * 100)


def create_full_mask(masking_array):
    """Concatenated flat mask of all surfaces (tm_func.py:502-510)."""
    full_mask = None
    for i in range(len(masking_array)):
        piece = masking_array[i][:, 0, 0] if masking_array[i].shape[2] == 1 else masking_array[i][masking_array[i] == True]  # noqa: E712
        full_mask = piece if full_mask is None else np.hstack((full_mask, piece))
    return full_mask


def merge_adjacency_array(adjacent_range, adjacency_array):
    """Block-diagonal merge of per-surface adjacency lists (tm_func.py:521-540); like the reference it
    always starts from adjacency_array[0] and offsets every later surface by the running vertex count."""
    v_count = 0
    if len(adjacent_range) == 1:
        adjacency = np.copy(adjacency_array[0])
    else:
        for e in adjacent_range:
            if v_count == 0:
                adjacency = np.copy(adjacency_array[0])
                v_count += len(adjacency_array[0])
            else:
                temp_adjacency = np.copy(adjacency_array[e])
                for i in range(len(adjacency_array[e])):
                    temp_adjacency[i] = np.add(list(temp_adjacency[i]), v_count).tolist()
                adjacency = np.hstack((adjacency, temp_adjacency))
                v_count += len(adjacency_array[e])
    return adjacency


def create_position_array(masking_array):
    """Start offsets of every surface in the concatenated data (tm_func.py:556-562)."""
    pointer = 0
    position_array = [0]
    for i in range(len(masking_array)):
        pointer += len(masking_array[i][masking_array[i] == True])  # noqa: E712
        position_array.append(pointer)
    return position_array


def calculate_tfce(merge_y, masking_array, pred_x, calcTFCE, vdensity, position_array, fullmask, perm_number=None,
                   randomise=False, verbose=False, no_intercept=True, set_surf_count=None, print_interation=False):
    """Non-low-RAM mmr (tm_func.py:54-123): ONE TFCE over the merged graph (so the threshold step is the
    global maximum / 100) followed by a per-surface rescale and per-surface nanmax rows."""
    X = np.column_stack([np.ones(merge_y.shape[0]), pred_x])
    if randomise:
        np.random.seed(_time_seed(perm_number))
        X = X[np.random.permutation(list(range(merge_y.shape[0])))]
    k = len(X.T)
    invXX = np.linalg.inv(np.dot(X.T, X))
    tvals = tval_int(X, invXX, merge_y, merge_y.shape[0], k, merge_y.shape[1])
    if no_intercept:
        tvals = tvals[1:, :]
    tvals = tvals.astype(np.float32, order="C")
    tfce_tvals = np.zeros_like(tvals).astype(np.float32, order="C")
    neg_tfce_tvals = np.zeros_like(tvals).astype(np.float32, order="C")
    for tstat_counter in range(tvals.shape[0]):
        tval_temp = np.zeros_like((fullmask)).astype(np.float32, order="C")
        tval_temp[fullmask == 1] = tvals[0] if tvals.shape[0] == 1 else tvals[tstat_counter]
        tval_temp = tval_temp.astype(np.float32, order="C")
        tfce_temp = np.zeros_like(tval_temp).astype(np.float32, order="C")
        neg_tfce_temp = np.zeros_like(tval_temp).astype(np.float32, order="C")
        calcTFCE.run(tval_temp, tfce_temp)
        calcTFCE.run((tval_temp * -1), neg_tfce_temp)
        tval_temp = tval_temp[fullmask == 1]
        tfce_temp = tfce_temp[fullmask == 1]
        neg_tfce_temp = neg_tfce_temp[fullmask == 1]
        for surf_count in range(len(masking_array)):
            start = position_array[surf_count]
            end = position_array[surf_count + 1]
            dens = vdensity if isinstance(vdensity, int) else vdensity[start:end]
            tfce_tvals[tstat_counter, start:end] = (tfce_temp[start:end] * (tval_temp[start:end].max() / 100) * dens)
            neg_tfce_tvals[tstat_counter, start:end] = (neg_tfce_temp[start:end] * ((tval_temp * -1)[start:end].max() / 100) * dens)
            label = int(set_surf_count[surf_count]) if set_surf_count is not None else surf_count
            pmax = np.nanmax(tfce_tvals[tstat_counter, start:end])
            nmax = np.nanmax(neg_tfce_tvals[tstat_counter, start:end])
            if randomise:
                _append("perm_maxTFCE_surf%d_tcon%d.csv" % (label, tstat_counter + 1), pmax)
                _append("perm_maxTFCE_surf%d_tcon%d.csv" % (label, tstat_counter + 1), nmax)
            else:
                print("Maximum (untransformed) postive tfce value for surface %s, tcon %d: %f" % (label, tstat_counter + 1, pmax))
                print("Maximum (untransformed) negative tfce value for surface %s, tcon %d: %f" % (label, tstat_counter + 1, nmax))
        if verbose:
            print("T-contrast: %d" % tstat_counter)
            print("Max tfce from all surfaces = %f" % tfce_tvals[tstat_counter].max())
            print("Max negative tfce from all surfaces = %f" % neg_tfce_tvals[tstat_counter].max())
    if randomise:
        if print_interation:
            print("Interation number: %d" % perm_number)
        return None
    return (tvals.astype(np.float32, order="C"), tfce_tvals.astype(np.float32, order="C"),
            neg_tfce_tvals.astype(np.float32, order="C"))


def low_ram_calculate_tfce(data, mask, pred_x, calcTFCE, vdensity, set_surf_count=0, perm_number=None, randomise=False,
                           no_intercept=True, output_dir=None, perm_seed=None):
    """mmr-lr, one surface (tm_func.py:144-185); deterministic when perm_seed is given."""
    X = np.column_stack([np.ones(data.shape[0]), pred_x])
    if randomise:
        np.random.seed(perm_number + perm_seed if perm_seed is not None else _time_seed(perm_number))
        X = X[np.random.permutation(list(range(data.shape[0])))]
    k = len(X.T)
    invXX = np.linalg.inv(np.dot(X.T, X))
    tvals = tval_int(X, invXX, data, data.shape[0], k, data.shape[1])
    if no_intercept:
        tvals = tvals[1:, :]
    tvals = tvals.astype(np.float32, order="C")
    tfce_tvals = np.zeros_like(tvals).astype(np.float32, order="C")
    neg_tfce_tvals = np.zeros_like(tvals).astype(np.float32, order="C")
    for tstat_counter in range(tvals.shape[0]):
        tval_temp = np.zeros_like((mask)).astype(np.float32, order="C")
        tval_temp[mask == 1] = tvals[0] if tvals.shape[0] == 1 else tvals[tstat_counter]
        tval_temp = tval_temp.astype(np.float32, order="C")
        tfce_temp = np.zeros_like(tval_temp).astype(np.float32, order="C")
        neg_tfce_temp = np.zeros_like(tval_temp).astype(np.float32, order="C")
        calcTFCE.run(tval_temp, tfce_temp)
        calcTFCE.run(-tval_temp, neg_tfce_temp)
        tfce_tvals[tstat_counter, :] = (tfce_temp[mask == 1] * (tval_temp.max() / 100) * vdensity)
        neg_tfce_tvals[tstat_counter, :] = (neg_tfce_temp[mask == 1] * ((tval_temp * -1).max() / 100) * vdensity)
        if randomise:
            name = "perm_maxTFCE_surf%d_tcon%d.csv" % (int(set_surf_count), tstat_counter + 1)
            permfile = "%s/%s" % (output_dir, name) if output_dir is not None else name
            _append(permfile, np.nanmax(tfce_tvals[tstat_counter, :]))
            _append(permfile, np.nanmax(neg_tfce_tvals[tstat_counter, :]))
    if not randomise:
        return (tvals.astype(np.float32, order="C"), tfce_tvals.astype(np.float32, order="C"),
                neg_tfce_tvals.astype(np.float32, order="C"))


def _permute_mediation(medtype, pred_x, depend_y, n, perm_number, perm_seed):
    np.random.seed(perm_number + perm_seed if perm_seed is not None else _time_seed(perm_number))
    indices_perm = np.random.permutation(list(range(n)))
    if medtype in ("M", "I"):
        return pred_x[indices_perm], depend_y
    return pred_x[indices_perm], depend_y[indices_perm]


def calculate_mediation_tfce(medtype, merge_y, masking_array, pred_x, depend_y, calcTFCE, vdensity, position_array,
                             fullmask, perm_number=None, randomise=False, verbose=False, no_intercept=True,
                             print_interation=False):
    """Non-low-RAM mmr mediation (tm_func.py:207-249)."""
    if randomise:
        pred_x, depend_y = _permute_mediation(medtype, pred_x, depend_y, merge_y.shape[0], perm_number, None)
    SobelZ = calc_sobelz(medtype, pred_x, depend_y, merge_y, merge_y.shape[0], merge_y.shape[1])
    SobelZ = SobelZ.astype(np.float32, order="C")
    tfce_SobelZ = np.zeros_like(SobelZ).astype(np.float32, order="C")
    zval_temp = np.zeros_like((fullmask)).astype(np.float32, order="C")
    zval_temp[fullmask == 1] = SobelZ
    zval_temp = zval_temp.astype(np.float32, order="C")
    tfce_temp = np.zeros_like(zval_temp).astype(np.float32, order="C")
    calcTFCE.run(zval_temp, tfce_temp)
    zval_temp = zval_temp[fullmask == 1]
    tfce_temp = tfce_temp[fullmask == 1]
    for surf_count in range(len(masking_array)):
        start = position_array[surf_count]
        end = position_array[surf_count + 1]
        dens = vdensity if isinstance(vdensity, int) else vdensity[start:end]
        tfce_SobelZ[start:end] = (tfce_temp[start:end] * (zval_temp[start:end].max() / 100) * dens)
        if randomise:
            _append("perm_maxTFCE_surf%d_%s_zstat.csv" % (surf_count, medtype), np.nanmax(tfce_SobelZ[start:end]))
        else:
            print("Max Sobel Z tfce value for surface %s:\t %1.5f" % (surf_count, np.nanmax(tfce_SobelZ[start:end])))
    if verbose:
        print("Max Zstat tfce from all surfaces = %f" % tfce_SobelZ.max())
    if randomise:
        print("Interation number: %d" % perm_number)
        return None
    return (SobelZ.astype(np.float32, order="C"), tfce_SobelZ.astype(np.float32, order="C"))


def low_ram_calculate_mediation_tfce(medtype, data, mask, pred_x, depend_y, calcTFCE, vdensity, set_surf_count=0,
                                     perm_number=None, randomise=False, no_intercept=True, output_dir=None,
                                     perm_seed=None):
    """mmr-lr mediation, one surface (tm_func.py:269-305)."""
    if randomise:
        pred_x, depend_y = _permute_mediation(medtype, pred_x, depend_y, data.shape[0], perm_number, perm_seed)
    SobelZ = calc_sobelz(medtype, pred_x, depend_y, data, data.shape[0], data.shape[1])
    SobelZ = SobelZ.astype(np.float32, order="C")
    zval = np.zeros_like((mask)).astype(np.float32, order="C")
    zval[mask == 1] = SobelZ
    zval = zval.astype(np.float32, order="C")
    tfce_zval = np.zeros_like(zval).astype(np.float32, order="C")
    calcTFCE.run(zval, tfce_zval)
    zval = zval[mask == 1]
    tfce_zval = tfce_zval[mask == 1]
    tfce_zval = (tfce_zval * (zval.max() / 100) * vdensity)
    if randomise:
        name = "perm_maxTFCE_surf%d_%s_zstat.csv" % (set_surf_count, medtype)
        permfile = "%s/%s" % (output_dir, name) if output_dir is not None else name
        _append(permfile, np.nanmax(tfce_zval))
    else:
        return (zval.astype(np.float32, order="C"), tfce_zval.astype(np.float32, order="C"))


def calc_mixed_tfce(assigntfcesettings, merge_y, masking_array, position_array, vdensity, pred_x, calcTFCE,
                    perm_number=None, randomise=False, medtype=None, depend_y=None):
    """One calculate_tfce per (H, E) group (tm_func.py:327-378).  NB: the reference
    compares a Python list with an int here, so callers must pass an ndarray for it to select anything;
    the same requirement holds here."""
    assigntfcesettings = np.asarray(assigntfcesettings)
    tvals = tfce_tvals = neg_tfce_tvals = None
    for i in np.unique(assigntfcesettings):
        data_mask = np.zeros(merge_y.shape[1], dtype=bool)
        extract_range = np.argwhere(assigntfcesettings == i)
        for surface in extract_range:
            data_mask[position_array[int(surface)]:position_array[int(surface) + 1]] = True
        subset_merge_y = merge_y[:, data_mask]
        try:
            temp_vdensity = vdensity[data_mask]
        except Exception:
            temp_vdensity = 1
        sub_masks = [m for m, a in zip(masking_array, assigntfcesettings) if a == i]
        args = (subset_merge_y, sub_masks, pred_x, calcTFCE[i], temp_vdensity, create_position_array(sub_masks),
                create_full_mask(sub_masks))
        if randomise:
            calculate_tfce(*args, set_surf_count=extract_range, perm_number=perm_number, randomise=True,
                           print_interation=(i == 0))
        else:
            t, p, q = calculate_tfce(*args, set_surf_count=extract_range)
            if tvals is None:
                tvals = np.zeros((t.shape[0], merge_y.shape[1]))
                tfce_tvals = np.zeros_like(tvals)
                neg_tfce_tvals = np.zeros_like(tvals)
            tvals[:, data_mask] = t
            tfce_tvals[:, data_mask] = p
            neg_tfce_tvals[:, data_mask] = q
    if not randomise:
        return tvals, tfce_tvals, neg_tfce_tvals


def find_nearest(array, value, p_array):
    """tm_func.py:566-573, vectorised over `value` (including the reference's idx-1 == -1 wrap-around)."""
    array = np.asarray(array)
    value = np.asarray(value, dtype=np.float64)
    n = len(p_array)
    idx = np.searchsorted(array, value, side="left")
    at_end = idx == n
    idc = np.minimum(idx, n - 1)
    lower = np.abs(value - array[idc - 1]) < np.abs(value - array[idc])   # idc - 1 == -1 wraps like Python
    p_array = np.asarray(p_array)
    return np.where(at_end, p_array[n - 1], np.where(lower, p_array[idc - 1], p_array[idc]))


def lowest_length(num_contrasts, surface_range, tmifilename, medtype=None):
    """Shortest permutation file over all surfaces/contrasts (tm_func.py:544-552)."""
    lengths = []
    for contrast in range(num_contrasts):
        for surface in surface_range:
            if medtype is not None:
                f = 'output_%s/perm_maxTFCE_surf%d_%s_zstat.csv' % (tmifilename, surface, medtype)
            else:
                f = 'output_%s/perm_maxTFCE_surf%d_tcon%d.csv' % (tmifilename, surface, contrast + 1)
            lengths.append(np.genfromtxt(f).shape[0])
    return np.array(lengths).min()


def apply_mfwer(image_array, num_contrasts, surface_range, num_perm, num_surf, tminame, position_array, pos_range,
                neg_range=None, method='scale', weight=None, mediation=False, medtype=None):
    """Study-wide (multi-surface) FWER correction, tm_func.py:403-491: per surface the null maxima are
    log-transformed and z-standardised (+10), the per-permutation maximum is taken across surfaces (optionally
    choosing the surface by log-mask-size weights), sorted, and every vertex's equally transformed TFCE value is
    looked up with the reference's nearest-value rule.  Host numpy: a post-processing step on
    num_perm x num_surf numbers plus one vectorised lookup per image column."""
    maxvalue_array = np.zeros((num_perm, num_contrasts))
    temp_max = np.zeros((num_perm, num_surf))
    positive_data = np.zeros((image_array[0].shape[0], num_contrasts))
    negative_data = None if mediation else np.zeros((image_array[0].shape[0], num_contrasts))
    if weight == 'logmasksize':
        x = [position_array[i + 1] - position_array[i] for i in range(len(position_array) - 1)]
        weights = (np.log(x) / np.log(x).sum()) / np.mean(np.log(x) / np.log(x).sum())
        w_temp_max = np.zeros((num_perm, num_surf))
    for contrast in range(num_contrasts):
        for surface in surface_range:
            if not mediation:
                f = 'output_%s/perm_maxTFCE_surf%d_tcon%d.csv' % (tminame, surface, contrast + 1)
            else:
                f = 'output_%s/perm_maxTFCE_surf%d_%s_zstat.csv' % (tminame, surface, str(medtype))
            with np.errstate(divide="ignore"):
                log_perm_results = np.log(np.genfromtxt(f)[:num_perm])
            log_perm_results[np.isinf(log_perm_results)] = 0
            start, end = position_array[surface], position_array[surface + 1]
            for data, rng in ((positive_data, pos_range),) + (() if mediation else ((negative_data, neg_range),)):
                col = image_array[0][start:end, rng[contrast]]
                with np.errstate(divide="ignore", invalid="ignore"):
                    posvmask = np.log(col) > 0
                    temp_lt = np.log(col[posvmask])
                temp_lt -= log_perm_results.mean()
                temp_lt /= log_perm_results.std()
                temp_lt += 10
                data[start:end, contrast][posvmask] = temp_lt
            log_perm_results -= log_perm_results.mean()
            log_perm_results /= log_perm_results.std()
            if weight == 'logmasksize':
                w_temp_max[:, surface] = log_perm_results * weights[surface] + 10
            log_perm_results += 10
            temp_max[:, surface] = log_perm_results
        if weight is not None:
            w_temp_max[np.isnan(w_temp_max)] = 0
            max_index = np.argmax(w_temp_max, axis=1)
            max_value_list = temp_max[np.arange(len(temp_max)), max_index].astype(np.float32)
            maxvalue_array[:, contrast] = np.sort(max_value_list)
        else:
            maxvalue_array[:, contrast] = np.sort(temp_max.max(axis=1))
    p_array = np.arange(num_perm) / float(num_perm)
    for contrast in range(num_contrasts):
        srt = maxvalue_array[:, contrast]
        positive_data[:, contrast] = find_nearest(srt, positive_data[:, contrast], p_array)
        if not mediation:
            negative_data[:, contrast] = find_nearest(srt, negative_data[:, contrast], p_array)
    if mediation:
        return positive_data
    return positive_data, negative_data
