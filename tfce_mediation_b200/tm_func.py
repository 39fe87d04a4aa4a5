"""Drop-in for the hot-path subset of the reference's ``tfce_mediation.tm_func`` (mmr / mmr-lr glue).

    create_full_mask(masking_array)                      tm_func.py:502-510
    merge_adjacency_array(adjacent_range, adjacency)     tm_func.py:521-540
    create_position_array(masking_array)                 tm_func.py:556-562
    calculate_tfce(...)                                  tm_func.py:54-123
    low_ram_calculate_tfce(...)                          tm_func.py:144-185
    calculate_mediation_tfce(...)                        tm_func.py:207-249
    low_ram_calculate_mediation_tfce(...)                tm_func.py:269-305
    calc_mixed_tfce(...)                                 tm_func.py:327-378

Same names, arguments, return values and CSV side effects ('%f' rows, positive then negative) as the reference, but every
function is an ADAPTER over the batched engine (engine.PermutationEngine): the surfaces of the merged graph behind
`calcTFCE` become the surfaces of one TFCE plan whose threshold sequence is shared by the whole group (the reference runs
ONE TFCE over the merged image, so its threshold step is the maximum over all surfaces / 100, fast_tfce.hpp:32-36) while
every surface is rescaled with its own maximum (tm_func.py:83-91).  A single call is a block of one shuffle; the engine --
data resident in HBM, graphs, plan -- is built on the first call and reused for as long as the same data and CreateAdjSet
objects are passed, which is how the reference's permutation loops call these functions
(tm_multimodality_multisurface_regression.py:540-572).  Whole permutation ranges go through
tm_multisurface.mmr_randomise / mmr_lr_randomise, which batch shuffles x surfaces and write the same rows.
"""
import weakref
from time import time

import numpy as np

from ._graph import induced_subgraph


def _time_seed(perm_number):
    # tm_func.py:57,150,209,275: perm_number + the last six CHARACTERS of str(time()) times 100
    return perm_number + int(float(str(time())[-6:]) * 100)


def _mask_piece(m):
    """Flat mask of one surface as create_full_mask sees it: a vertex image's [V, 1, 1] column, or the True voxels."""
    m = np.asarray(m)
    return m[:, 0, 0] if m.shape[2] == 1 else m[m == True]  # noqa: E712


def create_full_mask(masking_array):
    """Concatenated flat mask of all surfaces (tm_func.py:502-510); float64 like the reference's hstack from []."""
    return np.concatenate([np.zeros(0)] + [np.asarray(_mask_piece(m), dtype=np.float64) for m in masking_array])


def create_position_array(masking_array):
    """Start offset of every surface in the concatenated data (tm_func.py:556-562)."""
    sizes = [int(np.count_nonzero(np.asarray(m) == True)) for m in masking_array]  # noqa: E712
    return [0] + np.cumsum(sizes).tolist()


def merge_adjacency_array(adjacent_range, adjacency_array):
    """Block-diagonal merge of per-surface adjacency lists (tm_func.py:521-540).  Like the reference, the first block is
    always adjacency_array[0] (whatever adjacent_range[0] is) while the running offset advances by len(adjacency_array[e]);
    a single-entry range returns adjacency_array[0] alone."""
    blocks, offset = [], 0
    for pos, e in enumerate(adjacent_range):
        src = adjacency_array[0] if pos == 0 else adjacency_array[e]
        blk = np.empty(len(src), dtype=object)
        for i, nb in enumerate(src):
            blk[i] = (np.asarray(list(nb), dtype=np.int64) + offset).tolist() if len(nb) else []
        blocks.append(blk)
        offset += len(adjacency_array[e])
        if len(adjacent_range) == 1:
            break
    return np.concatenate(blocks) if len(blocks) > 1 else blocks[0]


# ---------------------------------------------------------------------------------------------- engines
_ENGINES = []       # [(weakref(data), weakref(calcTFCE), key, engine)], most recent last
_MAX_ENGINES = 4


def _cached_engine(data, calcTFCE, key, build):
    for ent in _ENGINES:
        if ent[0]() is data and ent[1]() is calcTFCE and ent[2] == key:
            return ent[3]
    eng = build()
    _ENGINES.append((weakref.ref(data), weakref.ref(calcTFCE), key, eng))
    del _ENGINES[:-_MAX_ENGINES]
    return eng


def _split_merged_graph(calcTFCE, pieces):
    """Per-surface CSR of the kept vertices out of the merged (block-diagonal) graph of a CreateAdjSet."""
    indptr, indices = calcTFCE.indptr, calcTFCE.indices
    if sum(len(p) for p in pieces) != calcTFCE.num_vertices:
        raise ValueError("the masks cover %d vertices, the adjacency %d" % (sum(len(p) for p in pieces), calcTFCE.num_vertices))
    out, lo = [], 0
    for piece in pieces:
        hi = lo + len(piece)
        ip = indptr[lo:hi + 1] - indptr[lo]
        ix = indices[indptr[lo]:indptr[hi]].astype(np.int64) - lo
        if ix.size and (ix.min() < 0 or ix.max() >= hi - lo):
            raise ValueError("the adjacency of one surface points into another: not a merge_adjacency_array graph")
        out.append(induced_subgraph(ip.astype(np.int64), ix.astype(np.int32), np.asarray(piece) == 1))
        lo = hi
    return out


def _density_slices(vdensity, position_array, nsurf):
    if np.ndim(vdensity) == 0:
        return [None if float(vdensity) == 1.0 else float(vdensity)] * nsurf
    v = np.asarray(vdensity)
    return [v[position_array[i]:position_array[i + 1]] for i in range(nsurf)]


def _mmr_engine(merge_y, masking_array, calcTFCE, vdensity, position_array, two_sided):
    """The engine behind calculate_tfce / calculate_mediation_tfce: one surface per mask, one threshold group."""
    from .engine import PermutationEngine, Surface
    from .tfce import CreateAdjSet
    dens_key = None if np.ndim(vdensity) == 0 else id(vdensity)

    def build():
        graphs = _split_merged_graph(calcTFCE, [_mask_piece(m) for m in masking_array])
        dens = _density_slices(vdensity, position_array, len(masking_array))
        surfs = [Surface(CreateAdjSet(calcTFCE.H, calcTFCE.E, g), position_array[i], dens[i]) for i, g in enumerate(graphs)]
        return PermutationEngine(merge_y, surfs, two_sided=two_sided, threshold_groups=[list(range(len(surfs)))])

    return _cached_engine(merge_y, calcTFCE, ("mmr", two_sided, dens_key, len(masking_array)), build)


def _lowram_engine(data, mask, calcTFCE, vdensity, two_sided):
    from .engine import PermutationEngine, Surface
    from .tfce import CreateAdjSet
    dens_key = None if np.ndim(vdensity) == 0 else id(vdensity)

    def build():
        g = induced_subgraph(calcTFCE.indptr, calcTFCE.indices, np.asarray(mask) == 1)
        w = None if (np.ndim(vdensity) == 0 or np.size(vdensity) == 1) and float(np.ravel(vdensity)[0]) == 1.0 else vdensity
        if w is not None and np.size(w) == 1:
            w = float(np.ravel(w)[0])
        return PermutationEngine(data, [Surface(CreateAdjSet(calcTFCE.H, calcTFCE.E, g), 0, w)], two_sided=two_sided)

    return _cached_engine(data, calcTFCE, ("lr", two_sided, dens_key), build)


def _append(path, values):
    with open(path, "a") as f:
        for v in values:
            f.write("%f\n" % v)


def _scaled_maps(eng, stat, maps):
    """tfce * (max(stat over the surface) / 100) * density per surface, as float32 (tm_func.py:83-91,173-174)."""
    out = np.zeros_like(stat, dtype=np.float32)
    for s in eng.plan.surfaces:
        a, b = s.col_offset, s.col_offset + s.adjset.num_vertices
        w = 1 if s.weight is None else (s.weight64 if s.weight64 is not None else s.weight)
        for c in range(stat.shape[0]):
            out[c, a:b] = maps[c, a:b] * (stat[c, a:b].max() / 100) * w
    return out


# ---------------------------------------------------------------------------------------------- regression
def _regression(eng, pred_x, n, seed, randomise, no_intercept):
    """(X as fitted, maxima [C, S, 2] or observed (t, pos, neg))."""
    X = np.column_stack([np.ones(n), pred_x])
    perm = np.arange(n)
    if randomise:
        np.random.seed(seed)
        perm = np.random.permutation(list(range(n)))
    if randomise:
        return eng.regression_block(X, perm_idx=perm[None, :])[0]
    mx, t32, (pos, neg) = eng.regression_block(X, perm_idx=perm[None, :], want_maps=True)
    V = eng.Y.V
    t = t32[0, :, :V].cpu().numpy()
    return t, _scaled_maps(eng, t, pos[:, :V].cpu().numpy()), _scaled_maps(eng, -t, neg[:, :V].cpu().numpy())


def calculate_tfce(merge_y, masking_array, pred_x, calcTFCE, vdensity, position_array, fullmask, perm_number=None,
                   randomise=False, verbose=False, no_intercept=True, set_surf_count=None, print_interation=False):
    """Non-low-RAM mmr (tm_func.py:54-123): ONE TFCE over the merged graph -- threshold step = global maximum / 100 --
    then a per-surface rescale and per-surface nanmax rows in perm_maxTFCE_surf{s}_tcon{c}.csv."""
    if not no_intercept:
        raise NotImplementedError("no_intercept=False: the reference's callers never pass it (the intercept row is dropped)")
    eng = _mmr_engine(merge_y, masking_array, calcTFCE, vdensity, position_array, True)
    res = _regression(eng, pred_x, merge_y.shape[0], _time_seed(perm_number) if randomise else None, randomise, no_intercept)
    labels = [int(np.ravel(s)[0]) for s in set_surf_count] if set_surf_count is not None else list(range(len(masking_array)))
    if randomise:
        for c in range(res.shape[0]):
            for s, label in enumerate(labels):
                _append("perm_maxTFCE_surf%d_tcon%d.csv" % (label, c + 1), (res[c, s, 0], res[c, s, 1]))
        if print_interation:
            print("Interation number: %d" % perm_number)
        return None
    tvals, tfce_tvals, neg_tfce_tvals = res
    for c in range(tvals.shape[0]):
        for s, label in enumerate(labels):
            a, b = position_array[s], position_array[s + 1]
            print("Maximum (untransformed) postive tfce value for surface %s, tcon %d: %f" % (label, c + 1, np.nanmax(tfce_tvals[c, a:b])))
            print("Maximum (untransformed) negative tfce value for surface %s, tcon %d: %f" % (label, c + 1, np.nanmax(neg_tfce_tvals[c, a:b])))
        if verbose:
            print("T-contrast: %d" % c)
            print("Max tfce from all surfaces = %f" % tfce_tvals[c].max())
            print("Max negative tfce from all surfaces = %f" % neg_tfce_tvals[c].max())
    return tvals, tfce_tvals, neg_tfce_tvals


def low_ram_calculate_tfce(data, mask, pred_x, calcTFCE, vdensity, set_surf_count=0, perm_number=None, randomise=False,
                           no_intercept=True, output_dir=None, perm_seed=None):
    """mmr-lr, one surface (tm_func.py:144-185); deterministic when perm_seed is given."""
    if not no_intercept:
        raise NotImplementedError("no_intercept=False: the reference's callers never pass it")
    eng = _lowram_engine(data, mask, calcTFCE, vdensity, True)
    seed = None
    if randomise:
        seed = perm_number + perm_seed if perm_seed is not None else _time_seed(perm_number)
    res = _regression(eng, pred_x, data.shape[0], seed, randomise, no_intercept)
    if not randomise:
        return res
    for c in range(res.shape[0]):
        name = "perm_maxTFCE_surf%d_tcon%d.csv" % (int(set_surf_count), c + 1)
        _append("%s/%s" % (output_dir, name) if output_dir is not None else name, (res[c, 0, 0], res[c, 0, 1]))


# ---------------------------------------------------------------------------------------------- mediation
def _mediation(eng, medtype, pred_x, depend_y, n, seed, randomise):
    perm = np.arange(n)
    if randomise:
        np.random.seed(seed)
        perm = np.random.permutation(list(range(n)))
    if randomise:
        return eng.mediation_block(medtype, pred_x, depend_y, perm[None, :])[0]
    mx, z32, (pos, _) = eng.mediation_block(medtype, pred_x, depend_y, perm[None, :], want_maps=True)
    V = eng.Y.V
    z = z32[:, :V].cpu().numpy()
    return z[0], _scaled_maps(eng, z, pos[:, :V].cpu().numpy())[0]


def calculate_mediation_tfce(medtype, merge_y, masking_array, pred_x, depend_y, calcTFCE, vdensity, position_array,
                             fullmask, perm_number=None, randomise=False, verbose=False, no_intercept=True,
                             print_interation=False):
    """Non-low-RAM mmr mediation (tm_func.py:207-249): Sobel z, one-sided TFCE over the merged graph."""
    eng = _mmr_engine(merge_y, masking_array, calcTFCE, vdensity, position_array, False)
    res = _mediation(eng, medtype, pred_x, depend_y, merge_y.shape[0], _time_seed(perm_number) if randomise else None, randomise)
    if randomise:
        for s in range(len(masking_array)):
            _append("perm_maxTFCE_surf%d_%s_zstat.csv" % (s, medtype), (res[s],))
        print("Interation number: %d" % perm_number)
        return None
    SobelZ, tfce_SobelZ = res
    for s in range(len(masking_array)):
        a, b = position_array[s], position_array[s + 1]
        print("Max Sobel Z tfce value for surface %s:\t %1.5f" % (s, np.nanmax(tfce_SobelZ[a:b])))
    if verbose:
        print("Max Zstat tfce from all surfaces = %f" % tfce_SobelZ.max())
    return SobelZ, tfce_SobelZ


def low_ram_calculate_mediation_tfce(medtype, data, mask, pred_x, depend_y, calcTFCE, vdensity, set_surf_count=0,
                                     perm_number=None, randomise=False, no_intercept=True, output_dir=None,
                                     perm_seed=None):
    """mmr-lr mediation, one surface (tm_func.py:269-305)."""
    eng = _lowram_engine(data, mask, calcTFCE, vdensity, False)
    seed = None
    if randomise:
        seed = perm_number + perm_seed if perm_seed is not None else _time_seed(perm_number)
    res = _mediation(eng, medtype, pred_x, depend_y, data.shape[0], seed, randomise)
    if not randomise:
        return res
    name = "perm_maxTFCE_surf%d_%s_zstat.csv" % (set_surf_count, medtype)
    _append("%s/%s" % (output_dir, name) if output_dir is not None else name, (res[0],))


# ---------------------------------------------------------------------------------------------- mixed settings
def calc_mixed_tfce(assigntfcesettings, merge_y, masking_array, position_array, vdensity, pred_x, calcTFCE,
                    perm_number=None, randomise=False, medtype=None, depend_y=None):
    """One calculate_tfce per (H, E) group (tm_func.py:327-378): the surfaces of a group form their own merged graph
    (calcTFCE[i]) and hence their own threshold group.  Two notes on the reference: it compares `assigntfcesettings == i`,
    which selects nothing for a Python list -- lists are accepted here; and its three return values
    are one aliased array (`tvals = tfce_tvals = neg_tfce_tvals = np.zeros(...)`, :368), so it really returns the
    negative TFCE values three times -- here the three arrays are what their names say."""
    assign = np.asarray(assigntfcesettings)
    tvals = pos = neg = None
    for i in np.unique(assign):
        members = np.flatnonzero(assign == i)
        cols = np.concatenate([np.arange(position_array[s], position_array[s + 1]) for s in members])
        key = ("mixed", int(i), id(merge_y))
        sub = _SUBSETS.get(key)
        if sub is None or sub[0]() is not merge_y:
            arr = np.ascontiguousarray(merge_y[:, cols])
            dens = vdensity if np.ndim(vdensity) == 0 else np.ascontiguousarray(np.asarray(vdensity)[cols])
            sub = (weakref.ref(merge_y), arr, dens)
            _SUBSETS[key] = sub
        sub_masks = [masking_array[s] for s in members]
        args = (sub[1], sub_masks, pred_x, calcTFCE[int(i)], sub[2], create_position_array(sub_masks),
                create_full_mask(sub_masks))
        if randomise:
            calculate_tfce(*args, set_surf_count=members, perm_number=perm_number, randomise=True,
                           print_interation=(i == 0))
        else:
            t, p, q = calculate_tfce(*args, set_surf_count=members)
            if tvals is None:
                tvals = np.zeros((t.shape[0], merge_y.shape[1]))
                pos, neg = np.zeros_like(tvals), np.zeros_like(tvals)
            tvals[:, cols], pos[:, cols], neg[:, cols] = t, p, q
    if not randomise:
        return tvals, pos, neg


_SUBSETS = {}       # column subsets of calc_mixed_tfce, kept so that the engine cache recognises them on the next call


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


def _perm_file(tminame, surface, contrast, medtype):
    if medtype is not None:
        return 'output_%s/perm_maxTFCE_surf%d_%s_zstat.csv' % (tminame, surface, str(medtype))
    return 'output_%s/perm_maxTFCE_surf%d_tcon%d.csv' % (tminame, surface, contrast + 1)


def lowest_length(num_contrasts, surface_range, tmifilename, medtype=None):
    """Shortest permutation file over all surfaces/contrasts (tm_func.py:544-552)."""
    return min(np.genfromtxt(_perm_file(tmifilename, sf, c, medtype)).shape[0]
               for c in range(num_contrasts) for sf in surface_range)


def apply_mfwer(image_array, num_contrasts, surface_range, num_perm, num_surf, tminame, position_array, pos_range,
                neg_range=None, method='scale', weight=None, mediation=False, medtype=None):
    """Study-wide (multi-surface) FWER correction, tm_func.py:403-491.  Per contrast and surface the null maxima are
    log-transformed and z-standardised (+10); row i of every surface file is taken as the SAME permutation and the
    per-permutation maximum across surfaces -- with weight='logmasksize' the value of the surface that wins after its
    z-scores were scaled by the normalised log mask size -- forms the null distribution; every vertex's TFCE value gets
    the same transform with ITS surface's null mean / s.d. and is looked up with the reference's nearest-value rule
    (find_nearest).  Host numpy: num_perm x num_surf numbers plus one vectorised lookup per image column."""
    img = image_array[0]
    nvert = img.shape[0]
    columns = [(pos_range, np.zeros((nvert, num_contrasts)))]
    if not mediation:
        columns.append((neg_range, np.zeros((nvert, num_contrasts))))
    by_size = None
    if weight == 'logmasksize':
        logsize = np.log(np.diff(np.asarray(position_array)))
        share = logsize / logsize.sum()
        by_size = share / share.mean()
    null_sorted = np.zeros((num_perm, num_contrasts))
    for c in range(num_contrasts):
        z10 = np.zeros((num_perm, num_surf))
        z10_weighted = np.zeros((num_perm, num_surf))
        for sf in surface_range:
            with np.errstate(divide="ignore"):
                lognull = np.log(np.genfromtxt(_perm_file(tminame, sf, c, medtype if mediation else None))[:num_perm])
            lognull[np.isinf(lognull)] = 0
            mu, sd = lognull.mean(), lognull.std()
            lo, hi = position_array[sf], position_array[sf + 1]
            for rng, out in columns:
                vals = img[lo:hi, rng[c]]
                with np.errstate(divide="ignore", invalid="ignore"):
                    logv = np.log(vals)
                keep = logv > 0
                out[lo:hi, c][keep] = (logv[keep] - mu) / sd + 10
            z = (lognull - mu) / sd
            if by_size is not None:
                z10_weighted[:, sf] = z * by_size[sf] + 10
            z10[:, sf] = z + 10
        if weight is not None:
            z10_weighted[np.isnan(z10_weighted)] = 0
            winner = np.argmax(z10_weighted, axis=1)
            null_sorted[:, c] = np.sort(z10[np.arange(num_perm), winner].astype(np.float32))
        else:
            null_sorted[:, c] = np.sort(z10.max(axis=1))
    p_array = np.arange(num_perm) / float(num_perm)
    for c in range(num_contrasts):
        for _, out in columns:
            out[:, c] = find_nearest(null_sorted[:, c], out[:, c], p_array)
    return columns[0][1] if mediation else (columns[0][1], columns[1][1])
