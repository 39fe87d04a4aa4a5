"""Batched permutation engine: the B200-first form of the reference's per-shuffle loop.

The reference runs, per shuffle (vertex_tfce_multiple_regression_randomise.py:90-117,
voxel_...:91-119, tm_func.py:144-185):
    permute design -> tval_int -> for each contrast and sign: scatter, TFCE per surface, scaled max
Here a block of P shuffles is one fused fit launch (tmb_glm_tstat, t-maps [P, C, V] straight into
HBM, betas never written) followed by one TFCE launch over P*C*S work items (tmb_plan_run), with the
subject data resident in HBM for the whole job.  Permutation index rows come from the caller (the
drivers generate them with the reference's own numpy RNG calls), so results are reproducible.
"""
import ctypes

import os as _os
import time as _time

import numpy as np

from . import _lib, parallel
from ._device import DeviceMatrix, TILE_M, round_up, to_host
from .tfce import CreateAdjSet

_E2E_DEBUG = bool(_os.environ.get("TMB_E2E_DEBUG"))   # per-block host timings of the exact-libm round trip on stderr


MAX_REGRESSORS = 64       # non-intercept regressors per design (more than 8 take the stored-beta path, tmb_glm_*_beta)


def _row_stride(stat):
    """Leading dimension of a [B, ld] statistic tensor; a one-row tensor may report any stride for its first axis."""
    return int(stat.stride(0)) if stat.shape[0] > 1 else max(int(stat.stride(0)), int(stat.shape[1]))


class Surface(object):
    """One TFCE graph laid on columns [col_offset, col_offset + V) of a statistic row."""

    def __init__(self, adjset, col_offset, weight=None):
        if not isinstance(adjset, CreateAdjSet):
            raise TypeError("adjset must be a tfce_mediation_b200.tfce.CreateAdjSet")
        self.adjset = adjset
        self.col_offset = int(col_offset)
        if weight is not None and np.ndim(weight) == 0:
            weight = None if float(weight) == 1.0 else np.full(adjset.num_vertices, weight, dtype=np.float32)
        self.weight64 = None
        if weight is not None:
            w = np.asarray(weight)
            if w.shape != (adjset.num_vertices,):
                raise ValueError("weight must have one entry per vertex")
            if w.dtype == np.float64 and not np.array_equal(w.astype(np.float32).astype(np.float64), w):
                # float64 weights that float32 cannot hold: the reference then multiplies in double
                # (non-low-RAM mmr, tm_func.py:83-91); exactly representable ones give the same product either way
                self.weight64 = np.ascontiguousarray(w, dtype=np.float64)
            weight = np.ascontiguousarray(w, dtype=np.float32)
        self.weight = weight


class TfcePlan(object):
    """tmb_plan handle: S surfaces along one statistic row."""

    def __init__(self, surfaces, device=None, max_slots=0, threshold_groups=None):
        """threshold_groups: optional list of surface-index lists whose members share ONE threshold sequence, built from
        the largest statistic over the whole group -- the non-low-RAM mmr path runs a single TFCE over the merged graph of
        all surfaces (tm_func.py:77-78, so fast_tfce.hpp:32-36 sees the global maximum) and then rescales every surface
        with its own max/100 (tm_func.py:83-91)."""
        import torch
        _lib.require_device()
        self.surfaces = list(surfaces)
        S = len(self.surfaces)
        if S == 0:
            raise ValueError("no surfaces")
        self.groups = None
        if threshold_groups is not None:
            gid = -np.ones(S, dtype=np.int64)
            for gi, members in enumerate(threshold_groups):
                gid[np.asarray(list(members), dtype=np.int64)] = gi
            if (gid < 0).any():
                raise ValueError("threshold_groups must cover every surface")
            self.groups = gid
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        parallel.check_device(self.device.index)
        graphs = (ctypes.c_void_p * S)(*[s.adjset._handle for s in self.surfaces])
        offs = (ctypes.c_int64 * S)(*[s.col_offset for s in self.surfaces])
        wts = (ctypes.c_void_p * S)(*[(s.weight.ctypes.data if s.weight is not None else None) for s in self.surfaces])
        w64 = (ctypes.c_void_p * S)(*[(s.weight64.ctypes.data if getattr(s, "weight64", None) is not None else None)
                                      for s in self.surfaces])
        h = ctypes.c_void_p()
        _lib.check(_lib.lib().tmb_plan_create(self.device.index or 0, S, graphs, offs, wts, w64, int(max_slots),
                                              ctypes.byref(h)))
        self._handle = h
        self.S = S
        self.row_len = max(s.col_offset + s.adjset.num_vertices for s in self.surfaces)
        self._H = np.array([s.adjset.H for s in self.surfaces], dtype=np.float32)
        self.internal_order = False
        self.exact_pow = True
        self._tickets = {}
        self._flip = 0

    def __del__(self):
        h = getattr(self, "_handle", None)
        if h is not None and _lib._lib is not None:
            _lib._lib.tmb_plan_destroy(h)
            self._handle = None

    def column_permutation(self, row_len):
        """int64 [row_len]: position j of an internal-order row holds caller column perm[j] (identity outside the
        surfaces).  Used by the engine to permute the data columns once, so that t-maps come out of the fit in
        the graphs' internal (locality) order and the sweep reads them coalesced."""
        perm = np.arange(row_len, dtype=np.int64)
        for s in self.surfaces:
            V = s.adjset.num_vertices
            vm = np.empty(V, dtype=np.int32)
            _lib.check(_lib.lib().tmb_graph_vmap(s.adjset._handle, _lib.ptr(vm)))
            perm[s.col_offset:s.col_offset + V] = s.col_offset + vm.astype(np.int64)
        return perm

    def set_internal_order(self, on=True):
        _lib.check(_lib.lib().tmb_plan_set_internal_order(self._handle, 1 if on else 0))
        self.internal_order = bool(on)

    def run(self, stat, two_sided=True, want_maps=False, out_max=None, exact_pow=None):
        """stat: CUDA float32 [B, ld].  Returns (max [B, S, 2] CUDA float32, status [B,S,2] int32, maps).

        exact_pow (default: self.exact_pow = True): build the threshold/height tables on the host with the C
        library's powf -- the very call the reference makes (fast_tfce.hpp:70) -- so TFCE values are
        bit-identical to the reference; costs one tiny maxima kernel + a host round trip per block.
        exact_pow=False computes correctly rounded tables on the device (no host sync); values then differ
        from the reference by <= 1-2 ulp on the ~6% of maps where libm's powf(T, 2) != T*T."""
        import torch
        if not (stat.is_cuda and stat.dtype == torch.float32 and stat.dim() == 2 and stat.stride(1) == 1):
            raise ValueError("stat must be a CUDA float32 [B, ld] tensor with unit column stride")
        if exact_pow is None:
            exact_pow = self.exact_pow
        if self.groups is not None:
            exact_pow = True              # group thresholds are built on the host
        B, ld = int(stat.shape[0]), _row_stride(stat)
        mx = out_max if out_max is not None else torch.empty((B, self.S, 2), dtype=torch.float32, device=stat.device)
        status = torch.empty((B, self.S, 2), dtype=torch.int32, device=stat.device)
        pos = neg = None
        if want_maps:
            pos = torch.zeros((B, ld), dtype=torch.float32, device=stat.device)
            neg = torch.zeros((B, ld), dtype=torch.float32, device=stat.device) if two_sided else None
        L = _lib.lib()
        stream = _lib.current_stream()
        if not exact_pow:
            _lib.check(L.tmb_plan_run(self._handle, _lib.ptr(stat), ld, B, 1 if two_sided else 0, _lib.ptr(mx),
                                      _lib.ptr(pos), _lib.ptr(neg), _lib.ptr(status), stream))
            return mx, status, (pos, neg)
        ticket = self.prepare(stat)
        return self.finish(ticket, stat, two_sided=two_sided, out_max=mx, pos=pos, neg=neg, status=status)

    # -- exact-libm mode in two stages, so callers can overlap the host part with other GPU work ----------
    def prepare(self, stat):
        """Stage 1: per-(row, surface, sign) maxima on the device and an asynchronous copy to pinned host
        memory.  Returns a ticket for finish().  Enqueue further GPU work (e.g. the fit of the next block)
        before calling finish() and the host table building overlaps it."""
        import torch
        B, ld = int(stat.shape[0]), _row_stride(stat)
        cnt = B * self.S * 2
        slot = self._tickets.setdefault((cnt, self._flip), {})
        self._flip ^= 1
        if not slot:
            slot["dev"] = torch.empty((cnt,), dtype=torch.float32, device=stat.device)
            slot["host"] = torch.empty((cnt,), dtype=torch.float32).pin_memory()
            slot["event"] = torch.cuda.Event()
            slot["scale"] = torch.empty((cnt,), dtype=torch.float32).pin_memory()
            slot["tab"] = dict(ns=torch.empty((cnt,), dtype=torch.int32).pin_memory(),
                               delta=torch.empty((cnt,), dtype=torch.float32).pin_memory(),
                               T=torch.empty((cnt, 128), dtype=torch.float32).pin_memory(),
                               HH=torch.empty((cnt, 128), dtype=torch.float32).pin_memory(),
                               st=torch.empty((cnt,), dtype=torch.int32).pin_memory())
            slot["tab_event"] = None
            slot["Hs"] = np.ascontiguousarray(np.tile(np.repeat(self._H, 2), B), dtype=np.float32)
        _lib.check(_lib.lib().tmb_plan_maxima(self._handle, _lib.ptr(stat), ld, B, _lib.ptr(slot["dev"]),
                                              _lib.current_stream()))
        slot["host"].copy_(slot["dev"], non_blocking=True)
        slot["event"].record()
        slot["cnt"] = cnt
        return slot

    def finish(self, ticket, stat, two_sided=True, out_max=None, pos=None, neg=None, status=None):
        """Stage 2: wait for the maxima, build the threshold tables on the host with libm's powf
        (tmb_threshold_tables), upload them and launch the sweep."""
        import torch
        L = _lib.lib()
        B, ld = int(stat.shape[0]), _row_stride(stat)
        cnt = ticket["cnt"]
        mx = out_max if out_max is not None else torch.empty((B, self.S, 2), dtype=torch.float32, device=stat.device)
        if status is None:
            status = torch.empty((B, self.S, 2), dtype=torch.int32, device=stat.device)
        dbg = _E2E_DEBUG
        t0 = _time.perf_counter() if dbg else 0.0
        ticket["event"].synchronize()
        if ticket["tab_event"] is not None:
            ticket["tab_event"].synchronize()     # the previous upload out of this staging buffer is done
        t1 = _time.perf_counter() if dbg else 0.0
        tab = ticket["tab"]
        mh = ticket["host"].numpy()
        scale_d = None
        if self.groups is not None:
            # thresholds from the group's maximum, scale factor from the surface's own: fl32(max / 100) as numpy forms
            # `tval_temp[start:end].max() / 100` on float32 (tm_func.py:87-91)
            m3 = mh.reshape(B, self.S, 2)
            with np.errstate(invalid="ignore"):
                ticket["scale"].numpy()[...] = (m3 / np.float32(100)).reshape(-1)
            gm = np.empty_like(m3)
            for gi in np.unique(self.groups):
                sel = self.groups == gi
                gm[:, sel, :] = np.fmax.reduce(m3[:, sel, :], axis=1, keepdims=True)
            mh = np.ascontiguousarray(gm.reshape(-1))
            scale_d = ticket["scale"].to(stat.device, non_blocking=True)
        _lib.check(L.tmb_threshold_tables(_lib.ptr(mh), _lib.ptr(ticket["Hs"]), cnt, _lib.ptr(tab["ns"]),
                                          _lib.ptr(tab["delta"]), _lib.ptr(tab["T"]), _lib.ptr(tab["HH"]),
                                          _lib.ptr(tab["st"])))
        t2 = _time.perf_counter() if dbg else 0.0
        d = {k: v.to(stat.device, non_blocking=True) for k, v in tab.items()}
        if dbg:
            import sys
            sys.stderr.write("[tmb e2e] wait maxima %.2f ms, host tables %.2f ms, table upload issue %.2f ms\n"
                             % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, (_time.perf_counter() - t2) * 1e3))
        ticket["tab_event"] = torch.cuda.Event()
        ticket["tab_event"].record()
        _lib.check(L.tmb_plan_run_tables(self._handle, _lib.ptr(stat), ld, B, 1 if two_sided else 0, _lib.ptr(d["ns"]),
                                         _lib.ptr(d["delta"]), _lib.ptr(d["T"]), _lib.ptr(d["HH"]), _lib.ptr(d["st"]),
                                         _lib.ptr(scale_d), _lib.ptr(mx), _lib.ptr(pos), _lib.ptr(neg), _lib.ptr(status),
                                         _lib.current_stream()))
        ticket["inflight"] = (d, scale_d)   # keep the device tables alive until this ticket is reused
        return mx, status, (pos, neg)


# ------------------------------------------------------------------------------------------ designs

def has_intercept(X):
    X = np.asarray(X)
    return X.ndim >= 2 and X.shape[-1] >= 2 and bool(np.all(X[..., 0] == 1))


def design_stack(Xs, center=True):
    """Per-design normal-equation algebra on the host (k x k work only).

    Xs: float64 [P, n, k] permuted designs (column 0 = intercept when center=True).
    Returns dict(pinv [P, r, n], G [P, r, r], d [P, r], r, dof) where, with an intercept, the
    regressors are mean-centred (Frisch-Waugh-Lovell: slopes, residuals and diag(inv(X'X)) of
    the slopes are unchanged) and r = k - 1."""
    Xs = np.asarray(Xs, dtype=np.float64)
    P, n, k = Xs.shape
    if center:
        Z = Xs[:, :, 1:] - Xs[:, :, 1:].mean(axis=1, keepdims=True)
    else:
        Z = Xs
    G = np.einsum("pni,pnj->pij", Z, Z)
    Ginv = np.linalg.inv(G)
    pinv = np.einsum("pij,pnj->pin", Ginv, Z)
    d = np.einsum("pii->pi", Ginv).copy()
    return dict(pinv=pinv, G=np.ascontiguousarray(G), d=np.ascontiguousarray(d), r=Z.shape[2], dof=float(n - k),
                centered=bool(center))


def fstat_blocks(G, var_lo, var_k):
    """Per design the matrices inv(C[S_i, S_i]), C = inv(G), of every tested variable i (regressor rows
    [var_lo[i], var_lo[i] + var_k[i])), one after the other: float64 [P, sum k_i^2].  With them the extra sum of
    squares of dropping variable i is b_S' inv(C_SS) b_S (no second fit), which is what pyfunc.py:2346-2351 obtains
    from the residual SS of the reduced design."""
    G = np.asarray(G, dtype=np.float64)
    C = np.linalg.inv(G)
    parts = []
    for lo, k in zip(var_lo, var_k):
        lo, k = int(lo), int(k)
        parts.append(np.linalg.inv(C[:, lo:lo + k, lo:lo + k]).reshape(G.shape[0], k * k))
    return np.ascontiguousarray(np.concatenate(parts, axis=1)) if parts else np.zeros((G.shape[0], 0))


def row_permuted_stack(X, perm_idx, center=True):
    """Same as design_stack for designs X[perm_idx[p]] (whole rows permuted): X'X is invariant, so
    one pseudo-inverse is formed and its columns are gathered."""
    X = np.asarray(X, dtype=np.float64)
    perm_idx = np.asarray(perm_idx)
    P = perm_idx.shape[0]
    base = design_stack(X[None], center=center)
    pinv = base["pinv"][0][:, perm_idx]                    # [r, P, n]
    r = base["r"]
    return dict(pinv=np.ascontiguousarray(pinv.transpose(1, 0, 2)), G=np.repeat(base["G"], P, axis=0),
                d=np.repeat(base["d"], P, axis=0), r=r, dof=base["dof"], centered=bool(center))


def cosinor_design(time_var, period, exog=None, dmy_covariates=None):
    """The cosinor design of pyfunc.py:2432-2461: [1, cos(2 pi t / T_i), sin(2 pi t / T_i) per period, tested
    variables..., covariates].  Returns (float64 [n, k], number of periods, number of tested columns)."""
    time_var = np.asarray(time_var, dtype=np.float64).reshape(-1)
    n = time_var.shape[0]
    cols = [np.ones(n)]
    for T in period:
        angle = np.divide(2.0 * np.pi * time_var, T)
        cols += [np.cos(angle), np.sin(angle)]
    nexog = 0
    for var in (exog if exog is not None else []):
        var = np.asarray(var, dtype=np.float64).reshape(n, -1)
        nexog += var.shape[1]
        cols.append(var)
    if dmy_covariates is not None:
        cols.append(np.asarray(dmy_covariates, dtype=np.float64).reshape(n, -1))
    return np.column_stack(cols), len(period), nexog


def cosinor_amplitude_t(y, time_var, period):
    """|t| of the amplitude(s) of ONE variable y [n] under the plain cosinor model (no tested variables, no covariates):
    the un-permuted path A of the cosinor mediation (tm_models_randomise.py:398-403; pyfunc.py:2530-2551).  Host
    algebra, k x k."""
    X, nper, _ = cosinor_design(time_var, period)
    y = np.asarray(y, dtype=np.float64).reshape(-1)
    n, k = X.shape
    invXX = np.linalg.inv(X.T @ X)
    a = (invXX @ X.T) @ y
    sigma = np.sqrt(np.sum((y - X @ a) ** 2) / (n - k))
    out = np.empty(nper)
    for i in range(nper):
        c, s = 1 + 2 * i, 2 + 2 * i
        amp = np.sqrt(a[c] ** 2 + a[s] ** 2)
        acr = np.arctan(np.abs(-a[s] / a[c]))
        var = invXX[c, c] * np.cos(acr) ** 2 - 2 * invXX[c, s] * np.sin(acr) * np.cos(acr) + invXX[s, s] * np.sin(acr) ** 2
        out[i] = np.abs(amp / (sigma * np.sqrt(var)))
    return out


def linregress_t(x, y):
    """slope / stderr of scipy.stats.linregress(x[p], y[p]) for every row p at once (the scalar path A of medtype 'Y',
    pyfunc.py:142): scipy's own formulas -- biased covariances, r, slope = ssxym / ssxm,
    stderr = sqrt((1 - r^2) ssym / ssxm / (n - 2)) -- vectorised over the shuffles (a per-shuffle scipy call costs
    ~0.1 ms, more than the GPU needs for the whole shuffle); agrees with the per-call result to a few ulp."""
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    n = x.shape[1]
    xm = x - x.mean(axis=1, keepdims=True)
    ym = y - y.mean(axis=1, keepdims=True)
    ssxm, ssym, ssxym = (xm * xm).sum(axis=1) / n, (ym * ym).sum(axis=1) / n, (xm * ym).sum(axis=1) / n
    r = np.clip(ssxym / np.sqrt(ssxm * ssym), -1.0, 1.0)
    slope = ssxym / ssxm
    stderr = np.sqrt((1 - r ** 2) * ssym / ssxm / (n - 2))
    return slope / stderr


def pack_At(pinv, rp, layout=0, ldA=None):
    """[P, r, n] pseudo-inverse rows -> At float64 [n, ldA]; row i of design p goes to column p*rp + i (layout 0, the
    fp64 vector kernel) or (p // 8)*8*rp + i*8 + p % 8 (layout 1, "tile8", the tensor-core kernels: include/tfce_b200.h).
    ldA: columns the kernel will read (tmb_glm_packed_columns); default: the packed columns rounded up to 128."""
    P, r, n = pinv.shape
    if ldA is None:
        ldA = round_up(P * rp if layout == 0 else (P + 7) // 8 * 8 * rp, TILE_M)
    At = np.zeros((n, ldA), dtype=np.float64)
    if layout == 0:
        At[:, :P * rp].reshape(n, P, rp)[:, :, :r] = pinv.transpose(2, 0, 1)
    else:
        P8 = (P + 7) // 8 * 8
        blk = np.zeros((n, P8, rp), dtype=np.float64)
        blk[:, :P, :r] = pinv.transpose(2, 0, 1)
        # [n, P8/8, 8 (q), rp (i)] -> [n, P8/8, rp (i), 8 (q)]
        At[:, :P8 * rp] = blk.reshape(n, P8 // 8, 8, rp).transpose(0, 1, 3, 2).reshape(n, P8 * rp)
    return At, ldA


class PermutationEngine(object):
    """Data resident in HBM + a TFCE plan; runs blocks of shuffles."""

    def _layout(self, rp):
        """Column order of the stacked pseudo-inverses the fit kernels expect for this data type (tmb_glm_layout)."""
        return int(_lib.lib().tmb_glm_layout(self.Y.dtype_code, int(rp)))

    def _rp(self, r):
        """Padded regressors per design for the fused kernels (tmb_glm_rp); 0: more than 8 -> stored-beta path."""
        if r > MAX_REGRESSORS:
            raise ValueError("at most %d non-intercept regressors per design (got %d)" % (MAX_REGRESSORS, r))
        return int(_lib.lib().tmb_glm_rp(self.Y.dtype_code, int(r)))

    def _pack(self, pinv, rp):
        """(At device tensor, ldA, layout) for a stack of pseudo-inverse rows [P, r, n]."""
        layout = self._layout(rp)
        ldA = int(_lib.lib().tmb_glm_packed_columns(self.Y.dtype_code, int(pinv.shape[0]), int(rp)))
        At, ldA = pack_At(pinv, rp, layout, ldA)
        return At, ldA, layout

    def _betas_chunks(self, pinv, budget=1.5e9):
        """Stored-beta path (r > 8): yields (first design, designs, beta64 [designs * r, ld]) chunk by chunk; every
        pseudo-inverse row is one column of the left operand and one output row of the plain contraction."""
        import torch
        P, r, n = pinv.shape
        per = max(1, int(budget // (r * self.Y.ld * 8)))
        for a in range(0, P, per):
            b = min(P, a + per)
            rows = (b - a) * r
            ldA = round_up(rows, TILE_M)
            At = np.zeros((n, ldA), dtype=np.float64)
            At[:, :rows] = pinv[a:b].reshape(rows, n).T
            At_d = self._upload("beta_At", At)
            beta = torch.empty((rows, self.Y.ld), dtype=torch.float64, device=self.device)
            _lib.check(_lib.lib().tmb_glm_beta(_lib.ptr(self.Y.t), self.Y.dtype_code, self.Y.n, self.Y.V, self.Y.ld,
                                               _lib.ptr(At_d), ldA, rows, _lib.ptr(beta), self.Y.ld, _lib.current_stream()))
            yield a, b - a, beta

    def __init__(self, data, surfaces, two_sided=True, nan_to_zero=False, device=None, max_slots=0,
                 permute_columns=True, threshold_groups=None):
        import torch
        _lib.require_device()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        parallel.check_device(self.device.index)
        owned = not isinstance(data, DeviceMatrix)
        self.Y = DeviceMatrix(data, device=self.device) if owned else data
        self.plan = (TfcePlan(surfaces, device=self.device, max_slots=max_slots, threshold_groups=threshold_groups)
                     if surfaces is not None else None)
        if self.plan is not None and self.plan.row_len > self.Y.V:
            raise ValueError("surfaces cover %d columns but the data has %d" % (self.plan.row_len, self.Y.V))
        self.colperm = None
        if self.plan is not None and permute_columns:
            perm = self.plan.column_permutation(self.Y.V)
            if not np.array_equal(perm, np.arange(self.Y.V)):
                # one-time gather of the data columns into the graphs' internal vertex order
                self.colperm = torch.from_numpy(perm).to(self.device)
                if not owned:
                    # a caller-supplied DeviceMatrix stays in the caller's column order (it may feed another engine
                    # or be read afterwards): the engine works on its own permuted copy
                    self.Y = self.Y.permuted_copy(self.colperm)
                else:
                    self.Y.t[:, :self.Y.V] = self.Y.t[:, :self.Y.V].index_select(1, self.colperm)
                    self.Y._yy = {}
                self.plan.set_internal_order(True)
        self.two_sided = bool(two_sided)
        self.nan_to_zero = bool(nan_to_zero)
        self.h2d_bytes = 0
        self.d2h_bytes = 0
        self._pinned = {}
        self._rings = {}

    def to_caller_order(self, t):
        """Undo the one-time column permutation on a [..., ld] device tensor (maps returned to the user)."""
        if self.colperm is None:
            return t
        out = t.clone()
        out[..., :self.Y.V].index_copy_(t.dim() - 1, self.colperm, t[..., :self.Y.V])
        return out

    # -- staging ---------------------------------------------------------------------------------
    def _upload(self, name, host_array):
        """Host -> device through a reused pinned staging buffer; bytes are counted for bench e2e."""
        import torch
        a = np.ascontiguousarray(host_array)
        key = (name, a.shape, a.dtype.str)
        ent = self._pinned.get(key)
        if ent is None:
            ent = [torch.empty(a.shape, dtype=torch.from_numpy(a).dtype).pin_memory(), None]
            self._pinned[key] = ent
        pin, ev = ent
        if ev is not None:
            ev.synchronize()          # the previous async copy out of this staging buffer has finished
        pin.numpy()[...] = a
        self.h2d_bytes += a.nbytes
        dev = pin.to(self.device, non_blocking=True)
        ent[1] = torch.cuda.Event()
        ent[1].record()
        return dev

    def _ring(self, name, shape, dtype, depth=3):
        """Reused device buffers for the per-block operands and statistic maps (a ring of `depth` per shape): the
        pipelined block loop keeps two blocks in flight on one stream, so a buffer comes round again only after the
        kernels that read it.  Without it every block took ~1.2 GB from torch's caching allocator, whose occasional
        cudaMalloc/cudaFree (synchronous) made one bench run in four 30% slower end to end."""
        import torch
        key = (name, tuple(int(x) for x in shape), dtype)
        ent = self._rings.get(key)
        if ent is None:
            if len(self._rings) > 12:                        # shapes changed (another block size): drop the old rings
                self._rings.clear()
            ent = [[torch.empty(key[1], dtype=dtype, device=self.device) for _ in range(depth)], 0]
            self._rings[key] = ent
        ent[1] = (ent[1] + 1) % depth
        return ent[0][ent[1]]

    def _download(self, t):
        self.d2h_bytes += t.numel() * t.element_size()
        return to_host(t)

    # -- fit -------------------------------------------------------------------------------------
    def tstat(self, stack, rows=None, want_f64=False, caller_order=True):
        """Fused fit+t for a design stack (design_stack / row_permuted_stack output).
        Returns CUDA float32 [P, C, ld] (and float64 when want_f64)."""
        import torch
        P, r, n = stack["pinv"].shape
        if n != self.Y.n:
            raise ValueError("design has %d subjects, data has %d" % (n, self.Y.n))
        rp = self._rp(r)
        row0, nrows = (0, r) if rows is None else (int(rows[0]), int(rows[1]))
        G_d = self._upload("G", stack["G"])
        d_d = self._upload("d", stack["d"])
        centered = stack.get("centered", True)
        yy = self.Y.sumsq(centered)
        t32 = torch.empty((P, nrows, self.Y.ld), dtype=torch.float32, device=self.device)
        t64 = torch.empty((P, nrows, self.Y.ld), dtype=torch.float64, device=self.device) if want_f64 else None
        if rp == 0:
            for a, cnt, beta in self._betas_chunks(stack["pinv"]):
                _lib.check(_lib.lib().tmb_glm_tstat_beta(
                    _lib.ptr(beta), self.Y.ld, self.Y.V, _lib.ptr(G_d[a:a + cnt]), _lib.ptr(d_d[a:a + cnt]), cnt, r, row0,
                    nrows, stack["dof"], _lib.ptr(yy), _lib.ptr(t32[a:a + cnt]),
                    _lib.ptr(t64[a:a + cnt]) if t64 is not None else None, self.Y.ld, 1 if self.nan_to_zero else 0,
                    _lib.current_stream()))
        else:
            At, ldA, layout = self._pack(stack["pinv"], rp)
            At_d = self._upload("At", At)
            _lib.check(_lib.lib().tmb_glm_tstat(
                _lib.ptr(self.Y.t), self.Y.dtype_code, self.Y.n, self.Y.V, self.Y.ld, _lib.ptr(At_d), ldA,
                _lib.ptr(G_d), _lib.ptr(d_d), P, r, rp, row0, nrows, stack["dof"], _lib.ptr(yy), _lib.ptr(t32),
                _lib.ptr(t64), self.Y.ld, 1 if self.nan_to_zero else 0, layout, _lib.current_stream()))
        if caller_order and self.colperm is not None:
            t32 = self.to_caller_order(t32)
            t64 = self.to_caller_order(t64) if t64 is not None else None
        return (t32, t64) if want_f64 else t32

    def _rowperm_operands(self, X, perm_idx):
        """Device operands for the designs X[perm_idx[p]] (whole rows permuted): the host sends only the index rows;
        the stacked pseudo-inverses are gathered on the device (tmb_glm_pack_rowperm).  X'X and its inverse are
        invariant under row permutations, so G and diag(inv(X'X)) are one row repeated P times."""
        import torch
        X = np.ascontiguousarray(X, dtype=np.float64)
        perm_idx = np.asarray(perm_idx)
        P, n = perm_idx.shape
        if n != self.Y.n or X.shape[0] != n:
            raise ValueError("design has %d subjects, data has %d" % (n, self.Y.n))
        key = X.tobytes()
        base = self._rowperm_base.get(key) if hasattr(self, "_rowperm_base") else None
        if base is None:
            st = design_stack(X[None], center=True)
            r = st["r"]
            base = dict(r=r, rp=self._rp(r), dof=st["dof"], G=st["G"][0], d=st["d"][0],
                        pinv=torch.from_numpy(np.ascontiguousarray(st["pinv"][0])).to(self.device), rep={}, fmat={})
            self._rowperm_base = {key: base}            # one design at a time (the drivers' loop)
        r, rp = base["r"], base["rp"]
        rep = base["rep"].get(P)
        if rep is None:
            rep = (torch.from_numpy(np.repeat(base["G"][None], P, axis=0)).to(self.device),
                   torch.from_numpy(np.repeat(base["d"][None], P, axis=0)).to(self.device))
            base["rep"] = {P: rep}
            base["fmat"] = {}
        idx_d = self._upload("perm_idx", np.ascontiguousarray(perm_idx, dtype=np.int32))
        layout = base["layout"] = self._layout(rp)
        ldA = int(_lib.lib().tmb_glm_packed_columns(self.Y.dtype_code, P, rp))
        At_d = self._ring("At", (n, ldA), torch.float64)
        _lib.check(_lib.lib().tmb_glm_pack_rowperm(_lib.ptr(base["pinv"]), r, n, _lib.ptr(idx_d), P, rp, _lib.ptr(At_d),
                                                   ldA, layout, _lib.current_stream()))
        return base, rep, At_d, ldA, P

    def tstat_rowperm(self, X, perm_idx, rows=None):
        """Fused fit+t for the designs X[perm_idx[p]] (whole rows permuted).  X float64 [n, k] with the intercept in
        column 0; rows = (first, count) selects regressors (default: all k-1).  Returns CUDA float32 [P, count, ld] in
        the engine's internal column order."""
        import torch
        if np.asarray(X).shape[1] - 1 > 8:       # stored-beta path: host-built stack (k x k algebra only)
            return self.tstat(row_permuted_stack(X, perm_idx), rows=rows, caller_order=False)
        base, rep, At_d, ldA, P = self._rowperm_operands(X, perm_idx)
        r, rp = base["r"], base["rp"]
        row0, nrows = (0, r) if rows is None else (int(rows[0]), int(rows[1]))
        yy = self.Y.sumsq(True)
        t32 = self._ring("t32", (P, nrows, self.Y.ld), torch.float32)
        _lib.check(_lib.lib().tmb_glm_tstat(
            _lib.ptr(self.Y.t), self.Y.dtype_code, self.Y.n, self.Y.V, self.Y.ld, _lib.ptr(At_d), ldA,
            _lib.ptr(rep[0]), _lib.ptr(rep[1]), P, r, rp, row0, nrows, base["dof"], _lib.ptr(yy), _lib.ptr(t32),
            None, self.Y.ld, 1 if self.nan_to_zero else 0, base["layout"], _lib.current_stream()))
        return t32

    # -- F statistics (tm-models GLM branch) -------------------------------------------------------
    def fstat(self, stack, var_lo, var_k, want_model=False, want_f64=False, caller_order=True):
        """Model F and per-variable partial F of pyfunc.py:2282-2401 glm_typeI for a design stack (design_stack
        output, intercept centred away): variable i covers regressors [var_lo[i], var_lo[i] + var_k[i]).
        Returns CUDA float32 [P, nvar (+1 with the model F first), ld] (and float64 when want_f64)."""
        import torch
        P, r, n = stack["pinv"].shape
        if n != self.Y.n:
            raise ValueError("design has %d subjects, data has %d" % (n, self.Y.n))
        rp = self._rp(r)
        G_d = self._upload("G", stack["G"])
        M_d = self._upload("M", fstat_blocks(stack["G"], var_lo, var_k))
        if rp == 0:
            nvar = len(var_lo)
            nrows = nvar + (1 if want_model else 0)
            lo = np.ascontiguousarray(var_lo, dtype=np.int32)
            kk = np.ascontiguousarray(var_k, dtype=np.int32)
            yy = self.Y.sumsq(True)
            f32 = torch.empty((P, nrows, self.Y.ld), dtype=torch.float32, device=self.device)
            f64 = torch.empty((P, nrows, self.Y.ld), dtype=torch.float64, device=self.device) if want_f64 else None
            for a, cnt, beta in self._betas_chunks(stack["pinv"]):
                _lib.check(_lib.lib().tmb_glm_fstat_beta(
                    _lib.ptr(beta), self.Y.ld, self.Y.V, _lib.ptr(G_d[a:a + cnt]), _lib.ptr(M_d[a:a + cnt]), cnt, r, nvar,
                    lo.ctypes.data, kk.ctypes.data, 1 if want_model else 0, stack["dof"], _lib.ptr(yy),
                    _lib.ptr(self.Y.sstotal_reference()) if want_model else None, _lib.ptr(f32[a:a + cnt]), _lib.ptr(f64[a:a + cnt]) if f64 is not None else None, self.Y.ld,
                    1 if self.nan_to_zero else 0, _lib.current_stream()))
            out = (f32, f64)
        else:
            At, ldA, layout = self._pack(stack["pinv"], rp)
            At_d = self._upload("At", At)
            out = self._fstat_launch(At_d, ldA, G_d, M_d, P, r, rp, var_lo, var_k, want_model, stack["dof"], want_f64,
                                     layout=layout)
        if caller_order and self.colperm is not None:
            out = tuple(self.to_caller_order(o) if o is not None else None for o in out)
        return out if want_f64 else out[0]

    def _fstat_launch(self, At_d, ldA, G_d, M_d, P, r, rp, var_lo, var_k, want_model, dof, want_f64=False, layout=0):
        import torch
        nvar = len(var_lo)
        nrows = nvar + (1 if want_model else 0)
        lo = np.ascontiguousarray(var_lo, dtype=np.int32)
        kk = np.ascontiguousarray(var_k, dtype=np.int32)
        yy = self.Y.sumsq(True)
        f32 = torch.empty((P, nrows, self.Y.ld), dtype=torch.float32, device=self.device)
        f64 = torch.empty((P, nrows, self.Y.ld), dtype=torch.float64, device=self.device) if want_f64 else None
        _lib.check(_lib.lib().tmb_glm_fstat(
            _lib.ptr(self.Y.t), self.Y.dtype_code, self.Y.n, self.Y.V, self.Y.ld, _lib.ptr(At_d), ldA, _lib.ptr(G_d),
            _lib.ptr(M_d), P, r, rp, nvar, lo.ctypes.data, kk.ctypes.data, 1 if want_model else 0, dof, _lib.ptr(yy),
            _lib.ptr(self.Y.sstotal_reference()) if want_model else None, _lib.ptr(f32), _lib.ptr(f64), self.Y.ld, 1 if self.nan_to_zero else 0, layout, _lib.current_stream()))
        return f32, f64

    def fstat_rowperm(self, X, var_lo, var_k, perm_idx, want_model=False):
        """fstat for the designs X[perm_idx[p]] (glm_typeI's `exog_vars[rand_array]`, pyfunc.py:2317-2321): only the
        index rows travel; X'X, and with it every variable's inverse block, is the same for all permutations."""
        import torch
        if np.asarray(X).shape[1] - 1 > 8:       # stored-beta path
            return self.fstat(row_permuted_stack(X, perm_idx), var_lo, var_k, want_model=want_model, caller_order=False)
        base, rep, At_d, ldA, P = self._rowperm_operands(X, perm_idx)
        key = (tuple(int(a) for a in var_lo), tuple(int(a) for a in var_k), P)
        M_d = base["fmat"].get(key)
        if M_d is None:
            M1 = fstat_blocks(base["G"][None], var_lo, var_k)
            M_d = torch.from_numpy(np.repeat(M1, P, axis=0)).to(self.device)
            base["fmat"] = {key: M_d}
        return self._fstat_launch(At_d, ldA, rep[0], M_d, P, base["r"], base["rp"], var_lo, var_k, want_model,
                                  base["dof"], layout=base["layout"])[0]

    def glm_typeI_block(self, exog_vars, kvars, perm_idx, stat="f", download=True):
        """One block of the tm-models GLM permutation loop (tmanalysis/tm_models_randomise.py:197-272): per shuffle the
        design is exog_vars[perm_idx[p]] (intercept in column 0, then the variables of interest with kvars[i] columns
        each, then covariates).  stat 'f': per-variable F maps -> one-sided TFCE -> scaled max, float32 [P, nvar, S];
        't': t of the variables' columns, both signs, float32 [P, sum(kvars), S, 2]; 'both': (F result, t result)."""
        exog_vars = np.asarray(exog_vars, dtype=np.float64)
        if not has_intercept(exog_vars):
            raise ValueError("exog_vars must have the intercept in column 0")
        kvars = [int(a) for a in kvars]
        var_lo = np.concatenate([[0], np.cumsum(kvars)[:-1]]).astype(np.int32)   # regressor index = column - 1
        out_f = out_t = None
        if stat in ("f", "both"):
            f32 = self.fstat_rowperm(exog_vars, var_lo, kvars, perm_idx)
            P, nv, ld = f32.shape
            mx, status, _ = self.plan.run(f32.view(P * nv, ld), two_sided=False)
            out_f = mx.view(P, nv, self.plan.S, 2)[..., 0]
            out_f = self._download(out_f.contiguous()) if download else out_f
        if stat in ("t", "both"):
            ncon = int(sum(kvars))
            t32 = self.tstat_rowperm(exog_vars, perm_idx, rows=(0, ncon))
            P, C, ld = t32.shape
            mx, status, _ = self.plan.run(t32.view(P * C, ld), two_sided=True)
            out_t = mx.view(P, C, self.plan.S, 2)
            out_t = self._download(out_t) if download else out_t
        return out_f if stat == "f" else out_t if stat == "t" else (out_f, out_t)

    # -- designs of which only some columns change between shuffles (the drivers' -v mode) -----------
    def _partial_columns(self, designs):
        """Indices (among the non-intercept columns) of the columns that differ between the designs of a block, or None
        when the cross-product path does not apply (all columns change, float64 data kept on the established path,
        more than 16 regressors, TMB_GLM_PARTIAL=0)."""
        P, n, k = designs.shape
        r = k - 1
        if P < 2 or r < 2 or r > 16 or _os.environ.get("TMB_GLM_PARTIAL", "") == "0":
            return None
        changing = np.flatnonzero(np.any(designs[:, :, 1:] != designs[:1, :, 1:], axis=(0, 1)))
        if changing.size == 0 or changing.size == r:
            return None
        return changing

    def tstat_partial(self, designs, changing, want_f64=False):
        """t of every regressor for designs [P, n, k] whose columns `changing` (indices among the k-1 regressors) differ
        between shuffles while the others are fixed -- vertex_tfce_multiple_regression_randomise.py:84-97 (`-v first
        last`): only the changing columns are contracted with the data per shuffle, the fixed columns' cross-products are
        fitted once (tmb_glm_tstat_cross_rows).  CUDA float32 [P, k-1, ld] (internal column order)."""
        import torch
        P, n, k = designs.shape
        r = k - 1
        Z = designs[:, :, 1:] - designs[:, :, 1:].mean(axis=1, keepdims=True)             # centred regressors [P, n, r]
        C = np.ascontiguousarray(np.linalg.inv(np.einsum("pni,pnj->pij", Z, Z)))
        fixed_idx = np.setdiff1d(np.arange(r), changing)
        m, f = int(changing.size), int(fixed_idx.size)
        colmap = np.empty(r, dtype=np.int32)
        colmap[changing] = np.arange(m)
        colmap[fixed_idx] = m + np.arange(f)
        fixed = np.ascontiguousarray(Z[0][:, fixed_idx])
        key = fixed.tobytes()
        st = getattr(self, "_partial_fixed", None)
        if st is None or st[0] != key:
            ldF = round_up(f, TILE_M)
            At = np.zeros((n, ldF))
            At[:, :f] = fixed
            cfix = torch.empty((f, self.Y.ld), dtype=torch.float64, device=self.device)
            At_d = torch.from_numpy(At).to(self.device)
            _lib.check(_lib.lib().tmb_glm_beta(_lib.ptr(self.Y.t), self.Y.dtype_code, n, self.Y.V, self.Y.ld, _lib.ptr(At_d), ldF,
                                               f, _lib.ptr(cfix), self.Y.ld, _lib.current_stream()))
            st = (key, cfix)
            self._partial_fixed = st
        cfix = st[1]
        yy = self.Y.sumsq(True)
        t32 = torch.empty((P, r, self.Y.ld), dtype=torch.float32, device=self.device)
        t64 = torch.empty((P, r, self.Y.ld), dtype=torch.float64, device=self.device) if want_f64 else None
        colmap_d = self._upload("part_colmap", colmap)
        C_d = self._upload("part_C", C)
        lib, stream = _lib.lib(), _lib.current_stream()
        per = max(1, int(1.5e9 // (m * self.Y.ld * 8)))
        for a in range(0, P, per):
            b = min(P, a + per)
            rows = (b - a) * m
            ldA = round_up(rows, TILE_M)
            At = np.zeros((n, ldA))
            At[:, :rows] = Z[a:b][:, :, changing].transpose(1, 0, 2).reshape(n, rows)      # column p*m + i
            At_d = self._upload("part_At", At)
            cperm = self._ring("part_cperm", (rows, self.Y.ld), torch.float64)
            _lib.check(lib.tmb_glm_beta(_lib.ptr(self.Y.t), self.Y.dtype_code, n, self.Y.V, self.Y.ld, _lib.ptr(At_d), ldA, rows,
                                        _lib.ptr(cperm), self.Y.ld, stream))
            _lib.check(lib.tmb_glm_tstat_cross_rows(
                _lib.ptr(cperm), self.Y.ld, m, _lib.ptr(cfix), self.Y.ld, f, self.Y.V, _lib.ptr(C_d[a:b]), r, _lib.ptr(colmap_d),
                _lib.ptr(colmap), 0, r, float(n - k), _lib.ptr(yy), b - a, _lib.ptr(t32[a:b]),
                _lib.ptr(t64[a:b]) if t64 is not None else None, self.Y.ld, 1 if self.nan_to_zero else 0, stream))
        return (t32, t64) if want_f64 else t32

    # -- whole shuffles --------------------------------------------------------------------------
    def regression_block(self, X, perm_idx=None, designs=None, want_maps=False, download=True):
        """Regression + TFCE + scaled max for a block of shuffles.

        Either perm_idx int [P, n] (designs X[perm_idx[p]], whole rows permuted) or designs
        float64 [P, n, k] (arbitrary per-shuffle designs, e.g. the -v partial-column mode).
        X / designs must carry the intercept in column 0, as every reference driver builds them
        (np.column_stack([np.ones(n), pred_x])).
        Returns float32 [P, C, S, 2]: per shuffle, contrast (= regressor 1..k-1), surface and sign
        (+t, -t) the scaled TFCE maximum of pyfunc.py:116-118 / :125 / tm_func.py:173-182."""
        if designs is not None:
            designs = np.asarray(designs, dtype=np.float64)
            if not has_intercept(designs):
                raise ValueError("designs must have the intercept in column 0")
            part = self._partial_columns(designs)
            if part is not None:
                t32 = self.tstat_partial(designs, part)
                return self._finish_regression(t32, want_maps, download)
            stack = design_stack(designs, center=True)
        else:
            X = np.asarray(X, dtype=np.float64)
            if not has_intercept(X):
                raise ValueError("X must have the intercept in column 0")
            stack = None
        t32 = self.tstat(stack, caller_order=False) if stack is not None else self.tstat_rowperm(X, perm_idx)
        return self._finish_regression(t32, want_maps, download)

    def _finish_regression(self, t32, want_maps, download):
        """TFCE + scaled maxima of a block of t maps [P, C, ld] (internal column order)."""
        P, C, ld = t32.shape
        mx, status, maps = self.plan.run(t32.view(P * C, ld), two_sided=self.two_sided, want_maps=want_maps)
        mx = mx.view(P, C, self.plan.S, 2)
        self.last_status = status
        if want_maps:
            maps = tuple(self.to_caller_order(m) if m is not None else None for m in maps)
            return mx, self.to_caller_order(t32), maps
        return self._download(mx) if download else mx

    def regression_blocks(self, X, perm_idx, block=256):
        """Many shuffles, `block` at a time, software-pipelined on one stream: while the sweep of block i runs,
        the fit and the maxima of block i+1 are already queued and the host builds block i+1's threshold
        tables, so the GPU never waits for the host round trip of the exact-libm mode.
        perm_idx int [N, n] (whole-row permutations of X).  Returns float32 [N, C, S, 2] on the host."""
        import torch
        X = np.asarray(X, dtype=np.float64)
        if not has_intercept(X):
            raise ValueError("X must have the intercept in column 0")
        perm_idx = np.asarray(perm_idx)
        N = perm_idx.shape[0]
        chunks = [(a, min(N, a + block)) for a in range(0, N, block)]
        C = X.shape[1] - 1
        host = torch.empty((N, C, self.plan.S, 2), dtype=torch.float32).pin_memory()

        def stage1(a, b):
            t32 = self.tstat_rowperm(X, perm_idx[a:b])
            flat = t32.view(t32.shape[0] * t32.shape[1], t32.shape[2])
            tk = self.plan.prepare(flat) if self.plan.exact_pow else None
            return t32, flat, tk

        nxt = stage1(*chunks[0]) if chunks else None
        for ci, (a, b) in enumerate(chunks):
            cur = nxt
            nxt = stage1(*chunks[ci + 1]) if ci + 1 < len(chunks) else None   # queued before this block's sweep
            t32, flat, tk = cur
            if tk is not None:
                mx, status, _ = self.plan.finish(tk, flat, two_sided=self.two_sided)
            else:
                mx, status, _ = self.plan.run(flat, two_sided=self.two_sided, exact_pow=False)
            host[a:b].copy_(mx.view(b - a, C, self.plan.S, 2), non_blocking=True)   # no per-block host sync
            self.d2h_bytes += (b - a) * C * self.plan.S * 2 * 4
        torch.cuda.current_stream().synchronize()
        self.last_status = None
        return host.numpy().copy()

    def observed_statistics(self, X):
        """Un-permuted statistics with full TFCE maps -- the computation of the reference's step-1 writers
        (STEP_1_vertex_tfce_multiple_regression.py:354-398 -> pyfunc.py:80-91 write_vertStat_img,
        pyfunc.py:93-103 write_voxelStat_img; image file I/O is out of scope).  Returns a dict of host arrays:
        t [C, V] float32, tfce_pos / tfce_neg [C, V] float32 scaled per surface by max(stat)/100 and the
        vertex-density weights exactly like `vertStat_TFCE * (vertStat.max()/100) * density_corr`, and
        max_pos / max_neg [C, S] (the values the reference echoes to max_TFCE_contrast_values.csv)."""
        n = self.Y.n
        mx, t32, (pos, neg) = self.regression_block(X, perm_idx=np.arange(n)[None, :], want_maps=True)
        V = self.Y.V
        t = to_host(t32[0, :, :V])
        pos = to_host(pos[:, :V])
        neg = to_host(neg[:, :V]) if neg is not None else None
        C = t.shape[0]
        out_pos = np.zeros_like(pos)
        out_neg = np.zeros_like(pos) if neg is not None else None
        for s in self.plan.surfaces:
            a, b = s.col_offset, s.col_offset + s.adjset.num_vertices
            w = 1 if s.weight is None else s.weight
            for c in range(C):
                seg = t[c, a:b]
                out_pos[c, a:b] = pos[c, a:b] * (seg[np.isfinite(seg)].max() / 100) * w
                if neg is not None:
                    nseg = -seg
                    out_neg[c, a:b] = neg[c, a:b] * (nseg[np.isfinite(nseg)].max() / 100) * w
        mxh = to_host(mx[0])
        return dict(t=t, tfce_pos=out_pos, tfce_neg=out_neg, max_pos=mxh[:, :, 0], max_neg=mxh[:, :, 1])

    def sobelz(self, medtype, pred_x, depend_y, perm_idx, alg="aroian", want_f64=False, caller_order=True):
        """Fused two-fit Sobel-family z for a block of shuffles (pyfunc.py:130-162).
        Returns CUDA float32 [P, ld] (and float64 when want_f64)."""
        if self.sobelz_cross_ok(medtype):
            z32, z64 = self.sobelz_cross(medtype, pred_x, depend_y, perm_idx, alg, want_f64=want_f64)
            if caller_order and self.colperm is not None:
                z32 = self.to_caller_order(z32)
                z64 = self.to_caller_order(z64) if z64 is not None else None
            return (z32, z64) if want_f64 else z32
        XA, XB, ta_scalar = self.mediation_designs(medtype, pred_x, depend_y, perm_idx)
        return self.sobelz_designs(XA, XB, ta_scalar, alg, want_f64, caller_order)

    def sobelz_cross_ok(self, medtype):
        """Medtype 'M' / 'I' on float32 data with the tensor-core fit: one contraction row per shuffle (sobelz_cross)."""
        return (medtype in ("M", "I") and self.Y.dtype_code == 0 and self._layout(1) == 1
                and _os.environ.get("TMB_SOBEL", "") != "designs")

    def sobelz_cross(self, medtype, pred_x, depend_y, perm_idx, alg="aroian", want_f64=False, out=None):
        """Sobel z for medtype 'M' / 'I' from ONE fitted row per shuffle (tmb_sobelz_cross).  Only pred_x is permuted in
        these two types (vertex_tfce_mediation_randomise.py:82-90), so of the centred cross-products with the data,
        x_p'y and dep'y, only the first changes: dep'y is one row fitted once per (engine, depend_y), and path A's and
        path B's betas and residual sums of squares follow per vertex from the two cross-products and the shuffle's 2 x 2
        Gram matrix.  The host ships the index rows and 4 doubles per shuffle; the permuted predictor columns are gathered
        on the device.  Returns (CUDA float32 [P, ld], float64 or None), internal column order."""
        import torch
        n = self.Y.n
        x = np.asarray(pred_x, dtype=np.float64).reshape(n)
        dep = np.asarray(depend_y, dtype=np.float64).reshape(n)
        perm_idx = np.ascontiguousarray(perm_idx, dtype=np.int32)
        P = perm_idx.shape[0]
        algc = {"aroian": 0, "sobel": 1, "goodman": 2}.get(alg)
        if algc is None:
            raise ValueError("Unknown indirect test algorithm")
        xc, dc = x - x.mean(), dep - dep.mean()
        key = (x.tobytes(), dep.tobytes())
        fixed = getattr(self, "_sobel_fixed", None)
        if fixed is None or fixed[0] != key:
            At = np.zeros((n, TILE_M), dtype=np.float64)
            At[:, 0] = dc
            cd = torch.empty((1, self.Y.ld), dtype=torch.float64, device=self.device)
            At_d = torch.from_numpy(At).to(self.device)
            _lib.check(_lib.lib().tmb_glm_beta(_lib.ptr(self.Y.t), self.Y.dtype_code, n, self.Y.V, self.Y.ld, _lib.ptr(At_d),
                                               TILE_M, 1, _lib.ptr(cd), self.Y.ld, _lib.current_stream()))
            fixed = (key, cd, torch.from_numpy(np.ascontiguousarray(xc[None, :])).to(self.device))
            self._sobel_fixed = fixed
        _, cd, xc_d = fixed
        xx, dd = float(xc @ xc), float(dc @ dc)
        sxd = xc[perm_idx] @ dc                                        # [P]: x_p'dep, the only entry that changes
        xpos = 1 if medtype == "M" else 0
        G = np.empty((P, 2, 2))
        G[:, xpos, xpos], G[:, 1 - xpos, 1 - xpos] = xx, dd
        G[:, 0, 1] = G[:, 1, 0] = sxd
        CB = np.zeros((P, 8))
        CB[:, :4] = np.linalg.inv(G).reshape(P, 4)
        CB[:, 4] = CB[:, 0] / float(n - 3)                            # tested row 0: C[0][0] / dofB
        CB_d = self._upload("sobel_CB", CB)
        idx_d = self._upload("perm_idx", perm_idx)
        ldA = int(_lib.lib().tmb_glm_packed_columns(self.Y.dtype_code, P, 1))
        At_d = self._ring("At", (n, ldA), torch.float64)
        lib, st = _lib.lib(), _lib.current_stream()
        _lib.check(lib.tmb_glm_pack_rowperm(_lib.ptr(xc_d), 1, n, _lib.ptr(idx_d), P, 1, _lib.ptr(At_d), ldA, 1, st))
        yy = self.Y.sumsq(True)
        z32 = out if out is not None else torch.empty((P, self.Y.ld), dtype=torch.float32, device=self.device)
        z64 = torch.empty((P, self.Y.ld), dtype=torch.float64, device=self.device) if want_f64 else None
        _lib.check(lib.tmb_sobelz_cross(_lib.ptr(self.Y.t), self.Y.dtype_code, n, self.Y.V, self.Y.ld, _lib.ptr(At_d), ldA,
                                        _lib.ptr(cd), xx, float(n - 2), _lib.ptr(CB_d), xpos, 0, float(n - 3), _lib.ptr(yy),
                                        P, algc, _lib.ptr(z32), _lib.ptr(z64), self.Y.ld, st))
        return z32, z64

    def mediation_designs(self, medtype, pred_x, depend_y, perm_idx):
        """Per-shuffle designs of calc_sobelz (pyfunc.py:130-162) under the drivers' permutation rule
        (vertex_tfce_mediation_randomise.py:82-90): (XA [P, n, 2] or None, XB [P, n, 3], ta scalars or None)."""
        n = self.Y.n
        pred_x = np.asarray(pred_x, dtype=np.float64).reshape(n)
        depend_y = np.asarray(depend_y, dtype=np.float64).reshape(n)
        perm_idx = np.asarray(perm_idx)
        P = perm_idx.shape[0]
        xp = pred_x[perm_idx]                                       # [P, n]
        ones = np.ones((P, n))
        ta_scalar = None
        if medtype in ("M", "I"):
            dep = np.broadcast_to(depend_y, (P, n))
            XA = np.stack([ones, xp], axis=2)
            XB = np.stack([ones, dep, xp], axis=2) if medtype == "M" else np.stack([ones, xp, dep], axis=2)
        elif medtype == "Y":
            dep = depend_y[perm_idx]
            ta_scalar = linregress_t(xp, dep)                        # scalar path A, pyfunc.py:142
            XA = None
            XB = np.stack([ones, dep, xp], axis=2)
        else:
            raise ValueError("Invalid mediation type")
        return XA, XB, ta_scalar

    def sobelz_designs(self, XA, XB, ta_scalar=None, alg="aroian", want_f64=False, caller_order=True):
        """Sobel-family z from two stacks of per-shuffle designs (intercept in column 0): path A = t of XA's first
        regressor (or the per-shuffle scalars ta_scalar when XA is None), path B = t of XB's first regressor.
        XA float64 [P, n, kA], XB [P, n, kB].  Serves calc_sobelz (pyfunc.py:130-162) and the tm-models mediation
        branch (tm_models_randomise.py:430-503: glm_typeI t-values + calc_indirect)."""
        ops = self.sobelz_operands(XA, XB, ta_scalar, alg)
        z32, z64 = self.sobelz_launch(ops, want_f64=want_f64)
        if caller_order and self.colperm is not None:
            z32 = self.to_caller_order(z32)
            z64 = self.to_caller_order(z64) if z64 is not None else None
        return (z32, z64) if want_f64 else z32

    def sobelz_operands(self, XA, XB, ta_scalar=None, alg="aroian", resident=False):
        """Host algebra (k x k per design) + upload of the two-path operands; `resident` keeps private device copies
        (bench: operands prepared before the timed region) instead of the reused pinned staging buffers."""
        import torch
        P = XB.shape[0]
        sB = design_stack(XB, center=True)
        if XA is not None:
            sA = design_stack(XA, center=True)
            rA = sA["r"]
            pinv = np.concatenate([sA["pinv"], sB["pinv"]], axis=1)
        else:
            sA, rA = None, 0
            pinv = sB["pinv"]
        rB = sB["r"]
        rp = self._rp(rA + rB)
        if rp == 0:
            At, ldA, layout = None, 0, 0
        else:
            At, ldA, layout = self._pack(pinv, rp)
        algc = {"aroian": 0, "sobel": 1, "goodman": 2}.get(alg)
        if algc is None:
            raise ValueError("Unknown indirect test algorithm")
        if resident:
            up = lambda name, a: torch.from_numpy(np.ascontiguousarray(a)).to(self.device)   # noqa: E731
        else:
            up = self._upload
        return dict(P=P, rp=rp, ldA=ldA, rA=rA, rB=rB, alg=algc, layout=layout, pinv=pinv if rp == 0 else None,
                    At=up("med_At", At) if At is not None else None, GA=up("med_GA", sA["G"]) if sA else None,
                    dA=up("med_dA", sA["d"]) if sA else None, GB=up("med_GB", sB["G"]), dB=up("med_dB", sB["d"]),
                    ta=up("med_ta", ta_scalar) if ta_scalar is not None else None,
                    dofA=sA["dof"] if sA else 1.0, dofB=sB["dof"])

    def sobelz_launch(self, ops, want_f64=False, out=None):
        """The fused two-fit kernel (tmb_sobelz) on prepared operands: CUDA float32 [P, ld] (internal column order)."""
        import torch
        P = ops["P"]
        yy = self.Y.sumsq(True)
        z32 = out if out is not None else torch.empty((P, self.Y.ld), dtype=torch.float32, device=self.device)
        z64 = torch.empty((P, self.Y.ld), dtype=torch.float64, device=self.device) if want_f64 else None
        if ops["rp"] == 0:                      # more than 8 regressors over both paths: stored-beta path
            sl = lambda t, a, c: _lib.ptr(t[a:a + c]) if t is not None else None   # noqa: E731
            for a, cnt, beta in self._betas_chunks(ops["pinv"]):
                _lib.check(_lib.lib().tmb_sobelz_beta(
                    _lib.ptr(beta), self.Y.ld, self.Y.V, sl(ops["GA"], a, cnt), sl(ops["dA"], a, cnt), ops["rA"], 0,
                    ops["dofA"], sl(ops["GB"], a, cnt), sl(ops["dB"], a, cnt), ops["rB"], 0, ops["dofB"], _lib.ptr(yy),
                    sl(ops["ta"], a, cnt), cnt, ops["alg"], _lib.ptr(z32[a:a + cnt]), sl(z64, a, cnt), self.Y.ld,
                    _lib.current_stream()))
            return z32, z64
        _lib.check(_lib.lib().tmb_sobelz(
            _lib.ptr(self.Y.t), self.Y.dtype_code, self.Y.n, self.Y.V, self.Y.ld, _lib.ptr(ops["At"]), ops["ldA"],
            ops["rp"], _lib.ptr(ops["GA"]), _lib.ptr(ops["dA"]), ops["rA"], 0, ops["dofA"], _lib.ptr(ops["GB"]),
            _lib.ptr(ops["dB"]), ops["rB"], 0, ops["dofB"], _lib.ptr(yy), _lib.ptr(ops["ta"]), P, ops["alg"],
            _lib.ptr(z32), _lib.ptr(z64), self.Y.ld, ops["layout"], _lib.current_stream()))
        return z32, z64

    def tm_models_mediation_block(self, medtype, leftvar, rightvar, dmy_covariates, perm_idx, alg="aroian", download=True):
        """One block of the tm-models mediation loop (tmanalysis/tm_models_randomise.py:430-520): shuffle p uses
        leftvar[perm_idx[p]] (and rightvar[perm_idx[p]] for medtype 'Y'); path A / path B are glm_typeI t-values of the
        designs [1, left, cov] and [1, left, right, cov] ('I') or [1, right, left, cov] ('M', 'Y'); 'Y' takes path A from
        the regression of rightvar on [1, left, cov] (one scalar per shuffle).  Sobel z (calc_indirect, Aroian by
        default) -> one-sided TFCE -> scaled max: float32 [P, S]."""
        n = self.Y.n
        left = np.asarray(leftvar, dtype=np.float64).reshape(n, -1)
        right = np.asarray(rightvar, dtype=np.float64).reshape(n, -1)
        cov = None if dmy_covariates is None else np.asarray(dmy_covariates, dtype=np.float64).reshape(n, -1)
        perm_idx = np.asarray(perm_idx)
        P = perm_idx.shape[0]
        ones = np.ones((P, n, 1))
        lv = left[perm_idx]                                                  # [P, n, kL]
        rv = right[perm_idx] if medtype == "Y" else np.broadcast_to(right, (P,) + right.shape)
        tail = [np.broadcast_to(cov, (P,) + cov.shape)] if cov is not None else []
        XA = np.concatenate([ones, lv] + tail, axis=2)
        if medtype == "I":
            XB = np.concatenate([ones, lv, rv] + tail, axis=2)
        elif medtype in ("M", "Y"):
            XB = np.concatenate([ones, rv, lv] + tail, axis=2)
        else:
            raise ValueError("Invalid mediation type")
        ta_scalar = None
        if medtype == "Y":
            if right.shape[1] != 1:
                raise ValueError("medtype 'Y' needs a single-column dependent variable")
            # n x 1 regressions of the permuted right variable on [1, left_p, cov], all shuffles at once on the host
            # (tval_int's arithmetic, cynumstats.pyx:59-64: explicit residuals, se rounded to float32)
            k = XA.shape[2]
            invXX = np.linalg.inv(np.einsum("pni,pnj->pij", XA, XA))
            a = np.einsum("pij,pj->pi", invXX, np.einsum("pni,pn->pi", XA, rv[:, :, 0]))
            resid = rv[:, :, 0] - np.einsum("pni,pi->pn", XA, a)
            sigma2 = np.sum(resid ** 2, axis=1) / (n - k)
            se = np.sqrt(sigma2 * invXX[:, 1, 1]).astype(np.float32)
            ta_scalar = a[:, 1] / se
            XA = None
        kL, kR = left.shape[1], right.shape[1]
        if (_os.environ.get("TMB_SOBEL", "") != "designs" and kL + (kR if medtype == "Y" else 0) <= MAX_REGRESSORS
                and XB.shape[2] - 1 <= 16):
            z32 = self._tm_models_sobelz_cross(medtype, left, right, cov, perm_idx, XB, ta_scalar, alg)
        else:
            z32 = self.sobelz_designs(XA, XB, ta_scalar, alg, caller_order=False)
        mx, status, _ = self.plan.run(z32, two_sided=False)
        self.last_status = status
        mx = mx[:, :, 0]
        return self._download(mx.contiguous()) if download else mx

    def _tm_models_sobelz_cross(self, medtype, left, right, cov, perm_idx, XB, ta_scalar, alg, want_f64=False):
        """tm-models Sobel z from centred cross-products (tmb_sobelz_cross_rows): only the permuted columns -- the left
        variable; for 'Y' the right one too -- are contracted with the data per shuffle; the rows of the fixed columns
        (right variable, covariates) are fitted once.  XB float64 [P, n, 1 + rB]: path B's designs (path A's columns are
        a subset), used for the k x k Gram matrices only.  CUDA float32 [P, ld], internal column order."""
        import torch
        n = self.Y.n
        P = perm_idx.shape[0]
        kL, kR = left.shape[1], right.shape[1]
        centre = lambda a: a - a.mean(axis=0, keepdims=True)   # noqa: E731
        perm_cols = centre(left) if medtype != "Y" else np.column_stack([centre(left), centre(right)])
        fixed = [centre(right)] if medtype != "Y" else []
        if cov is not None:
            fixed.append(centre(cov))
        fixed = np.column_stack(fixed) if fixed else np.zeros((n, 0))
        m, f = perm_cols.shape[1], fixed.shape[1]
        L = list(range(kL))                                                   # sources: permuted rows first, then fixed
        R = list(range(kL, kL + kR)) if medtype == "Y" else list(range(m, m + kR))
        Cv = list(range(m + (0 if medtype == "Y" else kR), m + f))
        mapA = L + Cv
        mapB = (L + R + Cv) if medtype == "I" else (R + L + Cv)
        rA, rB = (0 if ta_scalar is not None else len(mapA)), len(mapB)
        colmap = np.asarray((mapA if rA else []) + mapB, dtype=np.int32)
        # Gram matrices from the centred designs of path B; path A's columns are a subset of them
        ZB = XB[:, :, 1:] - XB[:, :, 1:].mean(axis=1, keepdims=True)
        GB = np.einsum("pni,pnj->pij", ZB, ZB)
        CB = np.ascontiguousarray(np.linalg.inv(GB))
        CA = None
        if rA:
            posB = {src: i for i, src in enumerate(mapB)}
            ia = [posB[src] for src in mapA]
            CA = np.ascontiguousarray(np.linalg.inv(GB[:, ia][:, :, ia]))
        key = (medtype, perm_cols.tobytes(), fixed.tobytes())
        st = getattr(self, "_tmm_fixed", None)
        if st is None or st[0] != key:
            cfix = None
            if f:
                ldF = round_up(f, TILE_M)
                At = np.zeros((n, ldF))
                At[:, :f] = fixed
                cfix = torch.empty((f, self.Y.ld), dtype=torch.float64, device=self.device)
                At_d = torch.from_numpy(At).to(self.device)
                _lib.check(_lib.lib().tmb_glm_beta(_lib.ptr(self.Y.t), self.Y.dtype_code, n, self.Y.V, self.Y.ld, _lib.ptr(At_d),
                                                   ldF, f, _lib.ptr(cfix), self.Y.ld, _lib.current_stream()))
            st = (key, cfix, torch.from_numpy(np.ascontiguousarray(perm_cols.T)).to(self.device))
            self._tmm_fixed = st
        _, cfix, pc_d = st
        algc = {"aroian": 0, "sobel": 1, "goodman": 2}.get(alg)
        if algc is None:
            raise ValueError("Unknown indirect test algorithm")
        yy = self.Y.sumsq(True)
        z32 = torch.empty((P, self.Y.ld), dtype=torch.float32, device=self.device)
        z64 = torch.empty((P, self.Y.ld), dtype=torch.float64, device=self.device) if want_f64 else None
        colmap_d = self._upload("tmm_colmap", colmap)
        CA_d = self._upload("tmm_CA", CA) if CA is not None else None
        CB_d = self._upload("tmm_CB", CB)
        ta_d = self._upload("tmm_ta", ta_scalar) if ta_scalar is not None else None
        idx_all = np.ascontiguousarray(perm_idx, dtype=np.int32)
        lib, stream = _lib.lib(), _lib.current_stream()
        per = max(1, int(1.5e9 // (m * self.Y.ld * 8)))
        for a in range(0, P, per):
            b = min(P, a + per)
            cnt = b - a
            rows = cnt * m
            ldA = round_up(rows, TILE_M)
            idx_d = self._upload("perm_idx", idx_all[a:b])
            At_d = self._ring("tmm_At", (n, ldA), torch.float64)
            _lib.check(lib.tmb_glm_pack_rowperm(_lib.ptr(pc_d), m, n, _lib.ptr(idx_d), cnt, m, _lib.ptr(At_d), ldA, 0, stream))
            cperm = self._ring("tmm_cperm", (rows, self.Y.ld), torch.float64)
            _lib.check(lib.tmb_glm_beta(_lib.ptr(self.Y.t), self.Y.dtype_code, n, self.Y.V, self.Y.ld, _lib.ptr(At_d), ldA, rows,
                                        _lib.ptr(cperm), self.Y.ld, stream))
            _lib.check(lib.tmb_sobelz_cross_rows(
                _lib.ptr(cperm), self.Y.ld, m, _lib.ptr(cfix), self.Y.ld, f, self.Y.V,
                _lib.ptr(CA_d[a:b]) if CA_d is not None else None, rA, 0, float(n - 1 - len(mapA)), _lib.ptr(CB_d[a:b]), rB, 0,
                float(n - 1 - rB), _lib.ptr(colmap_d), _lib.ptr(colmap), _lib.ptr(yy),
                _lib.ptr(ta_d[a:b]) if ta_d is not None else None, cnt, algc, _lib.ptr(z32[a:b]),
                _lib.ptr(z64[a:b]) if z64 is not None else None, self.Y.ld, stream))
        return (z32, z64) if want_f64 else z32

    # -- tm-models cosinor ------------------------------------------------------------------------
    def cosinor_stats(self, X, nper, nexog, perm_idx, mediation_ta=None, alg="aroian", want_f64=False, caller_order=True):
        """Cosinor statistics (pyfunc.py:2406-2563) of the designs X[perm_idx[p]] (cosinor_design's column order): CUDA
        float32 [P, 1 + 2*nper + nexog, ld] with rows [model F, (|t amplitude|, |t acrophase|) per period, t of every
        tested column]; with mediation_ta (path A's amplitude t) the single row calc_indirect(ta, t of tested column
        0), shape [P, 1, ld].  One plain contraction for the betas, then tmb_glm_cosinor_beta per (design, vertex)."""
        import torch
        X = np.asarray(X, dtype=np.float64)
        if not has_intercept(X):
            raise ValueError("X must have the intercept in column 0")
        stack = row_permuted_stack(X, perm_idx)
        P, r, n = stack["pinv"].shape
        if n != self.Y.n:
            raise ValueError("design has %d subjects, data has %d" % (n, self.Y.n))
        if r > MAX_REGRESSORS:
            raise ValueError("at most %d non-intercept regressors per design (got %d)" % (MAX_REGRESSORS, r))
        nrows = 1 if mediation_ta is not None else 1 + 2 * nper + nexog
        G_d = self._upload("G", stack["G"])
        C_d = self._upload("cosC", np.ascontiguousarray(np.linalg.inv(stack["G"])))
        yy = self.Y.sumsq(True)
        sstot = self.Y.sstotal_reference()
        s32 = torch.empty((P, nrows, self.Y.ld), dtype=torch.float32, device=self.device)
        s64 = torch.empty((P, nrows, self.Y.ld), dtype=torch.float64, device=self.device) if want_f64 else None
        for a, cnt, beta in self._betas_chunks(stack["pinv"]):
            _lib.check(_lib.lib().tmb_glm_cosinor_beta(
                _lib.ptr(beta), self.Y.ld, self.Y.V, _lib.ptr(G_d[a:a + cnt]), _lib.ptr(C_d[a:a + cnt]), cnt, r, nper, nexog,
                stack["dof"], _lib.ptr(yy), _lib.ptr(sstot), 0 if mediation_ta is None else 1,
                0.0 if mediation_ta is None else float(mediation_ta), {"aroian": 0, "sobel": 1, "goodman": 2}[alg], _lib.ptr(s32[a:a + cnt]),
                _lib.ptr(s64[a:a + cnt]) if s64 is not None else None, self.Y.ld, 1 if self.nan_to_zero else 0,
                _lib.current_stream()))
        if caller_order and self.colperm is not None:
            s32 = self.to_caller_order(s32)
            s64 = self.to_caller_order(s64) if s64 is not None else None
        return (s32, s64) if want_f64 else s32

    def cosinor_block(self, time_var, period, exog, dmy_covariates, perm_idx, download=True):
        """One block of the tm-models cosinor loop (tmanalysis/tm_models_randomise.py:274-381): model F, per period
        |t amplitude| and |t acrophase| -> one-sided TFCE -> scaled max, float32 [P, 1 + 2*nper, S]; t of every tested
        column with both signs, float32 [P, nexog, S, 2] (None without tested variables)."""
        X, nper, nexog = cosinor_design(time_var, period, exog, dmy_covariates)
        s32 = self.cosinor_stats(X, nper, nexog, perm_idx, caller_order=False)
        P, nrows, ld = s32.shape
        npos = 1 + 2 * nper
        mx, status, _ = self.plan.run(s32[:, :npos].contiguous().view(P * npos, ld), two_sided=False)
        pos = mx.view(P, npos, self.plan.S, 2)[..., 0].contiguous()
        tex = None
        if nexog:
            mx, status, _ = self.plan.run(s32[:, npos:].contiguous().view(P * nexog, ld), two_sided=True)
            tex = mx.view(P, nexog, self.plan.S, 2)
        self.last_status = status
        if download:
            return self._download(pos), (self._download(tex) if tex is not None else None)
        return pos, tex

    def cosinor_mediation_block(self, time_var, period, mediator, perm_idx, alg="aroian", download=True):
        """One block of the cosinor mediation loop (tm_models_randomise.py:383-426): path A = |t amplitude| of the
        un-permuted mediator's own cosinor fit (first period), path B = t of the mediator as the one tested column of
        the data's row-permuted cosinor design; calc_indirect -> one-sided TFCE -> scaled max, float32 [P, S]."""
        ta = cosinor_amplitude_t(mediator, time_var, period)[0]
        X, nper, nexog = cosinor_design(time_var, period, [np.asarray(mediator, dtype=np.float64).reshape(-1)])
        z32 = self.cosinor_stats(X, nper, nexog, perm_idx, mediation_ta=ta, alg=alg, caller_order=False)
        P, _, ld = z32.shape
        mx, status, _ = self.plan.run(z32.view(P, ld), two_sided=False)
        self.last_status = status
        mx = mx[:, :, 0]
        return self._download(mx.contiguous()) if download else mx

    # -- tm-models repeated-measures ANCOVA ----------------------------------------------------------
    def rm_ancova_stats(self, model, shuffles, rand_arrays, want_f64=False, caller_order=True, budget=1.5e9):
        """F statistics of pyfunc.py:1712-2280 reg_rm_ancova_{one,two}_bs_factor for P shuffles of the long-format data
        this engine holds ([intervals*subjects, V]); model: rmancova.RmAncovaModel; shuffles / rand_arrays as in
        RmAncovaModel.operands.  CUDA float32 [P, model.nout, ld] (and float64 when want_f64).  Per chunk of shuffles:
        one plain contraction for the cross-products of the union design (tmb_glm_beta), the order-dependent totals
        (tmb_rm_totals) and the statistics program (tmb_rm_ancova_stats)."""
        import torch
        if model.N != self.Y.n:
            raise ValueError("model has %d long-format rows, data has %d" % (model.N, self.Y.n))
        P = len(shuffles) if shuffles is not None else len(rand_arrays)
        rU, N, ld = model.rU, model.N, self.Y.ld
        dev = getattr(model, "_device_state", None)
        if dev is None or dev[0] != self.device:
            dev = (self.device, torch.from_numpy(model.meta).to(self.device), torch.from_numpy(model.mats).to(self.device),
                   torch.from_numpy(model.consts).to(self.device), torch.from_numpy(model.group_sizes).to(self.device))
            model._device_state = dev
        _, meta_d, mats_d, consts_d, sizes_d = dev
        yy = self.Y.sumsq(True)
        f32 = torch.empty((P, model.nout, ld), dtype=torch.float32, device=self.device)
        f64 = torch.empty((P, model.nout, ld), dtype=torch.float64, device=self.device) if want_f64 else None
        per = max(1, int(budget // (rU * ld * 8)))
        for a in range(0, P, per):
            b = min(P, a + per)
            cnt = b - a
            A, order, grp_rows = model.operands(None if shuffles is None else shuffles[a:b],
                                                None if rand_arrays is None else rand_arrays[a:b])
            rows = cnt * rU
            ldA = round_up(rows, TILE_M)
            At = np.zeros((N, ldA), dtype=np.float64)
            At[:, :rows] = A
            At_d = self._upload("rm_At", At)
            order_d, grp_d = self._upload("rm_order", order), self._upload("rm_grp", grp_rows)
            cross = torch.empty((rows, ld), dtype=torch.float64, device=self.device)
            lib, st = _lib.lib(), _lib.current_stream()
            _lib.check(lib.tmb_glm_beta(_lib.ptr(self.Y.t), self.Y.dtype_code, N, self.Y.V, ld, _lib.ptr(At_d), ldA, rows,
                                        _lib.ptr(cross), ld, st))
            tot = torch.empty((2, cnt, ld), dtype=torch.float64, device=self.device)
            _lib.check(lib.tmb_rm_totals(_lib.ptr(self.Y.t), self.Y.dtype_code, N, self.Y.V, ld, _lib.ptr(order_d),
                                         _lib.ptr(grp_d), _lib.ptr(sizes_d), int(model.group_sizes.shape[0]), cnt,
                                         _lib.ptr(tot[0]), _lib.ptr(tot[1]), ld, st))
            _lib.check(lib.tmb_rm_ancova_stats(
                _lib.ptr(cross), ld, self.Y.V, cnt, _lib.ptr(meta_d), _lib.ptr(model.meta), _lib.ptr(mats_d),
                _lib.ptr(consts_d), _lib.ptr(yy), _lib.ptr(tot[0]), _lib.ptr(tot[1]), ld, _lib.ptr(f32[a:b]),
                _lib.ptr(f64[a:b]) if f64 is not None else None, ld, 1 if self.nan_to_zero else 0, st))
        if caller_order and self.colperm is not None:
            f32 = self.to_caller_order(f32)
            f64 = self.to_caller_order(f64) if f64 is not None else None
        return (f32, f64) if want_f64 else f32

    def rm_ancova_block(self, model, shuffles, rand_arrays, download=True):
        """One block of the repeated-measures ANCOVA permutation loop (tm_models_randomise.py:522-677): every F map ->
        one-sided TFCE -> scaled max, float32 [P, model.nout, S] (rows in the reference's return order, model.names)."""
        f32 = self.rm_ancova_stats(model, shuffles, rand_arrays, caller_order=False)
        P, nout, ld = f32.shape
        mx, status, _ = self.plan.run(f32.view(P * nout, ld), two_sided=False)
        self.last_status = status
        mx = mx.view(P, nout, self.plan.S, 2)[..., 0].contiguous()
        return self._download(mx) if download else mx

    def mediation_blocks(self, medtype, pred_x, depend_y, perm_idx, alg="aroian", block=256):
        """Many mediation shuffles, `block` at a time, software-pipelined like regression_blocks: while block i is swept
        the host builds block i+1's designs (k x k algebra per shuffle: only pred_x is permuted, so X'X changes) and its
        fit is already queued.  Returns float32 [N, S] on the host."""
        import torch
        perm_idx = np.asarray(perm_idx)
        N = perm_idx.shape[0]
        chunks = [(a, min(N, a + block)) for a in range(0, N, block)]
        host = torch.empty((N, self.plan.S, 2), dtype=torch.float32).pin_memory()

        def stage1(a, b):
            z32 = self._ring("z32", (b - a, self.Y.ld), torch.float32)
            if self.sobelz_cross_ok(medtype):
                self.sobelz_cross(medtype, pred_x, depend_y, perm_idx[a:b], alg, out=z32)
                ops = None
            else:
                XA, XB, ta = self.mediation_designs(medtype, pred_x, depend_y, perm_idx[a:b])
                ops = self.sobelz_operands(XA, XB, ta, alg)
                self.sobelz_launch(ops, out=z32)
            return z32, (self.plan.prepare(z32) if self.plan.exact_pow else None), ops

        nxt = stage1(*chunks[0]) if chunks else None
        for ci, (a, b) in enumerate(chunks):
            cur = nxt
            nxt = stage1(*chunks[ci + 1]) if ci + 1 < len(chunks) else None
            z32, tk, _ = cur
            if tk is not None:
                mx, status, _ = self.plan.finish(tk, z32, two_sided=False)
            else:
                mx, status, _ = self.plan.run(z32, two_sided=False, exact_pow=False)
            host[a:b].copy_(mx, non_blocking=True)
            self.d2h_bytes += (b - a) * self.plan.S * 2 * 4
        torch.cuda.current_stream().synchronize()
        self.last_status = None
        return host.numpy()[:, :, 0].copy()

    def mediation_block(self, medtype, pred_x, depend_y, perm_idx, alg="aroian", want_maps=False, download=True):
        """Sobel-z + one-sided TFCE + scaled max for a block of shuffles
        (vertex_tfce_mediation_randomise.py:80-91, pyfunc.py:130-162).  Returns float32 [P, S]."""
        z32 = self.sobelz(medtype, pred_x, depend_y, perm_idx, alg, caller_order=False)
        mx, status, maps = self.plan.run(z32, two_sided=False, want_maps=want_maps)
        self.last_status = status
        mx = mx[:, :, 0]
        if want_maps:
            maps = tuple(self.to_caller_order(m) if m is not None else None for m in maps)
            return mx, self.to_caller_order(z32), maps
        return self._download(mx) if download else mx


def sobelz_single(medtype, pred_x, depend_y, merge_y, alg="aroian"):
    """calc_sobelz drop-in body: one design (identity permutation), float64 [V] on the host."""
    merge_y = np.asarray(merge_y)
    eng = PermutationEngine(merge_y, None)
    n = eng.Y.n
    _, z64 = eng.sobelz(medtype, pred_x, depend_y, np.arange(n)[None, :], alg, want_f64=True)
    return to_host(z64[0, :eng.Y.V])
