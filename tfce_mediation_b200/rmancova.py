"""Host model of the repeated-measures ANCOVA statistics of the reference's tm-models
(pyfunc.py:1712-2280 reg_rm_ancova_one_bs_factor / reg_rm_ancova_two_bs_factor, permutation loop
tmanalysis/tm_models_randomise.py:522-677) for the device path in csrc/rmancova_kernels.cu.

The reference runs, per shuffle, a chain of long-format regressions (one row per interval and subject) and combines
their residual sums of squares into Type I F statistics.  Every design of that chain is a subset of the columns of one
UNION design -- between-subject factors, their interaction, the interval dummies, the factor x interval products, the
covariates and covariate x interval products -- and a shuffle only permutes rows, which leaves Z'Z (Z: the centred union
design) unchanged.  So this model holds, once: Z, for every design its column set S and inv(G_SS), and a short program
that restates the reference's sequence of subtractions and divisions over the residual sums of squares; per shuffle it
only builds the row permutations.  k x k host algebra only; the per-vertex work is on the GPU."""
import numpy as np

MAX_COLUMNS = 64            # kRmMaxCols in csrc/rmancova_kernels.cu
MAX_REGISTERS = 96          # kRmMaxRegs
_DESIGN_STRIDE = 2 + MAX_COLUMNS
OP_SUB, OP_ADD, OP_DIVC, OP_DIV, OP_ZERO = 0, 1, 2, 3, 4
T, W = 0, 1                 # registers: SS_Total, residual of the subject-dummy regression


def _two_d(x, rows):
    return np.asarray(x, dtype=np.float64).reshape(rows, -1)


def _products(a, b):
    """Every column of a times every column of b, a-major (pyfunc.py:2622-2660 column_product)."""
    return np.concatenate([a[:, i:i + 1] * b for i in range(a.shape[1])], axis=1)


def interval_dummies(n, s):
    """Long-format interval dummies [s*n, s-1], interval 0 the reference level (pyfunc.py:2156-2166)."""
    out = np.zeros((s * n, s - 1))
    for t in range(1, s):
        out[t * n:(t + 1) * n, t - 1] = 1.0
    return out


class _Program(object):
    """Straight-line program over float64 registers; see tmb_rm_ancova_stats in include/tfce_b200.h."""

    def __init__(self, designs):
        self.ops, self.consts, self.next = [], [], 2 + designs

    def _emit(self, op, a, b):
        dst = self.next
        self.next += 1
        if dst >= MAX_REGISTERS:
            raise ValueError("repeated-measures program needs more than %d registers" % MAX_REGISTERS)
        self.ops.append((op, dst, a, b))
        return dst

    def sub(self, a, b):
        return self._emit(OP_SUB, a, b)

    def div(self, a, b):
        return self._emit(OP_DIV, a, b)

    def divc(self, a, value):
        self.consts.append(float(value))
        return self._emit(OP_DIVC, a, len(self.consts) - 1)

    def zero(self):
        return self._emit(OP_ZERO, 0, 0)


class RmAncovaModel(object):
    """factors: [dmy_factor1] (one between-subject factor, outputs F_a, F_s, F_sa: pyfunc.py:2068-2280) or
    [dmy_factor1, dmy_factor2] (outputs F_a, F_b, F_ab, F_s, F_sa, F_sb, F_sab: pyfunc.py:1712-2052), each [n] or [n, k];
    dmy_subjects [n, ...] dummy-coded subjects (equal rows = same subject); dmy_covariates [n] / [n, c] or None;
    s: number of intervals.  Long-format rows are interval-major: row t*n + m is subject slot m at interval t."""

    def __init__(self, n, s, factors, dmy_subjects, dmy_covariates=None):
        n, s = int(n), int(s)
        if len(factors) not in (1, 2):
            raise ValueError("one or two between-subject factors")
        if s < 2:
            raise ValueError("repeated-measures ANCOVA needs at least two intervals")
        self.n, self.s, self.N = n, s, n * s
        self.two = len(factors) == 2
        long_ = lambda x: np.concatenate([x] * s, axis=0)  # noqa: E731
        itv = interval_dummies(n, s)
        blocks = {}
        f1 = _two_d(factors[0], n)
        blocks["f1"] = long_(f1)
        if self.two:
            f2 = _two_d(factors[1], n)
            blocks["f2"], blocks["f12"] = long_(f2), long_(_products(f1, f2))
        blocks["itv"] = itv
        blocks["f1i"] = _products(blocks["f1"], itv)
        if self.two:
            blocks["f2i"], blocks["f12i"] = _products(blocks["f2"], itv), _products(blocks["f12"], itv)
        has_cov = dmy_covariates is not None
        if has_cov:
            cov = _two_d(dmy_covariates, n)
            blocks["cov"] = long_(cov)
            blocks["covi"] = _products(blocks["cov"], itv)
        cols, at = {}, 0
        for name, b in blocks.items():
            cols[name] = list(range(at, at + b.shape[1]))
            at += b.shape[1]
        U = np.column_stack(list(blocks.values()))
        self.rU = U.shape[1]
        if self.rU > MAX_COLUMNS:
            raise ValueError("repeated-measures ANCOVA: %d design columns, at most %d" % (self.rU, MAX_COLUMNS))
        self.Z = np.ascontiguousarray(U - U.mean(axis=0))
        G = self.Z.T @ self.Z
        # subjects: equal rows of the dummy coding are the same subject
        _, gid = np.unique(_two_d(dmy_subjects, n), axis=0, return_inverse=True)
        gid_long = np.concatenate([np.asarray(gid).reshape(-1)] * s)
        self.group_positions = np.argsort(gid_long, kind="stable")
        self.group_sizes = np.bincount(gid_long).astype(np.int32)

        tail, tail_i = (["cov"], ["cov", "covi"]) if has_cov else ([], [])
        if not self.two:
            names = [["f1"] + tail, ["itv"], ["f1", "itv", "f1i"] + tail_i, ["f1", "itv"] + tail_i]
        else:
            names = [["f1", "f2", "f12"] + tail, ["f1", "f2"] + tail, ["f1"] + tail, ["itv"],
                     ["f1", "f2", "f12", "itv", "f1i", "f2i", "f12i"] + tail_i,
                     ["f1", "f2", "f12", "itv", "f1i", "f2i"] + tail_i,
                     ["f1", "itv", "f1i"] + tail_i, ["f1", "itv"] + tail_i,
                     ["f1", "f2", "itv", "f1i", "f2i"] + tail_i, ["f1", "f2", "itv", "f1i"] + tail_i]
        if has_cov:
            names += [["cov"], ["cov", "covi", "itv"], ["cov", "itv"]]
        D = len(names)
        R = lambda i: 2 + i  # noqa: E731
        meta = np.zeros(8 + D * _DESIGN_STRIDE, dtype=np.int32)
        mats = []
        moff = 0
        for d, blk in enumerate(names):
            S = [c for b in blk for c in cols[b]]
            base = 8 + d * _DESIGN_STRIDE
            meta[base], meta[base + 1] = len(S), moff
            meta[base + 2:base + 2 + len(S)] = S
            mats.append(np.linalg.inv(G[np.ix_(S, S)]).reshape(-1))
            moff += len(S) ** 2

        df_a = f1.shape[1] if np.ndim(factors[0]) > 1 else 1
        df_cov = 0 if not has_cov else (cov.shape[1] if np.ndim(dmy_covariates) > 1 else 1)
        df_s = s - 1
        p = _Program(D)
        between = p.sub(T, W)                                   # SS_BetweenSubjects
        if not self.two:                                        # pyfunc.py:2168-2233
            cells_a = p.sub(T, R(0))
            covars = p.sub(T, R(D - 3)) if has_cov else p.zero()
            ss_a = p.sub(cells_a, covars)
            within_f = p.sub(p.sub(between, ss_a), covars)
            ss_s = p.sub(T, R(1))
            if has_cov:
                scov = p.sub(p.sub(T, R(D - 2)), p.sub(T, R(D - 1)))
            else:
                scov = p.zero()
            ss_sa = p.sub(p.sub(T, R(2)), p.sub(T, R(3)))
            s_within_f = p.sub(p.sub(p.sub(W, ss_s), ss_sa), scov)
            df_wf = (n - 1) - df_a - df_cov
            ms_w, ms_sw = p.divc(within_f, df_wf), p.divc(s_within_f, df_wf * df_s)
            outs = [p.div(p.divc(ss_a, df_a), ms_w), p.div(p.divc(ss_s, df_s), ms_sw),
                    p.div(p.divc(ss_sa, df_a * df_s), ms_sw)]
            self.names = ("a", "s", "sa")
        else:                                                   # pyfunc.py:1856-2028
            df_b = f2.shape[1] if np.ndim(factors[1]) > 1 else 1
            df_ab = df_a * df_b
            cells_ab = p.sub(T, R(0))
            ss_ab = p.sub(cells_ab, p.sub(T, R(1)))
            covars = p.sub(T, R(D - 3)) if has_cov else p.zero()
            ss_a = p.sub(p.sub(T, R(2)), covars)
            ss_b = p.sub(p.sub(p.sub(cells_ab, covars), ss_ab), ss_a)
            within_f = p.sub(p.sub(p.sub(p.sub(between, ss_a), ss_b), ss_ab), covars)
            ss_s = p.sub(T, R(3))
            ss_sab = p.sub(p.sub(T, R(4)), p.sub(T, R(5)))
            if has_cov:
                scov = p.sub(p.sub(T, R(D - 2)), p.sub(T, R(D - 1)))
            else:
                scov = p.zero()
            ss_sa = p.sub(p.sub(T, R(6)), p.sub(T, R(7)))
            ss_sb = p.sub(p.sub(T, R(8)), p.sub(T, R(9)))
            s_within_f = p.sub(p.sub(p.sub(p.sub(p.sub(W, ss_s), ss_sa), ss_sb), ss_sab), scov)
            df_wf = (n - 1) - df_a - df_b - df_ab - df_cov
            ms_w, ms_sw = p.divc(within_f, df_wf), p.divc(s_within_f, df_wf * df_s)
            F = lambda ss, df, ms: p.div(p.divc(ss, df), ms)  # noqa: E731
            outs = [F(ss_a, df_a, ms_w), F(ss_b, df_b, ms_w), F(ss_ab, df_ab, ms_w), F(ss_s, df_s, ms_sw),
                    F(ss_sa, df_a * df_s, ms_sw), F(ss_sb, df_b * df_s, ms_sw), F(ss_sab, df_ab * df_s, ms_sw)]
            self.names = ("a", "b", "ab", "s", "sa", "sb", "sab")
        meta[0], meta[1], meta[2], meta[3] = D, len(p.ops), len(outs), self.rU
        self.nout = len(outs)
        self.meta = np.concatenate([meta, np.asarray(p.ops, dtype=np.int32).reshape(-1),
                                    np.asarray(outs, dtype=np.int32)]).astype(np.int32)
        self.mats = np.ascontiguousarray(np.concatenate(mats))
        self.consts = np.asarray(p.consts, dtype=np.float64)
        self.df = dict(a=df_a, s=df_s, within_factors=df_wf, covariates=df_cov)
        if self.two:
            self.df.update(b=df_b, ab=df_ab)

    # -- per shuffle ---------------------------------------------------------------------------------
    def operands(self, shuffles, rand_arrays):
        """shuffles int [P, N]: shuffled data row i = original row shuffles[p, i] (the cumulative state of the
        reference's in-place np.random.shuffle; None = identity); rand_arrays int [P, n]: the subject permutation
        applied to factors and covariates (None = identity).  Returns (A float64 [N, P*rU] with column p*rU + j =
        column j of the centred union design as the ORIGINAL data rows see it, order int32 [P, N],
        grp_rows int32 [P, N])."""
        n, s, N, rU = self.n, self.s, self.N, self.rU
        P = len(shuffles) if shuffles is not None else len(rand_arrays)
        A = np.empty((N, P * rU))
        order = np.empty((P, N), dtype=np.int32)
        grp_rows = np.empty((P, N), dtype=np.int32)
        offs = (np.arange(s) * n)[:, None]
        for p in range(P):
            pi = np.arange(N) if shuffles is None else np.asarray(shuffles[p])
            rand = np.arange(n) if rand_arrays is None else np.asarray(rand_arrays[p])
            tau = (offs + rand[None, :]).reshape(-1)             # design row i of the shuffled problem = base row tau[i]
            A[pi, p * rU:(p + 1) * rU] = self.Z[tau]
            order[p] = pi
            grp_rows[p] = pi[self.group_positions]
        return A, order, grp_rows
