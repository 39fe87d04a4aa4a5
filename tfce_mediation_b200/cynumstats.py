"""Drop-in for the reference extension module ``tfce_mediation.cynumstats`` (cynumstats.pyx).

Same names, argument meaning and return shapes/dtypes as the reference:

    cy_lin_lstsqr_mat(X, y)                  cynumstats.pyx:28-29
    calcF(X, y, n, k)                        cynumstats.pyx:31-36
    se_of_slope(num_voxel, invXX, sigma2, k) cynumstats.pyx:47-52
    resid_covars(x_covars, data)             cynumstats.pyx:54-57
    tval_int(X, invXX, y, n, k, numvoxel)    cynumstats.pyx:59-64
    calc_beta_se(x, y, n, num_voxel)         cynumstats.pyx:66-74
    cy_lin_lstsqr_mat_residual(exog, endog)  cynumstats.pyx:109-112

Host arrays in, host arrays out; the per-vertex arithmetic runs on the GPU through the C ABI
(tmb_glm_direct / tmb_glm_beta / tmb_se_of_slope).  Only the k x k normal-equation inverse of the
single design matrix is formed on the host (numpy), exactly where the reference forms it.
The permutation loop does not come through here: it uses the batched engine
(tfce_mediation_b200.engine), which keeps the data resident in HBM.
"""
import numpy as np

from . import _lib
from ._device import DeviceMatrix, dev_empty, to_host


def _pinv(X):
    X = np.asarray(X, dtype=np.float64)
    return np.linalg.inv(X.T.dot(X)).dot(X.T)


def _as_2d(y):
    y = np.asarray(y)
    if y.ndim == 1:
        return y.reshape(-1, 1), True
    return y, False


def _direct(X, y, d=None, dof=1.0, grand_mean=0.0, want=()):
    """Run tmb_glm_direct for one design; returns dict of host arrays named in `want`."""
    import torch
    _lib.require_device()
    X = np.ascontiguousarray(X, dtype=np.float64)
    n, k = X.shape
    y2, was1d = _as_2d(y)
    if y2.shape[0] != n:
        raise ValueError("shapes (%d,%d) and %s not aligned" % (n, k, (y2.shape,)))
    Y = DeviceMatrix(y2)
    V, ld = Y.V, Y.ld
    dev = Y.t.device
    Xd = torch.from_numpy(X).to(dev)
    Pd = torch.from_numpy(np.ascontiguousarray(_pinv(X))).to(dev)
    dd = torch.from_numpy(np.ascontiguousarray(d, dtype=np.float64)).to(dev) if d is not None else None
    out = {}
    beta = dev_empty((k, ld), torch.float64, dev) if "beta" in want else None
    t64 = dev_empty((k, ld), torch.float64, dev) if "t" in want else None
    se32 = dev_empty((k, ld), torch.float32, dev) if "se" in want else None
    r64 = dev_empty((n, ld), torch.float64, dev) if "resid" in want else None
    sse = dev_empty((V,), torch.float64, dev) if "sse" in want else None
    tss = dev_empty((V,), torch.float64, dev) if "tss" in want else None
    _lib.check(_lib.lib().tmb_glm_direct(
        _lib.ptr(Y.t), Y.dtype_code, n, V, ld, _lib.ptr(Xd), _lib.ptr(Pd), k, _lib.ptr(dd), float(dof),
        float(grand_mean), _lib.ptr(beta), _lib.ptr(t64), _lib.ptr(se32), ld, _lib.ptr(r64), None, ld,
        _lib.ptr(sse), _lib.ptr(tss), _lib.current_stream()))
    for name, buf in (("beta", beta), ("t", t64), ("se", se32), ("resid", r64)):
        if buf is not None:
            a = to_host(buf[:, :V])
            out[name] = a[:, 0] if was1d else a
    for name, buf in (("sse", sse), ("tss", tss)):
        if buf is not None:
            a = to_host(buf)
            out[name] = a[0] if was1d else a
    return out


def cy_lin_lstsqr_mat(X, y):
    """(inv(X'X) X') y  ->  float64 [k, V]   (cynumstats.pyx:28-29)."""
    import torch
    _lib.require_device()
    X = np.ascontiguousarray(X, dtype=np.float64)
    n, k = X.shape
    y2, was1d = _as_2d(y)
    Y = DeviceMatrix(y2)
    dev = Y.t.device
    ldA = (k + 127) // 128 * 128
    At = np.zeros((n, ldA), dtype=np.float64)
    At[:, :k] = _pinv(X).T
    Atd = torch.from_numpy(At).to(dev)
    beta = dev_empty((k, Y.ld), torch.float64, dev)
    _lib.check(_lib.lib().tmb_glm_beta(_lib.ptr(Y.t), Y.dtype_code, n, Y.V, Y.ld, _lib.ptr(Atd), ldA, k,
                                       _lib.ptr(beta), Y.ld, _lib.current_stream()))
    a = to_host(beta[:, :Y.V])
    return a[:, 0] if was1d else a


def calcF(X, y, n, k):
    """((TSS-RSS)/(k-1)) / (RSS/(n-k)); TSS about the GRAND mean like the reference (cynumstats.pyx:31-36)."""
    y2, _ = _as_2d(y)
    grand = float(np.mean(y2))
    r = _direct(X, y, grand_mean=grand, want=("sse", "tss"))
    RSS, TSS = r["sse"], r["tss"]
    with np.errstate(divide="ignore", invalid="ignore"):
        return ((TSS - RSS) / (k - 1)) / (RSS / (n - k))


def se_of_slope(num_voxel, invXX, sigma2, k):
    """float32 [k, num_voxel]: sqrt(diag(sigma2[j] * invXX))   (cynumstats.pyx:47-52)."""
    import torch
    _lib.require_device()
    dev = torch.device("cuda", torch.cuda.current_device())
    s = torch.from_numpy(np.ascontiguousarray(sigma2, dtype=np.float64)).to(dev)
    d = torch.from_numpy(np.ascontiguousarray(np.diag(np.asarray(invXX, dtype=np.float64))[:k])).to(dev)
    V = int(num_voxel)
    se = dev_empty((k, V), torch.float32, dev)
    _lib.check(_lib.lib().tmb_se_of_slope(_lib.ptr(s), V, _lib.ptr(d), int(k), _lib.ptr(se), V,
                                          _lib.current_stream()))
    return to_host(se)


def resid_covars(x_covars, data):
    """data.T - x_covars (pinv(x_covars) data.T)  ->  float64 [n, V]   (cynumstats.pyx:54-57)."""
    return _direct(x_covars, np.asarray(data).T, want=("resid",))["resid"]


def tval_int(X, invXX, y, n, k, numvoxel):
    """float64 [k, V] t-values with the reference's float32 rounding of se (cynumstats.pyx:59-64)."""
    d = np.diag(np.asarray(invXX, dtype=np.float64))
    return _direct(X, y, d=d, dof=float(n - k), want=("t",))["t"]


def calc_beta_se(x, y, n, num_voxel):
    """X = [1, x]; returns (beta[1] float64 [V], se float32 [k, V])   (cynumstats.pyx:66-74)."""
    X = np.column_stack([np.ones(n), x])
    invXX = np.linalg.inv(np.dot(X.T, X))
    k = X.shape[1]
    r = _direct(X, y, d=np.diag(invXX), dof=float(n - k), want=("beta", "se"))
    return r["beta"][1], r["se"]


def cy_lin_lstsqr_mat_residual(exog_vars, endog_arr):
    """(beta float64 [k, V], sum of squared residuals [V])   (cynumstats.pyx:109-112)."""
    r = _direct(exog_vars, endog_arr, want=("beta", "sse"))
    return r["beta"], r["sse"]
