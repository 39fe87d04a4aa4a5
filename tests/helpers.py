"""Shared builders for the parity tests (seeded synthetic inputs; oracle-side reference pipelines)."""
import functools

import numpy as np

import oracle
from tfce_mediation_b200 import synth


@functools.lru_cache(maxsize=None)
def ico(level):
    v, f = synth.icosphere(level)
    csr = synth.faces_to_csr(v.shape[0], f)
    return v, f, csr


def grid_csr(nx, ny, diag=True):
    """8-neighbour (or 4-neighbour) 2-D grid graph."""
    idx = np.arange(nx * ny).reshape(nx, ny)
    adj = [[] for _ in range(nx * ny)]
    for x in range(nx):
        for y in range(ny):
            for dx in (-1, 0, 1):
                for dy in (-1, 0, 1):
                    if (dx or dy) and (diag or dx == 0 or dy == 0):
                        xx, yy = x + dx, y + dy
                        if 0 <= xx < nx and 0 <= yy < ny:
                            adj[idx[x, y]].append(int(idx[xx, yy]))
    return adj


def smooth_map(csr, seed, rounds=3, scale=1.0):
    rs = np.random.RandomState(seed)
    V = csr[0].shape[0] - 1
    m = rs.standard_normal((1, V)).astype(np.float32)
    if rounds:
        m = synth.smooth_columns(m, csr, rounds)
        m = m / m.std()
    return np.ascontiguousarray(m[0] * scale, dtype=np.float32)


def oracle_run(H, E, csr):
    def run(image, enhn):
        oracle.tfce_run(H, E, csr, image, enhn)
    return run


def oracle_signed_max(H, E, csr, stat, weight=None):
    """(+max, -max) of fl32(fl32(tfce * fl32(max/100)) * w) exactly as pyfunc.py:116-118 composes it."""
    out = []
    for sgn in (1.0, -1.0):
        img = np.ascontiguousarray(stat * np.float32(sgn), dtype=np.float32)
        mx = img.max()
        if not (mx > 0):
            out.append(np.float32(0))
            continue
        tf = oracle.tfce_run(H, E, csr, img)
        val = tf * (mx / 100)
        if weight is not None:
            val = val * weight
        out.append(np.float32(val.max()))
    return out

