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


def write_tmi_binary(path, data, masks, masknames, adjacency, vertices=(), faces=(), surfnames=(), affines=(),
                     column_ids=None, history=("history mode_add 20261017000000 1 1 0 0 1",)):
    """Test-side writer of the binary TMI container, following the header grammar of the reference's
    tm_io.write_tm_filetype (tm_io.py:158-228) and storing the payloads in HEADER order, which is the order the
    reference reader consumes them in (tm_io.py:364-398).  data: float [n_vertices, n_subjects]; masks: bool 3-D arrays;
    adjacency: object arrays of lists."""
    import pickle
    head, payloads = ["tmi", "format binary_little_endian 0.1", "comment made by tests/helpers.py"], []
    d32 = np.asarray(data, dtype=np.float32)
    head += ["element data_array", "dtype float32", "nbytes %d" % d32.nbytes, "datashape %d %d" % d32.shape]
    payloads.append(np.ascontiguousarray(d32.T).tobytes())
    for m, name in zip(masks, masknames):
        m8 = np.asarray(m, dtype=np.uint8)
        head += ["element masking_array", "dtype uint8", "nbytes %d" % m8.nbytes, "nmasked %d" % int(m8.sum()),
                 "maskshape %d %d %d" % m8.shape, "maskname %s" % name]
        payloads.append(np.ascontiguousarray(m8.T).tobytes())
    for a in affines:
        a32 = np.asarray(a, dtype=np.float32)
        head += ["element affine", "dtype float32", "nbytes %d" % a32.nbytes, "affineshape %d %d" % a32.shape]
        payloads.append(np.ascontiguousarray(a32.T).tobytes())
    for v, f, name in zip(vertices, faces, surfnames):
        v32, f32 = np.asarray(v, dtype=np.float32), np.asarray(f, dtype=np.uint32)
        head += ["surfname %s" % name, "element vertex", "dtype float32", "nbytes %d" % v32.nbytes,
                 "vertexshape %d %d" % v32.shape, "element face", "dtype uint32", "nbytes %d" % f32.nbytes,
                 "faceshape %d %d" % f32.shape]
        payloads += [np.ascontiguousarray(v32.T).tobytes(), np.ascontiguousarray(f32.T).tobytes()]
    for adj in adjacency:
        blob = pickle.dumps(adj, protocol=pickle.HIGHEST_PROTOCOL)
        head += ["element adjacency_object", "dtype python_object", "nbytes %d" % len(blob), "adjlength %d" % len(adj)]
        payloads.append(blob)
    if column_ids is not None:
        c = np.asarray(column_ids)
        head += ["element column_id", "dtype %s" % c.dtype, "nbytes %d" % c.nbytes, "listlength %d" % len(c)]
        payloads.append(c.tobytes())
    head += list(history) + ["end_header"]
    with open(path, "wb") as o:
        o.write(("\n".join(head) + "\n").encode("UTF-8"))
        for p in payloads:
            o.write(p)
