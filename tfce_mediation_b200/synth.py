"""Synthetic inputs of the BASELINE.json shapes.

There is no network and the reference's large adjacency/surface blobs were stripped, so the
bench and the tests build stand-ins: icosphere meshes with fsaverage's topology
(level 5 = 10,242 vertices = fsaverage5, level 7 = 163,842 = fsaverage), 1-ring or k-ring
adjacency, skeleton-like voxel masks, smoothed white-noise subject data and N(0,1) designs.
Host-side numpy/scipy only; nothing here is on the timed path.
"""
import numpy as np


def icosphere(level):
    """Unit icosphere after `level` 4-to-1 subdivisions: (vertices float64[V,3], faces int32[F,3]).
    V = 10*4**level + 2 (level 5 -> 10,242; level 7 -> 163,842)."""
    t = (1.0 + 5 ** 0.5) / 2.0
    v = np.array([[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0], [0, -1, t], [0, 1, t],
                  [0, -1, -t], [0, 1, -t], [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]], dtype=np.float64)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    f = np.array([[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11], [1, 5, 9], [5, 11, 4],
                  [11, 10, 2], [10, 7, 6], [7, 1, 8], [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8],
                  [3, 8, 9], [4, 9, 5], [2, 4, 11], [6, 2, 10], [8, 6, 7], [9, 8, 1]], dtype=np.int64)
    for _ in range(level):
        nv = v.shape[0]
        e = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]], axis=0)
        es = np.sort(e, axis=1)
        key = es[:, 0] * nv + es[:, 1]
        uniq, inv = np.unique(key, return_inverse=True)
        mid = v[uniq // nv] + v[uniq % nv]
        mid /= np.linalg.norm(mid, axis=1, keepdims=True)
        v = np.concatenate([v, mid], axis=0)
        nf = f.shape[0]
        m01, m12, m20 = nv + inv[:nf], nv + inv[nf:2 * nf], nv + inv[2 * nf:]
        f = np.concatenate([np.stack([f[:, 0], m01, m20], 1), np.stack([f[:, 1], m12, m01], 1),
                            np.stack([f[:, 2], m20, m12], 1), np.stack([m01, m12, m20], 1)], axis=0)
    return v, f.astype(np.int32)


def faces_to_csr(num_vertices, faces):
    """1-ring adjacency of a triangle mesh as symmetric CSR (sorted, no self loops) --
    same neighbour sets as pyfunc.create_adjac_vertex (pyfunc.py:37-46)."""
    import scipy.sparse as sp
    f = np.asarray(faces, dtype=np.int64)
    r = np.concatenate([f[:, 0], f[:, 0], f[:, 1], f[:, 1], f[:, 2], f[:, 2]])
    c = np.concatenate([f[:, 1], f[:, 2], f[:, 0], f[:, 2], f[:, 0], f[:, 1]])
    m = sp.csr_matrix((np.ones(r.shape[0], dtype=np.int8), (r, c)), shape=(num_vertices, num_vertices))
    m.sum_duplicates()
    m.sort_indices()
    return m.indptr.astype(np.int64), m.indices.astype(np.int32)


def kring_csr(csr, rings):
    """k-ring neighbourhoods (stand-in for the geodesic '3 mm' adjacency sets)."""
    import scipy.sparse as sp
    indptr, indices = csr
    V = indptr.shape[0] - 1
    a = sp.csr_matrix((np.ones(indices.shape[0], dtype=np.float32), indices, indptr), shape=(V, V))
    a = a + sp.identity(V, dtype=np.float32, format="csr")
    m = a.copy()
    for _ in range(rings - 1):
        m = (m @ a)
        m.data[:] = 1
    m.setdiag(0)
    m.eliminate_zeros()
    m.sort_indices()
    return m.indptr.astype(np.int64), m.indices.astype(np.int32)


def csr_to_lists(csr):
    indptr, indices = csr
    return [indices[indptr[i]:indptr[i + 1]].tolist() for i in range(indptr.shape[0] - 1)]


def smooth_columns(data, csr, rounds):
    """data: float32 [n, V]; `rounds` passes of (self + 1-ring) averaging along V."""
    import scipy.sparse as sp
    indptr, indices = csr
    V = indptr.shape[0] - 1
    a = sp.csr_matrix((np.ones(indices.shape[0], dtype=np.float32), indices, indptr), shape=(V, V))
    a = a + sp.identity(V, dtype=np.float32, format="csr")
    deg = np.asarray(a.sum(axis=1)).ravel()
    w = sp.diags(1.0 / deg).dot(a).astype(np.float32).tocsr()
    out = np.ascontiguousarray(data, dtype=np.float32)
    for _ in range(rounds):
        out = np.ascontiguousarray(w.dot(out.T).T, dtype=np.float32)
    return out


def subject_data(n, csr, seed, rounds):
    """White N(0,1) fp32 [n, V], smoothed and standardised per vertex."""
    rs = np.random.RandomState(seed)
    V = csr[0].shape[0] - 1
    y = rs.standard_normal((n, V)).astype(np.float32)
    if rounds:
        y = smooth_columns(y, csr, rounds)
    y -= y.mean(axis=0, keepdims=True)
    y /= y.std(axis=0, keepdims=True)
    return np.ascontiguousarray(y, dtype=np.float32)


def cap_mask(vertices, keep):
    """Boolean mask keeping the `keep` vertices with the largest x coordinate: a 'medial wall'
    style cut-out (the shipped cortex masks keep 149,955 / 149,926 of 163,842)."""
    order = np.argsort(-vertices[:, 0], kind="stable")
    m = np.zeros(vertices.shape[0], dtype=bool)
    m[order[:keep]] = True
    return m


def skeleton_mask(shape=(91, 109, 91), frac=0.29, seed=2, sigma=3.0, margin=8):
    """Skeleton-like voxel mask: |gaussian_filter(randn)| below its `frac` quantile inside a box."""
    from scipy.ndimage import gaussian_filter
    rs = np.random.RandomState(seed)
    g = np.abs(gaussian_filter(rs.standard_normal(shape).astype(np.float32), sigma))
    inner = np.zeros(shape, dtype=bool)
    inner[margin:-margin, margin:-margin, margin:-margin] = True
    thr = np.quantile(g[inner], frac)
    return (g < thr) & inner


def vertex_density(csr):
    """STEP_1_vertex_tfce_multiple_regression.py:166-173 density weights from neighbour counts (float32)."""
    indptr, _ = csr
    d = np.diff(indptr).astype(np.float64)
    return np.array(1 - (d / d.max()) + (d.mean() / d.max()), dtype=np.float32)


def voxel_csr(mask, dirtype=26):
    """26-connectivity adjacency of a 3-D mask as CSR with the semantics of pyfunc.create_adjac_voxel
    (pyfunc.py:48-76): voxels labelled in np.where (C) order, sorted neighbour lists without self, voxel 0
    isolated in both directions.  Vectorised host builder for the synthetic workloads (the product builds it
    on the GPU: tmb_voxel_adjacency)."""
    if int(dirtype) != 26:
        raise ValueError("only 26-connectivity here")
    mask = np.asarray(mask).astype(bool)
    lab = -np.ones(mask.shape, dtype=np.int64)
    V = int(mask.sum())
    lab[mask] = np.arange(V)
    pad = -np.ones(tuple(d + 2 for d in mask.shape), dtype=np.int64)
    pad[1:-1, 1:-1, 1:-1] = lab
    src, dst = [], []
    c = pad[1:-1, 1:-1, 1:-1]
    for dx in (-1, 0, 1):
        for dy in (-1, 0, 1):
            for dz in (-1, 0, 1):
                if dx == 0 and dy == 0 and dz == 0:
                    continue
                nb = pad[1 + dx:pad.shape[0] - 1 + dx, 1 + dy:pad.shape[1] - 1 + dy, 1 + dz:pad.shape[2] - 1 + dz]
                ok = (c > 0) & (nb > 0)          # label 0 neither lists nor is listed (the reference's `> 0` test)
                src.append(c[ok]); dst.append(nb[ok])
    src = np.concatenate(src); dst = np.concatenate(dst)
    order = np.lexsort((dst, src))
    src, dst = src[order], dst[order]
    indptr = np.zeros(V + 1, dtype=np.int64)
    np.add.at(indptr, src + 1, 1)
    np.cumsum(indptr, out=indptr)
    return indptr, dst.astype(np.int32)
