"""TMI container reader -- drop-in for the reference's tm_io.read_tm_filetype (tm_io.py:284-444), the real-data entry of
mmr / mmr-lr.

File layout (as written by tm_io.write_tm_filetype, tm_io.py:72-282): a text header

    tmi
    format {binary_little_endian|ascii} <version>
    comment ...
    element <kind>          kind: data_array | masking_array | affine | vertex | face | adjacency_object | column_id
    dtype <numpy dtype | python_object>
    nbytes <payload bytes>
    <kind-specific keys: datashape r c | nmasked n, maskshape x y z, maskname s | affineshape r c |
                         vertexshape r c | faceshape r c | adjlength n | listlength n>      (surfname precedes its vertex element)
    history mode_add <time> <ndata> <nmask> <naffine> <nobject> <nadjacency>
    end_header

followed, for the binary format, by the payloads in header order occupying exactly the LAST sum(nbytes) bytes of the
file; 2-D / 3-D arrays are stored transposed (Fortran order of the logical shape), adjacency objects as pickles.
Host-side I/O only: nothing here touches the GPU."""
import os
import pickle

import numpy as np

_SHAPE_KEYS = {"datashape": 2, "maskshape": 3, "affineshape": 2, "vertexshape": 2, "faceshape": 2}


def _parse_header(f):
    """Returns (format, records, masknames, surfnames, history, header_lines).  A record is a dict with the element's
    kind and the keys that followed it up to the next element."""
    first = f.readline().decode("UTF-8").strip().split()
    if not first or first[0] != "tmi":
        raise ValueError("not a TFCE_mediation image (first header word %r)" % (first[0] if first else ""))
    second = f.readline().decode("UTF-8").strip().split()
    if not second or second[0] != "format":
        raise ValueError("unknown reading file format (second header line %r)" % " ".join(second))
    fmt = second[1]
    records, masknames, surfnames, history = [], [], [], []
    cur = None
    while True:
        raw = f.readline()
        if not raw:
            raise ValueError("end of file inside the TMI header (no end_header)")
        words = raw.decode("UTF-8").strip().split()
        if not words:
            continue
        key = words[0]
        if key == "end_header":
            break
        if key == "element":
            cur = {"kind": words[1]}
            records.append(cur)
        elif key == "history":
            history.append(" ".join(words))
        elif key == "surfname":
            surfnames.append(words[1])
        elif key == "maskname":
            masknames.append(words[1])
        elif cur is not None:
            if key in _SHAPE_KEYS:
                cur["shape"] = tuple(int(w) for w in words[1:1 + _SHAPE_KEYS[key]])
            elif key in ("nbytes", "nmasked", "adjlength", "listlength"):
                cur[key] = int(words[1])
            elif key == "dtype":
                cur["dtype"] = words[1]
    return fmt, records, masknames, surfnames, history


def _unflatten(flat, shape):
    """Payload arrays are the transposed logical array flattened in C order (tm_io.py:246-268 writes `.T`)."""
    n = int(np.prod(shape))
    return np.array(flat[:n]).reshape(shape[::-1]).T


def read_tm_filetype(tm_file, verbose=True):
    """tm_io.py:284-444.  Returns the reference's 11-tuple
    (element names, [data n_vertices x n_subjects], [bool masks], masknames, [affines], [vertices], [faces], surfnames,
     [adjacency object arrays], history lines, [column ids])."""
    filesize = os.stat(tm_file).st_size
    o_img, o_mask, o_affine, o_vertex, o_face, o_adj, o_cols = [], [], [], [], [], [], []
    with open(tm_file, "rb") as f:
        fmt, records, masknames, surfnames, history = _parse_header(f)
        elements = [r["kind"] for r in records]
        if fmt == "binary_little_endian":
            position = filesize - sum(r.get("nbytes", 0) for r in records)
            for r in records:
                kind = r["kind"]
                if verbose:
                    print(position)
                    print("reading %s" % kind)
                f.seek(position)
                if kind == "adjacency_object":
                    obj = pickle.load(f)
                    o_adj.append(np.array(obj[:r["adjlength"]]))
                else:
                    dt = np.dtype(r["dtype"])
                    flat = np.frombuffer(f.read(r["nbytes"]), dtype=dt)
                    if kind == "data_array":
                        o_img.append(_unflatten(flat, r["shape"]))
                    elif kind == "masking_array":
                        o_mask.append(np.array(_unflatten(flat, r["shape"]), dtype=bool))
                    elif kind == "affine":
                        o_affine.append(_unflatten(flat, r["shape"]))
                    elif kind == "vertex":
                        o_vertex.append(_unflatten(flat, r["shape"]))
                    elif kind == "face":
                        o_face.append(_unflatten(flat, r["shape"]))
                    elif kind == "column_id":
                        o_cols.append(np.array(flat[:r["listlength"]]))
                position += r.get("nbytes", 0)
        elif fmt == "ascii":
            def rows(count, dtype):
                return [np.array(f.readline().strip().split()).astype(dtype) for _ in range(count)]
            for r in records:
                kind = r["kind"]
                if kind == "data_array":
                    o_img.append(np.array(rows(r["shape"][0], "float32"), dtype="float32").reshape(r["shape"]))
                elif kind == "masking_array":
                    idx = np.array(rows(r["nmasked"], np.int32))
                    m = np.zeros(r["shape"], dtype=bool)
                    if idx.size:
                        m[idx[:, 0], idx[:, 1], idx[:, 2]] = True
                    o_mask.append(m)
                elif kind == "affine":
                    o_affine.append(np.array(rows(r["shape"][0], "float32"), dtype="float32"))
                elif kind == "vertex":
                    o_vertex.append(np.array(rows(r["shape"][0], "float32"), dtype="float32"))
                elif kind == "face":
                    o_face.append(np.array(rows(r["shape"][0], "int32"), dtype="int32"))
                elif kind == "column_id":
                    o_cols.append(np.array([f.readline().strip() for _ in range(r["listlength"])], dtype=r["dtype"]))
        else:
            raise ValueError("Error unknown filetype: %s" % fmt)
    return (elements, o_img, o_mask, masknames, o_affine, o_vertex, o_face, surfnames, o_adj, history, o_cols)
