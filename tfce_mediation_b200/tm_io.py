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


# ------------------------------------------------------------------------------------------------------------ writer
def tm_filetype_version():
    """tm_io.py:34-36."""
    return "0.1"


def check_outname(outname):
    """pyfunc.py:315-326: never overwrite silently -- an existing name becomes new_<name> (which IS overwritten)."""
    if os.path.exists(outname):
        outpath, name = os.path.split(outname)
        outname = ("new_%s" % name) if not outpath else ("%s/new_%s" % (outpath, name))
        print("Output file aleady exists. Renaming output file to %s" % outname)
        if os.path.exists(outname):
            print("%s also exists. Overwriting the file." % outname)
            os.remove(outname)
    return outname


def _as_list(x, single_ndim):
    """The reference accepts one array or a list / object array of arrays (tm_io.py:106-139)."""
    if x is None:
        return []
    if isinstance(x, np.ndarray) and x.dtype != object and x.ndim == single_ndim:
        return [x]
    return [np.asarray(a) if not isinstance(a, np.ndarray) else a for a in x]


def _history_counts(tmi_history):
    """Net (masks, affines, objects, adjacencies) already recorded by the history lines (tm_io.py:87-104)."""
    tot = [0, 0, 0, 0]
    for line in tmi_history:
        w = line.split(" ")
        if w[1] in ("mode_add", "mode_sub"):
            sign = 1 if w[1] == "mode_add" else -1
            for i in range(4):
                tot[i] += sign * int(w[4 + i])
        elif w[1] not in ("mode_replace", "mode_reorder"):
            print("Error reading history. Mode %s is not understood. Count is reflect number of element in current file" % w[1])
    return tot


def write_tm_filetype(outname, columnids=[], imgtype=[], checkname=True, output_binary=True, image_array=[],
                      masking_array=[], maskname=[], affine_array=[], vertex_array=[], face_array=[], surfname=[],
                      adjacency_array=[], tmi_history=[], append_history=True):
    """tm_io.py:72-282 write_tm_filetype, binary container: same header grammar, element order, dtypes (float32 data /
    affines / vertices, uint8 masks, uint32 faces, pickled adjacency objects), transposed payloads and history line.
    Two deliberate differences: (1) the reference cannot run under numpy >= 2 (`image_array == []` on arrays,
    ragged `np.array(masks)`), this one can; (2) the reference stores the column ids BEFORE the adjacency objects while
    its header -- and therefore its own reader, which consumes payloads in header order (tm_io.py:364-398) -- lists
    them AFTER; here payloads follow the header order, so files with both are readable by either reader.
    The ascii variant is not built.  Returns the file name written."""
    from time import gmtime, strftime
    if not output_binary:
        raise NotImplementedError("write_tm_filetype: only the binary container is built")
    image = None if (isinstance(image_array, list) and len(image_array) == 0) else np.asarray(image_array)
    masks = _as_list(masking_array, 3)
    affines = _as_list(affine_array, 2)
    vertices = _as_list(vertex_array, 2)
    faces = _as_list(face_array, 2)
    adjacency = list(adjacency_array)      # a list (or object array) of adjacency objects, one per surface (tm_io.py:131-139)
    cols = None if (isinstance(columnids, list) and len(columnids) == 0) else np.asarray(columnids)
    tmi_history = list(tmi_history)
    h_mask, h_affine, h_object, h_adj = _history_counts(tmi_history)
    if not outname.endswith("tmi"):
        outname += ".tmi"
    if checkname:
        outname = check_outname(outname)
    head = ["tmi", "format binary_little_endian %s" % tm_filetype_version(), "comment made with TFCE_mediation"]
    payloads = []
    num_data = 0
    if image is not None:
        num_data = 1
        nvert, nsub = (len(image), 1) if image.ndim == 1 else image.shape
        d32 = image.astype("float32")
        head += ["element data_array", "dtype float32", "nbytes %d" % d32.nbytes, "datashape %d %d" % (nvert, nsub)]
        payloads.append(np.array(d32.T, dtype="float32").tobytes())
    for i, m in enumerate(masks):
        m = np.asarray(m)
        head += ["element masking_array", "dtype uint8", "nbytes %d" % m.astype(np.uint8).nbytes,
                 "nmasked %d" % int((m == True).sum()), "maskshape %d %d %d" % m.shape,  # noqa: E712
                 "maskname %s" % (maskname[i] if len(maskname) > i else "unknown")]
        payloads.append(np.array((m * 1).T, dtype=np.uint8).tobytes())
    for a in affines:
        a = np.asarray(a)
        head += ["element affine", "dtype float32", "nbytes %d" % a.astype("float32").nbytes, "affineshape %d %d" % a.shape]
        payloads.append(np.array(a.T, dtype="float32").tobytes())
    for i, (v, f) in enumerate(zip(vertices, faces)):
        v, f = np.asarray(v), np.asarray(f)
        head += ["surfname %s" % (surfname[i] if len(surfname) > i else "unknown"),
                 "element vertex", "dtype float32", "nbytes %d" % v.astype("float32").nbytes, "vertexshape %d %d" % v.shape,
                 "element face", "dtype uint32", "nbytes %d" % f.astype("uint32").nbytes, "faceshape %d %d" % f.shape]
        payloads += [np.array(v.T, dtype="float32").tobytes(), np.array(f.T, dtype="uint32").tobytes()]
    for adj in adjacency:
        blob = pickle.dumps(adj, protocol=pickle.HIGHEST_PROTOCOL)
        head += ["element adjacency_object", "dtype python_object", "nbytes %d" % len(blob), "adjlength %d" % len(adj)]
        payloads.append(blob)
    if cols is not None:
        head += ["element column_id", "dtype %s" % cols.dtype, "nbytes %d" % cols.nbytes, "listlength %d" % len(cols)]
        payloads.append(cols.tobytes())
    if append_history:
        tmi_history.append("history mode_add %d %d %d %d %d %d" % (
            int(strftime("%Y%m%d%H%M%S", gmtime())), num_data, len(masks) - h_mask, len(affines) - h_affine,
            len(vertices) - h_object, len(adjacency) - h_adj))
    head += tmi_history + ["end_header"]
    with open(outname, "wb") as o:
        o.write(("\n".join(head) + "\n").encode("UTF-8"))
        for p in payloads:
            o.write(p)
    return outname
