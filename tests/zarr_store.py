"""Test infrastructure: writes a Zarr V3 directory store with the layout the reference's zarr backend produces
(data-beans/src/sparse_backend/zarr.rs:31-64 keys, :285-310 `new_filled_vector`: 1-D arrays, regular chunk grid of
`chunk_elems(n, elem_bytes)` elements (utilities/io_helpers.rs:105-115), bytes(little) + zstd level 5 without checksum,
fill NaN for f32 / 0 for u64; root attributes nrow / ncol / nnz :515-523).  zarrs 0.23 itself is not in this image, so
the metadata follows the Zarr V3 specification and zarrs' defaults (default chunk-key encoding, "/" separator; the final
chunk padded to full length with the fill value).  Compression through pyarrow's bundled libzstd."""
import json
import os

import numpy as np
import pyarrow as pa

TARGET_CHUNK_BYTES, MIN_CHUNK_ELEMS = 1024 * 1024, 8192


def chunk_elems(nelem, elem_bytes):
    return min(max(TARGET_CHUNK_BYTES // max(elem_bytes, 1), MIN_CHUNK_ELEMS), max(nelem, 1))


def write_array(path, vec, chunk=None, compress=True, skip_chunks=()):
    vec = np.ascontiguousarray(vec)
    dt = {"uint64": "uint64", "float32": "float32"}[vec.dtype.name]
    fill = "NaN" if dt == "float32" else 0
    c = chunk or chunk_elems(len(vec), vec.itemsize)
    os.makedirs(os.path.join(path, "c"), exist_ok=True)
    codecs = [{"name": "bytes", "configuration": {"endian": "little"}}]
    if compress:
        codecs.append({"name": "zstd", "configuration": {"level": 5, "checksum": False}})
    meta = {"zarr_format": 3, "node_type": "array", "shape": [int(len(vec))], "data_type": dt,
            "chunk_grid": {"name": "regular", "configuration": {"chunk_shape": [int(c)]}},
            "chunk_key_encoding": {"name": "default", "configuration": {"separator": "/"}},
            "fill_value": fill, "codecs": codecs, "attributes": {}}
    with open(os.path.join(path, "zarr.json"), "w") as f:
        json.dump(meta, f)
    pad = np.full(c, np.nan if dt == "float32" else 0, vec.dtype)
    for i in range((len(vec) + c - 1) // c):
        if i in skip_chunks:
            continue
        part = vec[i * c:(i + 1) * c]
        if len(part) < c:
            part = np.concatenate([part, pad[len(part):]])
        raw = part.tobytes()
        if compress:
            raw = pa.compress(raw, codec="zstd", asbytes=True)
        with open(os.path.join(path, "c", str(i)), "wb") as f:
            f.write(raw)


def write_store(root, indptr, indices, data, nrow, chunk=None, compress=True):
    """root/{zarr.json, by_column/{indptr,indices,data}}"""
    os.makedirs(os.path.join(root, "by_column"), exist_ok=True)
    ncol = len(indptr) - 1
    with open(os.path.join(root, "zarr.json"), "w") as f:
        json.dump({"zarr_format": 3, "node_type": "group",
                   "attributes": {"nrow": int(nrow), "ncol": int(ncol), "nnz": int(len(data))}}, f)
    with open(os.path.join(root, "by_column", "zarr.json"), "w") as f:
        json.dump({"zarr_format": 3, "node_type": "group", "attributes": {}}, f)
    write_array(os.path.join(root, "by_column", "indptr"), np.asarray(indptr, np.uint64), chunk, compress)
    write_array(os.path.join(root, "by_column", "indices"), np.asarray(indices, np.uint64), chunk, compress)
    write_array(os.path.join(root, "by_column", "data"), np.asarray(data, np.float32), chunk, compress)
