"""Column-block ingest: lg_zarr_* reads the store layout of the reference's zarr backend
(data-beans/src/sparse_backend/zarr.rs).  The host half (metadata, chunk arithmetic, zstd) runs without a GPU; the device
half checks that the block lg_zarr_read_columns hands to the path is the block a direct lg_csc_upload gives.
Parity unpinned for this row: no store written by the reference's own zarrs exists in this environment (tests/zarr_store.py
writes the layout from the Zarr V3 specification)."""
import json
import os

import numpy as np
import pytest

import legume_b200 as lg
from zarr_store import chunk_elems, write_array, write_store


def random_csc(D, N, density, seed):
    rng = np.random.default_rng(seed)
    cnt = rng.binomial(D, density, N)
    cnt[rng.integers(0, N, max(N // 10, 1))] = 0  # empty columns too
    indptr = np.zeros(N + 1, np.uint64)
    indptr[1:] = np.cumsum(cnt)
    idx = np.concatenate([np.sort(rng.choice(D, c, replace=False)) for c in cnt] + [np.zeros(0, np.int64)]).astype(np.uint64)
    val = rng.poisson(0.3, len(idx)).astype(np.float32) + 1
    return indptr, idx, val


def test_chunk_elems_follows_the_reference():
    # utilities/io_helpers.rs:105-115 and its tests: 1 MiB target, never below 8192 elements, never above the array
    assert chunk_elems(10_000_000, 4) == 262_144
    assert chunk_elems(10_000_000, 8) == 131_072
    assert chunk_elems(100, 8) == 100
    assert chunk_elems(0, 4) == 1
    assert chunk_elems(10_000_000, 1024) == 8192


@pytest.mark.parametrize("compress", [True, False])
def test_zarr_host_read_round_trip(tmp_path, compress):
    D, N = 700, 900
    ip, ix, v = random_csc(D, N, 0.05, 1)
    root = str(tmp_path / "m.zarr")
    write_store(root, ip, ix, v, D, chunk=1000, compress=compress)  # many chunks, ragged last one
    be = lg.SparseMtxData.open(root)
    assert (be.num_rows(), be.num_columns(), be.num_non_zeros()) == (D, N, len(v))
    assert be.csc_column_arrays() is None
    be.preload_columns()
    gip, gix, gv = be.csc_column_arrays()
    assert np.array_equal(gip, ip) and np.array_equal(gix, ix) and gv.tobytes() == v.tobytes()
    # column ranges that start / end inside chunks, the empty range, single columns
    for lo, hi in [(0, 1), (17, 403), (402, 403), (N - 5, N), (250, 250), (0, N)]:
        rip, rix, rv = be.read_columns_host(lo, hi)
        a, b = int(ip[lo]), int(ip[hi])
        assert np.array_equal(rip, ip[lo:hi + 1] - ip[lo])
        assert np.array_equal(rix, ix[a:b]) and rv.tobytes() == v[a:b].tobytes()
    be.close()


def _zip_store(root, zip_path, prefix, method, pad_entries=0):
    """the archive common_io.rs:591-640 writes from a store directory: entries sorted, directories included, under `prefix`"""
    import zipfile
    with zipfile.ZipFile(zip_path, "w", compression=method) as zf:
        for dirpath, dirnames, filenames in sorted(os.walk(root)):
            rel = os.path.relpath(dirpath, root)
            rel = "" if rel == "." else rel + "/"
            if rel:
                zf.writestr(prefix + rel, b"")  # add_directory
            for fn in sorted(filenames):
                zf.write(os.path.join(dirpath, fn), prefix + rel + fn)
        for i in range(pad_entries):  # more than 65 535 entries: the ZIP64 end-of-central-directory record
            zf.writestr(f"{prefix}pad/{i:06d}", b"")


@pytest.mark.parametrize("form", ["stem_stored", "legacy_deflated", "bare_stored", "zip64"])
def test_zarr_zip_store_reads_in_place(tmp_path, form):
    """`.zarr.zip` (zarr_io.rs:30-85: the reference opens the archive in place through zarrs' ZipStorageAdapter, entries
    under `<stem>/`, formerly `<stem>.zarr/`, or bare): the same arrays as the directory it was zipped from — stored
    entries as the reference writes them, deflated ones, and an archive large enough for the ZIP64 records"""
    import zipfile
    D, N = 300, 400
    ip, ix, v = random_csc(D, N, 0.06, 7)
    root = str(tmp_path / "m.zarr")
    write_store(root, ip, ix, v, D, chunk=500)
    zpath = str(tmp_path / "m.zarr.zip")
    prefix, method, pad = {"stem_stored": ("m/", zipfile.ZIP_STORED, 0), "legacy_deflated": ("m.zarr/", zipfile.ZIP_DEFLATED, 0),
                           "bare_stored": ("", zipfile.ZIP_STORED, 0), "zip64": ("m/", zipfile.ZIP_STORED, 66000)}[form]
    _zip_store(root, zpath, prefix, method, pad)
    import shutil
    shutil.rmtree(root)  # finalize_zarr_output removes the directory: only the archive is left
    be = lg.SparseMtxData.open(zpath)
    assert (be.num_rows(), be.num_columns(), be.num_non_zeros()) == (D, N, len(v))
    gip, gix, gv = be.read_columns_host()
    assert np.array_equal(gip, ip) and np.array_equal(gix, ix) and gv.tobytes() == v.tobytes()
    rip, rix, rv = be.read_columns_host(33, 257)
    a, b = int(ip[33]), int(ip[257])
    assert np.array_equal(rip, ip[33:258] - ip[33]) and np.array_equal(rix, ix[a:b]) and rv.tobytes() == v[a:b].tobytes()
    be.close()


def test_zarr_zip_errors(tmp_path):
    bad = tmp_path / "x.zarr.zip"
    bad.write_bytes(b"this is not an archive, only a file whose name ends in .zip")
    with pytest.raises(lg.LegumeError, match="zip"):
        lg.SparseMtxData.open(str(bad))
    import zipfile
    empty = str(tmp_path / "e.zarr.zip")
    with zipfile.ZipFile(empty, "w") as zf:
        zf.writestr("e/readme.txt", b"no hierarchy here")
    with pytest.raises(lg.LegumeError, match="zarr.json"):
        lg.SparseMtxData.open(empty)


def test_zarr_default_chunking_and_missing_chunk(tmp_path):
    # the reference's own chunk size (the whole array when it is below 1 MiB); an absent chunk reads as the fill value
    D, N = 300, 400
    ip, ix, v = random_csc(D, N, 0.1, 2)
    root = str(tmp_path / "d.zarr")
    write_store(root, ip, ix, v, D)
    be = lg.SparseMtxData.open(root)
    gip, gix, gv = be.read_columns_host()
    assert np.array_equal(gip, ip) and np.array_equal(gix, ix) and np.array_equal(gv, v)
    be.close()
    write_array(os.path.join(root, "by_column", "data"), v, chunk=4096, skip_chunks=(1,))
    for f in os.listdir(os.path.join(root, "by_column", "data", "c")):
        assert f != "1"
    be = lg.SparseMtxData.open(root)
    _, _, gv = be.read_columns_host()
    assert np.array_equal(gv[:4096], v[:4096]) and np.isnan(gv[4096:8192]).all() and np.array_equal(gv[8192:], v[8192:])
    be.close()


def test_zarr_open_errors(tmp_path):
    with pytest.raises(lg.LegumeError, match="zarr.json"):
        lg.SparseMtxData.open(str(tmp_path / "absent.zarr"))
    D, N = 50, 60
    ip, ix, v = random_csc(D, N, 0.2, 3)
    root = str(tmp_path / "e.zarr")
    write_store(root, ip, ix, v, D, chunk=64)
    meta_path = os.path.join(root, "by_column", "indices", "zarr.json")
    meta = json.load(open(meta_path))
    bad = dict(meta, data_type="int32")
    json.dump(bad, open(meta_path, "w"))
    with pytest.raises(lg.LegumeError, match="data_type"):
        lg.SparseMtxData.open(root)
    bad = dict(meta, codecs=meta["codecs"] + [{"name": "blosc"}])
    json.dump(bad, open(meta_path, "w"))
    with pytest.raises(lg.LegumeError, match="blosc"):
        lg.SparseMtxData.open(root)
    json.dump(meta, open(meta_path, "w"))
    # a chunk that does not inflate to the chunk shape
    with open(os.path.join(root, "by_column", "indices", "c", "0"), "wb") as f:
        f.write(b"not zstd")
    be = lg.SparseMtxData.open(root)
    with pytest.raises(lg.LegumeError, match="c/0"):
        be.read_columns_host()
    with pytest.raises(lg.LegumeError, match="range"):
        be.read_columns_host(0, N + 1)
    be.close()
    # lengths that contradict the attributes
    json.dump({"zarr_format": 3, "node_type": "group", "attributes": {"nrow": D, "ncol": N + 1, "nnz": len(v)}},
              open(os.path.join(root, "zarr.json"), "w"))
    with pytest.raises(lg.LegumeError, match="lengths"):
        lg.SparseMtxData.open(root)


@pytest.mark.gpu
def test_zarr_block_equals_direct_upload(tmp_path):
    import oracle as orc
    D, N, K = 3000, 5000, 50
    ip, ix, v = random_csc(D, N, 0.04, 4)
    root = str(tmp_path / "g.zarr")
    write_store(root, ip, ix, v, D, chunk=20000)
    ctx = lg.Context(0)
    be = lg.SparseMtxData.open(root)
    whole = be.read_columns_csc(ctx)
    gip, gix, gv = whole.download()
    assert np.array_equal(gip, ip) and np.array_equal(gix, ix) and gv.tobytes() == v.tobytes()
    lo, hi = 1024, 4096
    part = be.read_columns_csc(ctx, lo, hi)
    pip, pix, pv = part.download()
    a, b = int(ip[lo]), int(ip[hi])
    assert np.array_equal(pip, ip[lo:hi + 1] - ip[lo]) and np.array_equal(pix, ix[a:b]) and pv.tobytes() == v[a:b].tobytes()
    # and the path runs on it: projection of the ingested block against the oracle on the arrays that were written
    basis = np.random.default_rng(0).standard_normal((D, K)).astype(np.float32)
    data = lg.SparseIoVec.from_zarr(ctx, root)
    _, got = data.project_columns_with_batch_correction(K, basis=basis)
    want = orc.project(ip, ix, v, basis, None, 0, nthreads=4)  # no batch labels on either side: no centring
    got = np.asarray(got.cpu() if hasattr(got, "cpu") else got)
    assert got.shape == want.shape
    assert np.max(np.abs(got - want) / (1 + np.maximum(np.abs(got), np.abs(want)))) <= 1e-5
    be.close()


def test_zarr_edge_shapes(tmp_path):
    # a matrix without any stored entry, and ranges that end exactly on chunk boundaries
    root = str(tmp_path / "empty.zarr")
    write_store(root, np.zeros(6, np.uint64), np.zeros(0, np.uint64), np.zeros(0, np.float32), 9)
    be = lg.SparseMtxData.open(root)
    assert (be.num_rows(), be.num_columns(), be.num_non_zeros()) == (9, 5, 0)
    ip, ix, v = be.read_columns_host()
    assert np.array_equal(ip, np.zeros(6, np.uint64)) and len(ix) == 0 and len(v) == 0
    ip, ix, v = be.read_columns_host(2, 4)
    assert len(ip) == 3 and len(ix) == 0
    be.close()
    D, N = 64, 256
    ip = np.arange(0, (N + 1) * 4, 4, dtype=np.uint64)           # 4 entries per column: column j ends at entry 4 (j + 1)
    ix = np.tile(np.array([1, 7, 30, 63], np.uint64), N)
    v = np.arange(1, 4 * N + 1, dtype=np.float32)
    root = str(tmp_path / "aligned.zarr")
    write_store(root, ip, ix, v, D, chunk=64)                     # 16 columns of entries per chunk, 64 indptr entries per chunk
    be = lg.SparseMtxData.open(root)
    for lo, hi in [(0, 16), (16, 32), (15, 17), (63, 64), (64, 128), (240, 256)]:
        rip, rix, rv = be.read_columns_host(lo, hi)
        assert np.array_equal(rip, ip[lo:hi + 1] - ip[lo])
        assert np.array_equal(rix, ix[4 * lo:4 * hi]) and np.array_equal(rv, v[4 * lo:4 * hi])
    be.close()
