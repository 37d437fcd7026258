"""Column-block ingest timed on the GPU box: a store in the reference's zarr layout (tests/zarr_store.py writes it from
simulated counts of configs[1]'s shape) -> lg_zarr_read_columns (chunks inflated on the host cores, block through
lg_csc_upload) -> the hot path.  Prints one JSON line; stage times of the read on stderr."""
import json, os, shutil, sys, tempfile, time
from concurrent.futures import ThreadPoolExecutor
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "legume-rs_b200"), os.path.join(ROOT, "tests")]
import numpy as np, torch
import legume_b200 as lg
from legume_b200 import sim
from legume_b200.pipeline import HotPath
import zarr_store

N = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
D, K, kk = 30000, 50, 10
ctx = lg.Context(0); hp = HotPath(ctx)
tabs = sim.make_tables(D, ntopic=8, nbatch=1, depth=1500, seed=42)
blk, _, _ = sim.sim_block(ctx, tabs, 0, N)
ip, ix, v = blk.download(); blk.free()
root = os.path.join(tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None), "m.zarr")
t0 = time.perf_counter()
# the three arrays are compressed on a pool (pyarrow releases the GIL); same files as zarr_store.write_store
os.makedirs(os.path.join(root, "by_column"), exist_ok=True)
json.dump({"zarr_format": 3, "node_type": "group", "attributes": {"nrow": D, "ncol": N, "nnz": int(len(v))}}, open(os.path.join(root, "zarr.json"), "w"))
def write_sharded(name, vec):
    c = zarr_store.chunk_elems(len(vec), vec.itemsize)
    path = os.path.join(root, "by_column", name)
    zarr_store.write_array(path, vec[:0], chunk=c)            # metadata (rewritten below with the true shape)
    meta = json.load(open(os.path.join(path, "zarr.json"))); meta["shape"] = [int(len(vec))]
    json.dump(meta, open(os.path.join(path, "zarr.json"), "w"))
    import pyarrow as pa
    def one(i):
        part = vec[i * c:(i + 1) * c]
        if len(part) < c:
            part = np.concatenate([part, np.full(c - len(part), np.nan if vec.dtype == np.float32 else 0, vec.dtype)])
        open(os.path.join(path, "c", str(i)), "wb").write(pa.compress(part.tobytes(), codec="zstd", asbytes=True))
    with ThreadPoolExecutor(os.cpu_count()) as ex:
        list(ex.map(one, range((len(vec) + c - 1) // c)))
for name, vec in (("indptr", ip), ("indices", ix), ("data", v)):
    write_sharded(name, vec)
t_write = time.perf_counter() - t0
store_bytes = sum(os.path.getsize(os.path.join(dp, f)) for dp, _, fs in os.walk(root) for f in fs)
def now():
    torch.cuda.synchronize(); return time.perf_counter()
basis = torch.from_numpy(np.random.default_rng(0).standard_normal((D, K)).astype(np.float32)).cuda()
batch = torch.zeros(N, dtype=torch.int32, device="cuda")
os.environ["LG_INGEST_TRACE"] = "1"
rows = []
be = lg.SparseMtxData.open(root)  # one handle for every read, as a backend that lives as long as its SparseIoVec
for it in range(4):
    t0 = now(); b = be.read_columns_csc(ctx); t1 = now()
    o = hp.run(b, basis, batch, 1, kk); t2 = now()
    rows.append({"ingest_ms": 1e3 * (t1 - t0), "path_ms": 1e3 * (t2 - t1)})
    if it == 0:
        gip, gix, gv = b.download()
        same = bool(np.array_equal(gip, ip) and np.array_equal(gix, ix) and gv.tobytes() == v.tobytes())
    b.free()
be.close()
host_bytes = ip.nbytes + ix.nbytes + v.nbytes
best = min(r["ingest_ms"] for r in rows)
print(json.dumps({"cells": N, "genes": D, "nnz": int(len(v)), "store_bytes": store_bytes, "array_bytes": host_bytes,
                  "compression_ratio": host_bytes / store_bytes, "write_s": t_write, "host_cores": os.cpu_count(), "block_identical": same,
                  "runs": rows, "ingest_cells_per_s": N / (best * 1e-3), "ingest_array_GBps": host_bytes / (best * 1e-3) / 1e9}))
shutil.rmtree(os.path.dirname(root), ignore_errors=True)
