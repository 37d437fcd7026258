"""What bounds the end-to-end feed when N ranks share one host: every rank, at the same time,
  (a) copies a pinned host buffer to its GPU with plain cudaMemcpyAsync   -> the PCIe / host-read ceiling of the box,
  (b) reads a host buffer with the CPU (sum over a u64 array, one and two threads per rank) -> the host DRAM ceiling the
      narrowing threads of lg_csc_upload live under,
  (c) does both at once.
One JSON line from rank 0 with per-rank and aggregate GB/s.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29571 tools/bench_h2d_ceiling.py
"""
import json, os, sys, threading, time
sys.stdout.flush()
_REAL = os.dup(1)
os.dup2(2, 1)
import numpy as np, torch, torch.distributed as dist

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
GB = float(1 << 30)
nbytes = int(float(sys.argv[1]) * GB) if len(sys.argv) > 1 else 2 << 30
host = torch.empty(nbytes // 8, dtype=torch.int64).pin_memory()
host.fill_(1)
dev = torch.empty_like(host, device="cuda")
cold = np.ones(nbytes // 8, np.uint64)  # ordinary (pageable) memory for the CPU reads


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def h2d(reps=6):
    barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        dev.copy_(host, non_blocking=True)
    torch.cuda.synchronize()
    return reps * nbytes / (time.perf_counter() - t0) / 1e9


def cpu_read(nthreads, reps=3, out=None):
    parts = np.array_split(cold, nthreads)
    def work(a):
        for _ in range(reps):
            a.sum()  # numpy releases the GIL: a streaming read of the slice
    barrier()
    t0 = time.perf_counter()
    ts = [threading.Thread(target=work, args=(p,)) for p in parts]
    [t.start() for t in ts]
    [t.join() for t in ts]
    gbps = reps * nbytes / (time.perf_counter() - t0) / 1e9
    if out is not None:
        out.append(gbps)
    return gbps


def both(nthreads):
    res = []
    th = threading.Thread(target=cpu_read, args=(nthreads, 3, res))
    barrier()
    t0 = time.perf_counter()
    th.start()
    for _ in range(6):
        dev.copy_(host, non_blocking=True)
    torch.cuda.synchronize()
    g = 6 * nbytes / (time.perf_counter() - t0) / 1e9
    th.join()
    return g, res[0]


h2d(2)
rows = {"h2d_GBps": h2d(), "cpu_read_1t_GBps": cpu_read(1), "cpu_read_2t_GBps": cpu_read(2)}
g, c = both(2)
rows["both_h2d_GBps"], rows["both_cpu_read_2t_GBps"] = g, c
vals = torch.tensor([rows[k] for k in sorted(rows)], dtype=torch.float64, device="cuda")
if world > 1:
    allv = [torch.empty_like(vals) for _ in range(world)]
    dist.all_gather(allv, vals)
else:
    allv = [vals]
if rank == 0:
    keys = sorted(rows)
    per_rank = {k: [float(v[i]) for v in allv] for i, k in enumerate(keys)}
    out = {"ranks": world, "bytes_per_rank": nbytes, "host_threads": os.cpu_count(),
           "aggregate_GBps": {k: sum(v) for k, v in per_rank.items()}, "per_rank_GBps": per_rank}
    os.write(_REAL, (json.dumps(out) + "\n").encode())
if world > 1:
    dist.destroy_process_group()
