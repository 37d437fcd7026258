"""End-to-end parity of the hot path FROM COUNTS: the CUDA path and the CPU oracle each run
counts + basis -> projection -> codes -> groups -> sums -> posterior on their own, nothing is handed
from one side to the other between stages.  Used by tests/test_e2e_parity.py (the gate) and by
bench.py (the `parity` object of the JSON line, computed outside every timed region).

The oracle is the checker here, never the thing measured (oracle/oracle.h header)."""
from __future__ import annotations

import time

import numpy as np


def sim_counts_cpu(D, N, ntopic=8, nbatch=1, depth=1000, seed=42):
    """configs[0]-shaped counts from the oracle's twin of the generator (no GPU involved)"""
    import oracle as orc
    from legume_b200 import sim
    tabs = sim.make_tables(D, ntopic=ntopic, nbatch=nbatch, depth=depth, seed=seed)
    topic, batch = tabs.cell_labels(0, N)
    ip, ix, v = orc.sim_poisson_csc(tabs.seed, D, 0, N, topic, batch, ntopic, nbatch, tabs.lam, tabs.p0, tabs.npiece)
    return ip, ix, v, batch.astype(np.uint32)


def oracle_path(ip, ix, v, D, basis, batch, nbatch, kk, nthreads=0):
    import oracle as orc
    t0 = time.perf_counter()
    proj = orc.project(ip, ix, v, basis, batch, nbatch, nthreads=nthreads)
    codes = orc.binary_codes(proj, kk)
    grp, ng = orc.assign_groups(codes)
    sums, size = orc.collapse_basic(ip, ix, v, D, grp, ng)
    post = orc.optimize_single(sums, size, 1.0, 1.0, 0)
    return dict(proj=proj, codes=codes, group=grp, num_groups=ng, sum_ds=sums, size_s=size, posterior=post,
                seconds=time.perf_counter() - t0)


def gpu_path(ctx, ip, ix, v, D, basis, batch, nbatch, kk, exact):
    import torch

    import legume_b200 as lg
    from legume_b200.pipeline import HotPath
    blk = lg.CscBlock.upload(ctx, ip, ix, v, D)
    dev = f"cuda:{ctx.device}"
    out = HotPath(ctx).run(blk, torch.from_numpy(basis).to(dev), torch.from_numpy(batch.astype(np.int32)).to(dev), nbatch, kk,
                           exact=exact)
    torch.cuda.synchronize()
    host = lambda t: t.cpu().numpy()
    res = dict(proj=host(out["proj"]), codes=host(out["codes"]).astype(np.uint64), group=host(out["group"]).astype(np.uint32),
               num_groups=out["num_groups"], sum_ds=host(out["sum_ds"]), size_s=host(out["size_s"]),
               posterior={k: host(x) for k, x in out["posterior"].items() if x is not None})
    blk.free()
    return res


def _rel(x, y):
    x, y = np.asarray(x, np.float64), np.asarray(y, np.float64)
    return float(np.max(np.abs(x - y) / (1.0 + np.maximum(np.abs(x), np.abs(y))))) if x.size else 0.0


def compare(got, want, kk):
    """mismatch counts of one GPU run against the oracle's run (both from counts)"""
    n = len(want["codes"])
    x = got["codes"] ^ want["codes"]
    flipped = int(sum(int(np.count_nonzero((x >> np.uint64(b)) & np.uint64(1))) for b in range(kk)))
    same_shape = got["num_groups"] == want["num_groups"]
    out = dict(cells=n,
               proj_bit_identical=bool(got["proj"].tobytes() == want["proj"].tobytes()),
               proj_max_err=_rel(got["proj"], want["proj"]),
               codes_bits_flipped=flipped,
               cells_with_flipped_code=int(np.count_nonzero(x)),
               num_groups=[int(got["num_groups"]), int(want["num_groups"])],
               groups_mismatch=int(np.count_nonzero(got["group"] != want["group"])),
               sums_mismatch=int(np.count_nonzero(got["sum_ds"] != want["sum_ds"])) if same_shape else -1,
               sizes_mismatch=int(np.count_nonzero(got["size_s"] != want["size_s"])) if same_shape else -1)
    if same_shape and out["groups_mismatch"] == 0:
        out["posterior_mean_max_err"] = _rel(got["posterior"]["mean"], want["posterior"]["mean"])
        out["posterior_log_mean_max_err"] = _rel(got["posterior"]["log_mean"], want["posterior"]["log_mean"])
    return out


def sim_counts_gpu(ctx, D, N, ntopic=8, depth=1000, seed=42):
    """the same counts from the GPU generator (bit-identical twin, tests/test_gpu_parity.py::test_sim_matches_cpu_twin);
    used by bench.py, where the single-threaded CPU generator would cost 15 s"""
    from legume_b200 import sim
    tabs = sim.make_tables(D, ntopic=ntopic, nbatch=1, depth=depth, seed=seed)
    blk, _, batch = sim.sim_block(ctx, tabs, 0, N)
    ip, ix, v = blk.download()
    blk.free()
    return ip, ix, v, batch.astype(np.uint32)


def run_e2e_parity(ctx, D=20000, N=50000, K=50, kk=10, depth=1000, seed=42, modes=("exact", "fast"), nthreads=0, counts="cpu"):
    """configs[0] by default.  Returns {"config": ..., "oracle_seconds": ..., "exact": {...}, "fast": {...}}"""
    ip, ix, v, batch = sim_counts_cpu(D, N, depth=depth, seed=seed) if counts == "cpu" else sim_counts_gpu(ctx, D, N, depth=depth, seed=seed)
    basis = np.random.default_rng(seed).standard_normal((D, K)).astype(np.float32)
    want = oracle_path(ip, ix, v, D, basis, batch, 1, kk, nthreads)
    rep = dict(config=dict(D=D, N=N, nnz=int(len(v)), K=K, kk=kk, depth=depth, seed=seed, nbatch=1),
               oracle_seconds=round(want["seconds"], 2))
    for mode in modes:
        got = gpu_path(ctx, ip, ix, v, D, basis, batch, 1, kk, exact=(mode == "exact"))
        rep[mode] = compare(got, want, kk)
    return rep


if __name__ == "__main__":
    import json
    import os
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, "legume-rs_b200")):
        sys.path.insert(0, p)
    import legume_b200 as lg
    print(json.dumps(run_e2e_parity(lg.Context(0)), indent=1))
