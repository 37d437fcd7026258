"""Stage timings for BASELINE config 3: N cells across B batches with the cross-batch kNN (k = 10)
neighbourhood adjustment, both arms (per-cell: collapse_columns; pb-sample: collapse_columns_multilevel_vec).

    python tools/bench_adjust.py [cells=1000000] [batches=8] [genes=30000] [reps=2]
"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "legume-rs_b200")]
import numpy as np
import torch

import legume_b200 as lg
from legume_b200 import sim
from legume_b200._lib import lib
from legume_b200.pipeline import HotPath

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
D = int(sys.argv[3]) if len(sys.argv) > 3 else 30000
REPS = int(sys.argv[4]) if len(sys.argv) > 4 else 2
K, kk, knn = 50, 10, 10
dev = torch.device("cuda:0")
ctx = lg.Context(0)
hp = HotPath(ctx)
p = lg._ptr

tabs = sim.make_tables(D, ntopic=8, nbatch=B, depth=1500, pve_batch=0.3, seed=42)
blk, _, batch_h = sim.sim_block(ctx, tabs, 0, N)
batch = torch.from_numpy(batch_h.astype(np.int32)).to(dev)
basis = torch.from_numpy(np.random.default_rng(0).standard_normal((D, K)).astype(np.float32)).to(dev)
times = {}


def timed(name, fn, reps=REPS):
    out = fn()  # two untimed calls: the first grows the stream-ordered pool (0.5 s for the 4 GB pattern scratch of K1)
    out = fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        out = fn()
    b.record()
    torch.cuda.synchronize()
    times[name] = a.elapsed_time(b) / reps
    return out


proj = timed("project", lambda: hp.project(blk, basis, batch, B))
codes = timed("binary_codes", lambda: hp.binary_codes(proj, kk))
group, S = timed("assign_groups", lambda: hp.assign_groups(codes, kk))
sum_ds, size_s = timed("collapse_basic", lambda: hp.collapse_basic(blk, group, S))
sum_db, n_bs = timed("collapse_batch", lambda: hp.collapse_batch(blk, group, batch, S, B))

# ---- pb-sample arm ------------------------------------------------------------------------------------
cap = S * B
c2p = torch.empty(N, dtype=torch.int32, device=dev)
pg, pb = torch.empty(cap, dtype=torch.int32, device=dev), torch.empty(cap, dtype=torch.int32, device=dev)
cnt = torch.empty(cap, dtype=torch.float32, device=dev)
cen = torch.empty((cap, K), dtype=torch.float32, device=dev)
npb_c = C.c_uint32()


def pb_layout():
    ctx.check(lib.lg_pb_layout(ctx.h, p(proj), K, N, p(group), S, p(batch), B, None, p(c2p), p(pg), p(pb), p(cnt), p(cen),
                               C.byref(npb_c)))
    return npb_c.value


npb = timed("pb_layout", pb_layout)
gs = torch.empty((npb, D), dtype=torch.float32, device=dev)
gsize = torch.empty(npb, dtype=torch.float32, device=dev)
timed("pb_gene_sums", lambda: ctx.check(lib.lg_collapse_basic(ctx.h, blk.h, p(c2p), None, npb, p(gs), p(gsize))))
T = B * knn
mp = torch.empty((npb, T), dtype=torch.int32, device=dev)
md = torch.empty((npb, T), dtype=torch.float32, device=dev)
timed("pb_match", lambda: ctx.check(lib.lg_pb_match(ctx.h, p(proj), K, N, p(batch), B, p(c2p), p(cen), p(pb), npb, knn, p(mp), p(md))))
imp = torch.empty((S, D), dtype=torch.float32, device=dev)
res = torch.empty((S, D), dtype=torch.float32, device=dev)
timed("pb_matched_stat_coarse", lambda: ctx.check(lib.lg_collect_matched_stat_coarse(ctx.h, p(gs), D, npb, p(cnt), p(pg), S, p(mp), p(md),
                                                                                     T, p(imp), p(res))))
outs = [torch.empty((S, D), dtype=torch.float32, device=dev) for _ in range(5)]
delta = torch.empty((B, D), dtype=torch.float32, device=dev)
timed("optimize_batched_30it", lambda: ctx.check(lib.lg_optimize_batched(ctx.h, p(sum_ds), p(imp), p(res), p(size_s), p(sum_db), p(n_bs), D,
                                                                          S, B, 1.0, 1.0, 30, 0, p(outs[0]), p(outs[1]), p(outs[2]),
                                                                          p(outs[3]), p(delta), p(outs[4]))))
pb_arm = ["project", "binary_codes", "assign_groups", "collapse_basic", "collapse_batch", "pb_layout", "pb_gene_sums", "pb_match",
          "pb_matched_stat_coarse", "optimize_batched_30it"]
del gs

# ---- per-cell arm ---------------------------------------------------------------------------------------
order = np.empty((B, B), np.uint32)
timed("batch_proximity", lambda: ctx.check(lib.lg_batch_proximity(ctx.h, p(proj), K, N, p(batch), B, p(order), None)))
midx = torch.empty((N, T), dtype=torch.int32, device=dev)
mdist = torch.empty((N, T), dtype=torch.float32, device=dev)
timed("knn_match_batches", lambda: ctx.check(lib.lg_knn_match_batches(ctx.h, p(proj), K, N, p(batch), B, knn, p(order), B, p(midx), p(mdist))),
      reps=1)
timed("collect_matched_stat", lambda: ctx.check(lib.lg_collect_matched_stat(ctx.h, blk.h, p(group), S, p(midx), p(mdist), T, p(imp), p(res))),
      reps=1)
cell_arm = ["project", "binary_codes", "assign_groups", "collapse_basic", "collapse_batch", "batch_proximity", "knn_match_batches",
            "collect_matched_stat", "optimize_batched_30it"]
nq_total = sum(N - int((batch_h == b).sum()) for b in range(B))
knn_flop = sum(2.0 * K * (N - int((batch_h == b).sum())) * int((batch_h == b).sum()) for b in range(B))
print(json.dumps({
    "workload": f"{D} genes x {N} cells, {B} batches, k={knn}, 2^{kk} bins -> {S} groups, {npb} pb-samples, nnz={blk.nnz}",
    "stage_ms": times,
    "pb_sample_arm_ms": sum(times[k] for k in pb_arm), "pb_sample_arm_cells_per_s": N / sum(times[k] for k in pb_arm) * 1e3,
    "per_cell_arm_ms": sum(times[k] for k in cell_arm), "per_cell_arm_cells_per_s": N / sum(times[k] for k in cell_arm) * 1e3,
    "knn_queries": nq_total, "knn_algorithmic_TFLOPs": knn_flop / times["knn_match_batches"] / 1e9,
}))
