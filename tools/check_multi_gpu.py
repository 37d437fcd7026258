"""Multi-GPU parity of the cell-sharded path: N ranks (one per GPU, NCCL) must reproduce the single-GPU result
BIT FOR BIT — projections, codes, groups, gene x group sums, posterior planes and the sharded kNN merge.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/check_multi_gpu.py [cells=300000] [genes=8000]
"""
import json
import os
import sys

os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
# only the report reaches stdout: NCCL prints its version banner on fd 1
sys.stdout.flush()
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "legume-rs_b200")]
import numpy as np
import torch
import torch.distributed as dist

import legume_b200 as lg
from legume_b200 import sim
from legume_b200._lib import lib
from legume_b200.pipeline import HotPath, shard_range

N = int(sys.argv[1]) if len(sys.argv) > 1 else 300_000
D = int(sys.argv[2]) if len(sys.argv) > 2 else 8000
K, kk, B, k = 50, 10, 4, 10
world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device(f"cuda:{local}")
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
ctx = lg.Context(local)
hp = HotPath(ctx)
tabs = sim.make_tables(D, ntopic=6, nbatch=B, depth=600, pve_batch=0.3, seed=5)
basis = torch.from_numpy(np.random.default_rng(0).standard_normal((D, K)).astype(np.float32)).to(dev)

lo, hi = shard_range(N, rank, world)
blk, _, batch_h = sim.sim_block(ctx, tabs, lo, hi)
batch = torch.from_numpy(batch_h.astype(np.int32)).to(dev)
out = hp.run(blk, basis, batch, B, kk)
# the same path as ONE library call on the library's own NCCL communicator (lg_hotpath_run_sharded)
nat = hp.run_native(blk, basis, batch, B, kk)
native_same = all(bool(torch.equal(out[key], nat[key])) for key in ("proj", "codes", "group", "sum_ds", "size_s")) and \
    out["num_groups"] == nat["num_groups"] and \
    all(bool(torch.equal(out["posterior"][key], nat["posterior"][key])) for key in ("mean", "sd", "log_mean", "log_sd"))
sum_db, n_bs = hp.collapse_batch(blk, out["group"], batch, out["num_groups"], B)
# sharded kNN: this rank's cells are both its query shard and its reference shard; self excluded by global index
q_local = out["proj"][: min(4096, hi - lo)].contiguous()
excl = torch.arange(lo, lo + q_local.shape[0], dtype=torch.int32, device=dev)
kidx, kdist = hp.knn_topk_sharded(out["proj"], q_local, k, excl)
# pb-sample arm of the cross-batch adjustment over the shards
pbs = hp.pb_matched_stat(blk, out["proj"], out["group"], out["num_groups"], batch, B, 5)
# the nnz streams either side of the path: per-gene statistics (all-reduced) and the Nystrom re-projection (no exchange)
rs = hp.row_stats(blk)
ny_basis = torch.from_numpy(np.random.default_rng(1).standard_normal((20, D)).astype(np.float32)).to(dev)
ny_delta = (out["posterior"]["mean"] / out["posterior"]["mean"].mean(0, keepdim=True).clamp_min(1e-8)).contiguous()
ny = hp.nystrom_project(blk, ny_basis, ny_delta, out["group"])
torch.cuda.synchronize()

report = {"world": world, "cells": N, "genes": D}
if world > 1:
    flag = torch.tensor([1 if native_same else 0], dtype=torch.int32, device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    report["native_c_abi_path_bit_exact"] = bool(flag.item())
    # rank 0 recomputes everything unsharded on its own GPU (no collectives) and compares with the gathered shards
    gathered = {}
    for name, t in (("proj", out["proj"]), ("codes", out["codes"]), ("group", out["group"]), ("kidx", kidx), ("kdist", kdist),
                    ("c2p", pbs["cell_to_pb"]), ("ny", ny)):
        rows, cnt = hp.ex.all_gather_rows(t.contiguous())
        gathered[name] = (rows, cnt)
    if rank == 0:
        solo = HotPath.__new__(HotPath)
        solo.ctx, solo.dev = ctx, dev
        from legume_b200.exchange import Exchange
        solo.ex = Exchange.__new__(Exchange)
        solo.ex.on, solo.ex.dist, solo.ex.pg, solo.ex.world, solo.ex.rank = False, None, None, 1, 0
        solo.world, solo.rank = 1, 0
        fblk, _, fbatch_h = sim.sim_block(ctx, tabs, 0, N)
        fbatch = torch.from_numpy(fbatch_h.astype(np.int32)).to(dev)
        ref = solo.run(fblk, basis, fbatch, B, kk)
        rdb, rnbs = solo.collapse_batch(fblk, ref["group"], fbatch, ref["num_groups"], B)
        same = lambda a, b: bool(torch.equal(a, b))
        report["proj_bit_exact"] = same(gathered["proj"][0], ref["proj"])
        report["codes_bit_exact"] = same(gathered["codes"][0], ref["codes"])
        report["groups_bit_exact"] = same(gathered["group"][0], ref["group"]) and out["num_groups"] == ref["num_groups"]
        report["sum_ds_bit_exact"] = same(out["sum_ds"], ref["sum_ds"]) and same(out["size_s"], ref["size_s"])
        report["sum_db_bit_exact"] = same(sum_db, rdb) and same(n_bs, rnbs)
        report["posterior_bit_exact"] = all(same(out["posterior"][key], ref["posterior"][key]) for key in ("mean", "sd", "log_mean", "log_sd"))
        # kNN: every rank's queries against ALL cells in one search
        cnt = gathered["kidx"][1]
        qs, ex_all = [], []
        for r in range(world):
            rlo, _ = shard_range(N, r, world)
            qs.append(ref["proj"][rlo:rlo + cnt[r]])
            ex_all.append(torch.arange(rlo, rlo + cnt[r], dtype=torch.int32, device=dev))
        qall, exall = torch.cat(qs).contiguous(), torch.cat(ex_all).contiguous()
        widx = torch.empty((qall.shape[0], k), dtype=torch.int32, device=dev)
        wdist = torch.empty((qall.shape[0], k), dtype=torch.float32, device=dev)
        ctx.check(lib.lg_knn_topk(ctx.h, lg._ptr(ref["proj"]), N, lg._ptr(qall), qall.shape[0], K, k, lg._ptr(exall), lg._ptr(widx),
                                  lg._ptr(wdist)))
        rpb = solo.pb_matched_stat(fblk, ref["proj"], ref["group"], ref["num_groups"], fbatch, B, 5)
        report["pb_layout_bit_exact"] = (pbs["num_pb"] == rpb["num_pb"] and same(gathered["c2p"][0], rpb["cell_to_pb"])
                                         and same(pbs["centroids"], rpb["centroids"]) and same(pbs["pb_count"], rpb["pb_count"]))
        report["pb_gene_sums_bit_exact"] = same(pbs["gene_sums"], rpb["gene_sums"])
        report["pb_matches_bit_exact"] = same(pbs["matched_pb"], rpb["matched_pb"]) and same(pbs["matched_dist"], rpb["matched_dist"])
        report["pb_matched_stat_bit_exact"] = (same(pbs["imputed_sum_ds"], rpb["imputed_sum_ds"])
                                               and same(pbs["residual_sum_ds"], rpb["residual_sum_ds"]))
        rrs = solo.row_stats(fblk)
        report["row_stats_bit_exact"] = all(same(rs[i], rrs[i]) for i in range(3)) and rs[3] == rrs[3]
        report["nystrom_bit_exact"] = same(gathered["ny"][0], solo.nystrom_project(fblk, ny_basis, ny_delta, ref["group"]))
        report["knn_idx_bit_exact"] = same(gathered["kidx"][0], widx)
        report["knn_dist_bit_exact"] = same(gathered["kdist"][0], wdist)
        report["ok"] = all(v for key, v in report.items() if key.endswith("bit_exact"))
        os.write(_REAL_STDOUT, (json.dumps(report) + "\n").encode())
    dist.barrier()
    dist.destroy_process_group()
else:
    report["native_c_abi_path_bit_exact"] = native_same
    report["note"] = "single rank: only the one-call C-ABI path is compared with the staged one"
    os.write(_REAL_STDOUT, (json.dumps(report) + "\n").encode())
