"""Measurements of the BASELINE configs that are not the headline line, as functions bench.py calls after its timed
region (and tools/bench_adjust.py / bench_posterior_sweep.py print on their own):

    c3_adjust      configs[2]: N cells across B batches, cross-batch kNN (k = 10) neighbourhood adjustment, both arms
    c5_posterior   configs[4]: Poisson-Gamma posterior update swept over 2^8 .. 2^14 groups at D genes
    knn_sharded    the reference-cell-sharded kNN (all-gather queries, per-shard top-k, all-to-all, merge) at any world size
"""
from __future__ import annotations

import ctypes as C
import json
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _timed(fn, reps=2, warm=2):
    out = None
    for _ in range(warm):  # the first call grows the stream-ordered pool
        out = fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        out = fn()
    b.record()
    torch.cuda.synchronize()
    return out, a.elapsed_time(b) / reps


def c3_adjust(ctx, hp, N=1_000_000, B=8, D=30000, K=50, kk=10, knn=10, reps=2, per_cell=True, do_refine=True):
    """stage times (ms) of configs[2] on one GPU; returns a dict ready for the bench line"""
    import legume_b200 as lg
    from legume_b200 import sim
    from legume_b200._lib import lib
    p = lg._ptr
    dev = hp.dev
    tabs = sim.make_tables(D, ntopic=8, nbatch=B, depth=1500, pve_batch=0.3, seed=42)
    blk, _, batch_h = sim.sim_block(ctx, tabs, 0, N)
    blk.keep_pattern(True)  # as the SparseIoVec mirrors do: the three collapses below sum the projection's pattern
    batch = torch.from_numpy(batch_h.astype(np.int32)).to(dev)
    basis = torch.from_numpy(np.random.default_rng(0).standard_normal((D, K)).astype(np.float32)).to(dev)
    t = {}

    def run(name, fn, r=reps, w=2):
        out, t[name] = _timed(fn, r, w)
        return out

    proj = run("project", lambda: hp.project(blk, basis, batch, B))
    codes = run("binary_codes", lambda: hp.binary_codes(proj, kk))
    group, S = run("assign_groups", lambda: hp.assign_groups(codes, kk))
    sum_ds, size_s = run("collapse_basic", lambda: hp.collapse_basic(blk, group, S))
    sum_db, n_bs = run("collapse_batch", lambda: hp.collapse_batch(blk, group, batch, S, B))
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

    npb = run("pb_layout", pb_layout)
    gs = torch.empty((npb, D), dtype=torch.float32, device=dev)
    gsize = torch.empty(npb, dtype=torch.float32, device=dev)
    run("pb_gene_sums", lambda: ctx.check(lib.lg_collapse_basic(ctx.h, blk.h, p(c2p), None, npb, p(gs), p(gsize))))
    T = B * knn
    mp = torch.empty((npb, T), dtype=torch.int32, device=dev)
    md = torch.empty((npb, T), dtype=torch.float32, device=dev)
    run("pb_match", lambda: ctx.check(lib.lg_pb_match(ctx.h, p(proj), K, N, p(batch), B, p(c2p), p(cen), p(pb), npb, knn, p(mp), p(md))))
    imp = torch.empty((S, D), dtype=torch.float32, device=dev)
    res = torch.empty((S, D), dtype=torch.float32, device=dev)
    run("pb_matched_stat_coarse", lambda: ctx.check(lib.lg_collect_matched_stat_coarse(ctx.h, p(gs), D, npb, p(cnt), p(pg), S, p(mp),
                                                                                       p(md), T, p(imp), p(res))))
    outs = [torch.empty((S, D), dtype=torch.float32, device=dev) for _ in range(5)]
    delta = torch.empty((B, D), dtype=torch.float32, device=dev)
    run("optimize_batched_30it", lambda: ctx.check(lib.lg_optimize_batched(ctx.h, p(sum_ds), p(imp), p(res), p(size_s), p(sum_db), p(n_bs),
                                                                            D, S, B, 1.0, 1.0, 30, 0, p(outs[0]), p(outs[1]), p(outs[2]),
                                                                            p(outs[3]), p(delta), p(outs[4]))))
    pb_arm = ["project", "binary_codes", "assign_groups", "collapse_basic", "collapse_batch", "pb_layout", "pb_gene_sums", "pb_match",
              "pb_matched_stat_coarse", "optimize_batched_30it"]
    # the refinement of the pb-sample partition that MultilevelParams::new switches on (refine.rs:329-345): 20 Gibbs + 10 greedy
    # Jacobi sweeps per level, two levels, NB Fisher weights; wall clock (the picks and the candidate sets are host work)
    refine = None
    if do_refine:
        import time
        codes_h = codes.cpu().numpy().astype(np.uint64)
        c2p_h = c2p.cpu().numpy().astype(np.int64)
        first = np.full(npb, N, np.int64)
        np.minimum.at(first, c2p_h, np.arange(N))
        dims = lg.compute_level_sort_dims(kk, 2)
        init = lg.initial_per_level_from_hash(codes_h, first, dims)
        offs = lg.build_reproject_offsets(codes_h, first, dims)
        mp_h = mp.cpu().numpy().astype(np.uint32)
        bbknn = mp_h  # the (npb, B * knn) matrix of per_batch_sc_neighbors, as SparseIoVec._refine_and_collect passes it
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        lv, ks, moves = lg.refine_assignments(ctx, gs, bbknn, init, offs, lg.RefineParams())
        torch.cuda.synchronize()
        t["pb_refine"] = (time.perf_counter() - t0) * 1e3
        refine = {"levels": [int(d) for d in dims], "groups_before": [int(x.max()) + 1 for x in init], "groups_after": ks, "moves": moves,
                  "sweeps": "20 Gibbs (stagnation 0.005) + 10 greedy per level, Fisher weights"}
    del gs
    out = {"workload": f"{D} genes x {N} cells, {B} batches, k={knn}, 2^{kk} bins -> {S} groups, {npb} pb-samples, nnz={blk.nnz}",
           "pb_arm_ms": sum(t[k] for k in pb_arm), "pb_arm_cells_per_s": N / sum(t[k] for k in pb_arm) * 1e3}
    if refine is not None:
        out["pb_refine"] = refine
        out["pb_arm_with_refine_ms"] = out["pb_arm_ms"] + t["pb_refine"]
    if per_cell:
        order = np.empty((B, B), np.uint32)
        run("batch_proximity", lambda: ctx.check(lib.lg_batch_proximity(ctx.h, p(proj), K, N, p(batch), B, p(order), None)))
        midx = torch.empty((N, T), dtype=torch.int32, device=dev)
        mdist = torch.empty((N, T), dtype=torch.float32, device=dev)
        run("knn_match_batches", lambda: ctx.check(lib.lg_knn_match_batches(ctx.h, p(proj), K, N, p(batch), B, knn, p(order), B, p(midx),
                                                                            p(mdist))), 1, 1)
        run("collect_matched_stat", lambda: ctx.check(lib.lg_collect_matched_stat(ctx.h, blk.h, p(group), S, p(midx), p(mdist), T, p(imp),
                                                                                  p(res))), 1, 1)
        cell_arm = ["project", "binary_codes", "assign_groups", "collapse_basic", "collapse_batch", "batch_proximity",
                    "knn_match_batches", "collect_matched_stat", "optimize_batched_30it"]
        cnts = [int((batch_h == b).sum()) for b in range(B)]
        knn_flop = sum(2.0 * K * (N - c) * c for c in cnts)
        out.update(per_cell_arm_ms=sum(t[k] for k in cell_arm), per_cell_arm_cells_per_s=N / sum(t[k] for k in cell_arm) * 1e3,
                   knn_algorithmic_TFLOPs=knn_flop / t["knn_match_batches"] / 1e9)
        del midx, mdist
    out["stage_ms"] = {k: round(v, 3) for k, v in t.items()}
    blk.free()
    return out


def c5_posterior(ctx, D=30000, kks=range(8, 15), ncells=10_000_000, peak=6545.6):
    """configs[4]: lg_optimize_single (target All) over 2^kk groups; [{S, ms, GBps, frac}]"""
    from legume_b200._lib import lib
    g = torch.Generator(device="cuda").manual_seed(0)
    rate = 0.05 * torch.exp(torch.randn(D, device="cuda", generator=g) - 0.5)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    rows = []
    for kk in kks:
        S = 1 << kk
        lam = (rate * (ncells / S))[None, :].expand(S, D).contiguous()
        sums = torch.poisson(lam, generator=g)
        del lam
        size = torch.full((S,), ncells / S, device="cuda")
        outs = [torch.empty((S, D), device="cuda") for _ in range(4)]

        def run():
            ctx.check(lib.lg_optimize_single(ctx.h, sums.data_ptr(), size.data_ptr(), D, S, 1.0, 1.0, 0, outs[0].data_ptr(),
                                             outs[1].data_ptr(), outs[2].data_ptr(), outs[3].data_ptr()))
        for _ in range(2):
            run()
        ts = []
        for _ in range(5):
            flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            run()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ms = float(np.median(ts))
        nbytes = 4.0 * D * S * 5
        rows.append({"S": S, "ms": round(ms, 4), "GBps": round(nbytes / ms / 1e6, 1), "frac": round(nbytes / ms / 1e6 / peak, 4)})
        del sums, outs
    return rows


def knn_sharded(ctx, hp, n_per_rank=125_000, d=50, k=10, reps=2):
    """every rank holds n_per_rank reference cells and the same number of queries; one search over all ranks' cells with
    the reference cells sharded (HotPath.knn_topk_sharded).  Returns ms (device time, this rank)."""
    dev = hp.dev
    g = torch.Generator(device=dev).manual_seed(1234 + hp.rank)
    std = lambda x: (x - x.mean(1, keepdim=True)) / x.std(1, keepdim=True, unbiased=False)
    ref = std(torch.randn((n_per_rank, d), device=dev, generator=g))
    qry = std(torch.randn((n_per_rank, d), device=dev, generator=g))
    _, ms = _timed(lambda: hp.knn_topk_sharded(ref, qry, k), reps, 1)
    return {"ms": ms, "queries": n_per_rank * hp.world, "refs": n_per_rank * hp.world, "d": d, "k": k,
            "algorithmic_TFLOPs": 2.0 * d * (n_per_rank * hp.world) ** 2 / ms / 1e9}


if __name__ == "__main__":
    import sys
    sys.path[:0] = [ROOT, os.path.join(ROOT, "legume-rs_b200")]
    import legume_b200 as lg
    from legume_b200.pipeline import HotPath
    ctx = lg.Context(0)
    hp = HotPath(ctx)
    which = sys.argv[1] if len(sys.argv) > 1 else "c3"
    if which == "c3":
        print(json.dumps(c3_adjust(ctx, hp, int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000)))
    elif which == "c5":
        print(json.dumps(c5_posterior(ctx)))
    else:
        print(json.dumps(knn_sharded(ctx, hp)))
