"""Host logic of the cell-sharded path on CPU: two gloo ranks run the exchange steps of
legume_b200.exchange with the oracle standing in for the kernels, and must reproduce the unsharded answer."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, case):
    for p in (ROOT, os.path.join(ROOT, "legume-rs_b200"), os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.distributed.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from legume_b200.exchange import Exchange
        from legume_b200.pipeline import shard_range
        ex = Exchange()
        assert (ex.world, ex.rank) == (world, rank)
        globals()["_case_" + case](ex, shard_range)
    finally:
        torch.distributed.destroy_process_group()


def _run(case, world=2):
    mp.spawn(_worker, args=(world, _free_port(), case), nprocs=world, join=True)


# ---- cases (run inside the workers) ---------------------------------------------------------------------
def _case_scalars(ex, shard_range):
    n_local = 100 + 7 * ex.rank
    assert ex.counts(n_local, "cpu") == [100 + 7 * r for r in range(ex.world)]
    assert ex.total(n_local, "cpu") == sum(100 + 7 * r for r in range(ex.world))
    mn, mx = ex.minmax(torch.tensor([-1.0 - ex.rank, 2.0 + ex.rank]))
    assert (mn, mx) == (-float(ex.world), 1.0 + ex.world)
    flags = torch.zeros(16, dtype=torch.int32)
    flags[ex.rank * 3] = 1
    ex.max_(flags)
    assert flags.tolist() == [1 if i in {3 * r for r in range(ex.world)} else 0 for i in range(16)]
    first = torch.full((4, 3), float(ex.rank))
    ex.broadcast_(first, 0)
    assert torch.all(first == 0)


def _case_block_partials(ex, shard_range):
    """ragged shards: the gathered partials are every block in GLOBAL block order (plus exact +0.0 rows), so the
    ordered f64 sum is bit-identical to the unsharded one"""
    from legume_b200 import BLOCK_CELLS
    ncols, M = 5 * BLOCK_CELLS + 321, 11  # 6 blocks: rank 0 gets 3, rank 1 gets 3 (the last one ragged)
    rng = np.random.default_rng(0)
    allblocks = rng.standard_normal((6, M)) * 10.0 ** rng.integers(-8, 8, (6, M))
    lo, hi = shard_range(ncols, ex.rank, ex.world)
    b0, b1 = lo // BLOCK_CELLS, (hi + BLOCK_CELLS - 1) // BLOCK_CELLS
    mine = torch.from_numpy(allblocks[b0:b1].copy())
    got = ex.gather_block_partials(mine, b1 - b0).numpy()
    live = got[np.abs(got).sum(1) > 0]
    assert live.tobytes() == allblocks.tobytes()
    seq = np.zeros(M)
    for row in got:  # lg_block_partials_finalize: plain left fold in block order
        seq = seq + row
    want = np.zeros(M)
    for row in allblocks:
        want = want + row
    assert seq.tobytes() == want.tobytes()
    # an uneven split (rank 1 holds fewer blocks) must pad, not reorder
    sizes = [4, 2]
    start = sum(sizes[:ex.rank])
    got2 = ex.gather_block_partials(torch.from_numpy(allblocks[start:start + sizes[ex.rank]].copy()), sizes[ex.rank]).numpy()
    assert got2.shape[0] == ex.world * 4
    assert got2[np.abs(got2).sum(1) > 0].tobytes() == allblocks.tobytes()


def _case_collapse_allreduce(ex, shard_range):
    """K5: per-shard gene x group count sums, all-reduced, equal the unsharded sums bit for bit"""
    import oracle as orc
    from util import random_csc
    D, N, S = 120, 2500, 9
    rng = np.random.default_rng(1)
    ip, ix, v = random_csc(rng, D, N, 0.1)
    grp = rng.integers(0, S, N).astype(np.uint32)
    want, wsize = orc.collapse_basic(ip, ix, v, D, grp, S)
    lo, hi = shard_range(N, ex.rank, ex.world)
    sub_ip = ip[lo:hi + 1] - ip[lo]
    sl = slice(int(ip[lo]), int(ip[hi]))
    part, psize = orc.collapse_basic(sub_ip, ix[sl], v[sl], D, grp[lo:hi], S)
    t, ts = torch.from_numpy(part), torch.from_numpy(psize)
    ex.sum_(t)
    ex.sum_(ts)
    assert np.array_equal(t.numpy(), want) and np.array_equal(ts.numpy(), wsize)


def _case_row_stats_allreduce(ex, shard_range):
    """K10: per-shard (npos, s1, s2) held as f64 whole numbers, all-reduced, equal the unsharded statistics bit for bit
    (merge = add, sparse_stat.rs:183-196)"""
    import oracle as orc
    from util import random_csc
    D, N = 150, 2300
    rng = np.random.default_rng(3)
    ip, ix, v = random_csc(rng, D, N, 0.12)
    want = orc.row_stats(ip, ix, v, D)
    lo, hi = shard_range(N, ex.rank, ex.world)
    sl = slice(int(ip[lo]), int(ip[hi]))
    part = orc.row_stats(ip[lo:hi + 1] - ip[lo], ix[sl], v[sl], D)
    t = torch.from_numpy(np.stack(part).astype(np.float64))
    ex.sum_(t)
    assert all(np.array_equal(t[k].numpy().astype(np.float32), want[k]) for k in range(3))
    assert ex.total(hi - lo, "cpu") == N


def _case_knn_shard_merge(ex, shard_range):
    """K7: reference cells sharded, queries all-gathered, k-lists sent home and merged by (squared distance,
    lower global index): identical to one exact search over all reference cells, ties included"""
    import oracle as orc
    d, k, nr_tot = 12, 7, 400
    rng = np.random.default_rng(2)
    ref = rng.integers(-3, 4, (nr_tot, d)).astype(np.float32)  # small integers: plenty of exact distance ties
    qry = rng.integers(-3, 4, (90, d)).astype(np.float32)
    rcut = [0, 230, nr_tot]
    qcut = [0, 35, 90]
    my_ref = ref[rcut[ex.rank]:rcut[ex.rank + 1]]
    my_q = torch.from_numpy(qry[qcut[ex.rank]:qcut[ex.rank + 1]].copy())
    allq, qcnt = ex.all_gather_rows(my_q)
    assert qcnt == [35, 55] and np.array_equal(allq.numpy(), qry)
    # this shard's answers for every query: squared distances, local indices (what lg_knn_topk_sq returns)
    nq = allq.shape[0]
    sq = np.array([[orc.l2_sq(r, q) for r in my_ref] for q in allq.numpy()], np.float32)
    order = np.lexsort((np.broadcast_to(np.arange(len(my_ref)), sq.shape), sq), axis=1)[:, :k]
    lidx = order.astype(np.int32)
    lsq = np.take_along_axis(sq, order, 1)
    sidx = ex.exchange_query_lists(torch.from_numpy(lidx), qcnt).numpy()
    ssq = ex.exchange_query_lists(torch.from_numpy(lsq), qcnt).numpy()
    assert sidx.shape == (ex.world, qcnt[ex.rank], k)
    # merge (the arithmetic of k_knn_merge)
    gi = sidx.astype(np.int64) + np.array(rcut[:-1])[:, None, None]
    flat_i = gi.transpose(1, 0, 2).reshape(qcnt[ex.rank], -1)
    flat_s = ssq.transpose(1, 0, 2).reshape(qcnt[ex.rank], -1)
    pick = np.lexsort((flat_i, flat_s), axis=1)[:, :k]
    got_i = np.take_along_axis(flat_i, pick, 1).astype(np.uint32)
    got_d = np.sqrt(np.take_along_axis(flat_s, pick, 1))
    widx, wdist = orc.knn_topk(ref, my_q.numpy(), k)
    assert np.array_equal(got_i, widx) and got_d.astype(np.float32).tobytes() == wdist.tobytes()


def _case_centroid_fold_chain(ex, shard_range):
    """K9: the pb-sample centroids are serial f32 folds over cells; shards continue the fold one after the other in
    rank order (broadcast of the running sums), which reproduces the unsharded fold bit for bit — a plain all-reduce
    of per-shard partial sums would not"""
    import oracle as orc
    N, K, S, B = 2600, 7, 5, 3
    rng = np.random.default_rng(3)
    proj = (rng.standard_normal((N, K)) * 10.0 ** rng.integers(-3, 4, (N, 1))).astype(np.float32)
    grp = rng.integers(0, S, N).astype(np.uint32)
    bat = rng.integers(0, B, N).astype(np.uint32)
    want = orc.pb_layout(proj, grp, S, bat, B)
    npb, c2p = want["num_pb"], want["cell_to_pb"]
    lo, hi = shard_range(N, ex.rank, ex.world)
    csum, ccnt = torch.zeros((npb, K), dtype=torch.float32), torch.zeros(npb, dtype=torch.float32)
    for r in range(ex.world):
        if r == ex.rank:  # lg_pb_centroid_fold: continue the folds over this shard's cells, ascending
            a, c = csum.numpy(), ccnt.numpy()
            for j in range(lo, hi):
                a[c2p[j]] = a[c2p[j]] + proj[j] * np.float32(1.0)
                c[c2p[j]] = c[c2p[j]] + np.float32(1.0)
        ex.broadcast_(csum, r)
        ex.broadcast_(ccnt, r)
    cen = csum.numpy() * (np.float32(1.0) / ccnt.numpy())[:, None]  # lg_pb_centroid_finish
    assert cen.astype(np.float32).tobytes() == want["centroids"].tobytes()
    assert ccnt.numpy().tobytes() == want["pb_count"].tobytes()
    # the contrast: summing per-shard partial folds is NOT the same arithmetic
    part = np.zeros((npb, K), np.float32)
    for j in range(lo, hi):
        part[c2p[j]] = part[c2p[j]] + proj[j]
    tot = torch.from_numpy(part.copy())
    ex.sum_(tot)
    assert tot.numpy().tobytes() != csum.numpy().tobytes()


def _case_min_keys_allreduce(ex, shard_range):
    """K9 matches: (squared distance << 32 | global cell) keys, min-reduced over shards as int64 with the all-ones
    'nothing here' key moved out of the way, equal the unsharded minimum"""
    rng = np.random.default_rng(4)
    nq, npb, ncell = 6, 9, 500
    d2 = rng.random((nq, ncell)).astype(np.float32)
    owner = rng.integers(0, npb + 2, ncell)  # some pb-samples own no cell at all
    keys_all = (d2.view(np.uint32).astype(np.uint64) << np.uint64(32)) | np.arange(ncell, dtype=np.uint64)[None, :]
    full = np.full((nq, npb), np.uint64(0xFFFFFFFFFFFFFFFF))
    for p in range(npb):
        cols = np.flatnonzero(owner == p)
        if len(cols):
            full[:, p] = keys_all[:, cols].min(1)
    lo, hi = (0, 260) if ex.rank == 0 else (260, ncell)
    mine = np.full((nq, npb), np.uint64(0xFFFFFFFFFFFFFFFF))
    for p in range(npb):
        cols = np.flatnonzero(owner[lo:hi] == p) + lo
        if len(cols):
            mine[:, p] = keys_all[:, cols].min(1)
    t = torch.from_numpy(mine.view(np.int64).copy())
    t[t < 0] = torch.iinfo(torch.int64).max
    ex.dist.all_reduce(t, op=ex.dist.ReduceOp.MIN, group=ex.pg)
    t[t == torch.iinfo(torch.int64).max] = -1
    assert np.array_equal(t.numpy().view(np.uint64), full)


@pytest.mark.parametrize("case", ["scalars", "block_partials", "collapse_allreduce", "knn_shard_merge", "centroid_fold_chain",
                                  "min_keys_allreduce", "row_stats_allreduce"])
def test_two_gloo_ranks(case):
    _run(case)


def test_single_process_is_the_identity():
    for p in (ROOT, os.path.join(ROOT, "legume-rs_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    from legume_b200.exchange import Exchange
    ex = Exchange()
    assert ex.world == 1 and ex.rank == 0
    x = torch.arange(12, dtype=torch.float64).view(4, 3)
    assert ex.gather_block_partials(x, 3).shape == (3, 3)
    rows, cnt = ex.all_gather_rows(x)
    assert cnt == [4] and rows is x
    assert ex.exchange_query_lists(x, [4]).shape == (1, 4, 3)


# ---- host logic of the mirror that needs no GPU -----------------------------------------------------------------------
def test_downsample_groups_host_logic():
    """partition_by_membership with nelem_per_group (matrix-util/src/utils.rs:36-66): sizes, per-group seeding by the smallest
    member (the subset of one group does not depend on the other groups), stability from run to run"""
    import legume_b200 as lg
    rng = np.random.default_rng(0)
    g = rng.integers(0, 5, 400).astype(np.uint32)
    out = lg.downsample_groups(g, 5, 30)
    assert np.array_equal(out, lg.downsample_groups(g, 5, 30))
    for k in range(5):
        assert (out == k).sum() == min(30, (g == k).sum())
        assert np.all(g[out == k] == k)
    assert np.all((out == g) | (out == 0xFFFFFFFF))
    # the members of group 2 keep their draw when another group changes
    g2 = g.copy()
    g2[g2 == 4] = 3
    assert np.array_equal(lg.downsample_groups(g2, 5, 30) == 2, out == 2)
    # nothing to do below the target
    assert np.array_equal(lg.downsample_groups(g, 5, 1000), g)
    assert lg.mix_seed(lg.PARTITION_SHUFFLE_SEED, 7) == lg.mix_seed(0x5041525453485546, 7)
