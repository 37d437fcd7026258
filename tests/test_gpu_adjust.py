"""GPU parity tests for stage 7, the cross-batch neighbourhood adjustment (SURVEY.md §8a rows a14–a17):
index-valued results (proximity order, neighbour sets, pb-sample layout and matches, level maps) are
bit-exact against the oracle; the weighted sums (imputed / residual) are within 1e-5 (mixed form) and
bit-identical run to run.  Run on the B200 box: pytest -m gpu."""
import numpy as np
import pytest

import oracle as orc
from util import close, matched_stat_f64, max_err, random_csc

pytestmark = pytest.mark.gpu
TOL = 1e-5
NONE = 0xFFFFFFFF


@pytest.fixture(scope="module")
def lg():
    import legume_b200
    return legume_b200


@pytest.fixture(scope="module")
def ctx(lg):
    c = lg.Context(0)
    yield c
    c.close()


def make_case(seed, D, N, B, S, K, density=0.06, clustered=False):
    rng = np.random.default_rng(seed)
    ip, ix, v = random_csc(rng, D, N, density=density)
    proj = rng.standard_normal((N, K)).astype(np.float32)
    if clustered:  # a few tight clusters, so pb-samples of a group sit close together
        proj = (rng.standard_normal((8, K))[rng.integers(0, 8, N)] * 3 + 0.3 * proj).astype(np.float32)
    batch = rng.integers(0, B, N).astype(np.uint32)
    grp = rng.integers(0, S, N).astype(np.uint32)
    return ip, ix, v, proj, batch, grp


# ---- batch.rs:182-234 -----------------------------------------------------------------------------
@pytest.mark.parametrize("N,B,K", [(3000, 5, 50), (700, 3, 17), (40, 8, 50)])
def test_batch_proximity_bit_exact(lg, ctx, N, B, K):
    _, _, _, proj, batch, _ = make_case(1, 50, N, B, 4, K)
    order, cen = lg.sort_batch_proximity(ctx, proj, batch, B)
    worder, wcen = orc.batch_proximity(proj, batch, B)
    assert cen.tobytes() == wcen.tobytes()
    assert np.array_equal(order, worder)


# ---- matched.rs:173-260 ----------------------------------------------------------------------------
@pytest.mark.parametrize("N,B,K,knn,use_order", [(2500, 3, 50, 10, True), (1500, 4, 20, 5, False), (300, 2, 50, 10, False),
                                                  (60, 3, 8, 30, True)])
def test_knn_match_batches_bit_exact(lg, ctx, N, B, K, knn, use_order):
    _, _, _, proj, batch, _ = make_case(2, 50, N, B, 4, K)
    order = orc.batch_proximity(proj, batch, B)[0] if use_order else None
    idx, dist = lg.knn_match_batches(ctx, proj, batch, B, knn, order)
    widx, wdist = orc.knn_match_batches(proj, batch, B, knn, order)
    assert np.array_equal(idx, widx)
    assert dist.tobytes() == wdist.tobytes()


def test_knn_match_batches_tensor_path_bit_exact(lg, ctx):
    """large enough for the tcgen05 filter + exact refine (nr >= 4096, nq * nr >= 5e7)"""
    N, B, K, knn = 20000, 3, 50, 10
    _, _, _, proj, batch, _ = make_case(3, 50, 10, B, 4, K)
    rng = np.random.default_rng(5)
    proj = rng.standard_normal((N, K)).astype(np.float32)
    proj = ((proj - proj.mean(1, keepdims=True)) / proj.std(1, keepdims=True)).astype(np.float32)
    batch = rng.integers(0, B, N).astype(np.uint32)
    l0 = ctx.launch_count
    idx, dist = lg.knn_match_batches(ctx, proj, batch, B, knn)
    assert ctx.launch_count > l0
    widx, wdist = orc.knn_match_batches(proj, batch, B, knn)
    assert np.array_equal(idx, widx)
    assert dist.tobytes() == wdist.tobytes()


# ---- stats.rs:26-108 ---------------------------------------------------------------------------------
@pytest.mark.parametrize("D,N,B,S,knn", [(400, 1200, 3, 5, 4),      # groups of ~240 cells: several segments each
                                          (30000, 500, 2, 7, 3),    # two gene ranges
                                          (900, 300, 4, 64, 2)])    # tiny groups: one segment each
def test_collect_matched_stat_matches_oracle(lg, ctx, D, N, B, S, knn):
    ip, ix, v, proj, batch, grp = make_case(4, D, N, B, S, 16, density=min(0.06, 600.0 / D))
    midx, mdist = orc.knn_match_batches(proj, batch, B, knn)
    wimp, wres = orc.collect_matched_stat(ip, ix, v, D, grp, S, midx, mdist)
    blk = lg.CscBlock.upload(ctx, ip, ix, v, D)
    runs = []
    for _ in range(2):
        imp, res = np.empty((S, D), np.float32), np.empty((S, D), np.float32)
        ctx.check(lg.lib.lg_collect_matched_stat(ctx.h, blk.h, lg._ptr(grp), S, lg._ptr(midx), lg._ptr(mdist), midx.shape[1],
                                                 lg._ptr(imp), lg._ptr(res)))
        runs.append((imp, res))
    assert close(runs[0][0], wimp, TOL), max_err(runs[0][0], wimp)
    # the per-cell division scale sum(y1) / sum(y_hat) is a sequential f32 fold over every stored y_hat entry in
    # the reference (dmatrix_util.rs:149-150): ~10^3..10^4 terms whose own rounding noise (sqrt(n) * 6e-8, worst
    # case n * 6e-8) already reaches 1e-5.  The kernel sums with a fixed tree instead, so against the f32 oracle
    # the residual is held to 1e-4, and against the float64 restatement both sums are held to the 1e-5 contract.
    assert close(runs[0][1], wres, 1e-4), max_err(runs[0][1], wres)
    fimp, fres = matched_stat_f64(ip, ix, v, D, grp, S, midx, mdist)
    assert close(runs[0][0], fimp, TOL), max_err(runs[0][0], fimp)
    assert close(runs[0][1], fres, TOL), max_err(runs[0][1], fres)
    assert runs[0][0].tobytes() == runs[1][0].tobytes() and runs[0][1].tobytes() == runs[1][1].tobytes()
    # what the adjustment conserves: every cell's imputed column carries the weights' total mass
    assert wimp.sum() > 0 and abs(float(runs[0][0].sum()) / float(wimp.sum()) - 1.0) < 1e-5


def test_collect_matched_stat_empty_slots_and_unmatched_cells(lg, ctx):
    D, N, S = 300, 200, 6
    ip, ix, v, proj, batch, grp = make_case(5, D, N, 3, S, 12)
    midx, mdist = orc.knn_match_batches(proj, batch, 3, 3)
    midx[::5] = NONE  # every fifth cell has no match at all: its counts pass through to the residual
    mdist[::5] = np.inf
    wimp, wres = orc.collect_matched_stat(ip, ix, v, D, grp, S, midx, mdist)
    blk = lg.CscBlock.upload(ctx, ip, ix, v, D)
    imp, res = np.empty((S, D), np.float32), np.empty((S, D), np.float32)
    ctx.check(lg.lib.lg_collect_matched_stat(ctx.h, blk.h, lg._ptr(grp), S, lg._ptr(midx), lg._ptr(mdist), midx.shape[1],
                                             lg._ptr(imp), lg._ptr(res)))
    assert close(imp, wimp, TOL) and close(res, wres, 1e-4)


# ---- pb_samples.rs --------------------------------------------------------------------------------------
@pytest.mark.parametrize("N,B,S,K,weighted", [(3000, 3, 16, 50, False), (900, 4, 7, 12, True), (50, 2, 32, 50, False)])
def test_pb_layout_bit_exact(lg, ctx, N, B, S, K, weighted):
    _, _, _, proj, batch, grp = make_case(6, 50, N, B, S, K)
    w = np.random.default_rng(0).uniform(0.5, 3.0, N).astype(np.float32) if weighted else None
    lay = lg.build_pb_sample_layout(ctx, grp, S, batch, B, proj, w)
    want = orc.pb_layout(proj, grp, S, batch, B, w)
    assert lay.num_pb == want["num_pb"]
    assert np.array_equal(lay.cell_to_pbsamp, want["cell_to_pb"])
    assert np.array_equal(lay.pb_sample_to_group, want["pb_group"]) and np.array_equal(lay.pb_sample_to_batch, want["pb_batch"])
    assert lay.cell_counts.tobytes() == want["pb_count"].tobytes()
    assert lay.centroids.tobytes() == want["centroids"].tobytes()


@pytest.mark.parametrize("N,B,S,K,knn", [(3000, 3, 16, 50, 5), (1200, 4, 40, 20, 10), (200, 2, 8, 50, 10)])
def test_pb_match_bit_exact(lg, ctx, N, B, S, K, knn):
    _, _, _, proj, batch, grp = make_case(7, 50, N, B, S, K, clustered=True)
    want = orc.pb_layout(proj, grp, S, batch, B)
    lay = lg.build_pb_sample_layout(ctx, grp, S, batch, B, proj)
    mp, md = lg.per_batch_sc_neighbors(ctx, lay, proj, batch, B, knn)
    wmp, wmd = orc.pb_match(proj, batch, B, want, knn)
    assert np.array_equal(mp, wmp)
    assert md.tobytes() == wmd.tobytes()


def test_pb_match_adaptive_fixture(lg, ctx):
    """pb_samples_tests.rs:10-53 through the CUDA path"""
    feats, c2p = [], []
    for pb, n in [(p, 20) for p in range(3)] + [(p, 3) for p in range(3, 15)]:
        feats += [float(pb)] * n
        c2p += [pb] * n
    proj = np.array(feats + [0.0], np.float32)[:, None]
    batch = np.array([1] * len(feats) + [0], np.uint32)
    c2p = np.array(c2p + [15], np.uint32)
    cen = np.zeros((16, 1), np.float32)
    cen[:15, 0] = np.arange(15)
    pbb = np.ones(16, np.uint32)
    pbb[15] = 0
    lay = lg.PbSampleLayout(centroids=cen, cell_to_pbsamp=c2p, pb_sample_to_batch=pbb, num_pb=16)
    mp, md = lg.per_batch_sc_neighbors(ctx, lay, proj, batch, 2, 10)
    assert np.array_equal(mp[15, 10:20], np.arange(10, dtype=np.uint32))
    assert np.array_equal(md[15, 10:20], np.arange(10, dtype=np.float32))
    assert np.all(mp[15, :10] == NONE)


# ---- stats.rs:698-784 -----------------------------------------------------------------------------------
@pytest.mark.parametrize("D,N,B,S,knn", [(500, 2000, 3, 16, 4), (3000, 800, 4, 9, 10)])
def test_collect_matched_stat_coarse_matches_oracle(lg, ctx, D, N, B, S, knn):
    ip, ix, v, proj, batch, grp = make_case(8, D, N, B, S, 20, clustered=True)
    lay = orc.pb_layout(proj, grp, S, batch, B)
    npb = lay["num_pb"]
    wmp, wmd = orc.pb_match(proj, batch, B, lay, knn)
    gs, cnt = orc.collapse_basic(ip, ix, v, D, lay["cell_to_pb"], npb)
    wimp, wres = orc.collect_matched_stat_coarse(gs, lay["pb_count"], lay["pb_group"], S, wmp, wmd)
    stat = lg.CollapsedStat(D, S, B)
    glay = lg.PbSampleLayout(cell_counts=lay["pb_count"], num_pb=npb)
    lg.collect_matched_stat_coarse(ctx, glay, gs, lay["pb_group"], (wmp, wmd), stat)
    assert close(stat.imputed_sum_ds, wimp, TOL), max_err(stat.imputed_sum_ds, wimp)
    assert close(stat.residual_sum_ds, wres, TOL), max_err(stat.residual_sum_ds, wres)
    assert wimp.any() and wres.any()


# ---- refine.rs:741-769 ------------------------------------------------------------------------------------
def test_fine_to_coarse(lg, ctx):
    rng = np.random.default_rng(9)
    codes = rng.integers(0, 1 << 10, 5000).astype(np.uint64)
    grp, ng = orc.assign_groups(codes)
    gcode = np.zeros(ng, np.uint64)
    gcode[grp] = codes
    for dim in (10, 8, 7, 3):
        f2c, k = lg.compute_fine_to_coarse_mapping(ctx, codes, grp, ng, dim)
        wf2c, wk = orc.fine_to_coarse(gcode, dim)
        assert k == wk and np.array_equal(f2c, wf2c)


# ---- the two composed paths -----------------------------------------------------------------------------------
def test_collapse_columns_with_batches_matches_oracle(lg, ctx):
    """CollapsingOps::collapse_columns with B > 1 (collapse_data/mod.rs:384-475), per-cell matched stats"""
    D, N, B, K, knn = 600, 1500, 3, 20, 5
    ip, ix, v, proj, batch, _ = make_case(10, D, N, B, 4, K, clustered=True)
    data = lg.SparseIoVec.from_csc(ctx, ip, ix, v, D)
    data.build_hnsw_per_batch(proj, batch)
    data.partition_columns_to_groups(proj, 5)
    out, stat = data.collapse_columns(knn_batches=2, knn_cells=knn, num_opt_iter=15)
    grp, S = np.asarray(data.col_to_group), data.num_groups()
    # oracle composition on the same groups
    order, _ = orc.batch_proximity(proj, batch, B)
    midx, mdist = orc.knn_match_batches(proj, batch, B, knn, order)
    obs, size = orc.collapse_basic(ip, ix, v, D, grp, S)
    obs_db, n_bs = orc.collapse_batch(ip, ix, v, D, grp, batch, S, B)
    imp, res = orc.collect_matched_stat(ip, ix, v, D, grp, S, midx, mdist)
    assert np.array_equal(stat.observed_sum_ds, obs) and np.array_equal(stat.observed_sum_db, obs_db)
    assert close(stat.imputed_sum_ds, imp, TOL) and close(stat.residual_sum_ds, res, 1e-4)
    want = orc.optimize_batched(obs, imp, res, size, obs_db, n_bs, 1.0, 1.0, 15, 0)
    assert close(out.mu_adjusted["mean"], want["mu_adjusted"], 1e-4), max_err(out.mu_adjusted["mean"], want["mu_adjusted"])
    assert close(out.delta["mean"], want["delta"], 1e-4)


def test_collapse_columns_multilevel_matches_oracle(lg, ctx):
    """collapse_columns_multilevel_vec, un-refined path (collapse_data/mod.rs:867-1050)"""
    D, N, B, K = 500, 4000, 3, 20
    ip, ix, v, proj, batch, _ = make_case(11, D, N, B, 4, K, clustered=True)
    data = lg.SparseIoVec.from_csc(ctx, ip, ix, v, D)
    params = lg.MultilevelParams(K, knn_pb_samples=4, num_levels=2, sort_dim=8, num_opt_iter=12, refine=None)
    outs, stats = data.collapse_columns_multilevel_vec(proj, batch, params)
    dims = orc.level_sort_dims(8, 2)
    assert len(outs) == len(dims) == 2
    codes = orc.binary_codes(proj, dims[0])
    grp, S = orc.assign_groups(codes)
    assert np.array_equal(np.asarray(data.col_to_group), grp)
    lay = orc.pb_layout(proj, grp, S, batch, B)
    gs, _ = orc.collapse_basic(ip, ix, v, D, lay["cell_to_pb"], lay["num_pb"])
    mp, md = orc.pb_match(proj, batch, B, lay, 4)
    imp, res = orc.collect_matched_stat_coarse(gs, lay["pb_count"], lay["pb_group"], S, mp, md)
    obs, size = orc.collapse_basic(ip, ix, v, D, grp, S)
    obs_db, n_bs = orc.collapse_batch(ip, ix, v, D, grp, batch, S, B)
    assert np.array_equal(stats[0].observed_sum_ds, obs)
    assert close(stats[0].imputed_sum_ds, imp, TOL) and close(stats[0].residual_sum_ds, res, TOL)
    want = orc.optimize_batched(obs, imp, res, size, obs_db, n_bs, 1.0, 1.0, 12, 0)
    assert close(outs[0].mu_adjusted["mean"], want["mu_adjusted"], 1e-4)
    # coarser level: merge by masked code, then fit with max(iter / 2, 10) sweeps
    gcode = np.zeros(S, np.uint64)
    gcode[grp] = codes
    f2c, nc = orc.fine_to_coarse(gcode, dims[1])
    cobs, cimp, cres = (orc.merge_stat(x, f2c, nc) for x in (obs, imp, res))
    assert np.array_equal(stats[1].observed_sum_ds, cobs)
    assert close(stats[1].imputed_sum_ds, cimp, TOL) and close(stats[1].residual_sum_ds, cres, TOL)
    csize = np.zeros(nc, np.float32)
    cnbs = np.zeros((nc, B), np.float32)
    for f, c in enumerate(f2c):
        csize[c] += size[f]
        cnbs[c] += n_bs[f]
    assert np.array_equal(stats[1].size_s, csize) and np.array_equal(stats[1].n_bs, cnbs)
    want1 = orc.optimize_batched(cobs, cimp, cres, csize, obs_db, cnbs, 1.0, 1.0, 10, 0)
    assert close(outs[1].mu_adjusted["mean"], want1["mu_adjusted"], 1e-4)


def _oracle_refined_levels(ip, ix, v, D, proj, batch, B, codes, dims, p2g_levels, k_levels, lay, knn, iters, size_ds=None,
                           mask=None):
    """the tail of refine_and_collect_single_layer / ..._with_partition on the oracle (refine.rs:380-500, mod.rs:715-815)"""
    c2p = lay["cell_to_pb"].astype(np.int64)
    fine = p2g_levels[0][c2p].astype(np.uint32)
    S = k_levels[0]
    obs, size = orc.collapse_basic(ip, ix, v, D, fine, S)
    imp, res = np.zeros_like(obs), np.zeros_like(obs)
    obs_db = n_bs = None
    if B >= 2:
        obs_db, n_bs = orc.collapse_batch(ip, ix, v, D, fine, batch, S, B)
        gs, _ = orc.collapse_basic(ip, ix, v, D, lay["cell_to_pb"], lay["num_pb"])
        mp, md = orc.pb_match(proj, batch, B, lay, knn)
        imp, res = orc.collect_matched_stat_coarse(gs, lay["pb_count"], p2g_levels[0], S, mp, md)
    levels = []
    for level in range(len(k_levels)):
        it = iters if level == 0 else max(iters // 2, 10)
        if level > 0:
            f2c = orc.fine_to_coarse_from_refined(p2g_levels[level - 1], p2g_levels[level], k_levels[level - 1])
            nc = k_levels[level]
            obs, imp, res = (orc.merge_stat(x, f2c, nc) for x in (obs, imp, res))
            csize = np.zeros(nc, np.float32)
            cnbs = None if n_bs is None else np.zeros((nc, B), np.float32)
            for f, c in enumerate(f2c):
                csize[c] += size[f]
                if cnbs is not None:
                    cnbs[c] += n_bs[f]
            size, n_bs = csize, cnbs
            if size_ds is not None:
                size_ds = orc.merge_stat(size_ds, f2c, nc)
        if B >= 2:
            fit = orc.optimize_batched_obs(obs, imp, res, size, size_ds, obs_db, n_bs, mask, 1.0, 1.0, it, 0)
        else:
            fit = orc.optimize_single_obs(obs, size, size_ds, 1.0, 1.0, 0)
        levels.append(dict(obs=obs, imp=imp, res=res, size=size, n_bs=n_bs, fit=fit))
    return levels, fine


def test_multilevel_refine_arm_single_batch(lg, ctx):
    """MultilevelParams::new has refine = Some(..) (collapse_data/mod.rs:115-130): with one batch refine_or_identity keeps the
    compacted hash partition of every level (refine.rs:126-147, 68-88) and the statistics descend by merge_stat along
    fine_to_coarse_from_refined; collapse_columns_multilevel_with_hierarchy adds the per-level cell -> pb map"""
    D, N, K = 400, 3000, 20
    ip, ix, v, proj, _, _ = make_case(21, D, N, 1, 4, K, clustered=True)
    batch = np.zeros(N, np.uint32)
    data = lg.SparseIoVec.from_csc(ctx, ip, ix, v, D)
    params = lg.MultilevelParams(K, knn_pb_samples=4, num_levels=3, sort_dim=9, num_opt_iter=12)
    assert params.refine is not None and params.observe_panels
    out = data.collapse_columns_multilevel_with_hierarchy(proj, batch, params)
    dims = orc.level_sort_dims(9, 3)
    assert len(out["levels"]) == len(dims) == 3
    codes = orc.binary_codes(proj, dims[0])
    grp, S = orc.assign_groups(codes)
    lay = orc.pb_layout(proj, grp, S, batch, 1)
    cells = orc.pb_sample_to_cells(lay["cell_to_pb"], lay["num_pb"])
    init = orc.initial_per_level_from_hash(codes, cells, dims)
    p2g, k = zip(*(orc.compact_labels(l) for l in init))
    want, fine = _oracle_refined_levels(ip, ix, v, D, proj, batch, 1, codes, dims, list(p2g), list(k), lay, 4, 12)
    assert np.array_equal(np.asarray(data.col_to_group), fine)
    for level in range(3):
        st, w = out["stats"][level], want[level]
        assert np.array_equal(st.observed_sum_ds, w["obs"]) and np.array_equal(st.size_s, w["size"])
        assert close(out["levels"][level].mu_observed["mean"], w["fit"]["mean"], TOL)
        assert close(out["levels"][level].mu_observed["log_mean"], w["fit"]["log_mean"], TOL)
        assert np.array_equal(out["cell_to_pb_per_level"][level], p2g[level][lay["cell_to_pb"].astype(np.int64)])
    # the trait entry takes the same arm and returns the levels only
    outs, stats = lg.SparseIoVec.from_csc(ctx, ip, ix, v, D).collapse_columns_multilevel_vec(proj, batch, params)
    assert np.array_equal(stats[2].observed_sum_ds, want[2]["obs"])
    with pytest.raises(lg.LegumeError):
        data.collapse_columns_multilevel_with_hierarchy(proj, batch, lg.MultilevelParams(K, refine=None))


@pytest.mark.parametrize("B", [1, 3])
def test_multilevel_with_inherited_partition(lg, ctx, B):
    """collapse_columns_multilevel_with_partition (collapse_data/mod.rs:617-821): every level's pb-sample -> group is the
    majority of an inherited cell -> pb map over the pb-sample's cells, compacted; no refinement even with several batches"""
    D, N, K = 300, 2500, 16
    ip, ix, v, proj, batch, _ = make_case(22 + B, D, N, B, 4, K, clustered=True)
    rng = np.random.default_rng(3)
    lvl0 = rng.integers(0, 40, N)
    inherited = [lvl0, lvl0 // 4]  # a hierarchy: the coarse label is a function of the fine one
    data = lg.SparseIoVec.from_csc(ctx, ip, ix, v, D)
    params = lg.MultilevelParams(K, knn_pb_samples=3, num_levels=2, sort_dim=8, num_opt_iter=14)
    out = data.collapse_columns_multilevel_with_partition(proj, batch, params, inherited)
    dims = orc.level_sort_dims(8, 2)
    codes = orc.binary_codes(proj, dims[0])
    grp, S = orc.assign_groups(codes)
    lay = orc.pb_layout(proj, grp, S, batch, B)
    cells = orc.pb_sample_to_cells(lay["cell_to_pb"], lay["num_pb"])
    p2g, k = [], []
    for lvl in inherited:
        votes = [orc.modal_group(c, lvl) for c in cells]
        compact, kl = orc.compact_labels(votes)
        p2g.append(compact)
        k.append(kl)
    want, fine = _oracle_refined_levels(ip, ix, v, D, proj, batch, B, codes, dims, p2g, k, lay, 3, 14)
    assert np.array_equal(np.asarray(data.col_to_group), fine)
    for level in range(2):
        st, w = out["stats"][level], want[level]
        assert np.array_equal(st.observed_sum_ds, w["obs"]) and np.array_equal(st.size_s, w["size"])
        if B >= 2:
            assert close(st.imputed_sum_ds, w["imp"], TOL) and close(st.residual_sum_ds, w["res"], TOL)
            assert np.array_equal(st.n_bs, w["n_bs"])
            assert close(out["levels"][level].mu_adjusted["mean"], w["fit"]["mu_adjusted"], 1e-4)
            assert close(out["levels"][level].delta["mean"], w["fit"]["delta"], 1e-4)
        else:
            assert close(out["levels"][level].mu_observed["mean"], w["fit"]["mean"], TOL)
        assert np.array_equal(out["cell_to_pb_per_level"][level], p2g[level][lay["cell_to_pb"].astype(np.int64)])
    with pytest.raises(lg.LegumeError):
        data.collapse_columns_multilevel_with_partition(proj, batch, params, inherited[:1])


@pytest.mark.parametrize("B", [1, 2])
def test_panel_observability(lg, ctx, B):
    """attach_observability (collapse_data/mod.rs:221-301) + the size_ds / obs_mask_db arms of optimize_block
    (stats.rs:176-204, 299-322): two backends with different gene panels side by side"""
    rng = np.random.default_rng(31 + B)
    D, K = 260, 12
    panels = [np.sort(rng.choice(D, 200, replace=False)), np.sort(rng.choice(D, 170, replace=False))]
    backends, cols = [], []
    for pan, n in zip(panels, (900, 700)):
        ip, ix, v = random_csc(rng, len(pan), n, density=0.1)
        backends.append((ip, ix, v, pan.astype(np.uint32)))
        cols.append((ip, pan[ix.astype(np.int64)].astype(np.uint64), v))
    N = 1600
    ip = np.concatenate([cols[0][0], cols[1][0][1:] + cols[0][0][-1]]).astype(np.uint64)
    ix, v = np.concatenate([c[1] for c in cols]), np.concatenate([c[2] for c in cols])
    data = lg.SparseIoVec.from_backends(ctx, backends, D)
    cov = np.zeros((2, D), bool)
    cov[0, panels[0]], cov[1, panels[1]] = True, True
    assert np.array_equal(data.row_coverage_by_backend(), cov)
    source = np.repeat([0, 1], [900, 700]).astype(np.uint32)
    # batch 0 draws from backend 0 only when B == 2, so its delta mask has zeros
    batch = source.copy() if B == 2 else np.zeros(N, np.uint32)
    mult = rng.uniform(0.5, 2.0, N).astype(np.float32)
    S = 9
    grp = rng.integers(0, S, N).astype(np.uint32)
    data.register_batch_membership(batch)
    data.register_column_multiplicity(mult)
    data.assign_groups([f"{g:02d}" for g in grp])
    stat = lg.CollapsedStat(D, S, B)
    data.collect_basic_stat(stat)
    if B >= 2:
        data.collect_batch_stat(stat)
        stat.imputed_sum_ds = rng.uniform(0, 3, (S, D)).astype(np.float32)
        stat.residual_sum_ds = rng.uniform(0, 3, (S, D)).astype(np.float32)
    data.attach_observability(stat)
    wsize, wmask = orc.attach_observability(cov, source, grp, batch, mult, S, B)
    assert close(stat.size_ds, wsize, 1e-6)
    # genes that no backend of a batch measures (with one batch: the genes outside both panels) put zeros in the mask
    assert wmask is not None and np.array_equal(stat.obs_mask_db, wmask)
    if B == 2:
        assert np.array_equal(wmask == 0, ~cov)
    out = lg.optimize(ctx, stat, (1.0, 1.0), 15, 0)
    if B == 1:
        want = orc.optimize_single_obs(stat.observed_sum_ds, stat.size_s, stat.size_ds, 1.0, 1.0, 0)
        for plane in ("mean", "sd", "log_mean", "log_sd"):
            assert close(out.mu_observed[plane], want[plane], TOL), plane
        # an unmeasured gene keeps the prior: a denominator of b0 alone
        g = int(np.nonzero(~cov[0] & ~cov[1])[0][0]) if (~cov[0] & ~cov[1]).any() else None
        if g is not None:
            assert np.allclose(out.mu_observed["mean"][:, g], 1.0)
    else:
        want = orc.optimize_batched_obs(stat.observed_sum_ds, stat.imputed_sum_ds, stat.residual_sum_ds, stat.size_s,
                                        stat.size_ds, stat.observed_sum_db, stat.n_bs, stat.obs_mask_db, 1.0, 1.0, 15, 0)
        for name, key in (("mu_observed", "mu_observed"), ("mu_adjusted", "mu_adjusted"), ("gamma", "gamma"), ("delta", "delta")):
            assert close(out[name]["mean"], want[key], 1e-4), name
        masked = wmask == 0
        assert np.allclose(np.asarray(out.delta["mean"])[masked], 1.0)  # a0 / b0: no evidence, the prior
    # the coarse level inherits the effective sizes by merge_stat and the mask unchanged (stats.rs:820-829)
    f2c = (np.arange(S) // 3).astype(np.uint32)
    coarse = data._merge_level(stat, f2c, 3, B)
    assert close(coarse.size_ds, orc.merge_stat(wsize, f2c, 3), 1e-6)
    assert (coarse.obs_mask_db is None) == (wmask is None)
    sub = stat.select_rows(10, 50)
    assert np.array_equal(sub.size_ds, np.asarray(stat.size_ds)[:, 10:60])


# ---- edge cases: empty batches / groups / columns, one batch only ---------------------------------------------------
def test_adjustment_edge_cases(lg, ctx):
    D, N, B, S, K, knn = 200, 700, 4, 12, 10, 6
    rng = np.random.default_rng(12)
    ip, ix, v = random_csc(rng, D, N, density=0.08, empty_every=9)   # every ninth cell has no counts at all
    proj = rng.standard_normal((N, K)).astype(np.float32)
    batch = rng.choice([0, 1, 3], N).astype(np.uint32)                # batch 2 is empty
    batch[:3] = 3
    grp = rng.choice([0, 2, 3, 7, 11], N).astype(np.uint32)           # most group ids are empty
    # per-cell arm
    order, _ = orc.batch_proximity(proj, batch, B)
    gorder, _ = lg.sort_batch_proximity(ctx, proj, batch, B)
    assert np.array_equal(gorder, order)
    midx, mdist = lg.knn_match_batches(ctx, proj, batch, B, knn, order)
    widx, wdist = orc.knn_match_batches(proj, batch, B, knn, order)
    assert np.array_equal(midx, widx) and mdist.tobytes() == wdist.tobytes()
    wimp, wres = orc.collect_matched_stat(ip, ix, v, D, grp, S, widx, wdist)
    blk = lg.CscBlock.upload(ctx, ip, ix, v, D)
    imp, res = np.empty((S, D), np.float32), np.empty((S, D), np.float32)
    ctx.check(lg.lib.lg_collect_matched_stat(ctx.h, blk.h, lg._ptr(grp), S, lg._ptr(midx), lg._ptr(mdist), midx.shape[1],
                                             lg._ptr(imp), lg._ptr(res)))
    assert close(imp, wimp, TOL) and close(res, wres, 1e-4)
    assert not imp[[1, 4, 5, 6, 8, 9, 10]].any() and not res[[1, 4, 5, 6, 8, 9, 10]].any()  # empty groups stay zero
    # pb-sample arm
    lay = lg.build_pb_sample_layout(ctx, grp, S, batch, B, proj)
    want = orc.pb_layout(proj, grp, S, batch, B)
    assert lay.num_pb == want["num_pb"] and np.array_equal(lay.cell_to_pbsamp, want["cell_to_pb"])
    assert lay.centroids.tobytes() == want["centroids"].tobytes()
    mp, md = lg.per_batch_sc_neighbors(ctx, lay, proj, batch, B, knn)
    wmp, wmd = orc.pb_match(proj, batch, B, want, knn)
    assert np.array_equal(mp, wmp) and md.tobytes() == wmd.tobytes()
    assert np.all(mp[:, 2 * knn:3 * knn] == NONE)  # nothing can be matched in the empty batch
    # one batch only: no foreign pb-samples, the matched stats stay zero
    one = np.zeros(N, np.uint32)
    lay1 = lg.build_pb_sample_layout(ctx, grp, S, one, 1, proj)
    mp1, md1 = lg.per_batch_sc_neighbors(ctx, lay1, proj, one, 1, knn)
    assert np.all(mp1 == NONE) and np.all(np.isinf(md1))
    gs, _ = orc.collapse_basic(ip, ix, v, D, lay1.cell_to_pbsamp, lay1.num_pb)
    stat = lg.CollapsedStat(D, S, 1)
    lg.collect_matched_stat_coarse(ctx, lay1, gs, lay1.pb_sample_to_group, (mp1, md1), stat)
    assert not np.asarray(stat.imputed_sum_ds).any() and not np.asarray(stat.residual_sum_ds).any()


def test_merge_stat_many_groups(lg, ctx):
    rng = np.random.default_rng(13)
    fine = rng.integers(0, 50, (300, 64)).astype(np.float32)
    f2c = rng.integers(0, 37, 300).astype(np.uint32)
    assert np.array_equal(lg.merge_stat(ctx, fine, f2c, 37), orc.merge_stat(fine, f2c, 37))


def test_sharded_driver_pb_arm_matches_oracle(lg, ctx):
    """HotPath.pb_matched_stat (the staged, shard-able form of the pb-sample arm) on one rank against the oracle"""
    import torch
    from legume_b200.pipeline import HotPath
    D, N, B, S, K, knn = 400, 3000, 3, 16, 20, 4
    ip, ix, v, proj, batch, grp = make_case(21, D, N, B, S, K, clustered=True)
    blk = lg.CscBlock.upload(ctx, ip, ix, v, D)
    hp = HotPath(ctx)
    t = lambda a, dt: torch.from_numpy(a.astype(dt)).cuda()
    out = hp.pb_matched_stat(blk, t(proj, np.float32), t(grp, np.int32), S, t(batch, np.int32), B, knn)
    torch.cuda.synchronize()
    want = orc.pb_layout(proj, grp, S, batch, B)
    assert out["num_pb"] == want["num_pb"]
    assert np.array_equal(out["cell_to_pb"].cpu().numpy().astype(np.uint32), want["cell_to_pb"])
    assert out["centroids"].cpu().numpy().tobytes() == want["centroids"].tobytes()
    assert out["pb_count"].cpu().numpy().tobytes() == want["pb_count"].tobytes()
    wmp, wmd = orc.pb_match(proj, batch, B, want, knn)
    assert np.array_equal(out["matched_pb"].cpu().numpy().astype(np.uint32), wmp)
    assert out["matched_dist"].cpu().numpy().tobytes() == wmd.tobytes()
    gs, _ = orc.collapse_basic(ip, ix, v, D, want["cell_to_pb"], want["num_pb"])
    assert np.array_equal(out["gene_sums"].cpu().numpy(), gs)
    wimp, wres = orc.collect_matched_stat_coarse(gs, want["pb_count"], want["pb_group"], S, wmp, wmd)
    assert close(out["imputed_sum_ds"].cpu().numpy(), wimp, TOL) and close(out["residual_sum_ds"].cpu().numpy(), wres, TOL)
