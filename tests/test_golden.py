"""Committed golden vectors (tests/golden/hotpath_small.npz, made by tests/golden/make_golden.py).
CPU: the oracle still reproduces them bit for bit.  GPU (-m gpu): the CUDA path, through the C ABI, hits the same
vectors — bit-exact for codes, groups, count sums, neighbour sets and layouts; 1e-5 for projections, posteriors and
the weighted sums."""
import os
import sys

import numpy as np
import pytest

import oracle as orc
from util import close, max_err

HERE = os.path.dirname(os.path.abspath(__file__))
G = np.load(os.path.join(HERE, "golden", "hotpath_small.npz"))
TOL = 1e-5


def test_generator_reproduces_the_committed_vectors():
    sys.path.insert(0, os.path.join(HERE, "golden"))
    import make_golden
    fresh = make_golden.build()
    assert sorted(fresh) == sorted(G.files)
    for key in G.files:
        a, b = np.asarray(fresh[key]), G[key]
        assert a.shape == b.shape and a.tobytes() == b.tobytes(), key


def test_golden_is_self_consistent():
    D, N, B, S = int(G["D"]), int(G["N"]), int(G["B"]), int(G["S"])
    assert G["proj"].shape == (N, int(G["K"])) and G["codes"].max() < (1 << int(G["kk"]))
    assert G["sum_ds"].sum() == G["data"].sum() and G["size_s"].sum() == N
    assert np.array_equal(G["sum_db"].sum(0), G["sum_ds"].sum(0))
    assert G["pb_count"].sum() == N and len(G["pb_count"]) == int(G["npb"])
    live = G["matched_idx"] != 0xFFFFFFFF
    assert np.all(G["batch"][G["matched_idx"][live]] != np.repeat(G["batch"], G["matched_idx"].shape[1]).reshape(live.shape)[live])


@pytest.fixture(scope="module")
def lg():
    import legume_b200
    return legume_b200


@pytest.fixture(scope="module")
def ctx(lg):
    c = lg.Context(0)
    yield c
    c.close()


@pytest.mark.gpu
def test_cuda_path_hits_the_golden_vectors(lg, ctx):
    D, N, K, kk, B, knn, S = (int(G[k]) for k in ("D", "N", "K", "kk", "B", "knn", "S"))
    data = lg.SparseIoVec.from_csc(ctx, G["indptr"], G["indices"], G["data"], D)
    _, proj = data.project_columns_with_batch_correction(K, None, G["batch"], basis=G["basis"])
    assert close(proj, G["proj"], TOL), max_err(proj, G["proj"])
    # every later stage starts from the GOLDEN projection, so that index-valued results can be compared exactly
    P = G["proj"]
    codes = lg.binary_sort_columns(ctx, P, kk)
    assert np.array_equal(codes, G["codes"])
    data.register_batch_membership(G["batch"])
    data.assign_groups(G["codes"])
    assert np.array_equal(data.col_to_group, G["group"]) and data.num_groups() == S
    stat = lg.CollapsedStat(D, S, B)
    data.collect_basic_stat(stat)
    data.collect_batch_stat(stat)
    assert np.array_equal(stat.observed_sum_ds, G["sum_ds"]) and np.array_equal(stat.size_s, G["size_s"])
    assert np.array_equal(stat.observed_sum_db, G["sum_db"]) and np.array_equal(stat.n_bs, G["n_bs"])
    single = lg.CollapsedStat(D, S, 1)
    single.observed_sum_ds, single.size_s = G["sum_ds"], G["size_s"]
    out = lg.optimize(ctx, single)
    for key in ("mean", "sd", "log_mean", "log_sd"):
        assert close(out.mu_observed[key], G["post_" + key], TOL), key
    idx, dist = lg.ColumnDict(ctx, P[:300], list(range(300))).search_indices(P[300:], 5)
    assert np.array_equal(idx, G["knn_idx"]) and dist.tobytes() == G["knn_dist"].tobytes()
    # stage 7, per-cell arm
    order, cen = lg.sort_batch_proximity(ctx, P, G["batch"], B)
    assert np.array_equal(order, G["prox_order"]) and cen.tobytes() == G["prox_centroids"].tobytes()
    midx, mdist = lg.knn_match_batches(ctx, P, G["batch"], B, knn, order)
    assert np.array_equal(midx, G["matched_idx"]) and mdist.tobytes() == G["matched_dist"].tobytes()
    data.batch_proj, data.between_batch_proximity = P, order
    data.collect_matched_stat(2, knn, None, stat)
    assert close(stat.imputed_sum_ds, G["imputed_cell"], TOL) and close(stat.residual_sum_ds, G["residual_cell"], 1e-4)
    # stage 7, pb-sample arm
    lay = lg.build_pb_sample_layout(ctx, G["group"], S, G["batch"], B, P)
    assert lay.num_pb == int(G["npb"]) and np.array_equal(lay.cell_to_pbsamp, G["pb_cell_to_pb"])
    assert lay.cell_counts.tobytes() == G["pb_count"].tobytes() and lay.centroids.tobytes() == G["pb_centroids"].tobytes()
    gs = lg.CollapsedStat(D, lay.num_pb, 1)
    ctx.check(lg.lib.lg_collapse_basic(ctx.h, data.block.h, lg._ptr(lay.cell_to_pbsamp), None, lay.num_pb, lg._ptr(gs.observed_sum_ds),
                                       lg._ptr(gs.size_s)))
    assert np.array_equal(gs.observed_sum_ds, G["pb_gene_sums"])
    mp, md = lg.per_batch_sc_neighbors(ctx, lay, P, G["batch"], B, knn)
    assert np.array_equal(mp, G["pb_matched"]) and md.tobytes() == G["pb_matched_dist"].tobytes()
    lg.collect_matched_stat_coarse(ctx, lay, gs.observed_sum_ds, lay.pb_sample_to_group, (mp, md), stat)
    assert close(stat.imputed_sum_ds, G["imputed_pb"], TOL) and close(stat.residual_sum_ds, G["residual_pb"], TOL)
    stat.imputed_sum_ds, stat.residual_sum_ds = G["imputed_pb"], G["residual_pb"]
    fit = lg.optimize(ctx, stat, (1.0, 1.0), 12)
    assert close(fit.mu_adjusted["mean"], G["mu_adjusted"], TOL) and close(fit.delta["mean"], G["delta"], TOL)
    f2c, nc = lg.compute_fine_to_coarse_mapping(ctx, G["codes"], G["group"], S, 4)
    assert nc == int(G["ncoarse_dim4"]) and np.array_equal(f2c, G["f2c_dim4"])


# ---- the steps either side of the path: tests/golden/next_small.npz -------------------------------
GN = np.load(os.path.join(HERE, "golden", "next_small.npz"))


def test_generator_reproduces_the_committed_next_vectors():
    sys.path.insert(0, os.path.join(HERE, "golden"))
    import make_golden
    fresh = make_golden.build_next()
    assert sorted(fresh) == sorted(GN.files)
    for key in GN.files:
        a, b = np.asarray(fresh[key]), GN[key]
        if key.startswith("nystrom_exact"):  # float64 numpy restatement: BLAS / pairwise sums may differ in the last bits
            assert a.shape == b.shape and np.allclose(a, b, rtol=1e-11, atol=1e-11), key
        else:
            assert a.shape == b.shape and a.tobytes() == b.tobytes(), key
    # the oracle's f32 folds against exact arithmetic: within the reference's own conditioning
    assert close(GN["nystrom_oracle"], GN["nystrom_exact"], 2e-3) and close(GN["nystrom_oracle_delta"], GN["nystrom_exact_delta"], 2e-3)
    assert GN["s1"].sum() == GN["data"].sum() and GN["npos"].sum() == (GN["data"] > 0).sum()


@pytest.mark.gpu
def test_cuda_path_hits_the_golden_next_vectors(lg, ctx):
    D, N, P = (int(GN[k]) for k in ("D", "N", "P"))
    data = lg.SparseIoVec.from_csc(ctx, GN["indptr"], GN["indices"], GN["data"], D)
    st = data.streaming_sparse_running_stats()
    assert st.ncols_processed() == N
    assert np.array_equal(st.count_positives(), GN["npos"]) and np.array_equal(st.sum(), GN["s1"])
    assert np.array_equal(st._s2.astype(np.float32), GN["s2"])
    assert np.array_equal(st.mean(), GN["mean"]) and np.array_equal(st.variance(), GN["variance"])
    assert np.array_equal(st.std(), GN["sd"], equal_nan=True)
    cs = np.abs(GN["nystrom_exact"]).max(0) + 1e-30  # per-column scale: the basis columns span three decades
    got = data.nystrom_project(GN["basis_dk"], None, 1e4)
    assert close(got / cs, GN["nystrom_exact"] / cs, TOL), max_err(got / cs, GN["nystrom_exact"] / cs)
    data.col_to_group = GN["pb"]  # the pseudobulk of a cell is its group id
    got_d = data.nystrom_project(GN["basis_dk"], GN["delta_dp"], 1e4)
    csd = np.abs(GN["nystrom_exact_delta"]).max(0) + 1e-30
    assert close(got_d / csd, GN["nystrom_exact_delta"] / csd, TOL), max_err(got_d / csd, GN["nystrom_exact_delta"] / csd)
    assert close(got_d, GN["nystrom_oracle_delta"], 2e-3)
