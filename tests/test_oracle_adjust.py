"""CPU checks of the oracle's cross-batch adjustment stages (SURVEY.md §8a rows a14–a17): the
reference's own known-answer fixtures where it has them (pb_samples_tests.rs), independent float64
numpy restatements elsewhere."""
import numpy as np

import oracle as orc
from util import close, matched_stat_f64, random_csc


def make_case(seed=0, D=300, N=400, B=3, S=7, K=12):
    rng = np.random.default_rng(seed)
    ip, ix, v = random_csc(rng, D, N, density=0.08)
    proj = rng.standard_normal((N, K)).astype(np.float32)
    batch = rng.integers(0, B, N).astype(np.uint32)
    grp = rng.integers(0, S, N).astype(np.uint32)
    return D, N, B, S, K, ip, ix, v, proj, batch, grp


def dense(ip, ix, v, D):
    n = len(ip) - 1
    out = np.zeros((n, D))
    for j in range(n):
        sl = slice(int(ip[j]), int(ip[j + 1]))
        out[j, ix[sl].astype(np.int64)] = v[sl]
    return out


# ---- batch.rs:182-234 ---------------------------------------------------------------------------
def test_batch_proximity_orders_by_centroid_distance():
    _, N, B, _, K, _, _, _, proj, batch, _ = make_case(1, B=5)
    order, cen = orc.batch_proximity(proj, batch, B)
    want = np.stack([proj[batch == b].astype(np.float64).mean(0) for b in range(B)])
    assert close(cen, want, 1e-5)
    for b in range(B):
        assert order[b, 0] == b  # a batch is nearest to itself (exclude_same = false)
        d = ((want - want[b]) ** 2).sum(1)
        assert np.array_equal(order[b], np.argsort(d, kind="stable"))


# ---- matched.rs:173-260 --------------------------------------------------------------------------
def test_knn_match_batches_against_brute_force():
    _, N, B, _, K, _, _, _, proj, batch, _ = make_case(2)
    knn = 4
    order, _ = orc.batch_proximity(proj, batch, B)
    idx, dist = orc.knn_match_batches(proj, batch, B, knn, order)
    assert idx.shape == (N, B * knn)
    for j in range(0, N, 7):
        s = batch[j]
        for i in range(B):
            b = order[s, i]
            got = idx[j, i * knn:(i + 1) * knn]
            if b == s:
                assert np.all(got == 0xFFFFFFFF)  # skip_same_batch
                continue
            members = np.flatnonzero(batch == b)
            d2 = np.array([orc.l2_sq(proj[m], proj[j]) for m in members], np.float32)
            pick = members[np.lexsort((members, d2))[:knn]]
            assert np.array_equal(got, pick.astype(np.uint32))
            assert np.array_equal(dist[j, i * knn:(i + 1) * knn], np.sqrt(np.sort(d2)[:knn]))
            assert np.all(batch[got] == b) and j not in got


# ---- stats.rs:26-108 -------------------------------------------------------------------------------
def test_collect_matched_stat_against_float64():
    D, N, B, S, K, ip, ix, v, proj, batch, grp = make_case(3)
    idx, dist = orc.knn_match_batches(proj, batch, B, 3)
    imp, res = orc.collect_matched_stat(ip, ix, v, D, grp, S, idx, dist)
    wimp, wres = matched_stat_f64(ip, ix, v, D, grp, S, idx, dist)
    assert close(imp, wimp, 1e-5) and close(res, wres, 1e-5)


def test_collect_matched_stat_cell_without_matches_keeps_its_counts():
    D, N, B, S, K, ip, ix, v, proj, batch, grp = make_case(4, N=60)
    idx = np.full((N, 4), 0xFFFFFFFF, np.uint32)
    dist = np.full((N, 4), np.inf, np.float32)
    imp, res = orc.collect_matched_stat(ip, ix, v, D, grp, S, idx, dist)
    obs, _ = orc.collapse_basic(ip, ix, v, D, grp, S)
    assert not imp.any() and np.array_equal(res, obs)


# ---- pb_samples.rs ------------------------------------------------------------------------------------
def test_pb_layout_blocks_counts_and_centroids():
    D, N, B, S, K, ip, ix, v, proj, batch, grp = make_case(5)
    lay = orc.pb_layout(proj, grp, S, batch, B)
    keys = sorted({(int(g), int(b)) for g, b in zip(grp, batch)})
    assert lay["num_pb"] == len(keys)
    assert [(int(g), int(b)) for g, b in zip(lay["pb_group"], lay["pb_batch"])] == keys
    for p, (g, b) in enumerate(keys):
        cells = np.flatnonzero((grp == g) & (batch == b))
        assert np.all(lay["cell_to_pb"][cells] == p) and lay["pb_count"][p] == len(cells)
        assert close(lay["centroids"][p], proj[cells].astype(np.float64).mean(0), 1e-5)
    w = np.random.default_rng(0).uniform(0.5, 3.0, N).astype(np.float32)
    layw = orc.pb_layout(proj, grp, S, batch, B, mult=w)
    for p, (g, b) in enumerate(keys):
        cells = np.flatnonzero((grp == g) & (batch == b))
        assert close(layw["pb_count"][p], w[cells].sum(), 1e-5)
        assert close(layw["centroids"][p], (proj[cells] * w[cells, None]).astype(np.float64).sum(0) / w[cells].sum(), 1e-5)


def _one_d_fixture(pb_sizes):
    """pb_samples_tests.rs: a 1-D target batch (batch 1) whose cells sit at x = pb id, queried from a
    single-cell pb-sample at x = 0 in batch 0"""
    feats, c2p = [], []
    for pb, (x, n) in enumerate(pb_sizes):
        feats += [float(x)] * n
        c2p += [pb] * n
    q = len(pb_sizes)
    proj = np.array(feats + [0.0], np.float32)[:, None]
    batch = np.array([1] * len(feats) + [0], np.uint32)
    c2p = np.array(c2p + [q], np.uint32)
    cen = np.zeros((q + 1, 1), np.float32)
    pbb = np.ones(q + 1, np.uint32)
    for pb, (x, _) in enumerate(pb_sizes):
        cen[pb, 0] = x
    pbb[q] = 0
    return proj, batch, dict(cell_to_pb=c2p, centroids=cen, pb_batch=pbb, num_pb=q + 1), q


def test_adaptive_recovers_knn_distinct_pbsamples():
    """pb_samples_tests.rs:10-53: 60 dense near cells in 3 pb-samples, 12 sparse far ones; knn = 10"""
    proj, batch, lay, q = _one_d_fixture([(pb, 20) for pb in range(3)] + [(pb, 3) for pb in range(3, 15)])
    mp, md = orc.pb_match(proj, batch, 2, lay, 10)
    hits = mp[q, 10:20]
    assert len(set(hits.tolist())) == 10 and np.all(hits != 0xFFFFFFFF)
    assert any(h >= 3 for h in hits)
    assert np.all(np.diff(md[q, 10:20]) >= 0)
    assert np.array_equal(hits, np.arange(10, dtype=np.uint32)) and np.array_equal(md[q, 10:20], np.arange(10, dtype=np.float32))
    assert np.all(mp[q, :10] == 0xFFFFFFFF)  # own batch is never searched


def test_adaptive_returns_all_when_fewer_than_knn():
    """pb_samples_tests.rs:55-79"""
    proj, batch, lay, q = _one_d_fixture([(pb, 5) for pb in range(3)])
    mp, _ = orc.pb_match(proj, batch, 2, lay, 10)
    hits = mp[q, 10:20]
    assert set(hits[hits != 0xFFFFFFFF].tolist()) == {0, 1, 2}


def test_pb_match_is_first_k_distinct_in_distance_order():
    D, N, B, S, K, ip, ix, v, proj, batch, grp = make_case(6, N=900, S=16)
    lay = orc.pb_layout(proj, grp, S, batch, B)
    knn = 5
    mp, md = orc.pb_match(proj, batch, B, lay, knn)
    for p in range(0, lay["num_pb"], 5):
        for b in range(B):
            got = mp[p, b * knn:(b + 1) * knn]
            if b == lay["pb_batch"][p]:
                assert np.all(got == 0xFFFFFFFF)
                continue
            members = np.flatnonzero(batch == b)
            d2 = np.array([orc.l2_sq(proj[m], lay["centroids"][p]) for m in members], np.float32)
            want = []
            for m in members[np.lexsort((members, d2))]:
                o = lay["cell_to_pb"][m]
                if o != p and o not in want:
                    want.append(int(o))
                if len(want) == knn:
                    break
            assert got[got != 0xFFFFFFFF].tolist() == want


# ---- stats.rs:698-784 ---------------------------------------------------------------------------------
def test_collect_matched_stat_coarse_against_float64():
    D, N, B, S, K, ip, ix, v, proj, batch, grp = make_case(7, N=900, S=16)
    lay = orc.pb_layout(proj, grp, S, batch, B)
    npb = lay["num_pb"]
    mp, md = orc.pb_match(proj, batch, B, lay, 4)
    gs, cnt = orc.collapse_basic(ip, ix, v, D, lay["cell_to_pb"], npb)
    assert np.array_equal(cnt, lay["pb_count"])
    imp, res = orc.collect_matched_stat_coarse(gs, lay["pb_count"], lay["pb_group"], S, mp, md)
    wimp, wres = np.zeros((S, D)), np.zeros((S, D))
    G = gs.astype(np.float64)
    for p in range(npb):
        live = mp[p] != 0xFFFFFFFF
        if not live.any():
            continue
        m, d = mp[p][live].astype(np.int64), md[p][live].astype(np.float64)
        w = np.exp(-d - (-d).max())
        w /= w.sum()
        yhat = (w[:, None] * G[m] / lay["pb_count"][m][:, None]).sum(0)
        s = lay["pb_group"][p]
        wimp[s] += lay["pb_count"][p] * yhat
        pos = (yhat > 0) & (G[p] > 0)
        wres[s][pos] += G[p][pos] / yhat[pos]
    assert close(imp, wimp, 1e-5) and close(res, wres, 1e-5)


# ---- refine.rs:741-769 --------------------------------------------------------------------------------
def test_fine_to_coarse_mapping():
    codes = np.array([0b0000, 0b0100, 0b1000, 0b1100, 0b0101, 0b1111], np.uint64)
    f2c, k = orc.fine_to_coarse(codes, 2)
    assert k == 3 and f2c.tolist() == [0, 0, 0, 0, 1, 2]
    f2c, k = orc.fine_to_coarse(codes, 4)
    assert k == 6 and f2c.tolist() == [0, 1, 3, 4, 2, 5]


# ---- the refinement arm's bookkeeping and panel observability (oracle side; the GPU checks are in test_gpu_adjust.py) ----
def test_refine_bookkeeping_restatements():
    """dc_poisson.rs:493-509 (first-appearance compaction), refine.rs:43-62, 68-88, collapse_data/mod.rs:823-841"""
    c, k = orc.compact_labels([7, 7, 2, 9, 2, 7])
    assert c.tolist() == [0, 0, 1, 2, 1, 0] and k == 3
    cells = orc.pb_sample_to_cells(np.array([1, 0, 1, 0xFFFFFFFF, 2], np.uint32), 3)
    assert cells == [[1], [0, 2], [4]]
    codes = np.array([0b1011, 0b0011, 0b1011, 0, 0b0111], np.uint64)
    init = orc.initial_per_level_from_hash(codes, cells, [4, 2])
    assert init[0].tolist() == [0, 1, 2] and init[1].tolist() == [0, 0, 0]  # masked to 2 bits every pb-sample reads 0b11
    f2c = orc.fine_to_coarse_from_refined(np.array([0, 1, 1, 2]), np.array([0, 1, 1, 0]), 3)
    assert f2c.tolist() == [0, 1, 0]
    assert orc.modal_group([], [1, 2]) == 0 and orc.modal_group([1], [4, 6]) == 6
    assert orc.modal_group([0, 1, 2], [3, 5, 5]) == 5


def test_full_observability_is_bitwise_the_plain_fit():
    """stats.rs:172-175: 'full observability is bitwise-identical by construction' - size_ds = 1_d * size_s' and a mask of
    ones must reproduce optimize_block without them; a zero of the mask sends delta to a0 / b0"""
    rng = np.random.default_rng(8)
    S, D, B = 5, 40, 3
    obs, imp, res = (rng.integers(0, 9, (S, D)).astype(np.float32) for _ in range(3))
    size = rng.integers(1, 30, S).astype(np.float32)
    obs_db = rng.integers(0, 50, (B, D)).astype(np.float32)
    n_bs = rng.integers(0, 9, (S, B)).astype(np.float32)
    size_ds = np.repeat(size[:, None], D, axis=1)
    a = orc.optimize_batched(obs, imp, res, size, obs_db, n_bs, 1.0, 1.0, 7, 0)
    b = orc.optimize_batched_obs(obs, imp, res, size, size_ds, obs_db, n_bs, np.ones((B, D), np.float32), 1.0, 1.0, 7, 0)
    for key in a:
        assert a[key].tobytes() == b[key].tobytes(), key
    s0 = orc.optimize_single(obs, size, 1.0, 1.0, 0)
    s1 = orc.optimize_single_obs(obs, size, size_ds, 1.0, 1.0, 0)
    for key in s0:
        assert s0[key].tobytes() == s1[key].tobytes(), key
    mask = np.ones((B, D), np.float32)
    mask[1, :7] = 0.0
    c = orc.optimize_batched_obs(obs, imp, res, size, None, obs_db, n_bs, mask, 1.0, 1.0, 7, 0)
    assert np.all(c["delta"][1, :7] == 1.0) and np.array_equal(c["delta"][0], a["delta"][0])
    # attach_observability: the mass of a sample counts for a gene only where the column's backend measures it
    cov = np.array([[1, 1, 0], [0, 1, 1]], bool)
    size_ds, m = orc.attach_observability(cov, [0, 0, 1], [0, 1, 1], [0, 0, 1], None, 2, 2)
    assert size_ds.tolist() == [[1.0, 1.0, 0.0], [1.0, 2.0, 1.0]]
    assert m.tolist() == [[1.0, 1.0, 0.0], [0.0, 1.0, 1.0]]
