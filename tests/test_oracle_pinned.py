"""Pins the CPU oracle against every known-answer test the reference holds for the hot path
(SURVEY.md §8c).  Each test names the reference test it restates."""
import math

import numpy as np
import pytest

import oracle as orc


def close(x, y, tol=1e-5):
    """stats_tests.rs:20-30 assert_mat_close: |x-y| <= tol * (1 + max(|x|,|y|))"""
    x = np.asarray(x, np.float64)
    y = np.asarray(y, np.float64)
    return np.all(np.abs(x - y) <= tol * (1.0 + np.maximum(np.abs(x), np.abs(y))))


# ---- matrix-param/src/dmatrix_gamma_tests.rs:9-32 -------------------------------------------
def test_log_sd_is_sqrt_trigamma_of_the_shape():
    cases = [(1.0, math.sqrt(math.pi ** 2 / 6)), (2.0, math.sqrt(math.pi ** 2 / 6 - 1.0)),
             (0.5, math.sqrt(math.pi ** 2 / 2))]
    for a, want in cases:
        out = orc.gamma_calibrate(np.array([a], np.float32), np.array([3.0], np.float32), a0=0.0, b0=0.0)
        assert abs(out["log_sd"][0] - want) < 1e-4


# ---- matrix-param/src/dmatrix_gamma_tests.rs:34-60 ------------------------------------------
def test_an_unobserved_feature_has_the_largest_log_sd_not_zero():
    out = orc.gamma_calibrate(np.array([0.0, 500.0], np.float32), np.array([0.0, 500.0], np.float32), 1.0, 1.0)
    sd = out["log_sd"]
    assert sd[0] > 1.0 and sd[1] < 0.1 and sd[0] > sd[1] * 10.0


def test_digamma_trigamma_against_scipy():
    from scipy.special import digamma, polygamma
    xs = np.concatenate([np.linspace(0.01, 12, 400), np.array([1e-6, 1e-5, 1e-4, 50.0, 1e3, 1e5])])
    for x in xs.astype(np.float32):
        assert close(orc.digamma(float(x)), digamma(float(x)), 2e-6), x
        assert close(orc.trigamma(float(x)), polygamma(1, float(x)), 2e-6), x


# ---- data-beans-alg/tests/weighted_columns.rs:76-121 ----------------------------------------
def _dense_cells_to_csc(cells, D):
    indptr, idx, val = [0], [], []
    for col in cells:
        for g, v in enumerate(col):
            if v != 0.0:
                idx.append(g)
                val.append(v)
        indptr.append(len(idx))
    return np.array(indptr, np.uint64), np.array(idx, np.uint64), np.array(val, np.float32)


def _one_group(cells, D, weights=None):
    ip, ix, v = _dense_cells_to_csc(cells, D)
    n = len(cells)
    s, size = orc.collapse_basic(ip, ix, v, D, np.zeros(n, np.uint32), 1, mult=weights)
    out = orc.optimize_single(s, size, 1.0, 1.0, orc.TARGET_ALL)
    return out["mean"][0], s[0], size[0]


def test_unit_weights_change_nothing():
    D = 6
    cells = [[float((g + j) % 5) for g in range(D)] for j in range(8)]
    plain, s, size = _one_group(cells, D)
    weighted, _, _ = _one_group(cells, D, np.ones(8, np.float32))
    assert plain.tobytes() == weighted.tobytes()
    # mu = (a0 + sum y) / (b0 + n) exactly under (1, 1)
    want = (1.0 + np.array(cells, np.float32).sum(0)) / np.float32(1.0 + 8.0)
    assert np.array_equal(plain, want.astype(np.float32))


def test_one_weighted_column_equals_the_cells_it_summarizes():
    D, M = 6, 20
    profile = [1.0 + g for g in range(D)]
    from_cells, _, n_cells = _one_group([profile] * M, D)
    from_summary, _, n_sum = _one_group([profile], D, np.array([M], np.float32))
    assert n_cells == M and n_sum == M  # size_s carries the multiplicity
    assert np.all(np.abs(from_cells - from_summary) < 1e-4)


# ---- data-beans-alg/src/collapse_data/stats_tests.rs ------------------------------------------
def toy_stat(G, S, B):
    f = lambda a, b: 1.0 + ((a * 7 + b * 13) % 11)
    obs = np.array([[f(g, c) for g in range(G)] for c in range(S)], np.float32)
    imp = np.array([[0.5 * f(g + 1, c + 2) for g in range(G)] for c in range(S)], np.float32)
    res = np.array([[0.3 * f(g + 2, c + 1) for g in range(G)] for c in range(S)], np.float32)
    size = np.array([2.0 + (c % 3) for c in range(S)], np.float32)
    obs_db = np.array([[f(g, b) + 0.7 for g in range(G)] for b in range(B)], np.float32)
    n_bs = np.array([[1.0 + ((b + c) % 4) for b in range(B)] for c in range(S)], np.float32)
    return obs, imp, res, size, obs_db, n_bs


def test_blocked_optimize_matches_single_block():
    G, S, B = 10, 4, 2
    obs, imp, res, size, obs_db, n_bs = toy_stat(G, S, B)
    full = orc.optimize_batched(obs, imp, res, size, obs_db, n_bs, 1.0, 1.0, 25, orc.TARGET_ALL)
    parts = []
    for r0, nr in [(0, 3), (3, 4), (7, 3)]:
        sl = slice(r0, r0 + nr)
        parts.append(orc.optimize_batched(obs[:, sl], imp[:, sl], res[:, sl], size, obs_db[:, sl], n_bs,
                                          1.0, 1.0, 25, orc.TARGET_ALL))
    for key in ["mu_observed", "mu_adjusted", "mu_residual", "gamma", "delta", "mu_adjusted_log_mean"]:
        blk = np.concatenate([p[key] for p in parts], axis=1)
        assert close(full[key], blk), key
    # float64 restatement of stats.rs:249-285 as an independent check of the sweep
    o, i_, r_ = obs.astype(np.float64), imp.astype(np.float64), res.astype(np.float64)
    sz = size.astype(np.float64)[:, None]
    m_res = (1 + r_) / (1 + sz)
    gam = np.zeros_like(o)
    for _ in range(25):
        mu = (1 + o + i_) / (1 + (m_res + gam) * sz)
        gam = (1 + i_) / (1 + mu * sz)
    assert close(full["mu_adjusted"], mu, 1e-5) and close(full["gamma"], gam, 1e-5)
    delta = (1 + obs_db.astype(np.float64)) / (1 + (mu.T @ n_bs.astype(np.float64)).T)
    assert close(full["delta"], delta, 1e-5)


def test_mean_only_sparsifies_unobserved_cells():
    G, S, B = 4, 3, 2
    obs = np.zeros((S, G), np.float32)
    imp = np.zeros((S, G), np.float32)
    res = np.zeros((S, G), np.float32)
    obs[0, 0], obs[1, 1], imp[2, 2] = 5.0, 3.0, 2.0
    size = np.full(S, 10.0, np.float32)
    n_bs = np.full((S, B), 5.0, np.float32)
    obs_db = np.ones((B, G), np.float32)
    out = orc.optimize_batched(obs, imp, res, size, obs_db, n_bs, 1.0, 1.0, 10, orc.TARGET_MEAN_ONLY)
    m = out["mu_adjusted"]
    assert m[0, 0] > 0 and m[1, 1] > 0 and m[2, 2] > 0
    assert m[0, 3] == 0.0 and m[1, 0] == 0.0
    out_all = orc.optimize_batched(obs, imp, res, size, obs_db, n_bs, 1.0, 1.0, 10, orc.TARGET_ALL)
    assert out_all["mu_adjusted"][0, 3] > 0.0


# ---- matrix-util/src/knn/tests.rs ------------------------------------------------------------
def random_points(n, d, seed):
    return np.random.default_rng(seed).uniform(-1.0, 1.0, size=(n, d)).astype(np.float32)


def brute(points, q, k, exclude=None):
    d = np.array([orc.l2_sq(p, q) for p in points], np.float32)
    order = np.lexsort((np.arange(len(points)), d))  # total_cmp on distance, lower index on ties
    if exclude is not None:
        order = order[order != exclude]
    return order[:k], np.sqrt(d[order[:k]])


def test_l2_simd_matches_scalar():
    a = np.array([i * 0.3 for i in range(37)], np.float32)
    b = np.array([math.sin(np.float32(i) * np.float32(0.1)) for i in range(37)], np.float32)
    scalar = math.sqrt(float(np.sum((a.astype(np.float64) - b) ** 2)))
    assert abs(math.sqrt(orc.l2_sq(a, b)) - scalar) < 1e-4


def test_exact_path_is_perfect():
    pts = random_points(500, 16, 2)
    k = 8
    idx, dist = orc.knn_topk(pts, pts, k, exclude=np.arange(500, dtype=np.uint32))
    for q in range(500):
        truth, td = brute(pts, pts[q], k, exclude=q)
        assert np.array_equal(idx[q], truth)
        assert np.all(np.diff(dist[q]) >= 0) and q not in idx[q]
        assert np.array_equal(dist[q], td)


def test_cross_dict_match_and_query_by_slice():
    a, b = random_points(300, 12, 4), random_points(400, 12, 5)
    idx, dist = orc.knn_topk(b, a[7:8], 5)
    truth, _ = brute(b, a[7], 5)
    assert np.array_equal(idx[0], truth) and dist.shape == (1, 5)
    pts = random_points(400, 16, 6)
    query = np.array([d * 0.05 - 0.4 for d in range(16)], np.float32)
    idx, dist = orc.knn_topk(pts, query[None, :], 6)
    truth, _ = brute(pts, query, 6)
    assert np.array_equal(idx[0], truth) and np.all(np.diff(dist[0]) >= 0)


def test_knn_l2_lane_order_d50():
    """metric.rs:23-44 at d=50: 3 full 16-lane chunks, lanes left-folded, then a 2-element tail."""
    rng = np.random.default_rng(0)
    a, b = rng.normal(size=50).astype(np.float32), rng.normal(size=50).astype(np.float32)
    acc = np.zeros(16, np.float32)
    for c in range(3):
        d = a[16 * c:16 * c + 16] - b[16 * c:16 * c + 16]
        acc = acc + d * d
    s = np.float32(0)
    for l in range(16):
        s = np.float32(s + acc[l])
    for c in (48, 49):
        d = np.float32(a[c] - b[c])
        s = np.float32(s + np.float32(d * d))
    assert orc.l2_sq(a, b) == float(s)


# ---- data-beans-alg/src/random_projection.rs:571-646 (tiny 8×12 fixture) ---------------------
def tiny_fixture():
    d, n = 8, 12
    indptr, idx, val = [0], [], []
    for j in range(n):
        for i in range(d):
            if (i * 7 + j * 3) % 5 < 3:
                idx.append(i)
                val.append(1.0 + ((i + j) % 4))
        indptr.append(len(idx))
    return d, n, np.array(indptr, np.uint64), np.array(idx, np.uint64), np.array(val, np.float32)


def test_seeded_projection_is_reproducible_and_matches_f64():
    d, n, ip, ix, v = tiny_fixture()
    basis = np.random.default_rng(123).normal(size=(d, 4)).astype(np.float32)
    a = orc.project(ip, ix, v, basis)
    b = orc.project(ip, ix, v, basis, nthreads=4)
    assert a.tobytes() == b.tobytes()
    # independent float64 restatement of :169-199 + :399-407
    X = np.zeros((n, d))
    for j in range(n):
        for t in range(int(ip[j]), int(ip[j + 1])):
            X[j, int(ix[t])] = math.log1p(float(v[t]))
    X /= np.maximum(np.sqrt((X ** 2).sum(1, keepdims=True)), 1e-8)
    P = X @ basis.astype(np.float64)
    P = (P - P.mean(1, keepdims=True)) / P.std(1, keepdims=True)
    assert P.max() <= 4 and P.min() >= -4
    assert close(a, P, 1e-5)
    other = orc.project(ip, ix, v, np.random.default_rng(124).normal(size=(d, 4)).astype(np.float32))
    assert not np.array_equal(a, other)


def test_batch_centring_and_clamp():
    rng = np.random.default_rng(5)
    raw = rng.normal(size=(200, 6)).astype(np.float32)
    raw[3, 2] = 40.0  # forces the clamp branch (:401-407)
    batch = (np.arange(200) % 3).astype(np.uint32)
    got = orc.project_finish(raw, batch, 3)
    P = raw.astype(np.float64)
    for b in range(3):
        P[batch == b] -= P[batch == b].mean(0, keepdims=True)
    P = (P - P.mean(1, keepdims=True)) / P.std(1, keepdims=True)
    # sqrt(K-1) is the largest |z| a K-vector can reach: K=6 gives 2.24 < 4, so no clamp here
    assert close(got, P, 1e-5)
    raw2 = rng.normal(size=(50, 50)).astype(np.float32)
    raw2[0, 0] = 1000.0
    got2 = orc.project_finish(raw2)
    Q = raw2.astype(np.float64)
    Q = (Q - Q.mean(1, keepdims=True)) / Q.std(1, keepdims=True)
    assert Q.max() > 4
    Q = np.clip(Q, -4, 4)
    Q = (Q - Q.mean(1, keepdims=True)) / Q.std(1, keepdims=True)
    assert close(got2, Q, 1e-5)


# ---- groups.rs:13-37, refine.rs:21-35, refine.rs:718-734 ---------------------------------------
def test_assign_groups_is_lexicographic_on_decimal_strings():
    codes = np.array([10, 2, 2, 33, 10, 7, 100], np.uint64)
    g, ng = orc.assign_groups(codes)
    # keys sorted as strings: "10" < "100" < "2" < "33" < "7"
    assert ng == 5 and list(g) == [0, 2, 2, 3, 0, 4, 1]
    g2, ng2 = orc.assign_groups_padded(codes, 101)
    assert ng2 == 5 and list(g2) == [2, 0, 0, 3, 2, 1, 4]


def test_level_sort_dims():
    assert orc.level_sort_dims(10, 3) == [10, 9, 7]  # 8.5 rounds half away from zero
    assert orc.level_sort_dims(10, 1) == [10]
    assert orc.level_sort_dims(10, 2) == [10, 7]
    assert orc.level_sort_dims(5, 3) == [5]
    assert orc.level_sort_dims(12, 4) == [12, 10, 9, 7]


# ---- matrix-util/tests/rsvd_tests.rs:22-32, 75-93 + binary_sort_columns properties -------------
def test_householder_q_is_orthonormal_and_spans_input():
    rng = np.random.default_rng(1)
    a = rng.normal(size=(15, 50)).astype(np.float32)
    q = orc.householder_q(a)
    assert np.allclose(q @ q.T, np.eye(15), atol=1e-5)
    # same column space: projecting a onto span(q) reproduces a
    assert np.allclose((a @ q.T) @ q, a, atol=1e-4)
    # first kk columns of Q span the first kk input columns (Gram-Schmidt property)
    assert np.allclose((a[:10] @ q[:10].T) @ q[:10], a[:10], atol=1e-4)


def test_jacobi_eig():
    rng = np.random.default_rng(2)
    m = rng.normal(size=(10, 10))
    g = m @ m.T
    ev, vec = orc.jacobi_eig(g)
    w = np.linalg.eigvalsh(g)[::-1]
    assert np.allclose(ev, w, rtol=1e-12)
    assert np.allclose(vec @ vec.T, np.eye(10), atol=1e-12)
    assert np.allclose(vec @ g @ vec.T, np.diag(ev), atol=1e-9 * ev[0])


def test_binary_codes_match_f64_svd_up_to_bit_complement():
    rng = np.random.default_rng(3)
    n, K, kk = 5000, 50, 10
    z = rng.normal(size=(n, 6)) @ rng.normal(size=(6, K)) + 0.3 * rng.normal(size=(n, K))
    proj = orc.project_finish(z.astype(np.float32))
    codes, q, u, sig, mean = orc.binary_codes(proj, kk, details=True)
    assert codes.max() < (1 << kk)
    again = orc.binary_codes(proj, kk)
    assert np.array_equal(codes, again)  # rsvd_tests.rs:22-32 reproducible
    assert np.allclose(u @ u.T, np.eye(kk), atol=1e-5)  # rsvd_tests.rs:75-93 orthonormal
    # independent f64 path: Q from numpy QR of the first kk+5 cells, SVD of B = Q^T X
    X = proj.astype(np.float64).T  # K × n
    Qn, _ = np.linalg.qr(X[:, :kk + 5])
    B = Qn[:, :kk].T @ X
    _, s, vt = np.linalg.svd(B, full_matrices=False)
    assert np.allclose(s, sig, rtol=1e-4)
    V = vt.T
    ref_bits = (V - V.mean(0)) > 0
    got_bits = ((codes[:, None] >> np.arange(kk, dtype=np.uint64)) & 1).astype(bool)
    for k in range(kk):
        agree = np.mean(ref_bits[:, k] == got_bits[:, k])
        assert max(agree, 1 - agree) > 0.999, (k, agree)


def test_independent_svd_oracle_against_numpy_and_the_mirror():
    """oracle_svd.cpp (f32 Householder bidiagonalisation + implicit-shift QR, the reference's route) is a second,
    independent statement of binary_sort_columns: its singular values match numpy's f64 SVD, its factor is
    orthonormal (rsvd_tests.rs:75-93), and the partitions it defines agree with numpy's and with the mirror oracle's
    (the one the GPU is held to bit for bit) up to per-bit complement on all but the cells whose standardised
    coordinate is rounding noise away from zero."""
    rng = np.random.default_rng(31)
    for n, K, kk in ((5000, 50, 10), (20000, 50, 10), (777, 24, 7)):
        z = rng.normal(size=(n, 6)) @ rng.normal(size=(6, K)) + 0.3 * rng.normal(size=(n, K))
        proj = orc.project_finish(z.astype(np.float32))
        codes, v, sig = orc.binary_codes_svd(proj, kk, details=True)
        assert np.array_equal(codes, orc.binary_codes_svd(proj, kk))  # reproducible, rsvd_tests.rs:22-32
        mirror, q, _, msig, _ = orc.binary_codes(proj, kk, details=True)
        X = proj.astype(np.float64)
        B = X @ q.astype(np.float64).T
        U, S, _ = np.linalg.svd(B, full_matrices=False)
        assert np.allclose(sig, S, rtol=1e-4) and np.allclose(msig, S, rtol=1e-4)
        # v holds the standardised columns: unit variance, zero mean, mutually orthogonal
        g = (v.astype(np.float64) @ v.astype(np.float64).T) / n
        assert np.allclose(g, np.eye(kk), atol=2e-2)  # centring costs a little orthogonality
        ref = np.zeros(n, np.uint64)
        for k in range(kk):
            ref |= ((U[:, k] - U[:, k].mean()) > 0).astype(np.uint64) << np.uint64(k)
        for other in (ref, mirror):
            agree = orc.partition_agreement(codes, other, kk)
            assert min(agree) > 0.998, agree
        # the cells that differ are exactly the ones sitting on the boundary
        diff = codes ^ mirror
        for k in range(kk):
            a = float(np.mean(((diff >> np.uint64(k)) & np.uint64(1)) == 0))
            flipped = (((diff >> np.uint64(k)) & np.uint64(1)) == (1 if a > 0.5 else 0))
            if flipped.any():
                assert np.abs(v[k][flipped]).max() < 1e-2, (k, np.abs(v[k][flipped]).max())


def test_collapse_basic_and_batch_against_dense():
    rng = np.random.default_rng(4)
    D, N, S, B = 40, 300, 7, 3
    dense = rng.poisson(0.2, size=(N, D)).astype(np.float32)
    ip, ix, v = _dense_cells_to_csc(dense, D)
    grp = rng.integers(0, S, N).astype(np.uint32)
    bat = rng.integers(0, B, N).astype(np.uint32)
    s, size = orc.collapse_basic(ip, ix, v, D, grp, S)
    for g in range(S):
        assert np.array_equal(s[g], dense[grp == g].sum(0))
        assert size[g] == (grp == g).sum()
    sdb, nbs = orc.collapse_batch(ip, ix, v, D, grp, bat, S, B)
    for b in range(B):
        assert np.array_equal(sdb[b], dense[bat == b].sum(0))
        for g in range(S):
            assert nbs[g, b] == ((grp == g) & (bat == b)).sum()
    f2c = (np.arange(S) % 3).astype(np.uint32)
    coarse = orc.merge_stat(s, f2c, 3)
    for c in range(3):
        assert np.array_equal(coarse[c], s[f2c == c].sum(0))
