"""Pins the oracle's restatement of the steps either side of the hot path on the reference's own
known-answer tests, and checks it against independent float64 numpy restatements.  CPU only."""
import numpy as np

import oracle as orc
from util import close, nystrom_f64, random_csc


def csc_from_dense(a):
    a = np.asarray(a, np.float32)
    ip, ix, v = [0], [], []
    for j in range(a.shape[1]):
        rows = np.nonzero(a[:, j])[0]
        ix += list(rows)
        v += list(a[rows, j])
        ip.append(len(ix))
    return np.array(ip, np.uint64), np.array(ix, np.uint64), np.array(v, np.float32)


def test_row_stats_reference_known_answers():
    """matrix-util/src/sparse_stat.rs:671-728"""
    # test_sparse_running_stat_basic: columns [1,0,2,0] and [0,3,0,4]
    ip, ix, v = csc_from_dense([[1, 0], [0, 3], [2, 0], [0, 4]])
    npos, s1, s2 = orc.row_stats(ip, ix, v, 4)
    assert npos.tolist() == [1, 1, 1, 1] and s1.tolist() == [1, 3, 2, 4]
    mean, _, _ = orc.row_stats_moments(s1, s2, 2)
    assert np.allclose(mean, [0.5, 1.5, 1.0, 2.0], atol=1e-6)
    # test_sparse_running_stat_csc: 3 x 2 matrix [[1,0],[0,2],[3,0]]
    ip, ix, v = csc_from_dense([[1, 0], [0, 2], [3, 0]])
    npos, s1, _ = orc.row_stats(ip, ix, v, 3)
    assert npos.tolist() == [1, 1, 1] and s1.tolist() == [1, 2, 3]
    # test_sparse_running_stat_f64: columns {0:1, 1:2} and {0:3}
    ip, ix, v = csc_from_dense([[1, 3], [2, 0]])
    _, s1, s2 = orc.row_stats(ip, ix, v, 2)
    assert s1.tolist() == [4, 2]
    assert np.allclose(orc.row_stats_moments(s1, s2, 2)[0], [2.0, 1.0])


def test_row_stats_channelized_fixture_and_blocks():
    """data-beans-alg/tests/sparse_streaming_folded.rs:26-41: the 7 x 6 fixture; blocks merge by addition"""
    a = np.array([[3, 0, 1, 0, 5, 2], [4, 2, 0, 7, 1, 0], [0, 6, 3, 0, 0, 8], [1, 5, 0, 2, 0, 0], [2, 9, 0, 0, 4, 0],
                  [0, 0, 7, 3, 0, 1], [5, 0, 0, 6, 2, 0]], np.float32)
    ip, ix, v = csc_from_dense(a)
    npos, s1, s2 = orc.row_stats(ip, ix, v, 7)
    assert np.array_equal(npos, (a > 0).sum(1)) and np.array_equal(s1, a.sum(1)) and np.array_equal(s2, (a * a).sum(1))
    mean, var, sd = orc.row_stats_moments(s1, s2, 6)
    assert close(mean, a.mean(1)) and close(var, a.var(1)) and close(sd, a.std(1))
    halves = [csc_from_dense(a[:, :3]), csc_from_dense(a[:, 3:])]
    parts = [orc.row_stats(*h, 7) for h in halves]
    assert all(np.array_equal(parts[0][k] + parts[1][k], (npos, s1, s2)[k]) for k in range(3))


def test_row_stats_skips_non_finite_and_counts_only_positive():
    """sparse_stat.rs:68-77: non-finite values are skipped; a stored zero or a negative value is not a detection"""
    ip = np.array([0, 3, 5], np.uint64)
    ix = np.array([0, 1, 2, 0, 2], np.uint64)
    v = np.array([np.inf, 0.0, -2.0, np.nan, 3.0], np.float32)
    npos, s1, s2 = orc.row_stats(ip, ix, v, 3)
    assert npos.tolist() == [0, 0, 1] and s1.tolist() == [0, 0, 1] and s2.tolist() == [0, 0, 13]
    assert orc.row_stats_moments(np.zeros(2, np.float32), np.zeros(2, np.float32), 0)[0].tolist() == [0, 0]  # safe_denom


def test_nystrom_matches_float64_restatement():
    rng = np.random.default_rng(2)
    D, N, K, P = 400, 300, 20, 7
    ip, ix, v = random_csc(rng, D, N, 0.08, empty_every=41)
    basis = rng.standard_normal((K, D)).astype(np.float32)
    delta = np.exp(0.4 * rng.standard_normal((P, D))).astype(np.float32)
    delta[:, ::9] = 0.0  # non-positive divisors leave the entry alone
    pb = rng.integers(0, P, N).astype(np.uint32)
    for dl, pp in ((None, None), (delta, pb)):
        got = orc.nystrom_project(ip, ix, v, D, basis, dl, pp, 1e4)
        want = nystrom_f64(ip, ix, v, D, basis, dl, pp, 1e4)
        # the reference's f32 variance s2/n - mean^2 cancels about three digits (z ~ 5.5, sd ~ 0.2), so its own
        # result sits ~5e-4 from exact arithmetic; the CUDA path mirrors the f32 folds and is held to 1e-5 in test_gpu_next
        assert got.shape == (N, K) and close(got, want, 2e-3)
    # a one-entry column: sd = 0 -> z - mean = 0 -> a zero row; an empty column -> zeros
    assert not np.any(orc.nystrom_project(np.array([0, 1, 1], np.uint64), np.array([3], np.uint64), np.array([5.0], np.float32),
                                          D, basis))
