"""GPU parity for the nnz streams either side of the hot path: per-gene running statistics
(K10) and the Nystrom re-projection (K11), through the C ABI against the CPU oracle on identical inputs."""
import numpy as np
import pytest

import oracle as orc
from util import close, max_err, nystrom_f64, random_csc

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.fixture(scope="module")
def lg():
    import legume_b200
    return legume_b200


@pytest.fixture(scope="module")
def ctx(lg):
    c = lg.Context(0)
    yield c
    c.close()


# ---- K10 ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("D,N,dens", [(700, 900, 0.1), (30000, 2000, 0.05), (64, 5000, 0.5), (70000, 300, 0.02)])
def test_row_stats_counts_bit_exact(lg, ctx, D, N, dens):
    """count data: npos, s1, s2 equal the reference's f32 folds exactly (all partial sums < 2^24); D = 70000 walks the
    gene axis in two shared-memory windows"""
    rng = np.random.default_rng(D + N)
    ip, ix, v = random_csc(rng, D, N, dens, empty_every=53)
    v[::17] = 0.0  # stored zeros: finite, not positive
    data = lg.SparseIoVec.from_csc(ctx, ip, ix, v, D)
    st = data.streaming_sparse_running_stats()
    npos, s1, s2 = orc.row_stats(ip, ix, v, D)
    assert st.ncols_processed() == N and st.nrows() == D
    assert np.array_equal(st.count_positives(), npos) and np.array_equal(st.sum(), s1)
    assert np.array_equal(st._s2.astype(np.float32), s2)
    mean, var, sd = orc.row_stats_moments(s1, s2, N)
    assert np.array_equal(st.mean(), mean) and np.array_equal(st.variance(), var)
    assert np.array_equal(st.std(), sd, equal_nan=True)


def test_row_stats_large_counts_and_block_merge(lg, ctx):
    """counts of 2^15 and above bypass the packed accumulator; two blocks merge to the whole (sparse_stat.rs:183-196)"""
    rng = np.random.default_rng(8)
    D, N = 500, 1200
    ip, ix, v = random_csc(rng, D, N, 0.2)
    v[::11] = 40000.0
    v[5] = 1048575.0
    whole = lg.SparseIoVec.from_csc(ctx, ip, ix, v, D).streaming_sparse_running_stats()
    want = [np.zeros(D), np.zeros(D), np.zeros(D)]
    rows = ix.astype(np.int64)
    np.add.at(want[0], rows, (v > 0).astype(np.float64))
    np.add.at(want[1], rows, v.astype(np.float64))
    np.add.at(want[2], rows, v.astype(np.float64) ** 2)
    assert all(np.array_equal(a, b) for a, b in zip((whole._npos, whole._s1, whole._s2), want))  # exact integers in f64
    parts = lg.SparseRunningStatistics(D)
    for lo, hi in ((0, 500), (500, N)):
        parts.add_block(ctx, lg.CscBlock.upload(ctx, ip, ix, v, D, lo, hi))
    assert parts.ncols_processed() == N
    assert all(np.array_equal(a, b) for a, b in zip((parts._npos, parts._s1, parts._s2), (whole._npos, whole._s1, whole._s2)))


def test_row_stats_general_values(lg, ctx):
    """fractional, negative and non-finite values (sparse_stat.rs:68-77): non-finite skipped, only v > 0 detected"""
    rng = np.random.default_rng(9)
    D, N = 300, 800
    ip, ix, v = random_csc(rng, D, N, 0.15)
    v = (v * rng.normal(size=len(v))).astype(np.float32)
    v[3], v[10], v[20] = np.inf, np.nan, -np.inf
    st = lg.SparseIoVec.from_csc(ctx, ip, ix, v, D).streaming_sparse_running_stats()
    npos, s1, s2 = orc.row_stats(ip, ix, v, D)
    assert np.array_equal(st.count_positives(), npos)
    assert close(st.sum(), s1, TOL) and close(st._s2.astype(np.float32), s2, TOL)
    # empty block
    e = lg.SparseIoVec.from_csc(ctx, np.zeros(4, np.uint64), np.zeros(0, np.uint64), np.zeros(0, np.float32), 7)
    s = e.streaming_sparse_running_stats()
    assert s.ncols_processed() == 3 and not s.sum().any() and not s.mean().any()


# ---- K11 ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("D,N,K,P", [(2000, 1500, 50, 0), (2000, 1500, 50, 9), (600, 700, 33, 4), (900, 400, 100, 0),
                                     (30000, 300, 50, 16)])
def test_nystrom_matches_oracle(lg, ctx, D, N, K, P):
    rng = np.random.default_rng(D + K + P)
    ip, ix, v = random_csc(rng, D, N, 0.05, empty_every=67)
    basis = rng.standard_normal((K, D)).astype(np.float32)
    delta = pb = None
    data = lg.SparseIoVec.from_csc(ctx, ip, ix, v, D)
    if P:
        delta = np.exp(0.5 * rng.standard_normal((P, D))).astype(np.float32)
        delta[:, ::13] = 0.0
        pb = rng.integers(0, P, N)
        data.assign_groups(pb)
        pb = data.get_group_membership()
    got = data.nystrom_project(basis, delta, 1e4)
    assert got.shape == (N, K)
    # the per-cell variance s2/n - mean^2 cancels ~3 digits in f32, so the reference's serial f32 folds (the oracle)
    # sit up to ~5e-4 from exact arithmetic and one ulp of libm difference moves them by 1e-4: the CUDA path
    # accumulates those sums in f64 and is held to 1e-5 against the float64 restatement of the reference's formulas,
    # and to the reference's own error against the f32 oracle
    exact = nystrom_f64(ip, ix, v, D, basis, delta, pb, 1e4)
    assert close(got, exact, TOL), max_err(got, exact)
    want = orc.nystrom_project(ip, ix, v, D, basis, delta, pb, 1e4)
    assert close(got, want, 2e-3), max_err(got, want)


def test_nystrom_edge_cases(lg, ctx):
    """empty and one-entry columns give zero rows; a constant column has sd = 0 (z - mean); cells whose pseudobulk is
    out of range are left unadjusted; run-to-run identical; both the tensor path (no divisor) and the CUDA-core path"""
    rng = np.random.default_rng(4)
    D, K = 50, 6
    ip = np.array([0, 0, 1, 5, 9], np.uint64)
    ix = np.array([3, 0, 1, 2, 3, 4, 5, 6, 7], np.uint64)
    v = np.array([5, 2, 2, 2, 2, 1, 3, 1, 2], np.float32)
    basis = rng.standard_normal((K, D)).astype(np.float32)
    blk = lg.CscBlock.upload(ctx, ip, ix, v, D)
    got = lg.nystrom_project(ctx, blk, basis)
    # zero rows: exactly zero on the CUDA-core path, zero up to the basis quantisation (~1e-7) on the tensor path
    assert np.abs(got[:3]).max() < 1e-5 and close(got, nystrom_f64(ip, ix, v, D, basis, None, None, 1e4), TOL)
    delta = np.exp(rng.standard_normal((2, D))).astype(np.float32)
    pb = np.array([0, 1, 7, 1], np.uint32)  # 7 >= P: unadjusted
    g2 = lg.nystrom_project(ctx, blk, basis, delta, pb)
    assert close(g2, nystrom_f64(ip, ix, v, D, basis, delta, pb, 1e4), TOL)
    assert close(g2[2], got[2], TOL) and not close(g2[3], got[3], 1e-3)  # cell 2 unadjusted, cell 3 adjusted
    assert g2.tobytes() == lg.nystrom_project(ctx, blk, basis, delta, pb).tobytes()
    with pytest.raises(lg.LegumeError):
        lg.nystrom_project(ctx, blk, basis[:, :10])


def test_nystrom_tensor_and_cuda_core_paths_agree(lg, ctx, monkeypatch):
    """without a divisor the counts of one share a value, so the pass runs as K1 does (0/1 pattern on tcgen05 + the
    counts above one on CUDA cores); LG_K11_CUDA_CORES=1 forces the warp-per-cell kernel: both within 1e-5 of float64"""
    rng = np.random.default_rng(31)
    D, N, K = 5000, 3000, 50
    ip, ix, v = random_csc(rng, D, N, 0.04, empty_every=101)
    v[::29] = 0.0      # stored zeros are entries of the column: z = 0 counts in the moments
    v[7::31] = 2.5     # fractional values are simply exceptions
    basis = (rng.standard_normal((K, D)) * (10.0 ** rng.uniform(-3, 1, K))[:, None]).astype(np.float32)  # U / sigma-like column scales
    blk = lg.CscBlock.upload(ctx, ip, ix, v, D)
    exact = nystrom_f64(ip, ix, v, D, basis, None, None, 1e4)
    cs = np.abs(exact).max(0) + 1e-30
    tensor = lg.nystrom_project(ctx, blk, basis)
    monkeypatch.setenv("LG_K11_CUDA_CORES", "1")
    cores = lg.nystrom_project(ctx, blk, basis)
    assert close(tensor / cs, exact / cs, TOL), max_err(tensor / cs, exact / cs)
    assert close(cores / cs, exact / cs, TOL), max_err(cores / cs, exact / cs)
