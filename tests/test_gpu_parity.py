"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on identical seeded
inputs.  Bit-exact for codes, groups, sums and neighbour sets; 1e-5 (the reference's mixed abs/rel
form) for projections and posteriors.  Run on the B200 box: pytest -m gpu."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle as orc
from util import close, max_err, random_csc, tiny_fixture, toy_stat

pytestmark = pytest.mark.gpu

TOL = 1e-5  # north_star: projections and posterior means within 1e-5 relative (mixed form)


@pytest.fixture(scope="module")
def lg():
    import legume_b200
    return legume_b200


@pytest.fixture(scope="module")
def ctx(lg):
    c = lg.Context(0)
    yield c
    c.close()


def basis_for(D, K, seed=7):
    return np.random.default_rng(seed).standard_normal((D, K)).astype(np.float32)


# ---- data feed ---------------------------------------------------------------------------------
def test_csc_upload_round_trip_and_range_check(lg, ctx):
    rng = np.random.default_rng(0)
    ip, ix, v = random_csc(rng, 300, 200, 0.1, empty_every=7)
    blk = lg.CscBlock.upload(ctx, ip, ix, v, 300)
    assert (blk.nrows, blk.ncols, blk.nnz) == (300, 200, len(v))
    ip2, ix2, v2 = blk.download()
    assert np.array_equal(ip, ip2) and np.array_equal(ix, ix2) and np.array_equal(v, v2)
    # a column sub-range, as SparseIoVec::read_columns_csc(lb..ub) slices (read.rs:204-220)
    sub = lg.CscBlock.upload(ctx, ip, ix, v, 300, 50, 120)
    sp, sx, sv = sub.download()
    assert np.array_equal(sp, ip[50:121] - ip[50]) and np.array_equal(sx, ix[ip[50]:ip[120]])
    bad = ix.copy()
    bad[3] = 300
    with pytest.raises(lg.LegumeError):
        lg.CscBlock.upload(ctx, ip, bad, v, 300)
    back = ip.copy()
    back[10], back[11] = back[11], back[10] - 1 if back[10] > 0 else 0  # a column pointer running backwards
    if back[11] < back[10]:
        with pytest.raises(lg.LegumeError):
            lg.CscBlock.upload(ctx, back, ix, v, 300)


@pytest.mark.parametrize("mode", ["default", "host_only", "host_no_pack", "host_no_gaps", "device_only"])
def test_csc_upload_large_block_host_narrowing(lg, ctx, mode, monkeypatch):
    """blocks of >= 8 Mi non-zeros are narrowed by host threads into a pinned ring — row indices as one-byte gaps to the
    previous entry (patch list for column starts and wide gaps, u32 when the list overflows), count values packed to
    bytes with a patch list — with a wide share sent through the device narrowing when the cores cannot keep up;
    every mix must give the same device arrays, bit for bit"""
    if mode == "host_only":
        monkeypatch.setenv("LG_UPLOAD_NO_WIDE", "1")
        monkeypatch.setenv("LG_UPLOAD_THREADS", "3")
    elif mode == "host_no_pack":
        monkeypatch.setenv("LG_UPLOAD_NO_PACK", "1")
    elif mode == "host_no_gaps":
        monkeypatch.setenv("LG_UPLOAD_NO_GAPS", "1")
        monkeypatch.setenv("LG_UPLOAD_NO_WIDE", "1")
    elif mode == "device_only":
        monkeypatch.setenv("LG_UPLOAD_THREADS", "0")
    rng = np.random.default_rng(5)
    D, N, per = 30000, 9000, 1000
    nnz = N * per + 5  # ragged last chunk
    ip = np.minimum(np.arange(N + 1, dtype=np.uint64) * per, nnz).astype(np.uint64)
    ip[-1] = nnz
    # canonical CSC (rows strictly ascending inside a column): position p of a column lands in [29 p, 29 p + 28]
    ix = ((np.arange(nnz, dtype=np.uint64) - np.repeat(ip[:-1], np.diff(ip).astype(np.int64))) * np.uint64(29)
          + rng.integers(0, 29, nnz, dtype=np.uint64))
    # ... gaps of 1..57 travel as one byte; every 37th column gets a gap of 256..800 somewhere (patch list), one column a gap
    # of exactly 255 and one of exactly 256 (the largest byte / the smallest patch)
    for c in range(3, N, 37):
        ix[int(ip[c]) + 100 + (c % 800):int(ip[c + 1])] += np.uint64(256 + c % 545)
    ix[int(ip[11]) + 500:int(ip[12])] += np.uint64(255) - (ix[int(ip[11]) + 500] - ix[int(ip[11]) + 499])
    ix[int(ip[13]) + 500:int(ip[14])] += np.uint64(256) - (ix[int(ip[13]) + 500] - ix[int(ip[13]) + 499])
    assert ix.max() < D
    v = rng.integers(0, 256, nnz).astype(np.float32)  # whole numbers 0..254 travel as bytes ...
    v[3_000_000:3_000_010] = 0.5                      # ... anything else (255 too) through the chunk's patch list ...
    v[5_000_001] = 70000.0
    v[(2 << 20) * 3 - 1] = -0.0
    v[6_500_000:6_600_000] = 1000.5                   # ... and a chunk with more than 65536 of those as raw f32
    blk = lg.CscBlock.upload(ctx, ip, ix, v, D)
    ip2, ix2, v2 = blk.download()
    blk.free()
    assert np.array_equal(ip, ip2) and np.array_equal(ix, ix2) and v.tobytes() == v2.tobytes()
    for pos, val in ((7, D), (nnz - 2, D + 5), (nnz // 2, 1 << 40)):
        bad = ix.copy()
        bad[pos] = val
        with pytest.raises(lg.LegumeError):
            lg.CscBlock.upload(ctx, ip, bad, v, D)
    # rows running backwards / repeated inside a column are not canonical CSC: rejected, not silently mis-computed
    for pos in (5, nnz // 3):
        bad = ix.copy()
        bad[pos] = bad[pos - 1] if pos % per else bad[pos + 1]
        with pytest.raises(lg.LegumeError, match="canonical"):
            lg.CscBlock.upload(ctx, ip, bad, v, D)


def test_csc_upload_many_short_columns_overflow_the_gap_patches(lg, ctx, monkeypatch):
    """400k columns of 21 entries about 1400 rows apart: every gap is a patch, every chunk's list overflows and the chunk
    travels as u32 — same arrays"""
    monkeypatch.setenv("LG_UPLOAD_NO_WIDE", "1")
    rng = np.random.default_rng(15)
    D, N, per = 30000, 400_000, 21
    nnz = N * per
    ip = (np.arange(N + 1, dtype=np.uint64) * np.uint64(per))
    ix = (np.tile(np.arange(per, dtype=np.uint64) * np.uint64(1400), N) + rng.integers(0, 1400, nnz, dtype=np.uint64))
    ix[::per][: N // 2] = 0  # half of the columns start at row 0
    v = rng.integers(1, 5, nnz).astype(np.float32)
    blk = lg.CscBlock.upload(ctx, ip, ix, v, D)
    ip2, ix2, v2 = blk.download()
    blk.free()
    assert np.array_equal(ip, ip2) and np.array_equal(ix, ix2) and v.tobytes() == v2.tobytes()


def test_sparse_io_vec_from_several_backends_with_row_remaps(lg, ctx):
    """SparseIoVec::read_columns_csc (read.rs:202-219): every backend's local rows are mapped into the union of rows;
    the joined block equals the matrix assembled by hand"""
    rng = np.random.default_rng(6)
    D = 120
    parts, dense = [], []
    for nloc, ncol in ((80, 40), (120, 25), (55, 60)):
        remap = rng.permutation(D)[:nloc].astype(np.uint32)  # local row -> union row
        ip, ix, v = random_csc(rng, nloc, ncol, 0.15)
        parts.append((ip, ix, v, remap))
        a = np.zeros((D, ncol), np.float32)
        for j in range(ncol):
            a[remap[ix[ip[j]:ip[j + 1]].astype(np.int64)], j] = v[ip[j]:ip[j + 1]]
        dense.append(a)
    whole = np.concatenate(dense, axis=1)
    data = lg.SparseIoVec.from_backends(ctx, parts, D)
    assert (data.num_rows(), data.num_columns()) == (D, 125)
    st = data.streaming_sparse_running_stats()
    assert np.array_equal(st.sum(), whole.sum(1)) and np.array_equal(st.count_positives(), (whole > 0).sum(1))
    one = lg.SparseIoVec.from_backends(ctx, parts[:1], D)  # a single backend
    assert np.array_equal(one.streaming_sparse_running_stats().sum(), dense[0].sum(1))
    bad = (parts[0][0], parts[0][1], parts[0][2], parts[0][3][:10])
    with pytest.raises(lg.LegumeError):
        lg.SparseIoVec.from_backends(ctx, [bad, parts[1]], D)


def _canonical_from_dense(a):
    """dense (D, N) -> canonical CSC, the form read_columns_csc hands out"""
    ip, ix, v = [0], [], []
    for j in range(a.shape[1]):
        r = np.nonzero(a[:, j])[0]
        ix.extend(r.tolist())
        v.extend(a[r, j].tolist())
        ip.append(len(ix))
    return np.array(ip, np.uint64), np.array(ix, np.uint64), np.array(v, np.float32)


def test_remapped_backends_are_canonical_for_the_batch_arm(lg, ctx):
    """read.rs:246-281: a permuted remap leaves a backend's columns unsorted, a many-to-one remap leaves duplicate rows
    (summed), a row the shared axis lacks is dropped.  The block must come out canonical — the matched-column kernel
    binary-searches a column's rows — and collapse_columns (batch arm) on it must equal the oracle on the matrix
    assembled by hand."""
    rng = np.random.default_rng(16)
    D, K, knn = 150, 20, 5
    parts, dense = [], []
    for nloc, ncol, kind in ((100, 260, "perm"), (150, 240, "many_to_one"), (90, 200, "drop")):
        ip, ix, v = random_csc(rng, nloc, ncol, 0.2)
        if kind == "perm":
            remap = rng.permutation(D)[:nloc].astype(np.uint32)
        elif kind == "many_to_one":
            remap = (rng.permutation(nloc) // 2).astype(np.uint32)          # two local rows per union row
        else:
            remap = rng.permutation(D)[:nloc].astype(np.uint32)
            remap[rng.choice(nloc, 20, replace=False)] = 0xFFFFFFFF         # g2c == None
        parts.append((ip, ix, v, remap))
        a = np.zeros((D, ncol), np.float32)
        for j in range(ncol):
            for t in range(int(ip[j]), int(ip[j + 1])):
                g = remap[int(ix[t])]
                if g != 0xFFFFFFFF:
                    a[g, j] += v[t]
        dense.append(a)
    whole = np.concatenate(dense, axis=1)
    N = whole.shape[1]
    data = lg.SparseIoVec.from_backends(ctx, parts, D)
    ip2, ix2, v2 = data.block.download()
    wip, wix, wv = _canonical_from_dense(whole)
    assert np.array_equal(ip2, wip) and np.array_equal(ix2, wix) and np.array_equal(v2, wv)
    # a backend row outside the remap is an error when the remap's length is known
    with pytest.raises(lg.LegumeError):
        lg.CscBlock.upload(ctx, parts[0][0], parts[0][1], parts[0][2], D, row_remap=parts[0][3][:10])
    # the batch arm (per-cell matched statistics) on the joined block
    proj = (rng.standard_normal((6, K))[rng.integers(0, 6, N)] * 3 + 0.3 * rng.standard_normal((N, K))).astype(np.float32)
    batch = np.concatenate([np.full(a.shape[1], b, np.uint32) for b, a in enumerate(dense)])
    data.build_hnsw_per_batch(proj, batch)
    data.partition_columns_to_groups(proj, 4)
    out, stat = data.collapse_columns(knn_batches=2, knn_cells=knn, num_opt_iter=10)
    grp, S = np.asarray(data.col_to_group), data.num_groups()
    order, _ = orc.batch_proximity(proj, batch, 3)
    midx, mdist = orc.knn_match_batches(proj, batch, 3, knn, order)
    obs, _ = orc.collapse_basic(wip, wix, wv, D, grp, S)
    imp, res = orc.collect_matched_stat(wip, wix, wv, D, grp, S, midx, mdist)
    assert np.array_equal(stat.observed_sum_ds, obs)
    assert close(stat.imputed_sum_ds, imp, 1e-5) and close(stat.residual_sum_ds, res, 1e-4)
    # wrapping device arrays that are not canonical is caught by the first kernel that relies on the order
    import torch
    bad_ix = ix2.astype(np.int32).copy()
    lo, hi = int(ip2[3]), int(ip2[4])
    if hi - lo >= 2:
        bad_ix[lo], bad_ix[lo + 1] = bad_ix[lo + 1], bad_ix[lo]
        blk = lg.CscBlock.wrap_device(ctx, torch.from_numpy(ip2.astype(np.int64)).cuda(), torch.from_numpy(bad_ix).cuda(),
                                      torch.from_numpy(v2).cuda(), D)
        with pytest.raises(lg.LegumeError, match="canonical"):
            lg.SparseIoVec(ctx, blk).streaming_sparse_running_stats()


def test_sim_matches_cpu_twin(lg, ctx):
    from legume_b200 import sim
    tabs = sim.make_tables(700, ntopic=4, nbatch=2, depth=400, seed=11)
    blk, topic, batch = sim.sim_block(ctx, tabs, 100, 420)
    ip, ix, v = blk.download()
    oip, oix, ov = orc.sim_poisson_csc(tabs.seed, 700, 100, 420, topic, batch, 4, 2, tabs.lam, tabs.p0, tabs.npiece)
    assert np.array_equal(ip, oip) and np.array_equal(ix, oix) and np.array_equal(v, ov)
    assert 0.2 < len(v) / (700 * 320) < 0.8 and v.min() >= 1.0


# ---- stage 1 -----------------------------------------------------------------------------------
@pytest.mark.parametrize("D,N,K,nb", [(2000, 3000, 50, 0), (2000, 3000, 50, 3), (500, 1500, 20, 1), (800, 700, 64, 2),
                                      (300, 100, 100, 0)])
def test_project_matches_oracle(lg, ctx, D, N, K, nb):
    rng = np.random.default_rng(D + N + K)
    ip, ix, v = random_csc(rng, D, N, 0.05, empty_every=97)
    basis = basis_for(D, K)
    batch = None if nb == 0 else rng.integers(0, nb, N).astype(np.uint32)
    data = lg.SparseIoVec.from_csc(ctx, ip, ix, v, D)
    _, got = data.project_columns_with_batch_correction(K, None, batch, basis=basis)
    want = orc.project(ip, ix, v, basis, batch, nb, nthreads=4)
    assert got.shape == (N, K)
    assert close(got, want, TOL), max_err(got, want)
    # raw stage alone
    raw = np.empty((N, K), np.float32)
    ctx.check(lg.lib.lg_project_raw(ctx.h, data.block.h, basis.ctypes.data, K, raw.ctypes.data))
    assert close(raw, orc.project_raw(ip, ix, v, basis, 4), TOL)


def test_project_tiny_fixture_reproducible(lg, ctx):
    """random_projection.rs:607-645: same basis -> byte-identical, different basis -> different"""
    d, n, ip, ix, v = tiny_fixture()
    data = lg.SparseIoVec.from_csc(ctx, ip, ix, v, d)
    b1, b2 = basis_for(d, 4, 123), basis_for(d, 4, 124)
    _, a = data.project_columns(4, None, basis=b1)
    _, b = data.project_columns(4, None, basis=b1)
    _, c = data.project_columns(4, None, basis=b2)
    assert a.tobytes() == b.tobytes() and not np.array_equal(a, c)
    assert close(a, orc.project(ip, ix, v, b1), TOL)


def test_project_clamp_branch_and_weights(lg, ctx):
    rng = np.random.default_rng(3)
    D, N, K = 400, 600, 50
    ip, ix, v = random_csc(rng, D, N, 0.08)
    basis = basis_for(D, K)
    basis[:, 0] *= 60.0  # one dominant dim pushes standardised values beyond 4 -> clamp + re-standardise
    data = lg.SparseIoVec.from_csc(ctx, ip, ix, v, D)
    _, got = data.project_columns(K, None, basis=basis)
    want = orc.project(ip, ix, v, basis)
    pre = orc.project_raw(ip, ix, v, basis)
    z = (pre - pre.mean(1, keepdims=True)) / pre.std(1, keepdims=True)
    assert np.abs(z).max() > 4.0, "fixture must exercise the clamp branch"
    assert close(got, want, TOL), max_err(got, want)
    # weighted variant (:417-495): zero weights drop rows from the geometry, not from the norm
    w = np.ones(D, np.float32)
    w[::3] = 0.0
    w[1::3] = 2.5
    _, gw = data.project_columns_weighted(K, None, None, w, basis=basis_for(D, K, 9))
    bw = basis_for(D, K, 9).copy()
    bw[::3] = 0.0
    bw[1::3] *= 2.5
    assert close(gw, orc.project(ip, ix, v, bw), TOL)


def test_project_raw_columns_of_very_different_scale(lg, ctx):
    """the tensor path quantises every basis column on its own power-of-two grid, so a column 1e6 times smaller or 1e4
    times larger than its neighbours keeps the same RELATIVE accuracy (row-weighted bases, U / sigma of a Nystrom basis)"""
    rng = np.random.default_rng(12)
    D, N, K = 3000, 2048, 50
    ip, ix, v = random_csc(rng, D, N, 0.05)
    col_scale = (10.0 ** rng.uniform(-6, 4, K)).astype(np.float32)
    basis = basis_for(D, K) * col_scale[None, :]
    data = lg.SparseIoVec.from_csc(ctx, ip, ix, v, D)
    raw = np.empty((N, K), np.float32)
    ctx.check(lg.lib.lg_project_raw(ctx.h, data.block.h, basis.ctypes.data, K, raw.ctypes.data))
    want = orc.project_raw(ip, ix, v, basis, 4)
    assert close(raw / col_scale[None, :], want / col_scale[None, :], TOL), max_err(raw / col_scale, want / col_scale)
    # a non-finite column cannot be quantised: the call falls back to the CUDA-core kernel and propagates it
    basis[7, 3] = np.inf
    ctx.check(lg.lib.lg_project_raw(ctx.h, data.block.h, basis.ctypes.data, K, raw.ctypes.data))
    assert not np.isfinite(raw[:, 3]).all() and np.isfinite(np.delete(raw, 3, axis=1)).all()


def test_project_raw_bits_do_not_depend_on_where_a_column_starts(lg, ctx):
    """a cell's raw projection must be bit-identical whether its block starts at column 0 or anywhere else (any alignment
    of its first entry in the index / value arrays): that is what makes the cell-sharded run reproduce the single-GPU
    bits.  The same for the Nystrom pass."""
    rng = np.random.default_rng(77)
    D, N, K = 4000, 1500, 50
    ip, ix, v = random_csc(rng, D, N, 0.05, empty_every=89)
    basis = basis_for(D, K)
    whole = lg.CscBlock.upload(ctx, ip, ix, v, D)
    raw = np.empty((N, K), np.float32)
    ctx.check(lg.lib.lg_project_raw(ctx.h, whole.h, basis.ctypes.data, K, raw.ctypes.data))
    ny = lg.nystrom_project(ctx, whole, np.ascontiguousarray(basis.T))
    for lo, hi in ((1, 700), (2, 903), (3, 1500), (517, 1499), (1000, 1001)):
        sub = lg.CscBlock.upload(ctx, ip, ix, v, D, lo, hi)
        part = np.empty((hi - lo, K), np.float32)
        ctx.check(lg.lib.lg_project_raw(ctx.h, sub.h, basis.ctypes.data, K, part.ctypes.data))
        assert part.tobytes() == raw[lo:hi].tobytes(), (lo, hi)
        assert lg.nystrom_project(ctx, sub, np.ascontiguousarray(basis.T)).tobytes() == ny[lo:hi].tobytes(), (lo, hi)


def test_project_fused_form_is_bit_identical(lg, ctx, monkeypatch):
    """LG_K1_FUSED=1 runs K1 as the one-kernel form (producer warps build the pattern bitmap in shared memory from per-cell
    cursors, k_project_finalize folds the exception lists): an independent second scan of the same stream whose raw
    projection must equal the default two-kernel form bit for bit — counts, empty columns, a ragged last supertile, blocks
    that start at any alignment, and non-count values (every entry an exception: the list overflows and the cell is
    re-scanned)."""
    rng = np.random.default_rng(5)
    K = 50
    cases = []
    for D, N, dens, ee in ((4000, 1500, 0.05, 89), (9000, 700, 0.2, 0), (300, 257, 0.5, 7), (20000, 513, 0.02, 0)):
        cases.append((D, N) + random_csc(rng, D, N, dens, empty_every=ee))
    D, N = 5000, 600
    ip, ix, v = random_csc(rng, D, N, 0.08)
    cases.append((D, N, ip, ix, (v + rng.uniform(0.1, 0.9, v.size)).astype(np.float32)))  # no entry equals one
    for D, N, ip, ix, v in cases:
        basis = basis_for(D, K, 9)
        for lo, hi in ((0, N), (3, N - 1)):
            blk = lg.CscBlock.upload(ctx, ip, ix, v, D, lo, hi)
            got = {}
            for mode in ("0", "1"):
                monkeypatch.setenv("LG_K1_FUSED", mode)
                raw = np.full((hi - lo, K), np.nan, np.float32)
                ctx.check(lg.lib.lg_project_raw(ctx.h, blk.h, basis.ctypes.data, K, raw.ctypes.data))
                got[mode] = raw
            assert np.isfinite(got["1"]).all()
            assert got["0"].tobytes() == got["1"].tobytes(), (D, N, lo, hi)


def test_assign_groups_downsampled_and_gamma_vconcat(lg, ctx):
    """assign_groups(.., ncolumns_per_group) (groups.rs:13-37 + utils.rs:36-66): the columns left out of a down-sampled group
    belong to no group and every statistic skips them; GammaMatrix::vconcat (dmatrix_gamma.rs:301-326) stacks gene blocks
    of calibrated planes into the planes of the whole fit"""
    rng = np.random.default_rng(12)
    D, N, S = 300, 2000, 6
    ip, ix, v = random_csc(rng, D, N, 0.08)
    labels = rng.integers(0, S, N)
    data = lg.SparseIoVec.from_csc(ctx, ip, ix, v, D)
    data.assign_groups(labels, ncolumns_per_group=150)
    grp = np.asarray(data.get_group_membership())
    assert all((grp == k).sum() == min(150, (labels == k).sum()) for k in range(S)) and (grp == 0xFFFFFFFF).sum() == N - 6 * 150
    stat = lg.CollapsedStat(D, S, 1)
    data.collect_basic_stat(stat)
    wsum, wsize = orc.collapse_basic(ip, ix, v, D, grp, S)
    assert np.array_equal(stat.observed_sum_ds, wsum) and np.array_equal(stat.size_s, wsize) and np.all(wsize == 150)
    # vconcat: fit gene blocks separately, stack, compare with the whole fit
    whole = lg.GammaMatrix(ctx, (D, S), 1.0, 1.0)
    whole.update_stat(stat.observed_sum_ds, np.repeat(stat.size_s[:, None], D, axis=1))
    whole.calibrate()
    blocks = []
    for r0 in range(0, D, 110):
        sub = stat.select_rows(r0, min(110, D - r0))
        b = lg.GammaMatrix(ctx, (sub.num_genes(), S), 1.0, 1.0)
        b.update_stat(sub.observed_sum_ds, np.repeat(sub.size_s[:, None], sub.num_genes(), axis=1))
        b.calibrate()
        blocks.append(b)
    cat = lg.GammaMatrix.vconcat(blocks, stack_stats=False)
    assert (cat.nrows(), cat.ncols()) == (D, S) and cat.a_stat is None
    for plane in ("estimated_mean", "estimated_sd", "estimated_log_mean", "estimated_log_sd"):
        assert np.array_equal(getattr(cat, plane), getattr(whole, plane)), plane
    assert np.array_equal(lg.GammaMatrix.vconcat(blocks, True).a_stat, whole.a_stat)


def test_tensor_path_fallbacks_are_counted_not_silent(lg, ctx, capfd):
    """a call the tcgen05 paths cannot take (K > 53 for the projection) runs on the CUDA-core kernel with the same contract,
    and says so: once on stderr per distinct reason, and in lg_ctx_fallback_count / lg_ctx_last_fallback"""
    rng = np.random.default_rng(3)
    D, N, K = 600, 300, 60
    ip, ix, v = random_csc(rng, D, N, 0.1)
    basis = basis_for(D, K, 2)
    blk = lg.CscBlock.upload(ctx, ip, ix, v, D)
    before = lg.lib.lg_ctx_fallback_count(ctx.h)
    raw = np.empty((N, K), np.float32)
    for _ in range(2):
        ctx.check(lg.lib.lg_project_raw(ctx.h, blk.h, basis.ctypes.data, K, raw.ctypes.data))
    assert lg.lib.lg_ctx_fallback_count(ctx.h) == before + 2
    assert b"K = 60" in lg.lib.lg_ctx_last_fallback(ctx.h)
    assert capfd.readouterr().err.count("outside the tensor-core path") == 1  # said once, counted every time
    want = orc.project_raw(ip, ix, v, basis) if hasattr(orc, "project_raw") else None
    if want is not None:
        assert close(raw, want, TOL)
    # a call the tensor path takes does not count
    b50, r50 = np.ascontiguousarray(basis[:, :50]), np.empty((N, 50), np.float32)  # kept alive across the call
    ctx.check(lg.lib.lg_project_raw(ctx.h, blk.h, b50.ctypes.data, 50, r50.ctypes.data))
    assert lg.lib.lg_ctx_fallback_count(ctx.h) == before + 2


def test_hotpath_run_sharded_one_rank_equals_staged_path(lg, ctx):
    """lg_hotpath_run_sharded (lg_comm.cu) with a world of one: the whole single-batch arm as ONE C-ABI call must equal the
    staged calls of legume_b200.pipeline bit for bit (the multi-rank comparison is tools/check_multi_gpu.py), with and
    without batch labels; lg_comm_init / lg_comm_info / lg_allreduce_stats are no-ops on one rank"""
    import torch
    from legume_b200 import sim
    from legume_b200.pipeline import HotPath
    D, N, K, kk = 3000, 5000, 50, 8
    tabs = sim.make_tables(D, ntopic=5, nbatch=3, depth=300, seed=9)
    blk, _, batch_h = sim.sim_block(ctx, tabs, 0, N)
    basis = torch.from_numpy(basis_for(D, K, 4)).cuda()
    batch = torch.from_numpy(batch_h.astype(np.int32)).cuda()
    hp = HotPath(ctx)
    for b, nb in ((batch, 3), (None, 0)):
        a = hp.run(blk, basis, b, nb, kk)
        c = hp.run_native(blk, basis, b, nb, kk)
        assert a["num_groups"] == c["num_groups"]
        for key in ("proj", "codes", "group", "sum_ds", "size_s"):
            assert torch.equal(a[key], c[key]), key
        for key in ("mean", "sd", "log_mean", "log_sd"):
            assert torch.equal(a["posterior"][key], c["posterior"][key]), key
    r, w = C.c_int(-1), C.c_int(-1)
    ctx.check(lg.lib.lg_comm_info(ctx.h, C.byref(r), C.byref(w)))
    assert (r.value, w.value) == (0, 1)
    s = a["sum_ds"].clone()
    ctx.check(lg.lib.lg_allreduce_stats(ctx.h, lg._ptr(s), None, None, None, D, a["num_groups"], 0))
    assert torch.equal(s, a["sum_ds"])


def _pattern_case(rng, D, N, density, p_geom, max_count, zeros=0, always=(), empty_every=0):
    ip, ix, v = random_csc(rng, D, N, density, max_count=max_count, empty_every=empty_every)
    ip, ix, v = ip.astype(np.int64), ix.astype(np.int64), v.copy()
    v[:] = np.minimum(rng.geometric(p_geom, size=v.size), max_count).astype(np.float32)
    if zeros:
        v[rng.choice(v.size, size=zeros, replace=False)] = 0.0  # stored zeros: bit set, count 0
    if len(always):  # genes present in every non-empty cell: a chunk of 256 cells of one group counts to 256 (the ninth plane)
        cols = []
        for j in range(N):
            rows, vals = ix[ip[j]:ip[j + 1]], v[ip[j]:ip[j + 1]]
            if rows.size:
                extra = np.setdiff1d(np.asarray(always), rows)
                rows = np.concatenate([rows, extra])
                vals = np.concatenate([vals, np.ones(extra.size, np.float32)])
                o = np.argsort(rows, kind="stable")
                rows, vals = rows[o], vals[o]
            cols.append((rows, vals))
        ip = np.concatenate([[0], np.cumsum([c[0].size for c in cols])])
        ix = np.concatenate([c[0] for c in cols])
        v = np.concatenate([c[1] for c in cols]).astype(np.float32)
    return ip.astype(np.uint64), ix.astype(np.uint64), v


@pytest.mark.parametrize("case", ["ones_mostly", "many_listed", "stored_zeros", "huge_count", "too_many_genes", "fractional", "list_overflows"])
def test_collapse_from_the_projection_pattern_equals_the_csc_collapse(lg, ctx, case):
    """Inside lg_hotpath_run_sharded K5 sums the groups from the 1-bit pattern + list of counts != 1 that K1's scan left
    behind (k_collapse_pattern, lg_collapse.cu) instead of streaming the CSC arrays again.  Whole-number sums: they must
    equal the CSC kernel's (the staged path) bit for bit — ragged cell / gene counts, empty cells, a gene present in all 256
    cells of a chunk, cells with more than 128 listed entries, stored zeros — and a block the pattern cannot express
    (a count above 32 767, a fractional value, more than 32 768 genes, more than half of a cell's entries not ones)
    must quietly take the CSC kernel."""
    import torch
    from legume_b200.pipeline import HotPath
    rng = np.random.default_rng(77)
    K = 50
    if case == "ones_mostly":
        D, N, kk = 5000, 1300, 1
        ip, ix, v = _pattern_case(rng, D, N, 0.08, 0.92, 9, always=(7, 4999), empty_every=17)
    elif case == "many_listed":
        D, N, kk = 4100, 700, 6
        ip, ix, v = _pattern_case(rng, D, N, 0.2, 0.75, 40)  # ~820 entries per cell, ~205 of them listed (room for 411)
    elif case == "list_overflows":
        D, N, kk = 4100, 700, 6
        ip, ix, v = _pattern_case(rng, D, N, 0.2, 0.3, 40)  # 70 % of the entries are not ones: the lists (half a cell's entries) do not fit
    elif case == "stored_zeros":
        D, N, kk = 2048, 513, 3
        ip, ix, v = _pattern_case(rng, D, N, 0.1, 0.8, 5, zeros=900)
    elif case == "huge_count":
        D, N, kk = 3000, 600, 4
        ip, ix, v = _pattern_case(rng, D, N, 0.05, 0.8, 5)
        v[v.size // 2] = 40000.0
    elif case == "fractional":
        D, N, kk = 3000, 600, 4
        ip, ix, v = _pattern_case(rng, D, N, 0.05, 0.8, 5)
        v[v.size // 3] = 2.5
    else:
        D, N, kk = 33000, 300, 3
        ip, ix, v = _pattern_case(rng, D, N, 0.01, 0.8, 5)
    blk = lg.CscBlock.upload(ctx, ip, ix, v, D)
    basis = torch.from_numpy(basis_for(D, K, 5)).cuda()
    hp = HotPath(ctx)
    before = lg.lib.lg_ctx_pattern_collapse_count(ctx.h)
    c = hp.run_native(blk, basis, None, 0, kk)
    took = lg.lib.lg_ctx_pattern_collapse_count(ctx.h) - before
    assert took == (0 if case in ("huge_count", "too_many_genes", "fractional", "list_overflows") else 1)
    a = hp.run(blk, basis, None, 0, kk)  # staged calls: lg_collapse_basic on the CSC arrays
    assert a["num_groups"] == c["num_groups"]
    for key in ("proj", "group", "sum_ds", "size_s"):
        assert torch.equal(a[key], c[key]), key
    # and against the arrays themselves
    dense = np.zeros((c["num_groups"], D), np.float64)
    g = c["group"].cpu().numpy()
    for j in range(N):
        np.add.at(dense[g[j]], ix[ip[j]:ip[j + 1]].astype(np.int64), v[ip[j]:ip[j + 1]])
    assert np.array_equal(c["sum_ds"].cpu().numpy().astype(np.float64), dense)
    os.environ["LG_COLLAPSE_PATTERN"] = "0"
    try:
        off = hp.run_native(blk, basis, None, 0, kk)
    finally:
        del os.environ["LG_COLLAPSE_PATTERN"]
    assert lg.lib.lg_ctx_pattern_collapse_count(ctx.h) - before == took
    assert torch.equal(off["sum_ds"], c["sum_ds"]) and torch.equal(off["size_s"], c["size_s"])


def test_block_that_keeps_its_pattern_collapses_from_it(lg, ctx):
    """lg_csc_keep_pattern: after a projection of the block, lg_collapse_basic and lg_collapse_batch (unit multiplicities) sum
    the pattern; with multiplicities, before any projection, after an exact-order projection (no scan ran) or once the
    buffers are released they stream the arrays.  The sums are the same bits every time."""
    import torch
    from legume_b200.pipeline import HotPath
    rng = np.random.default_rng(5)
    D, N, K, kk, B = 4500, 1100, 50, 5, 3
    ip, ix, v = _pattern_case(rng, D, N, 0.07, 0.9, 12, zeros=40, empty_every=23)
    blk = lg.CscBlock.upload(ctx, ip, ix, v, D)
    basis = torch.from_numpy(basis_for(D, K, 6)).cuda()
    batch = torch.from_numpy(rng.integers(0, B, N).astype(np.int32)).cuda()
    hp = HotPath(ctx)
    count = lambda: lg.lib.lg_ctx_pattern_collapse_count(ctx.h)
    ref = hp.run(blk, basis, batch, B, kk)  # no pattern kept: the CSC kernels
    ref_db, ref_nbs = hp.collapse_batch(blk, ref["group"], batch, ref["num_groups"], B)
    blk.keep_pattern()
    c0 = count()
    s0, z0 = hp.collapse_basic(blk, ref["group"], ref["num_groups"])  # nothing projected yet: arrays
    assert count() == c0 and torch.equal(s0, ref["sum_ds"]) and torch.equal(z0, ref["size_s"])
    got = hp.run(blk, basis, batch, B, kk)  # projection fills the pattern, the collapse of the same pass sums it
    assert count() == c0 + 1
    for key in ("proj", "group", "sum_ds", "size_s"):
        assert torch.equal(got[key], ref[key]), key
    db, nbs = hp.collapse_batch(blk, ref["group"], batch, ref["num_groups"], B)  # label = batch
    assert count() == c0 + 2 and torch.equal(db, ref_db) and torch.equal(nbs, ref_nbs)
    coarse = (ref["group"] // 4).contiguous()  # another level of the multilevel scheme: same block, coarser labels
    ng = int(coarse.max().item()) + 1
    s1, z1 = hp.collapse_basic(blk, coarse, ng)
    assert count() == c0 + 3
    dense = np.zeros((ng, D), np.float64)
    g = coarse.cpu().numpy()
    for j in range(N):
        np.add.at(dense[g[j]], ix[ip[j]:ip[j + 1]].astype(np.int64), v[ip[j]:ip[j + 1]])
    assert np.array_equal(s1.cpu().numpy().astype(np.float64), dense)
    assert np.array_equal(z1.cpu().numpy(), np.bincount(g, minlength=ng).astype(np.float32))
    blk.keep_pattern(False)
    s2, _ = hp.collapse_basic(blk, coarse, ng)
    assert count() == c0 + 3 and torch.equal(s2, s1)


def test_sparse_io_stack_projects_every_modality_and_stacks(lg, ctx):
    """RandProjOps for SparseIoStack (random_projection.rs:200-340): per-modality projection with its own basis, vertical
    concatenation of bases (rows) and projections (dims); batch labels cut to the shared column count; one set of codes"""
    rng = np.random.default_rng(21)
    N, K = 900, 12
    mods = [random_csc(rng, D, N, 0.1) + (D,) for D in (300, 170)]
    vecs = [lg.SparseIoVec.from_csc(ctx, ip, ix, v, D) for ip, ix, v, D in mods]
    bases = [basis_for(D, K, 40 + m) for m, (_, _, _, D) in enumerate(mods)]
    batch = rng.integers(0, 3, N + 5)  # longer than the columns: `.get(0..ncols)`
    stack = lg.SparseIoStack(vecs)
    assert (stack.num_columns(), stack.num_rows()) == (N, 470)
    basis, proj = stack.project_columns_with_batch_correction(K, None, batch, bases=bases)
    assert basis.shape == (470, K) and proj.shape == (N, 2 * K)
    for m, (ip, ix, v, D) in enumerate(mods):
        want = orc.project(ip, ix, v, bases[m], batch[:N].astype(np.uint32), 3, nthreads=4)
        assert close(proj[:, m * K:(m + 1) * K], want, TOL)
    w = rng.uniform(0.2, 1.0, 470).astype(np.float32)
    _, pw = stack.project_columns_weighted(K, None, None, w, bases=bases)
    assert close(pw[:, K:], orc.project(*mods[1][:3], bases[1] * w[300:, None]), TOL)
    with pytest.raises(lg.LegumeError):
        stack.project_columns_weighted(K, None, None, w[:-1], bases=bases)
    ncode = stack.partition_columns_to_groups(proj, 6)
    assert ncode <= 64 and np.array_equal(vecs[0].get_group_membership(), vecs[1].get_group_membership())


def test_project_batch_label_mismatch_skips_centring(lg, ctx):
    rng = np.random.default_rng(4)
    ip, ix, v = random_csc(rng, 200, 300, 0.1)
    basis = basis_for(200, 10)
    data = lg.SparseIoVec.from_csc(ctx, ip, ix, v, 200)
    _, a = data.project_columns_with_batch_correction(10, None, ["x"] * 299, basis=basis)  # wrong length (:389-395)
    _, b = data.project_columns(10, None, basis=basis)
    assert a.tobytes() == b.tobytes()


# ---- stage 2 / 3 -------------------------------------------------------------------------------
def make_proj(rng, n, K, rank=6):
    z = rng.normal(size=(n, rank)) @ rng.normal(size=(rank, K)) + 0.3 * rng.normal(size=(n, K))
    return orc.project_finish(z.astype(np.float32))


@pytest.mark.parametrize("n,K,kk", [(5000, 50, 10), (1024, 50, 10), (1025, 50, 7), (333, 20, 12), (40, 50, 10),
                                    (20000, 50, 14), (16, 8, 8)])
def test_binary_codes_bit_exact(lg, ctx, n, K, kk):
    proj = make_proj(np.random.default_rng(n + kk), n, K)
    got = lg.binary_sort_columns(ctx, proj, kk)
    want = orc.binary_codes(proj, kk)
    assert got.dtype == np.uint64 and np.array_equal(got, want), int((got != want).sum())
    again = lg.binary_sort_columns(ctx, proj, kk)
    assert np.array_equal(got, again)  # rsvd_tests.rs:22-32


def test_binary_codes_staged_factors_match_oracle(lg, ctx):
    """the two small factorisations are CUDA kernels (k_codes_basis, k_codes_factor): their Q, U and sigma against the
    mirror oracle's host loops, bit for bit — part of the parity surface"""
    rng = np.random.default_rng(8)
    for n, K, kk in ((3000, 50, 10), (500, 24, 16), (64, 50, 3)):
        proj = make_proj(rng, n, K)
        _, q, u, sig, mean = orc.binary_codes(proj, kk, details=True)
        r = min(kk + 5, n) if min(K, n) > kk else min(K, n)
        q_got = np.empty((kk, K), np.float32)
        ctx.check(lg.lib.lg_codes_basis(ctx.h, np.ascontiguousarray(proj[:r]).ctypes.data, K, r, kk, q_got.ctypes.data))
        assert q_got.tobytes() == q.tobytes()
        # Gram sums from the device's own B = Q^T X (block partials in f64), then the factor kernel
        import torch
        d_q = torch.from_numpy(q).cuda()
        d_b = torch.empty((n, kk), dtype=torch.float32, device="cuda")
        nblk = (n + 1023) // 1024
        M = kk * (kk + 1) // 2
        d_part = torch.empty((nblk, M), dtype=torch.float64, device="cuda")
        d_sums = torch.empty(M, dtype=torch.float64, device="cuda")
        d_u = torch.empty((kk, kk), dtype=torch.float32, device="cuda")
        d_sig = torch.empty(kk, dtype=torch.float32, device="cuda")
        d_proj = torch.from_numpy(proj).cuda()
        ctx.check(lg.lib.lg_codes_gram(ctx.h, d_proj.data_ptr(), K, n, d_q.data_ptr(), kk, d_b.data_ptr(), d_part.data_ptr()))
        ctx.check(lg.lib.lg_block_partials_finalize(ctx.h, d_part.data_ptr(), nblk, M, d_sums.data_ptr()))
        ctx.check(lg.lib.lg_codes_factor(ctx.h, d_sums.data_ptr(), d_q.data_ptr(), K, kk, d_u.data_ptr(), d_sig.data_ptr()))
        ctx.sync()
        torch.cuda.synchronize()
        assert d_u.cpu().numpy().tobytes() == u.tobytes()
        assert d_sig.cpu().numpy().tobytes() == sig.tobytes()


def test_binary_codes_against_the_independent_svd_oracle(lg, ctx):
    """GPU codes against oracle_svd.cpp — the reference's own route (f32 bidiagonalisation + implicit QR), which shares
    no factorisation code with the product: same partition per bit up to complement, except for cells whose
    standardised coordinate is rounding noise away from zero.  The agreement is printed (DESIGN.md quotes it)."""
    rng = np.random.default_rng(21)
    worst = 1.0
    for n, K, kk in ((50000, 50, 10), (8000, 50, 8), (3000, 32, 12)):
        proj = make_proj(rng, n, K)
        got = lg.binary_sort_columns(ctx, proj, kk)
        want, v, _ = orc.binary_codes_svd(proj, kk, details=True)
        agree = orc.partition_agreement(got, want, kk)
        worst = min(worst, min(agree))
        assert min(agree) > 0.998, agree
        diff = got ^ want
        for k in range(kk):
            bit = (diff >> np.uint64(k)) & np.uint64(1)
            flipped = bit == (1 if np.mean(bit == 0) > 0.5 else 0)
            if flipped.any():
                assert np.abs(v[k][flipped]).max() < 1e-2
    print(f"\nworst per-bit partition agreement, GPU vs independent f32 SVD: {worst:.6f}")


@pytest.mark.parametrize("padded", [False, True])
def test_assign_groups_lexicographic(lg, ctx, padded):
    rng = np.random.default_rng(5)
    codes = rng.integers(0, 1 << 10, 5000).astype(np.uint64)
    codes[codes % 7 == 0] = 33  # leave holes so that rank != code
    got, ng = lg.assign_groups_from_codes(ctx, codes, 10, padded)
    want, wng = orc.assign_groups_padded(codes, 1 << 10) if padded else orc.assign_groups(codes)
    assert ng == wng and np.array_equal(got, want)
    if not padded:  # "10" < "2"
        g = {int(c): int(x) for c, x in zip(codes, got)}
        if 10 in g and 2 in g:
            assert g[10] < g[2]


def test_partition_columns_to_groups(lg, ctx):
    rng = np.random.default_rng(6)
    D, N, K = 600, 2500, 50
    ip, ix, v = random_csc(rng, D, N, 0.06)
    data = lg.SparseIoVec.from_csc(ctx, ip, ix, v, D)
    _, proj = data.project_columns(K, None, basis=basis_for(D, K))
    nmax = data.partition_columns_to_groups(proj, 8)
    codes = orc.binary_codes(proj, 8)
    want, ng = orc.assign_groups(codes)
    assert nmax == int(codes.max()) + 1 and data.num_groups() == ng
    assert np.array_equal(data.get_group_membership(), want)


# ---- stage 4 -----------------------------------------------------------------------------------
@pytest.mark.parametrize("D,N,S", [(1000, 4000, 37), (30000, 600, 5), (70000, 300, 3), (50, 2000, 1024)])
def test_collapse_basic_bit_exact(lg, ctx, D, N, S):
    rng = np.random.default_rng(D + S)
    ip, ix, v = random_csc(rng, D, N, min(0.05, 200.0 / D), empty_every=53)
    grp = rng.integers(0, S + 2, N).astype(np.uint32)  # ids >= S are skipped
    data = lg.SparseIoVec.from_csc(ctx, ip, ix, v, D)
    data.col_to_group = grp
    stat = lg.CollapsedStat(D, S, 1)
    data.collect_basic_stat(stat)
    ws, wsize = orc.collapse_basic(ip, ix, v, D, grp, S)
    assert np.array_equal(stat.observed_sum_ds, ws) and np.array_equal(stat.size_s, wsize)
    assert float(stat.observed_sum_ds.sum(dtype=np.float64)) == float(v[np.repeat(grp < S, np.diff(ip).astype(np.int64))].sum(dtype=np.float64))


def test_collapse_batch_and_multiplicity(lg, ctx):
    rng = np.random.default_rng(12)
    D, N, S, B = 800, 3000, 20, 4
    ip, ix, v = random_csc(rng, D, N, 0.05)
    grp = rng.integers(0, S, N).astype(np.uint32)
    bat = rng.integers(0, B, N)
    data = lg.SparseIoVec.from_csc(ctx, ip, ix, v, D)
    data.col_to_group = grp
    data.register_batch_membership([f"b{b}" for b in bat])
    stat = lg.CollapsedStat(D, S, B)
    data.collect_basic_stat(stat)
    data.collect_batch_stat(stat)
    wdb, wnbs = orc.collapse_batch(ip, ix, v, D, grp, bat.astype(np.uint32), S, B)
    assert np.array_equal(stat.observed_sum_db, wdb) and np.array_equal(stat.n_bs, wnbs)
    # unit multiplicity is bit-for-bit inert (weighted_columns.rs:76-89)
    base = stat.observed_sum_ds.copy()
    data.register_column_multiplicity(np.ones(N, np.float32))
    data.collect_basic_stat(stat)
    assert stat.observed_sum_ds.tobytes() == base.tobytes()
    # fractional weights: order of the float adds is free, so 1e-5
    w = rng.uniform(0.5, 3.0, N).astype(np.float32)
    data.register_column_multiplicity(w)
    data.collect_basic_stat(stat)
    ws, wsz = orc.collapse_basic(ip, ix, v, D, grp, S, mult=w)
    assert close(stat.observed_sum_ds, ws, TOL) and close(stat.size_s, wsz, TOL)
    with pytest.raises(lg.LegumeError):
        data.register_column_multiplicity(np.zeros(N, np.float32))


@pytest.mark.parametrize("D,N,S", [(800, 3000, 20), (70_000, 1500, 7)])
def test_fractional_collapse_is_deterministic(lg, ctx, D, N, S):
    """SURVEY section 8b "Determinism": with fractional multiplicities or non-integer values the sums go through 64-bit fixed
    point (integer adds, any order), so they are bit-identical run to run, the sizes are the reference's own serial fold, and the
    sums sit within 1e-5 of the reference's f32 folds; D = 70 000 exceeds one shared-memory window of the 64-bit accumulators"""
    rng = np.random.default_rng(31)
    ip, ix, v = random_csc(rng, D, N, 0.05 if D < 10_000 else 0.004)
    grp = rng.integers(0, S, N).astype(np.uint32)
    w = rng.uniform(1e-4, 3.0, N).astype(np.float32)
    for values, mult in ((v, w), ((v * np.float32(0.37)).astype(np.float32), None), (-v * np.float32(1.5), w)):
        data = lg.SparseIoVec.from_csc(ctx, ip, ix, values, D)
        data.col_to_group = grp
        if mult is not None:
            data.register_column_multiplicity(mult)
        runs = []
        for _ in range(3):
            stat = lg.CollapsedStat(D, S, 0)
            data.collect_basic_stat(stat)
            runs.append((np.asarray(stat.observed_sum_ds).copy(), np.asarray(stat.size_s).copy()))
        assert all(r[0].tobytes() == runs[0][0].tobytes() and r[1].tobytes() == runs[0][1].tobytes() for r in runs[1:])
        ws, wsz = orc.collapse_basic(ip, ix, values, D, grp, S, mult=mult)
        assert close(runs[0][0], ws, TOL)
        assert runs[0][1].tobytes() == wsz.tobytes()  # the serial fold over the group's cells, as the reference makes it
        # and against exact arithmetic: the fixed-point sums are the better ones
        cols = np.repeat(np.arange(N), np.diff(ip).astype(np.int64))
        exact = np.zeros((S, D))
        np.add.at(exact, (grp[cols], ix.astype(np.int64)), values.astype(np.float64) * (1.0 if mult is None else mult[cols].astype(np.float64)))
        assert np.max(np.abs(runs[0][0] - exact) / (1 + np.abs(exact))) < 1e-6


def test_weighted_columns_identity(lg, ctx):
    """data-beans-alg/tests/weighted_columns.rs:91-121"""
    D, M = 6, 20
    profile = np.array([1.0 + g for g in range(D)], np.float32)

    def one_group(cells, weights=None):
        ip = np.arange(0, (len(cells) + 1) * D, D, dtype=np.uint64)
        ix = np.tile(np.arange(D, dtype=np.uint64), len(cells))
        data = lg.SparseIoVec.from_csc(ctx, ip, ix, np.concatenate(cells), D)
        data.register_batch_membership(["b0"] * len(cells))
        if weights is not None:
            data.register_column_multiplicity(weights)
        data.assign_groups(["g0"] * len(cells))
        out, stat = data.collapse_columns(None, None, None, 1)
        return out.mu_observed["mean"][0]

    many = one_group([profile] * M)
    one = one_group([profile], np.array([M], np.float32))
    assert np.all(np.abs(many - one) < 1e-4)
    assert np.array_equal(many, ((1.0 + M * profile) / np.float32(1.0 + M)).astype(np.float32))


def test_merge_stat(lg, ctx):
    rng = np.random.default_rng(2)
    fine = rng.poisson(3, size=(12, 500)).astype(np.float32)
    f2c = (np.arange(12) % 5).astype(np.uint32)
    got = lg.merge_stat(ctx, fine, f2c, 5)
    assert np.array_equal(got, orc.merge_stat(fine, f2c, 5))


# ---- stage 5 -----------------------------------------------------------------------------------
@pytest.mark.parametrize("target", [0, 1, 2])
def test_gamma_calibrate_and_optimize_single(lg, ctx, target):
    rng = np.random.default_rng(target)
    S, D = 40, 700
    sums = rng.poisson(0.7, size=(S, D)).astype(np.float32) * rng.integers(0, 30, size=(S, D))
    size = rng.integers(1, 2000, S).astype(np.float32)
    stat = lg.CollapsedStat(D, S, 1)
    stat.observed_sum_ds, stat.size_s = sums, size
    out = lg.optimize(ctx, stat, (1.0, 1.0), 10, target).mu_observed
    want = orc.optimize_single(sums, size, 1.0, 1.0, target)
    assert close(out["mean"], want["mean"], TOL), max_err(out["mean"], want["mean"])
    if target != 1:
        assert close(out["log_mean"], want["log_mean"], TOL), max_err(out["log_mean"], want["log_mean"])
    if target == 0:
        assert close(out["sd"], want["sd"], TOL) and close(out["log_sd"], want["log_sd"], TOL)
    if target == 1:
        assert np.all(out["mean"][sums == 0] == 0.0)  # stats_tests.rs:99-130


def test_gamma_matrix_trigamma_identities(lg, ctx):
    """matrix-param/src/dmatrix_gamma_tests.rs:9-60"""
    import math
    for a, want in [(1.0, math.sqrt(math.pi ** 2 / 6)), (2.0, math.sqrt(math.pi ** 2 / 6 - 1)), (0.5, math.sqrt(math.pi ** 2 / 2))]:
        p = lg.GammaMatrix.new(ctx, (1, 1), 0.0, 0.0)
        p.update_stat(np.full((1, 1), a, np.float32), np.full((1, 1), 3.0, np.float32))
        p.calibrate()
        assert abs(p.posterior_log_sd()[0, 0] - want) < 1e-4
    p = lg.GammaMatrix.new(ctx, (2, 1), 1.0, 1.0)
    p.update_stat(np.array([[0.0, 500.0]], np.float32), np.array([[0.0, 500.0]], np.float32))
    p.calibrate()
    sd = p.posterior_log_sd()[0]
    assert sd[0] > 1.0 and sd[1] < 0.1 and sd[0] > 10 * sd[1]
    xs = np.concatenate([np.linspace(0.01, 30, 3000), [1e-6, 1e-5, 1e-4, 1e3, 1e5]]).astype(np.float32)
    g = lg.GammaMatrix.new(ctx, (len(xs), 1), 0.0, 0.0)
    g.update_stat(xs[None, :], np.ones((1, len(xs)), np.float32))
    g.calibrate()
    want = orc.gamma_calibrate(xs, np.ones_like(xs), 0.0, 0.0)
    assert close(g.posterior_log_mean()[0], want["log_mean"], TOL) and close(g.posterior_log_sd()[0], want["log_sd"], TOL)


@pytest.mark.parametrize("target,iters", [(0, 25), (1, 10), (2, 30)])
def test_optimize_batched_matches_oracle(lg, ctx, target, iters):
    """stats_tests.rs toy_stat + blocked == whole"""
    G, S, B = 10, 4, 2
    obs, imp, res, size, obs_db, n_bs = toy_stat(G, S, B)
    stat = lg.CollapsedStat(G, S, B)
    stat.observed_sum_ds, stat.imputed_sum_ds, stat.residual_sum_ds = obs, imp, res
    stat.size_s, stat.observed_sum_db, stat.n_bs = size, obs_db, n_bs
    out = lg.optimize(ctx, stat, (1.0, 1.0), iters, target)
    want = orc.optimize_batched(obs, imp, res, size, obs_db, n_bs, 1.0, 1.0, iters, target)
    for key, ok in [("mu_observed", "mu_observed"), ("mu_adjusted", "mu_adjusted"), ("mu_residual", "mu_residual"),
                    ("gamma", "gamma"), ("delta", "delta")]:
        assert close(out[key]["mean"], want[ok], TOL), (key, max_err(out[key]["mean"], want[ok]))
    if target != 1:
        assert close(out.mu_adjusted["log_mean"], want["mu_adjusted_log_mean"], TOL)
    parts = [lg.optimize(ctx, stat.select_rows(r0, nr), (1.0, 1.0), iters, target) for r0, nr in [(0, 3), (3, 4), (7, 3)]]
    for key in ["mu_observed", "mu_adjusted", "mu_residual", "gamma", "delta"]:
        blk = np.concatenate([p[key]["mean"] for p in parts], axis=1)
        assert close(out[key]["mean"], blk, TOL), key


# ---- stage 6 -----------------------------------------------------------------------------------
def random_points(n, d, seed):
    return np.random.default_rng(seed).uniform(-1, 1, size=(n, d)).astype(np.float32)


@pytest.mark.parametrize("nr,nq,d,k", [(500, 500, 16, 8), (2000, 300, 50, 10), (9000, 64, 50, 11), (37, 10, 5, 50),
                                       (1000, 100, 32, 41)])
def test_knn_index_sets_bit_exact(lg, ctx, nr, nq, d, k):
    ref = random_points(nr, d, nr + d)
    qry = ref[:nq] if nq <= nr else random_points(nq, d, 1)
    dct = lg.ColumnDict.from_dmatrix(ctx, ref, list(range(nr)))
    ex = np.arange(nq, dtype=np.uint32)
    idx, dist = dct.search_indices(qry, k, ex)
    widx, wdist = orc.knn_topk(ref, qry, k, ex, nthreads=4)
    assert np.array_equal(idx, widx) and dist.tobytes() == wdist.tobytes()
    idx2, dist2 = dct.search_indices(qry, k)
    widx2, wdist2 = orc.knn_topk(ref, qry, k, nthreads=4)
    assert np.array_equal(idx2, widx2) and dist2.tobytes() == wdist2.tobytes()


def test_knn_ties_and_api(lg, ctx):
    """duplicate points: equal distances ordered by lower index; reference API behaviour (knn/tests.rs)"""
    pts = random_points(300, 12, 4)
    pts[50] = pts[10]
    pts[200] = pts[10]
    a = lg.ColumnDict.from_dmatrix(ctx, pts, list(range(300)))
    names, d = a.search_others(10, 5)
    assert names[:2] == [50, 200] and d[0] == 0.0 and d[1] == 0.0 and 10 not in names
    widx, _ = orc.knn_topk(pts, pts[10:11], 5, np.array([10], np.uint32))
    assert names == [int(i) for i in widx[0]]
    b = lg.ColumnDict.from_dmatrix(ctx, random_points(400, 12, 5), [f"n{i}" for i in range(400)])
    got, dd = a.match_by_query_name_against(7, 5, b)
    widx, wd = orc.knn_topk(b.data, pts[7:8], 5)
    assert got == [f"n{i}" for i in widx[0]] and dd == [float(x) for x in wd[0]]
    q = np.array([t * 0.05 - 0.4 for t in range(12)], np.float32)
    got, dd = a.search_by_query_data(q, 6)
    assert got == [int(i) for i in orc.knn_topk(pts, q[None], 6)[0][0]] and all(x <= y for x, y in zip(dd, dd[1:]))
    with pytest.raises(lg.LegumeError):
        a.search_by_query_data(q[:5], 3)
    with pytest.raises(lg.LegumeError):
        a.search_others("missing", 3)
    few = lg.ColumnDict.from_dmatrix(ctx, pts[:4], list(range(4)))
    names, _ = few.search_others(0, 10)  # fewer than k+1 points: returns what exists
    assert sorted(names) == [1, 2, 3]


# ---- whole path --------------------------------------------------------------------------------
def test_hot_path_device_pipeline_matches_oracle(lg, ctx):
    """sim -> project -> codes -> groups -> collapse -> posterior, all device-resident, one GPU"""
    import torch
    from legume_b200 import sim
    from legume_b200.pipeline import HotPath
    D, N, K, kk = 3000, 6000, 50, 8
    tabs = sim.make_tables(D, ntopic=6, nbatch=1, depth=300, seed=5)
    blk, topic, batch = sim.sim_block(ctx, tabs, 0, N)
    ip, ix, v = blk.download()
    basis = basis_for(D, K)
    hp = HotPath(ctx)
    out = hp.run(blk, torch.from_numpy(basis).cuda(), torch.zeros(N, dtype=torch.int32, device="cuda"), 1, kk)
    torch.cuda.synchronize()
    proj = out["proj"].cpu().numpy()
    want_proj = orc.project(ip, ix, v, basis, np.zeros(N, np.uint32), 1, nthreads=4)
    assert close(proj, want_proj, TOL), max_err(proj, want_proj)
    # downstream stages are checked on the GPU's own projection (identical stage inputs)
    codes = out["codes"].cpu().numpy().astype(np.uint64)
    assert np.array_equal(codes, orc.binary_codes(proj, kk))
    want_grp, ng = orc.assign_groups(codes)
    grp = out["group"].cpu().numpy().astype(np.uint32)
    assert out["num_groups"] == ng and np.array_equal(grp, want_grp)
    ws, wsize = orc.collapse_basic(ip, ix, v, D, grp, ng)
    assert np.array_equal(out["sum_ds"].cpu().numpy(), ws) and np.array_equal(out["size_s"].cpu().numpy(), wsize)
    post = orc.optimize_single(ws, wsize, 1.0, 1.0, 0)
    for key in ["mean", "sd", "log_mean", "log_sd"]:
        assert close(out["posterior"][key].cpu().numpy(), post[key], TOL), key
    # size-independent properties: every cell lands in exactly one group, counts are conserved
    assert float(out["size_s"].sum().item()) == N
    assert float(out["sum_ds"].double().sum().item()) == float(v.astype(np.float64).sum())
    assert ctx.launch_count > 0


def test_composite_and_staged_paths_agree(lg, ctx):
    import torch
    from legume_b200.pipeline import HotPath
    rng = np.random.default_rng(21)
    D, N, K, kk = 900, 2600, 50, 9
    ip, ix, v = random_csc(rng, D, N, 0.05)
    basis = basis_for(D, K)
    batch = rng.integers(0, 3, N).astype(np.uint32)
    data = lg.SparseIoVec.from_csc(ctx, ip, ix, v, D)
    _, proj = data.project_columns_with_batch_correction(K, None, batch, basis=basis)
    hp = HotPath(ctx)
    p2 = hp.project(data.block, torch.from_numpy(basis).cuda(), torch.from_numpy(batch.astype(np.int32)).cuda(), 3)
    assert proj.tobytes() == p2.cpu().numpy().tobytes()
    c1 = lg.binary_sort_columns(ctx, proj, kk)
    c2 = hp.binary_codes(p2, kk).cpu().numpy().astype(np.uint64)
    assert np.array_equal(c1, c2)


# ---- stage 6 on the tensor path (lg_knn_umma.cu): large enough to take the tcgen05 filter + exact refine ----
@pytest.mark.parametrize("nr,nq,d,k,kind", [(20000, 3000, 50, 10, "gauss"), (9000, 6000, 32, 5, "uniform"),
                                            (16384, 4096, 50, 12, "clustered"), (12000, 5000, 17, 10, "dupes")])
def test_knn_tensor_path_is_exact(lg, ctx, nr, nq, d, k, kind):
    rng = np.random.default_rng(nr + d)
    if kind == "gauss":
        ref = orc.project_finish(rng.normal(size=(nr, d)).astype(np.float32))
    elif kind == "uniform":
        ref = rng.uniform(-1, 1, size=(nr, d)).astype(np.float32)
    elif kind == "clustered":  # near-degenerate gaps: exercises the verified fallback
        centres = rng.normal(size=(40, d))
        ref = (centres[rng.integers(0, 40, nr)] + 1e-4 * rng.normal(size=(nr, d))).astype(np.float32)
    else:  # exact duplicates: ties must resolve by lower index
        ref = rng.normal(size=(nr, d)).astype(np.float32)
        ref[nr // 2:] = ref[: nr - nr // 2]
    qry = ref[:nq].copy()
    ex = np.arange(nq, dtype=np.uint32)
    dct = lg.ColumnDict.from_dmatrix(ctx, ref, list(range(nr)))
    idx, dist = dct.search_indices(qry, k, ex)
    widx, wdist = orc.knn_topk(ref, qry, k, ex, nthreads=8)
    assert np.array_equal(idx, widx), int((idx != widx).sum())
    assert dist.tobytes() == wdist.tobytes()
    qry2 = (qry + 0.01 * rng.normal(size=qry.shape)).astype(np.float32)
    idx, dist = dct.search_indices(qry2, k)
    widx, wdist = orc.knn_topk(ref, qry2, k, nthreads=8)
    assert np.array_equal(idx, widx) and dist.tobytes() == wdist.tobytes()


def test_knn_tensor_path_wide_dynamic_range(lg, ctx):
    """one huge outlier sets the global f16 scale; a tight cluster near the origin then lives in the subnormal range of
    the filter's "lo" halves, where its error is absolute, not relative: the acceptance test must still either prove
    the answer or hand the query to the brute-force kernel.  Checked against the oracle (index sets and bytes)."""
    rng = np.random.default_rng(99)
    nr, nq, d, k = 20000, 4000, 50, 10
    ref = (1e-4 * rng.normal(size=(nr, d))).astype(np.float32)       # tight cluster, norms ~ 7e-4
    ref[nr - 1] = 50.0                                                 # the outlier: absmax 50 -> scale 2^-3
    ref[nr // 2: nr // 2 + 500] *= 30.0                                # a shell further out, so that gaps differ in scale
    qry = ref[:nq].copy()
    ex = np.arange(nq, dtype=np.uint32)
    dct = lg.ColumnDict.from_dmatrix(ctx, ref, list(range(nr)))
    idx, dist = dct.search_indices(qry, k, ex)
    widx, wdist = orc.knn_topk(ref, qry, k, ex, nthreads=8)
    assert np.array_equal(idx, widx), int((idx != widx).sum())
    assert dist.tobytes() == wdist.tobytes()


# ---- BASELINE-sized inputs: size-independent properties (the oracle does not finish at these sizes) ---------------
def test_full_size_properties(lg, ctx):
    """200k cells x 30k genes of the configs[1] generator (a fifth of it, same per-cell shape): conservation of
    counts through the collapse, per-cell standardisation of the projection, code range, group ids dense and
    ordered, posterior mean == (1 + sum) / (1 + n), kNN lists ascending with the query itself excluded"""
    import torch
    from legume_b200 import sim
    from legume_b200.pipeline import HotPath
    D, N, K, kk = 30000, 200_000, 50, 10
    tabs = sim.make_tables(D, ntopic=8, nbatch=1, depth=1500, seed=42)
    blk, _, _ = sim.sim_block(ctx, tabs, 0, N)
    basis = torch.from_numpy(basis_for(D, K)).cuda()
    batch = torch.zeros(N, dtype=torch.int32, device="cuda")
    hp = HotPath(ctx)
    out = hp.run(blk, basis, batch, 1, kk)
    proj, codes, group = out["proj"], out["codes"], out["group"]
    ip, ix, v = blk.download()
    # projection: every cell standardised (population variance), nothing outside the clamp after re-scaling by much
    assert float(proj.mean(1).abs().max()) < 1e-5 and float((proj.var(1, unbiased=False) - 1).abs().max()) < 1e-4
    # codes and groups
    assert int(codes.min()) >= 0 and int(codes.max()) < (1 << kk)
    ng = out["num_groups"]
    assert int(group.max()) == ng - 1 and len(torch.unique(group)) == ng
    keys = sorted({str(int(c)) for c in torch.unique(codes).cpu().numpy()}, key=lambda s: s.encode())
    lut = {int(k): i for i, k in enumerate(keys)}
    sample = np.random.default_rng(0).integers(0, N, 2000)
    assert all(lut[int(codes[j])] == int(group[j]) for j in sample)
    # collapse: integer counts are conserved exactly, gene by gene and group by group
    sum_ds, size_s = out["sum_ds"], out["size_s"]
    assert float(size_s.sum()) == N
    gene_tot = np.bincount(ix.astype(np.int64), weights=v.astype(np.float64), minlength=D)
    assert np.array_equal(sum_ds.sum(0).double().cpu().numpy(), gene_tot)
    cell_tot = np.add.reduceat(v.astype(np.float64), ip[:-1].astype(np.int64))
    cell_tot[np.diff(ip.astype(np.int64)) == 0] = 0.0
    grp_tot = np.bincount(group.cpu().numpy(), weights=cell_tot, minlength=ng)
    assert np.array_equal(sum_ds.sum(1).double().cpu().numpy(), grp_tot)
    # posterior (a0, b0) = (1, 1): mean = (1 + sum) / (1 + n)   (weighted_columns.rs:76-121)
    want = (1.0 + sum_ds) / (1.0 + size_s[:, None])
    assert close(out["posterior"]["mean"].cpu().numpy(), want.cpu().numpy(), TOL)
    # kNN on the tensor path at a size the oracle cannot check: ascending, self excluded, exact distances
    q = proj[:20000].contiguous()
    excl = torch.arange(20000, dtype=torch.int32, device="cuda")
    idx = torch.empty((20000, 10), dtype=torch.int32, device="cuda")
    dist = torch.empty((20000, 10), dtype=torch.float32, device="cuda")
    ctx.check(lg.lib.lg_knn_topk(ctx.h, proj.data_ptr(), N, q.data_ptr(), 20000, K, 10, excl.data_ptr(), idx.data_ptr(), dist.data_ptr()))
    assert bool((dist[:, 1:] >= dist[:, :-1]).all()) and not bool((idx == excl[:, None]).any())
    pick = np.random.default_rng(1).integers(0, 20000, 50)
    P = proj.cpu().numpy()
    for qi in pick:
        d2 = np.array([orc.l2_sq(P[int(j)], P[qi]) for j in idx[qi].cpu().numpy()], np.float32)
        assert np.sqrt(d2).tobytes() == dist[qi].cpu().numpy().tobytes()
        # nothing closer was missed among a random sample of the other cells
        others = np.random.default_rng(int(qi)).integers(0, N, 300)
        far = np.array([orc.l2_sq(P[int(j)], P[qi]) for j in others if j != qi and j not in idx[qi].cpu().numpy()], np.float32)
        assert far.min() >= d2.max()
    blk.free()
