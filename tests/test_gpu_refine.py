"""GPU parity tests for the BBKNN + DC-Poisson refinement of the pb-sample partition (SURVEY.md section 8f rank 3;
refine_multilevel.rs:170-298, dc_poisson.rs:128-915): Fisher weights, weighted profiles, one level of Jacobi sweeps and the
top-down driver against oracle/oracle_refine.cpp — labels and move counts equal, weights and size factors bit for bit.
Run on the B200 box: pytest -m gpu."""
import ctypes as C

import numpy as np
import pytest

import oracle as orc
from test_gpu_adjust import _oracle_refined_levels, close, make_case

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


def planted(n, m, nblocks, seed, noise=0.35):
    """count-like pb-sample profiles with block structure: entity e prefers the genes of block e % nblocks"""
    rng = np.random.default_rng(seed)
    truth = (np.arange(n) % nblocks).astype(np.uint32)
    per = m // nblocks
    lam = np.full((n, m), noise)
    for e in range(n):
        lam[e, truth[e] * per:(truth[e] + 1) * per] += 4.0
    P = rng.poisson(lam * rng.gamma(2.0, 1.0, (n, 1))).astype(np.float32)
    return P, truth


def csr(sets):
    ptr = np.zeros(len(sets) + 1, np.uint32)
    ptr[1:] = np.cumsum([len(s) for s in sets])
    return ptr, np.fromiter((g for s in sets for g in s), np.uint32, int(ptr[-1]))


# ---- host bookkeeping of the mirror against the oracle (no kernel involved, but the package needs the library) ----
def test_label_bookkeeping_matches_oracle(lg):
    rng = np.random.default_rng(0)
    for n in (1, 7, 300):
        child, parent = rng.integers(0, 9, n), rng.integers(0, 4, n)
        a, ka = lg.project_to_refinement(child, parent)
        b, kb = orc.project_to_refinement(child, parent)
        assert ka == kb and np.array_equal(a, b)
        assert np.array_equal(lg.child_offset_within_parent(child, parent), orc.child_offset_within_parent(child, parent))
        fine, k = lg.compact_labels(child * 4 + parent)  # a strict refinement of `parent`
        coarse, _ = lg.compact_labels(parent)
        assert lg.compute_sibling_sets([fine, coarse], 0, k) == orc.sibling_sets(fine, coarse, k)
        assert lg.compute_sibling_sets([fine, coarse], 1, 4) == orc.sibling_sets(coarse, None, 4)
        bb = [rng.choice(n, min(n, 5), replace=False).tolist() for _ in range(n)]
        sib = lg.compute_sibling_sets([fine, coarse], 0, k)
        assert lg.build_candidate_sets(sib, bb, fine) == orc.candidate_sets(sib, bb, fine)
    codes = rng.integers(0, 1 << 9, 50).astype(np.uint64)
    off = lg.build_reproject_offsets(codes, np.arange(50), [9, 6, 3])
    assert np.array_equal(off[0], (codes >> np.uint64(6)) & np.uint64(7)) and np.array_equal(off[1], (codes >> np.uint64(3)) & np.uint64(7))
    assert len(off[2]) == 0  # refine_tests.rs: the coarsest level has no parent


@pytest.mark.gpu
@pytest.mark.parametrize("n,m", [(300, 96), (257, 131), (40, 4000)])
def test_fisher_weights_and_profiles_bit_exact(lg, ctx, n, m):
    from legume_b200._lib import lib
    P, _ = planted(n, m, 4, 1)
    P[:, : m // 8] *= np.random.default_rng(2).gamma(0.4, 2.5, (n, 1)).astype(np.float32).round()  # over-dispersed genes
    w = np.empty(m, np.float32)
    ctx.check(lib.lg_dcp_fisher_weights(ctx.h, P.ctypes.data, m, n, w.ctypes.data))
    want = orc.dcp_fisher_weights(P)
    assert w.tobytes() == want.tobytes() and (want < 1).any()
    got, sf = P.copy(), np.empty(n, np.float32)
    ctx.check(lib.lg_dcp_profiles(ctx.h, got.ctypes.data, m, n, w.ctypes.data, sf.ctypes.data))
    wp, wsf = orc.dcp_profiles(P, want)
    assert got.tobytes() == wp.tobytes() and sf.tobytes() == wsf.tobytes()
    got2, sf2 = P.copy(), np.empty(n, np.float32)
    ctx.check(lib.lg_dcp_profiles(ctx.h, got2.ctypes.data, m, n, None, sf2.ctypes.data))
    wp2, wsf2 = orc.dcp_profiles(P, None)
    assert got2.tobytes() == wp2.tobytes() and sf2.tobytes() == wsf2.tobytes()


@pytest.mark.gpu
@pytest.mark.parametrize("n,m,k,gibbs,greedy,seed", [(256, 96, 4, 0, 10, 1), (256, 96, 4, 6, 4, 12345), (500, 131, 8, 20, 10, 0x9E3779B97F4A7C15),
                                                      (64, 4000, 4, 3, 3, 7), (1200, 400, 16, 5, 5, 99)])
def test_refine_level_matches_oracle(lg, ctx, n, m, k, gibbs, greedy, seed):
    from legume_b200._lib import lib
    P, truth = planted(n, m, k, 3)
    P, sf = orc.dcp_profiles(P, orc.dcp_fisher_weights(P))
    rng = np.random.default_rng(4)
    start = truth.copy()
    wrong = rng.choice(n, n // 3, replace=False)
    start[wrong] = rng.integers(0, k, len(wrong))
    # candidate sets of mixed size: everything, two groups, a single group (never moves)
    cand = []
    for e in range(n):
        r = e % 3
        cand.append(list(range(k)) if r == 0 else sorted({int(start[e]), int((start[e] + 1 + e) % k)}) if r == 1 else [int(start[e])])
    cp, cf = csr(cand)
    labels = start.copy()
    moves = C.c_uint64(0)
    ctx.check(lib.lg_dcp_refine_level(ctx.h, P.ctypes.data, sf.ctypes.data, m, n, cp.ctypes.data, cf.ctypes.data, k, gibbs, greedy,
                                      seed | 1, 0.005, labels.ctypes.data, C.byref(moves)))
    want, wmoves = orc.dcp_refine_level(P, cand, k, start, gibbs, greedy, seed | 1, 0.005)
    assert np.array_equal(labels, want) and moves.value == wmoves and wmoves > 0
    single = np.array([len(c) == 1 for c in cand])
    assert np.array_equal(labels[single], start[single])


@pytest.mark.gpu
def test_refine_level_edge_cases(lg, ctx):
    from legume_b200._lib import lib
    P, truth = planted(50, 40, 4, 5)
    P, sf = orc.dcp_profiles(P, None)
    cand = [[int(t)] for t in truth]
    cp, cf = csr(cand)
    labels = truth.copy()
    moves = C.c_uint64(7)
    ctx.check(lib.lg_dcp_refine_level(ctx.h, P.ctypes.data, sf.ctypes.data, 40, 50, cp.ctypes.data, cf.ctypes.data, 4, 5, 5, 1, 0.005,
                                      labels.ctypes.data, C.byref(moves)))
    assert moves.value == 0 and np.array_equal(labels, truth)
    bad = truth.copy()
    bad[3] = 9
    rc = lib.lg_dcp_refine_level(ctx.h, P.ctypes.data, sf.ctypes.data, 40, 50, cp.ctypes.data, cf.ctypes.data, 4, 5, 5, 1, 0.005,
                                 bad.ctypes.data, C.byref(moves))
    assert rc != 0 and b"label out of range" in lib.lg_last_error(ctx.h)
    with pytest.raises(lg.LegumeError):
        lg.RefineParams(parallel=False)
    with pytest.raises(lg.LegumeError):
        lg.RefineParams(profile_source="Spatial")


@pytest.mark.gpu
@pytest.mark.parametrize("fisher,gibbs", [(False, 0), (True, 20)])
def test_refine_assignments_matches_oracle(lg, ctx, fisher, gibbs):
    n, m = 600, 240
    P, truth = planted(n, m, 8, 6)
    rng = np.random.default_rng(7)
    codes = (truth.astype(np.uint64) ^ (rng.random(n) < 0.25).astype(np.uint64))  # a quarter starts in the sibling group
    first = np.arange(n)
    dims = [3, 2, 1]
    init = lg.initial_per_level_from_hash(codes, first, dims)
    off = lg.build_reproject_offsets(codes, first, dims)
    bbknn = [rng.choice(n, 8, replace=False).tolist() for _ in range(n)]
    params = lg.RefineParams(num_gibbs=gibbs, num_greedy=10, feature_weighting="FisherInfoNb" if fisher else "None", seed=42)
    levels, ks, moves = lg.refine_assignments(ctx, P, bbknn, init, off, params)
    wl, wk, wm = orc.refine_assignments(P, bbknn, init, off, num_gibbs=gibbs, num_greedy=10, fisher=fisher, seed=42)
    assert ks == wk and moves == wm and moves > 0
    for a, b in zip(levels, wl):
        assert np.array_equal(a, b)
    for lvl in range(2):  # the hierarchy survives: every group of a level has one parent
        for g in range(ks[lvl]):
            assert len(set(levels[lvl + 1][levels[lvl] == g].tolist())) == 1
    # without reprojection offsets the positional offset takes over on both sides (refine_multilevel.rs:266-272)
    l2, k2, m2 = lg.refine_assignments(ctx, P, bbknn, init, None, params)
    w2, wk2, wm2 = orc.refine_assignments(P, bbknn, init, None, num_gibbs=gibbs, num_greedy=10, fisher=fisher, seed=42)
    assert k2 == wk2 and m2 == wm2 and all(np.array_equal(a, b) for a, b in zip(l2, w2))


@pytest.mark.gpu
def test_multilevel_collapse_with_refinement_over_batches(lg, ctx):
    """collapse_columns_multilevel_vec with MultilevelParams::new's defaults (refine = Some) and three batches: the whole arm
    of refine_and_collect_single_layer (refine.rs:264-500) against the oracle composed stage by stage"""
    D, N, B, K = 400, 4000, 3, 20
    ip, ix, v, proj, batch, _ = make_case(31, D, N, B, 4, K, clustered=True)
    data = lg.SparseIoVec.from_csc(ctx, ip, ix, v, D)
    params = lg.MultilevelParams(K, knn_pb_samples=4, num_levels=2, sort_dim=8, num_opt_iter=12, observe_panels=False)
    assert params.refine is not None and params.refine.num_gibbs == 20 and params.refine.feature_weighting == "FisherInfoNb"
    out = data.collapse_columns_multilevel_with_hierarchy(proj, batch, params)
    dims = orc.level_sort_dims(8, 2)
    codes = orc.binary_codes(proj, dims[0])
    grp, S = orc.assign_groups(codes)
    lay = orc.pb_layout(proj, grp, S, batch, B)
    cells = orc.pb_sample_to_cells(lay["cell_to_pb"], lay["num_pb"])
    first = np.array([c[0] for c in cells])
    init = orc.initial_per_level_from_hash(codes, cells, dims)
    gs, _ = orc.collapse_basic(ip, ix, v, D, lay["cell_to_pb"], lay["num_pb"])
    mp, _ = orc.pb_match(proj, batch, B, lay, 4)
    bbknn = [row[row != 0xFFFFFFFF].tolist() for row in mp]
    off = lg.build_reproject_offsets(codes, first, dims)
    p2g, k, moves = orc.refine_assignments(gs, bbknn, init, off, seed=42)
    assert data.refine_moves == moves
    want, fine = _oracle_refined_levels(ip, ix, v, D, proj, batch, B, codes, dims, p2g, k, lay, 4, 12)
    assert np.array_equal(np.asarray(data.col_to_group), fine)
    for level in range(2):
        st, w = out["stats"][level], want[level]
        assert np.array_equal(st.observed_sum_ds, w["obs"]) and np.array_equal(st.size_s, w["size"])
        assert close(st.imputed_sum_ds, w["imp"], TOL) and close(st.residual_sum_ds, w["res"], TOL)
        assert close(out["levels"][level].mu_adjusted["mean"], w["fit"]["mu_adjusted"], 1e-4)
        assert np.array_equal(out["cell_to_pb_per_level"][level], p2g[level][lay["cell_to_pb"].astype(np.int64)])


@pytest.mark.gpu
def test_refine_level_tiled_and_untiled_forms_agree(lg, ctx, monkeypatch):
    """the shared-memory form of the score kernel keeps every pair's order of additions: same labels as the one-thread-per-pair
    form (LG_DCP_UNTILED=1) and as the oracle; a level whose candidate sets exceed 128 groups takes the untiled form by itself"""
    from legume_b200._lib import lib
    n, m = 700, 517  # a feature count that is neither a multiple of 4 nor of the 128-feature tile
    for k, full in ((24, False), (160, True)):
        P, truth = planted(n, m, min(k, 8), 11)
        P, sf = orc.dcp_profiles(P, None)
        rng = np.random.default_rng(12)
        start = rng.integers(0, k, n).astype(np.uint32)
        cand = [list(range(k)) if full or e % 2 else sorted({int(start[e]), int((start[e] + 3) % k), int((start[e] + 7) % k)}) for e in range(n)]
        cp, cf = csr(cand)
        got = {}
        for form in ("tiled", "untiled"):
            if form == "untiled":
                monkeypatch.setenv("LG_DCP_UNTILED", "1")
            else:
                monkeypatch.delenv("LG_DCP_UNTILED", raising=False)
            labels, moves = start.copy(), C.c_uint64(0)
            ctx.check(lib.lg_dcp_refine_level(ctx.h, P.ctypes.data, sf.ctypes.data, m, n, cp.ctypes.data, cf.ctypes.data, k, 4, 6, 77, 0.0,
                                              labels.ctypes.data, C.byref(moves)))
            got[form] = (labels, moves.value)
        want, wmoves = orc.dcp_refine_level(P, cand, k, start, 4, 6, 77, 0.0)
        for form in got:
            assert np.array_equal(got[form][0], want) and got[form][1] == wmoves


def test_vectorised_candidate_sets_equal_the_list_form(lg):
    from legume_b200 import _candidate_csr
    rng = np.random.default_rng(21)
    for n, kf, kc, T in ((1, 1, 1, 3), (50, 12, 3, 6), (400, 64, 8, 14)):
        coarse = rng.integers(0, kc, n)
        fine, k = lg.compact_labels(rng.integers(0, max(kf // kc, 1), n) * kc + coarse)  # a strict refinement of `coarse`
        coarse, kcc = lg.compact_labels(coarse)
        bb = rng.integers(0, n, (n, T)).astype(np.uint32)
        bb[rng.random((n, T)) < 0.3] = 0xFFFFFFFF
        bb[0] = 0xFFFFFFFF  # an entity without any neighbour falls back to its siblings
        lists = [row[row != 0xFFFFFFFF].tolist() for row in bb]
        for level, kk_ in ((0, k), (1, kcc)):
            refined = [fine, coarse]
            want = lg.build_candidate_sets(lg.compute_sibling_sets(refined, level, kk_), lists, refined[level])
            ptr, flat = _candidate_csr(refined, level, kk_, bb)
            got = [flat[ptr[e]:ptr[e + 1]].tolist() for e in range(n)]
            assert got == want
            assert got == orc.candidate_sets(orc.sibling_sets(refined[level], refined[1] if level == 0 else None, kk_), lists, refined[level])


@pytest.mark.gpu
@pytest.mark.parametrize("refine", [True, False])
def test_stack_multilevel_collapse_first_layer_owns_the_partition(lg, ctx, refine):
    """MultilevelCollapsingOps for SparseIoStack (collapse_data/mod.rs:1050-1260, refine.rs:503-716): the first modality's gene
    sums drive the refinement, every modality is then collapsed on that one partition"""
    N, B, K = 3000, 3, 20
    ip0, ix0, v0, proj, batch, _ = make_case(41, 350, N, B, 4, K, clustered=True)
    ip1, ix1, v1, _, _, _ = make_case(42, 220, N, B, 4, K)
    layers = [(ip0, ix0, v0, 350), (ip1, ix1, v1, 220)]
    stack = lg.SparseIoStack([lg.SparseIoVec.from_csc(ctx, ip, ix, v, D) for ip, ix, v, D in layers])
    params = lg.MultilevelParams(K, knn_pb_samples=4, num_levels=2, sort_dim=8, num_opt_iter=12, refine="default" if refine else None)
    outs, stats = stack.collapse_columns_multilevel_vec(proj, batch, params)
    dims = orc.level_sort_dims(8, 2)
    assert len(outs) == len(dims) and all(len(lv) == 2 for lv in outs)
    codes = orc.binary_codes(proj, dims[0])
    grp, S = orc.assign_groups(codes)
    lay = orc.pb_layout(proj, grp, S, batch, B)
    if refine:
        cells = orc.pb_sample_to_cells(lay["cell_to_pb"], lay["num_pb"])
        first = np.array([c[0] for c in cells])
        init = orc.initial_per_level_from_hash(codes, cells, dims)
        gs0, _ = orc.collapse_basic(ip0, ix0, v0, 350, lay["cell_to_pb"], lay["num_pb"])
        mp, _ = orc.pb_match(proj, batch, B, lay, 4)
        bbknn = [row[row != 0xFFFFFFFF].tolist() for row in mp]
        p2g, k, _ = orc.refine_assignments(gs0, bbknn, init, lg.build_reproject_offsets(codes, first, dims), seed=42)
        for d, (ip, ix, v, D) in enumerate(layers):
            want, fine = _oracle_refined_levels(ip, ix, v, D, proj, batch, B, codes, dims, p2g, k, lay, 4, 12)
            assert np.array_equal(np.asarray(stack.stack[d].col_to_group), fine)
            for level in range(2):
                st, w = stats[level][d], want[level]
                assert np.array_equal(st.observed_sum_ds, w["obs"]) and np.array_equal(st.size_s, w["size"])
                assert close(st.imputed_sum_ds, w["imp"], TOL) and close(st.residual_sum_ds, w["res"], TOL)
                assert close(outs[level][d].mu_adjusted["mean"], w["fit"]["mu_adjusted"], 1e-4)
    else:
        for d, (ip, ix, v, D) in enumerate(layers):
            assert np.array_equal(np.asarray(stack.stack[d].col_to_group), grp)
            obs, size = orc.collapse_basic(ip, ix, v, D, grp, S)
            assert np.array_equal(stats[0][d].observed_sum_ds, obs) and np.array_equal(stats[0][d].size_s, size)
    finest = stack.collapse_columns_multilevel(proj, batch, params)
    assert len(finest) == 2


@pytest.mark.gpu
def test_multilevel_collapse_with_projected_profiles(lg, ctx):
    """RefineParams.profile_source = Projected (dc_poisson.rs:39-49, 164-195): the refinement scores on the pb-samples' summed
    projection columns (entries > 0 only, no feature weighting)"""
    D, N, B, K = 300, 3000, 3, 16
    ip, ix, v, proj, batch, _ = make_case(51, D, N, B, 4, K, clustered=True)
    data = lg.SparseIoVec.from_csc(ctx, ip, ix, v, D)
    rp = lg.RefineParams(num_gibbs=5, num_greedy=5, profile_source="Projected", seed=7)
    params = lg.MultilevelParams(K, knn_pb_samples=4, num_levels=2, sort_dim=8, num_opt_iter=12, refine=rp, observe_panels=False)
    out = data.collapse_columns_multilevel_with_hierarchy(proj, batch, params)
    dims = orc.level_sort_dims(8, 2)
    codes = orc.binary_codes(proj, dims[0])
    grp, S = orc.assign_groups(codes)
    lay = orc.pb_layout(proj, grp, S, batch, B)
    cells = orc.pb_sample_to_cells(lay["cell_to_pb"], lay["num_pb"])
    first = np.array([c[0] for c in cells])
    prof = np.zeros((lay["num_pb"], K), np.float32)
    np.add.at(prof, lay["cell_to_pb"].astype(np.int64), proj)  # unbuffered: one f32 add per cell, ascending — the reference's fold
    mp, _ = orc.pb_match(proj, batch, B, lay, 4)
    bbknn = [row[row != 0xFFFFFFFF].tolist() for row in mp]
    init = orc.initial_per_level_from_hash(codes, cells, dims)
    p2g, k, moves = orc.refine_assignments(prof, bbknn, init, lg.build_reproject_offsets(codes, first, dims), num_gibbs=5, num_greedy=5,
                                           fisher=False, seed=7)
    assert data.refine_moves == moves
    for level in range(2):
        assert np.array_equal(out["cell_to_pb_per_level"][level], p2g[level][lay["cell_to_pb"].astype(np.int64)])
