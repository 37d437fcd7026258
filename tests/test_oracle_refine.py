"""The refinement oracle (oracle/oracle_refine.cpp) pinned on the reference's own tests for the BBKNN + DC-Poisson
refinement: data-beans-alg/src/dc_poisson_tests.rs and refine_multilevel_tests.rs.  CPU only."""
import numpy as np

import oracle as orc


def toy_profiles(n, m, labels, seed):
    """the shape of dc_poisson_tests.rs:4-31: a block of strong features per label plus three weak random ones"""
    rng = np.random.default_rng(seed)
    nb = max(labels) + 1
    per = m // nb
    P = np.zeros((n, m), np.float32)
    for e in range(n):
        c = labels[e]
        P[e, c * per:min((c + 1) * per, m)] = 5.0 + 3.0 * rng.random(min((c + 1) * per, m) - c * per)
        for _ in range(3):
            g = rng.integers(0, m)
            if P[e, g] == 0:
                P[e, g] = rng.random()
    return P


def test_compact_labels():  # dc_poisson_tests.rs:93-98
    c, k = orc.compact_labels([5, 5, 2, 7, 2, 7, 5])
    assert k == 3 and c.tolist() == [0, 0, 1, 2, 1, 2, 0]


def test_sibling_sets():  # dc_poisson_tests.rs:100-130
    refined = [[0, 1, 2, 3], [0, 0, 1, 1]]
    assert all(s == [0, 1] for s in orc.sibling_sets(refined[1], None, 2))
    sib = orc.sibling_sets(refined[0], refined[1], 4)
    assert sib == [[0, 1], [0, 1], [2, 3], [2, 3]]


def test_candidate_sets_fallback():  # refine_multilevel_tests.rs:4-16
    cand = orc.candidate_sets([[0, 1], [0, 1]], [[1], [0]], [0, 1])
    assert cand == [[0, 1], [0, 1]]
    # an empty intersection falls back to the siblings; a single sibling is returned as it is (dc_poisson.rs:604-617)
    assert orc.candidate_sets([[0, 1], [2]], [[1], [0]], [0, 2]) == [[0, 1], [2]]
    assert orc.candidate_sets([[0, 1, 2]], [[0]], [2]) == [[2]]  # the only neighbour group is the current one
    assert orc.candidate_sets([[3, 4, 5], [3, 4, 5]], [[1], [0]], [5, 3]) == [[3, 5], [3, 5]]


def test_child_offset_and_projection():  # refine_multilevel_tests.rs:18-56
    parent, child = [0, 0, 0, 1, 1, 1], [10, 10, 11, 22, 23, 22]
    off = orc.child_offset_within_parent(child, parent)
    assert off.tolist() == [0, 0, 1, 0, 1, 0]
    _, k = orc.project_to_refinement(off, [3, 3, 3, 8, 8, 8])
    assert k <= 4
    reproj, k = orc.project_to_refinement([0, 0, 0, 1, 1], [0, 0, 1, 1, 1])
    assert k == 3 and reproj[0] == reproj[1] != reproj[2] and reproj[3] == reproj[4] != reproj[2]
    assert reproj.tolist() == [0, 0, 1, 2, 2]  # first-appearance order of the (child, parent) pairs


def test_delta_moves_match_recompute():  # dc_poisson_tests.rs:38-66
    n, m = 24, 16
    labels = [i % 4 for i in range(n)]
    P = toy_profiles(n, m, labels, 1)
    rng = np.random.default_rng(7)
    moves = [(int(rng.integers(0, n)), int(rng.integers(0, 4))) for _ in range(50)]
    st = orc.dcp_stats(P, 4, labels, moves)
    fresh = orc.dcp_stats(P, 4, st["membership"])
    assert np.abs(st["gene_sum"] - fresh["gene_sum"]).max() < 1e-6
    assert np.abs(st["log_gene"] - fresh["log_gene"]).max() < 1e-6
    assert np.abs(st["size_sum"] - fresh["size_sum"]).max() < 1e-6
    assert np.abs(st["log_size_offset"] - fresh["log_size_offset"]).max() < 1e-6


def test_scores_against_float64_restatement():  # the formula of dc_poisson.rs:405-431
    n, m = 20, 12
    labels = [i % 4 for i in range(n)]
    P = toy_profiles(n, m, labels, 2)
    st = orc.dcp_stats(P, 4, labels)
    got = orc.dcp_scores(P, 4, labels, 3)
    sf = np.float32(0)
    for v in P[3][P[3] > 0]:
        sf = np.float32(sf + v)
    want = float(sf) * st["log_size_offset"].astype(np.float64) + st["log_gene"].astype(np.float64) @ P[3].astype(np.float64)
    assert np.allclose(got, want, rtol=1e-12, atol=1e-9)
    # and the statistics themselves
    gs = np.zeros((4, m))
    for e in range(n):
        gs[labels[e]] += P[e]
    assert np.allclose(st["gene_sum"], gs, rtol=1e-12)
    assert np.allclose(st["log_gene"], np.log(gs + 1e-9), rtol=1e-6, atol=1e-6)


def test_empty_block_finite():  # dc_poisson_tests.rs:84-91
    P = np.array([[3.0, 0.0], [0.0, 4.0]], np.float32)
    st = orc.dcp_stats(P, 3, [1, 2])
    assert np.isfinite(st["log_size_offset"][0]) and st["size_sum"][0] == 0.0


def test_smallrng_is_xoshiro256pp_seeded_by_splitmix64():
    # SplitMix64's published first outputs for state 0 are the xoshiro state words of seed_from_u64(0); one xoshiro256++
    # step from them, computed here in plain Python, must be what the oracle's generator returns
    M = (1 << 64) - 1
    s, st = 0, []
    for _ in range(4):
        s = (s + 0x9E3779B97F4A7C15) & M
        z = s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M
        st.append(z ^ (z >> 31))
    assert st[0] == 0xE220A8397B1DCDAF and st[1] == 0x6E789E6AA1B965F4  # SplitMix64 known answers
    rot = lambda x, k: ((x << k) | (x >> (64 - k))) & M
    assert orc.smallrng_u64(0) == (rot((st[0] + st[3]) & M, 23) + st[0]) & M
    u = orc.smallrng_range_f64(0, 1e-12, 1.0)
    assert 1e-12 <= u < 1.0
    bits = ((rot((st[0] + st[3]) & M, 23) + st[0]) & M) >> 12
    assert u == (bits / 2.0 ** 52) * (1.0 - 1e-12) + 1e-12


def test_fisher_weights_bounds_and_poisson_limit():  # nb_dispersion.rs: w in (0, 1], w = 1 when no gene is over-dispersed
    rng = np.random.default_rng(3)
    P = rng.poisson(5.0, (200, 40)).astype(np.float32)
    w = orc.dcp_fisher_weights(P)
    assert np.all(w > 0) and np.all(w <= 1)
    Q = P.copy()
    Q[:, :10] *= rng.gamma(0.5, 2.0, (200, 1)).astype(np.float32)  # over-dispersed genes
    w2 = orc.dcp_fisher_weights(np.round(Q))
    assert w2[:10].mean() < w2[10:].mean()
    const = np.full((50, 8), 3.0, np.float32)  # variance 0 < mean: nothing enters the fit -> phi = 0 -> w = 1
    assert np.array_equal(orc.dcp_fisher_weights(const), np.ones(8, np.float32))


def test_greedy_refinement_recovers_planted_blocks():
    # two levels; strong block structure, a quarter of the entities start in the wrong sibling group
    n, m = 64, 32
    truth = np.array([i % 4 for i in range(n)], np.uint32)
    P = toy_profiles(n, m, truth.tolist(), 5)
    start = truth.copy()
    rng = np.random.default_rng(6)
    wrong = rng.choice(n, 16, replace=False)
    start[wrong] = (start[wrong] + 1) % 4
    cand = [[0, 1, 2, 3]] * n
    got, moves = orc.dcp_refine_level(P, cand, 4, start, 0, 10, 1)
    assert np.array_equal(got, truth) and moves >= 16
    # entities with a single candidate never move; Gibbs sweeps are reproducible for one seed and differ between seeds
    got1, _ = orc.dcp_refine_level(P, [[int(s)] for s in start], 4, start, 5, 5, 1)
    assert np.array_equal(got1, start)
    a, _ = orc.dcp_refine_level(P, cand, 4, start, 3, 0, 12345, 0.0)
    b, _ = orc.dcp_refine_level(P, cand, 4, start, 3, 0, 12345, 0.0)
    assert np.array_equal(a, b)


def test_refine_assignments_keeps_the_hierarchy():
    n, m = 96, 48
    fine_truth = np.array([i % 8 for i in range(n)], np.uint32)
    P = toy_profiles(n, m, fine_truth.tolist(), 8)
    coarse = fine_truth // 2
    init_fine = fine_truth.copy()
    rng = np.random.default_rng(9)
    flip = rng.choice(n, 20, replace=False)
    init_fine[flip] ^= 1  # wrong sibling inside the same parent
    bbknn = [[int(j) for j in rng.choice(n, 6, replace=False)] for _ in range(n)]
    levels, ks, moves = orc.refine_assignments(P, bbknn, [init_fine, coarse], None, num_gibbs=0, num_greedy=10, fisher=False)
    assert len(levels) == 2 and ks[1] == 4 and ks[0] <= 8
    # strict refinement: every fine group has one parent
    for f in range(ks[0]):
        assert len(set(levels[1][levels[0] == f].tolist())) == 1
    # zero sweeps: the compacted initial labels come back (refine_multilevel.rs:215-222)
    lv0, k0, mv0 = orc.refine_assignments(P, bbknn, [init_fine, coarse], None, num_gibbs=0, num_greedy=0)
    assert mv0 == 0 and np.array_equal(lv0[0], orc.compact_labels(init_fine)[0])
