// C++ host-API parity tests: the reference's own tests for the path, restated against include/legume_b200.hpp
// (the CUDA library) with the CPU oracle (oracle/oracle.h, test infrastructure) as the checker.
//
//   random_projection.rs:578-645     tiny 8 x 12 fixture: seeded projection is reproducible
//   weighted_columns.rs:76-121       mu = (1 + sum y) / (1 + n) under (a0, b0) = (1, 1)
//   dmatrix_gamma_tests.rs:9-32      log_sd = sqrt(trigamma(a))
//   knn/tests.rs:74-150              exact search == brute force, ascending, self excluded, cross-dict
//   groups.rs:20-24                  groups ordered by the byte-wise order of code.to_string()
//   collapse_data/mod.rs:867-1050    collapse_columns_multilevel_vec, un-refined, two batches-plus
#include <cmath>
#include <cstdio>
#include <cstring>
#include <random>
#include <string>
#include <vector>

#include "legume_b200.hpp"
#include "oracle.h"

static int g_checks = 0, g_fail = 0;
#define CHECK(cond)                                                          \
    do {                                                                     \
        ++g_checks;                                                          \
        if (!(cond)) {                                                       \
            ++g_fail;                                                        \
            std::fprintf(stderr, "FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond); \
        }                                                                    \
    } while (0)

// stats_tests.rs:20-30 assert_mat_close
static bool close_all(const float* x, const float* y, size_t n, double tol) {
    for (size_t i = 0; i < n; ++i) {
        const double a = x[i], b = y[i];
        if (!(std::fabs(a - b) <= tol * (1.0 + std::max(std::fabs(a), std::fabs(b))))) return false;
    }
    return true;
}

struct Csc {
    std::vector<uint64_t> indptr{0}, indices;
    std::vector<float> data;
    size_t nrows = 0;
    size_t ncols() const { return indptr.size() - 1; }
};

static Csc tiny_fixture() {  // random_projection.rs:578-605
    Csc m;
    m.nrows = 8;
    for (int j = 0; j < 12; ++j) {
        for (int i = 0; i < 8; ++i)
            if ((i * 7 + j * 3) % 5 < 3) {
                m.indices.push_back(i);
                m.data.push_back(1.0f + (float)((i + j) % 4));
            }
        m.indptr.push_back(m.indices.size());
    }
    return m;
}

static Csc random_counts(std::mt19937& rng, size_t D, size_t N, double density) {
    Csc m;
    m.nrows = D;
    std::uniform_real_distribution<double> u(0.0, 1.0);
    std::geometric_distribution<int> geo(0.7);
    for (size_t j = 0; j < N; ++j) {
        for (size_t i = 0; i < D; ++i)
            if (u(rng) < density) {
                m.indices.push_back(i);
                m.data.push_back((float)std::min(1 + geo(rng), 6));
            }
        m.indptr.push_back(m.indices.size());
    }
    return m;
}

static legume::DMatrix gaussian(std::mt19937& rng, size_t r, size_t c) {
    legume::DMatrix a(r, c);
    std::normal_distribution<float> n(0.0f, 1.0f);
    for (auto& x : a.data) x = n(rng);
    return a;
}

int main() {
    using namespace legume;
    Context ctx(0);
    std::mt19937 rng(20240517);

    {  // ---- seeded projection is reproducible and matches the reference arithmetic ----
        Csc m = tiny_fixture();
        SparseIoVec data(ctx, m.indptr, m.indices, m.data, m.nrows);
        DMatrix basis = gaussian(rng, 8, 3);
        auto a = data.project_columns(basis), b = data.project_columns(basis);
        CHECK(a.proj.nrows == 3 && a.proj.ncols == 12 && a.basis.nrows == 8);
        CHECK(std::memcmp(a.proj.data.data(), b.proj.data.data(), a.proj.data.size() * 4) == 0);
        DMatrix basis_kd(3, 8);
        for (int g = 0; g < 8; ++g)
            for (int k = 0; k < 3; ++k) basis_kd(k, g) = basis(g, k);
        std::vector<float> want(3 * 12);
        orc_project_raw(m.indptr.data(), m.indices.data(), m.data.data(), 12, basis_kd.data.data(), 3, want.data(), 1);
        orc_project_finish(want.data(), 3, 12, nullptr, 0);
        CHECK(close_all(a.proj.data.data(), want.data(), want.size(), 1e-5));
    }

    Csc m = random_counts(rng, 150, 600, 0.1);
    const size_t D = m.nrows, N = m.ncols(), K = 12;
    SparseIoVec data(ctx, m.indptr, m.indices, m.data, D);
    DMatrix basis = gaussian(rng, D, K);
    std::vector<std::string> batch(N);
    for (size_t j = 0; j < N; ++j) batch[j] = "b" + std::to_string((j * 7 + j / 13) % 3);
    auto rp = data.project_columns_with_batch_correction(basis, std::nullopt, &batch);
    {  // ---- projection with batch centring vs the oracle ----
        DMatrix basis_kd(K, D);
        for (size_t g = 0; g < D; ++g)
            for (size_t k = 0; k < K; ++k) basis_kd(k, g) = basis(g, k);
        auto ranks = rank_labels(batch);
        std::vector<float> want(K * N);
        orc_project_raw(m.indptr.data(), m.indices.data(), m.data.data(), N, basis_kd.data.data(), (int)K, want.data(), 1);
        orc_project_finish(want.data(), (int)K, N, ranks.first.data(), (uint32_t)ranks.second.size());
        CHECK(close_all(rp.proj.data.data(), want.data(), want.size(), 1e-5));
    }
    {  // ---- groups: lexicographic order of the decimal code strings; sums conserved; mu = (1 + sum) / (1 + n) ----
        const size_t ncode = data.partition_columns_to_groups(rp.proj, 5);
        CHECK(ncode <= 32 && data.num_groups() >= 2);
        std::vector<uint64_t> codes(N);
        CHECK(orc_binary_codes(rp.proj.data.data(), (int)K, N, 5, codes.data(), nullptr, nullptr, nullptr, nullptr) == 0);
        CHECK(codes == data.binary_codes());
        std::vector<uint32_t> grp(N);
        const uint32_t ng = orc_assign_groups(codes.data(), N, grp.data());
        CHECK(ng == data.num_groups() && grp == data.get_group_membership());
        SparseIoVec single(ctx, m.indptr, m.indices, m.data, D);
        single.assign_groups(codes);
        CHECK(single.get_group_membership() == grp);
        CollapsedStat stat(D, ng, 0);
        CollapsedOut out = single.collapse_columns(std::nullopt, std::nullopt, nullptr, std::nullopt, &stat);
        std::vector<float> ws(D * ng), wn(ng);
        orc_collapse_basic(m.indptr.data(), m.indices.data(), m.data.data(), D, N, grp.data(), nullptr, ng, ws.data(), wn.data());
        CHECK(stat.observed_sum_ds.data == ws && stat.size_s == wn);
        bool mu_ok = true;
        for (uint32_t s = 0; s < ng; ++s)
            for (size_t g = 0; g < D; ++g) {
                const float want = (1.0f + ws[s * D + g]) / (1.0f + wn[s]);
                mu_ok &= std::fabs(out.mu_observed.mean(g, s) - want) <= 1e-5f * (1.0f + std::fabs(want));
            }
        CHECK(mu_ok);
    }
    {  // ---- GammaMatrix: log_sd = sqrt(trigamma(a)) ----
        GammaMatrix gm(ctx, 3, 1, 1.0f, 1.0f);
        DMatrix a(3, 1), b(3, 1, 1.0f);
        a(0, 0) = 0.0f;
        a(1, 0) = 1.0f;
        a(2, 0) = 4.0f;
        gm.update_stat(a, b);
        gm.calibrate();
        const double pi2_6 = M_PI * M_PI / 6.0;
        CHECK(std::fabs(gm.posterior_log_sd()(0, 0) - std::sqrt(pi2_6)) < 1e-4);
        CHECK(std::fabs(gm.posterior_log_sd()(1, 0) - std::sqrt(pi2_6 - 1.0)) < 1e-4);
        CHECK(std::fabs(gm.posterior_mean()(2, 0) - 2.5f) < 1e-6 && gm.posterior_log_sd()(2, 0) < gm.posterior_log_sd()(0, 0));
    }
    {  // ---- ColumnDict, exact backend ----
        DMatrix pts = gaussian(rng, 16, 500);
        ColumnDict dict(ctx, pts);
        bool same = true, ascending = true, no_self = true;
        for (size_t q = 0; q < 500; q += 37) {
            auto res = dict.search_others(q, 8);
            std::vector<uint32_t> widx(8), ex{(uint32_t)q};
            std::vector<float> wd(8);
            orc_knn_topk(pts.data.data(), 500, pts.column(q), 1, 16, 8, ex.data(), widx.data(), wd.data(), 1);
            for (size_t i = 0; i < 8; ++i) {
                same &= res.first[i] == widx[i] && res.second[i] == wd[i];
                no_self &= res.first[i] != q;
                if (i) ascending &= res.second[i] >= res.second[i - 1];
            }
        }
        CHECK(same && ascending && no_self);
        DMatrix other = gaussian(rng, 16, 40);
        ColumnDict small(ctx, other);
        auto cross = dict.match_by_query_name_against(3, 50, small);  // fewer points than knn: all of them, nearest first
        CHECK(cross.first.size() == 40);
        auto by_data = dict.search_by_query_data(std::vector<float>(pts.column(7), pts.column(7) + 16), 1);
        CHECK(by_data.first.size() == 1 && by_data.first[0] == 7 && by_data.second[0] == 0.0f);
    }
    {  // ---- collapse_columns_multilevel_vec (pb-sample matched stats) vs the oracle's composition ----
        MultilevelParams params(K);
        params.refine = false;  // the legacy un-refined descent (MultilevelParams::new has refine = Some)
        params.sort_dim = 5;
        params.num_levels = 2;
        params.knn_pb_samples = 3;
        params.num_opt_iter = 12;
        std::vector<CollapsedStat> stats;
        auto levels = data.collapse_columns_multilevel_vec(rp.proj, batch, params, &stats);
        const uint32_t B = (uint32_t)data.num_batches(), S = (uint32_t)stats[0].num_samples();
        CHECK(levels.size() == 1 && B == 3);  // dims(5, 2) = [5, 5] dedups to one level (refine.rs:718-734)
        const std::vector<uint32_t>& grp = data.get_group_membership();
        const std::vector<uint32_t>& bat = data.col_to_batch();
        std::vector<uint32_t> c2p(N), pg(S * B), pb(S * B);
        std::vector<float> cnt(S * B), cen((size_t)S * B * K);
        const uint32_t npb = orc_pb_layout(rp.proj.data.data(), (int)K, N, grp.data(), S, bat.data(), B, nullptr, c2p.data(), pg.data(),
                                           pb.data(), cnt.data(), cen.data());
        std::vector<float> gs((size_t)D * npb), gsz(npb);
        orc_collapse_basic(m.indptr.data(), m.indices.data(), m.data.data(), D, N, c2p.data(), nullptr, npb, gs.data(), gsz.data());
        std::vector<uint32_t> mp((size_t)npb * B * 3);
        std::vector<float> md((size_t)npb * B * 3);
        orc_pb_match(rp.proj.data.data(), (int)K, N, bat.data(), B, c2p.data(), cen.data(), pb.data(), npb, 3, mp.data(), md.data(), 1);
        std::vector<float> imp((size_t)D * S), res((size_t)D * S);
        orc_collect_matched_stat_coarse(gs.data(), D, npb, cnt.data(), pg.data(), S, mp.data(), md.data(), B * 3, imp.data(), res.data());
        CHECK(close_all(stats[0].imputed_sum_ds.data.data(), imp.data(), imp.size(), 1e-5));
        CHECK(close_all(stats[0].residual_sum_ds.data.data(), res.data(), res.size(), 1e-5));
        std::vector<float> mu_obs((size_t)D * S), mu_adj((size_t)D * S), mu_res((size_t)D * S), gam((size_t)D * S), delta((size_t)D * B),
            lm((size_t)D * S);
        orc_optimize_batched(stats[0].observed_sum_ds.data.data(), imp.data(), res.data(), stats[0].size_s.data(),
                             stats[0].observed_sum_db.data.data(), stats[0].n_bs.data.data(), D, S, B, 1.0f, 1.0f, 12, 0, mu_obs.data(),
                             mu_adj.data(), mu_res.data(), gam.data(), delta.data(), lm.data());
        CHECK(close_all(levels[0].mu_adjusted->mean.data.data(), mu_adj.data(), mu_adj.size(), 1e-4));
        CHECK(close_all(levels[0].delta->mean.data.data(), delta.data(), delta.size(), 1e-4));
    }
    {  // ---- the refinement arm with ONE batch (refine.rs:126-147, 264-500): hash partition compacted per level, merge descent ----
        SparseIoVec one(ctx, m.indptr, m.indices, m.data, m.nrows);
        MultilevelParams params(K);
        params.sort_dim = 9;  // levels 9, 8, 7: below 8 compute_level_sort_dims gives one level only
        params.num_levels = 3;
        params.num_opt_iter = 12;
        std::vector<uint32_t> b1(N, 0u);
        MultilevelCollapseOut out = one.collapse_columns_multilevel_with_hierarchy(rp.proj, b1, params);
        const std::vector<size_t> dims = compute_level_sort_dims(9, 3);
        CHECK(dims.size() == 3);
        CHECK(params.refine && out.levels.size() == dims.size() && out.cell_to_pb_per_level.size() == dims.size());
        std::vector<uint64_t> codes(N);
        CHECK(orc_binary_codes(rp.proj.data.data(), (int)K, N, (int)dims[0], codes.data(), nullptr, nullptr, nullptr, nullptr) == 0);
        // one batch: a pb-sample is a hash group, so level l is the compacted masked code of the group's first cell
        std::vector<uint32_t> hash_grp(N);
        const uint32_t ng = orc_assign_groups(codes.data(), N, hash_grp.data());
        std::vector<uint32_t> c2p(N), pg(ng), pb(ng);
        std::vector<float> cnt(ng), cen((size_t)ng * K);
        const uint32_t npb = orc_pb_layout(rp.proj.data.data(), (int)K, N, hash_grp.data(), ng, b1.data(), 1, nullptr, c2p.data(), pg.data(),
                                           pb.data(), cnt.data(), cen.data());
        bool maps_ok = npb == ng;
        for (size_t level = 0; level < dims.size() && maps_ok; ++level) {
            std::vector<uint64_t> first_code(npb, ~0ull);
            for (size_t c = N; c-- > 0;) first_code[c2p[c]] = codes[c] & ((1ull << dims[level]) - 1);
            auto cl = compact_labels(first_code);
            for (size_t c = 0; c < N; ++c) maps_ok = maps_ok && out.cell_to_pb_per_level[level][c] == cl.first[c2p[c]];
            CHECK(out.stats[level].num_samples() == cl.second);
        }
        CHECK(maps_ok);
        const std::vector<uint32_t>& fine = out.cell_to_pb_per_level[0];
        const uint32_t S0 = (uint32_t)out.stats[0].num_samples();
        std::vector<float> ws((size_t)D * S0), wn(S0);
        orc_collapse_basic(m.indptr.data(), m.indices.data(), m.data.data(), D, N, fine.data(), nullptr, S0, ws.data(), wn.data());
        CHECK(out.stats[0].observed_sum_ds.data == ws && out.stats[0].size_s == wn);
        std::vector<float> mean((size_t)D * S0), sd((size_t)D * S0), lm((size_t)D * S0), ls((size_t)D * S0);
        orc_optimize_single(ws.data(), wn.data(), D, S0, 1.0f, 1.0f, 0, mean.data(), sd.data(), lm.data(), ls.data());
        CHECK(close_all(out.levels[0].mu_observed.mean.data.data(), mean.data(), mean.size(), 1e-5));
        CHECK(close_all(out.levels[0].mu_observed.log_mean.data.data(), lm.data(), lm.size(), 1e-5));
        // the coarsest level's sums are the finest ones merged: totals per gene agree exactly (whole numbers)
        const CollapsedStat& last = out.stats.back();
        bool totals = true;
        for (size_t g = 0; g < D && totals; ++g) {
            double a = 0, b = 0;
            for (size_t sidx = 0; sidx < S0; ++sidx) a += out.stats[0].observed_sum_ds(g, sidx);
            for (size_t sidx = 0; sidx < last.num_samples(); ++sidx) b += last.observed_sum_ds(g, sidx);
            totals = a == b;
        }
        CHECK(totals);
        // inheriting the finest map as a one-level partition reproduces the finest statistics
        MultilevelParams p1(K);
        p1.sort_dim = 9;
        p1.num_levels = 1;
        SparseIoVec again(ctx, m.indptr, m.indices, m.data, m.nrows);
        MultilevelCollapseOut inh = again.collapse_columns_multilevel_with_partition(rp.proj, b1, p1, {fine});
        CHECK(inh.levels.size() == 1 && inh.stats[0].observed_sum_ds.data == out.stats[0].observed_sum_ds.data);
        bool threw = false;
        try {
            again.collapse_columns_multilevel_with_partition(rp.proj, b1, params, {fine});  // 1 level given, 3 asked for
        } catch (const Error&) {
            threw = true;
        }
        CHECK(threw);
    }
    {  // ---- the refinement arm with THREE batches: BBKNN candidates + DC-Poisson sweeps (refine_multilevel.rs:170-298) ----
        SparseIoVec three(ctx, m.indptr, m.indices, m.data, m.nrows);
        MultilevelParams params(K);
        params.sort_dim = 9;
        params.num_levels = 2;
        params.knn_pb_samples = 3;
        params.num_opt_iter = 12;
        MultilevelCollapseOut out = three.collapse_columns_multilevel_with_hierarchy(rp.proj, batch, params);
        const std::vector<size_t> dims = compute_level_sort_dims(9, 2);
        const uint32_t B = (uint32_t)three.num_batches();
        CHECK(B == 3 && out.levels.size() == dims.size());
        std::vector<uint64_t> codes(N);
        CHECK(orc_binary_codes(rp.proj.data.data(), (int)K, N, (int)dims[0], codes.data(), nullptr, nullptr, nullptr, nullptr) == 0);
        std::vector<uint32_t> hash_grp(N);
        const uint32_t ng = orc_assign_groups(codes.data(), N, hash_grp.data());
        const std::vector<uint32_t>& bat = three.col_to_batch();
        std::vector<uint32_t> c2p(N), pg((size_t)ng * B), pb((size_t)ng * B);
        std::vector<float> cnt((size_t)ng * B), cen((size_t)ng * B * K);
        const uint32_t npb = orc_pb_layout(rp.proj.data.data(), (int)K, N, hash_grp.data(), ng, bat.data(), B, nullptr, c2p.data(), pg.data(),
                                           pb.data(), cnt.data(), cen.data());
        std::vector<float> gs((size_t)D * npb), gsz(npb);
        orc_collapse_basic(m.indptr.data(), m.indices.data(), m.data.data(), D, N, c2p.data(), nullptr, npb, gs.data(), gsz.data());
        std::vector<uint32_t> mp((size_t)npb * B * 3), bb_ptr(npb + 1, 0), bb;
        std::vector<float> md((size_t)npb * B * 3);
        orc_pb_match(rp.proj.data.data(), (int)K, N, bat.data(), B, c2p.data(), cen.data(), pb.data(), npb, 3, mp.data(), md.data(), 1);
        for (uint32_t p = 0; p < npb; ++p) {
            for (uint32_t i = 0; i < B * 3; ++i)
                if (mp[(size_t)p * B * 3 + i] != 0xFFFFFFFFu) bb.push_back(mp[(size_t)p * B * 3 + i]);
            bb_ptr[p + 1] = (uint32_t)bb.size();
        }
        const size_t L = dims.size();
        std::vector<uint64_t> first_code(npb, 0);
        for (size_t c = N; c-- > 0;) first_code[c2p[c]] = codes[c];
        std::vector<uint32_t> init(L * npb), offs(L * npb, 0u), want(L * npb), want_k(L);
        for (size_t level = 0; level < L; ++level) {
            std::vector<uint64_t> masked(npb);
            for (uint32_t p = 0; p < npb; ++p) masked[p] = first_code[p] & ((1ull << dims[level]) - 1);
            auto cl = compact_labels(masked);
            std::copy(cl.first.begin(), cl.first.end(), init.begin() + level * npb);
            if (level + 1 < L)
                for (uint32_t p = 0; p < npb; ++p)
                    offs[level * npb + p] = (uint32_t)((first_code[p] >> dims[level + 1]) & ((1ull << (dims[level] - dims[level + 1])) - 1));
        }
        const uint64_t wmoves = orc_refine_assignments(gs.data(), npb, D, bb_ptr.data(), bb.data(), (int)L, init.data(), offs.data(), 20, 10, 1, 42,
                                                       0.005, want.data(), want_k.data());
        CHECK(three.refine_moves() == wmoves);
        bool maps_ok = true;
        for (size_t level = 0; level < L; ++level) {
            CHECK(out.stats[level].num_samples() == want_k[level]);
            for (size_t c = 0; c < N; ++c) maps_ok = maps_ok && out.cell_to_pb_per_level[level][c] == want[level * npb + c2p[c]];
        }
        CHECK(maps_ok);
        const std::vector<uint32_t>& fine = out.cell_to_pb_per_level[0];
        const uint32_t S0 = (uint32_t)out.stats[0].num_samples();
        std::vector<float> ws((size_t)D * S0), wn(S0);
        orc_collapse_basic(m.indptr.data(), m.indices.data(), m.data.data(), D, N, fine.data(), nullptr, S0, ws.data(), wn.data());
        CHECK(out.stats[0].observed_sum_ds.data == ws && out.stats[0].size_s == wn);
    }
    {  // ---- either side of the path: running statistics (sparse_stat.rs:671-728) and the Nystrom pass ----
        std::mt19937 rng21(21);
        Csc m = random_counts(rng21, 300, 500, 0.1);
        SparseIoVec x(ctx, m.indptr, m.indices, m.data, 300);
        SparseRunningStatistics st = x.streaming_sparse_running_stats();
        std::vector<float> npos(300), s1(300), s2(300), mean(300), var(300), sd(300);
        orc_row_stats(m.indptr.data(), m.indices.data(), m.data.data(), 300, 500, npos.data(), s1.data(), s2.data());
        orc_row_stats_moments(s1.data(), s2.data(), 300, 500, mean.data(), var.data(), sd.data());
        CHECK(st.ncols_processed() == 500 && st.count_positives() == npos && st.sum() == s1);
        CHECK(st.mean() == mean && st.variance() == var);
        // reference's own known answer: columns [1,0,2,0] and [0,3,0,4] -> npos 1,1,1,1, sum 1,3,2,4, mean .5,1.5,1,2
        SparseIoVec tiny(ctx, std::vector<uint64_t>{0, 2, 4}, std::vector<uint64_t>{0, 2, 1, 3}, std::vector<float>{1, 2, 3, 4}, 4);
        SparseRunningStatistics ts = tiny.streaming_sparse_running_stats();
        CHECK((ts.count_positives() == std::vector<float>{1, 1, 1, 1}) && (ts.sum() == std::vector<float>{1, 3, 2, 4}) &&
              (ts.mean() == std::vector<float>{0.5f, 1.5f, 1.0f, 2.0f}));
        DMatrix basis(300, 12);
        for (size_t i = 0; i < basis.data.size(); ++i) basis.data[i] = std::sin(0.37f * (float)i);
        DMatrix ny = x.nystrom_project(basis);
        std::vector<float> want(12 * 500);
        orc_nystrom_project(m.indptr.data(), m.indices.data(), m.data.data(), 300, 500, basis.data.data(), 12, nullptr, nullptr, 0, 1e4f,
                            want.data());
        // the reference's f32 folds (the oracle) sit up to ~5e-4 from exact arithmetic here; see tests/test_gpu_next.py
        CHECK(ny.nrows == 12 && ny.ncols == 500 && close_all(ny.data.data(), want.data(), want.size(), 2e-3));
    }
    {  // ---- SparseIoStack: per-modality projection stacked vertically (random_projection.rs:200-260) ----
        std::mt19937 r1(31), r2(32);
        Csc a = random_counts(r1, 200, 300, 0.1), b = random_counts(r2, 90, 300, 0.15);
        SparseIoVec va(ctx, a.indptr, a.indices, a.data, 200), vb(ctx, b.indptr, b.indices, b.data, 90);
        std::mt19937 rb(33);
        std::vector<DMatrix> bases{gaussian(rb, 200, 8), gaussian(rb, 90, 8)};
        SparseIoStack stack({&va, &vb});
        RandColProjOut st = stack.project_columns_with_batch_correction<uint32_t>(bases, std::nullopt, nullptr);
        RandColProjOut pa = va.project_columns(bases[0]), pb2 = vb.project_columns(bases[1]);
        bool same = st.proj.nrows == 16 && st.proj.ncols == 300 && st.basis.nrows == 290 && stack.num_rows() == 290;
        for (size_t j = 0; j < 300 && same; ++j)
            for (size_t k = 0; k < 8; ++k) same = same && st.proj(k, j) == pa.proj(k, j) && st.proj(8 + k, j) == pb2.proj(k, j);
        CHECK(same);
        CHECK(mix_seed(0, 1) == 0xE220A8397B1DCDAFull);  // first output of SplitMix64 seeded with 0
    }
    {  // ---- error behaviour: anyhow::Error -> legume::Error, never a crash ----
        bool threw = false;
        try {
            std::vector<std::string> bad(3, "x");
            data.register_batch_membership(bad);
        } catch (const Error&) {
            threw = true;
        }
        CHECK(threw);
        threw = false;
        try {
            Csc bad = tiny_fixture();
            bad.indices[2] = 99;  // row out of range
            SparseIoVec x(ctx, bad.indptr, bad.indices, bad.data, bad.nrows);
        } catch (const Error& e) {
            threw = std::string(e.what()).find("out of range") != std::string::npos;
        }
        CHECK(threw);
    }
    std::printf("%s: %d checks, %d failed, %llu kernel launches\n", g_fail ? "FAILED" : "ok", g_checks, g_fail,
                (unsigned long long)ctx.launch_count());
    return g_fail ? 1 : 0;
}
