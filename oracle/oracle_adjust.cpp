/*
 * oracle_adjust.cpp — CPU restatement of the cross-batch neighbourhood adjustment
 * (SURVEY.md §8a rows a14–a17).  TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Two paths, both restated from the reference:
 *   per-cell   matched.rs:173-260 (neighbouring_columns_triplets) + batch.rs:182-234
 *              (sort_batch_proximity) + collapse_data/stats.rs:26-108 (collect_matched_stat_visitor)
 *   pb-sample  collapse_data/pb_samples.rs:94-459 + collapse_data/stats.rs:698-784
 *              (collect_matched_stat_coarse) — what collapse_columns_multilevel_vec runs for B >= 2
 *   levels     collapse_data/refine.rs:741-769 (compute_fine_to_coarse_mapping)
 *
 * Third-party arithmetic restated here (not on disk; parity of these pieces is unpinned):
 *   nalgebra 0.34.2  column_mean  = fold of axpy(1/n, col, 1) over columns (two roundings per step)
 *   nalgebra-sparse 0.11.0 CSC*CSC = spmm_csr_prealloc on the transposes: for each output column j,
 *                    for each stored w[t,j] in ascending t, c[g] += (1*w)*y[g,t]  (product rounded, then sum)
 *   libm expf (Rust f32::exp)
 *
 * All citations are relative to /root/reference (causalpathlab/legume-rs v0.3.2).
 */
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <utility>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "oracle.h"

namespace {
const float INF = std::numeric_limits<float>::infinity();

/* (squared distance, index) of every point of `members` from q, ascending: the exact backend's order
 * (exact.rs:36-55) with the oracle's tie rule (lower index first). */
void scan_sorted(const float* proj, int K, const std::vector<uint32_t>& members, const float* q,
                 std::vector<std::pair<float, uint32_t>>& out) {
    out.resize(members.size());
    for (size_t i = 0; i < members.size(); ++i) out[i] = {orc_l2_sq(proj + (size_t)members[i] * K, q, K), members[i]};
    std::sort(out.begin(), out.end());
}
}  // namespace

/* ---- batch.rs:182-234 sort_batch_proximity -------------------------------------------------------
 * centroid_b = DMatrix(K x n_b).column_mean(); prox[b] = all batches by distance from centroid_b
 * (search_by_query_name(b, nbatches, exclude_same = false); exact backend since B <= 8192). */
extern "C" void orc_batch_proximity(const float* proj, int K, uint64_t N, const uint32_t* batch, uint32_t B,
                                    uint32_t* order, float* out_centroids) {
    std::vector<float> cen((size_t)B * K, 0.0f);
    std::vector<uint64_t> cnt(B, 0);
    for (uint64_t j = 0; j < N; ++j)
        if (batch[j] < B) cnt[batch[j]]++;
    for (uint64_t j = 0; j < N; ++j) {
        const uint32_t b = batch[j];
        if (b >= B) continue;
        const float denom = 1.0f / (float)(double)cnt[b];
        for (int k = 0; k < K; ++k) {
            const float ax = denom * proj[(size_t)j * K + k];
            cen[(size_t)b * K + k] = ax + cen[(size_t)b * K + k];
        }
    }
    if (out_centroids) std::memcpy(out_centroids, cen.data(), sizeof(float) * cen.size());
    for (uint32_t b = 0; b < B; ++b) {
        std::vector<std::pair<float, uint32_t>> sc(B);
        for (uint32_t o = 0; o < B; ++o) sc[o] = {orc_l2_sq(&cen[(size_t)o * K], &cen[(size_t)b * K], K), o};
        std::sort(sc.begin(), sc.end());
        for (uint32_t o = 0; o < B; ++o) order[(size_t)b * B + o] = sc[o].second;
    }
}

/* ---- matched.rs:173-260 neighbouring_columns_triplets (the kNN part) -------------------------------
 * For source cell j (batch s) and slot i: target batch b = target_order[s*nt + i] (or i when NULL);
 * b == s / b >= B leaves the slot empty (skip_same_batch = true).  Otherwise the knn nearest cells of
 * batch b (match_by_query_name_against, knn/mod.rs:230-241), nearest first, as GLOBAL cell indices. */
extern "C" void orc_knn_match_batches(const float* proj, int K, uint64_t N, const uint32_t* batch, uint32_t B, int knn,
                                      const uint32_t* target_order, uint32_t nt, uint32_t* out_idx, float* out_dist,
                                      int nthreads) {
    std::vector<std::vector<uint32_t>> members(B);
    for (uint64_t j = 0; j < N; ++j)
        if (batch[j] < B) members[batch[j]].push_back((uint32_t)j);
    const size_t T = (size_t)nt * knn;
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
    (void)nthreads;
#endif
#pragma omp parallel num_threads(nthreads > 0 ? nthreads : 1)
    {
        std::vector<std::pair<float, uint32_t>> sc;
#pragma omp for schedule(dynamic, 16)
        for (int64_t j = 0; j < (int64_t)N; ++j) {
            uint32_t* oi = out_idx + (size_t)j * T;
            float* od = out_dist + (size_t)j * T;
            for (size_t t = 0; t < T; ++t) {
                oi[t] = UINT32_MAX;
                od[t] = INF;
            }
            const uint32_t s = batch[j];
            if (s >= B) continue;
            for (uint32_t i = 0; i < nt; ++i) {
                const uint32_t b = target_order ? target_order[(size_t)s * nt + i] : i;
                if (b >= B || b == s) continue;
                scan_sorted(proj, K, members[b], proj + (size_t)j * K, sc);
                int w = 0;
                for (size_t r = 0; r < sc.size() && r < (size_t)knn; ++r) {
                    if (sc[r].second == (uint32_t)j) continue;  // matched.rs:241-243
                    oi[(size_t)i * knn + w] = sc[r].second;
                    od[(size_t)i * knn + w] = std::sqrt(sc[r].first);
                    ++w;
                }
            }
        }
    }
}

/* ---- stats.rs:26-108 collect_matched_stat_visitor ------------------------------------------------
 * per group s, per source cell j (ascending): W = normalize_exp_logits_columns of -d (dmatrix_util.rs:
 * 649-671: subtracts the MIN logit), y_hat = Y_matched * W, y1 adjusted by division
 * (dmatrix_util.rs:145-176), imputed[:, s] += y_hat, residual[:, s] += y1. */
extern "C" void orc_collect_matched_stat(const uint64_t* indptr, const uint64_t* indices, const float* data, uint64_t D,
                                         uint64_t N, const uint32_t* grp, uint32_t S, const uint32_t* midx,
                                         const float* mdist, uint32_t T, float* imputed_ds, float* residual_ds) {
    std::memset(imputed_ds, 0, sizeof(float) * (size_t)D * S);
    std::memset(residual_ds, 0, sizeof(float) * (size_t)D * S);
    std::vector<float> yhat(D, 0.0f);
    std::vector<uint8_t> present(D, 0);
    std::vector<uint32_t> touched;
    std::vector<float> w;
    for (uint64_t j = 0; j < N; ++j) {  // per (gene, group) entry the adds happen in ascending cell order
        const uint32_t s = grp[j];
        if (s >= S) continue;
        const uint32_t* mi = midx + (size_t)j * T;
        const float* md = mdist + (size_t)j * T;
        // softmax over the stored logits of this column, in ascending matched-column order
        w.assign(T, 0.0f);
        bool any = false;
        float log_max = 0.0f;
        for (uint32_t t = 0; t < T; ++t)
            if (mi[t] != UINT32_MAX) {
                const float l = -md[t];
                log_max = any ? std::min(log_max, l) : l;
                any = true;
            }
        float denom = 0.0f;
        for (uint32_t t = 0; t < T; ++t)
            if (mi[t] != UINT32_MAX) denom += std::exp(-md[t] - log_max);
        for (uint32_t t = 0; t < T; ++t)
            if (mi[t] != UINT32_MAX) w[t] = std::exp(-md[t] - log_max) / denom;
        // y_hat[:, j] = sum_t w_t * y[:, m_t]
        touched.clear();
        for (uint32_t t = 0; t < T; ++t) {
            if (mi[t] == UINT32_MAX) continue;
            const uint64_t m = mi[t];
            for (uint64_t e = indptr[m]; e < indptr[m + 1]; ++e) {
                const uint64_t g = indices[e];
                if (!present[g]) {
                    present[g] = 1;
                    touched.push_back((uint32_t)g);
                }
                const float prod = w[t] * data[e];
                yhat[g] += prod;
            }
        }
        std::sort(touched.begin(), touched.end());
        float dsum = 0.0f, xsum = 0.0f;
        for (uint32_t g : touched) dsum += yhat[g];
        for (uint64_t e = indptr[j]; e < indptr[j + 1]; ++e) xsum += data[e];
        const float scale = dsum > 0.0f ? xsum / dsum : 1.0f;
        float* imp = imputed_ds + (size_t)s * D;
        float* res = residual_ds + (size_t)s * D;
        for (uint32_t g : touched) imp[g] += yhat[g];
        for (uint64_t e = indptr[j]; e < indptr[j + 1]; ++e) {
            const uint64_t g = indices[e];
            float x = data[e];
            const float d = present[g] ? yhat[g] : 0.0f;
            if (d > 0.0f) x /= d * scale;
            res[g] += x;
        }
        for (uint32_t g : touched) {
            yhat[g] = 0.0f;
            present[g] = 0;
        }
    }
}

/* ---- pb_samples.rs:94-219 build_pb_sample_layout (no anchor / bulk batches) -------------------------
 * pb-sample = non-empty (group, batch) block.  The reference enumerates a group's blocks in HashMap
 * order (unspecified); the oracle fixes ascending batch.  Returns the number of pb-samples. */
extern "C" uint32_t orc_pb_layout(const float* proj, int K, uint64_t N, const uint32_t* grp, uint32_t S,
                                  const uint32_t* batch, uint32_t B, const float* mult, uint32_t* cell_to_pb,
                                  uint32_t* pb_group, uint32_t* pb_batch, float* pb_count, float* centroids) {
    std::vector<uint32_t> id((size_t)S * B, UINT32_MAX);
    std::vector<uint8_t> seen((size_t)S * B, 0);
    for (uint64_t j = 0; j < N; ++j)
        if (grp[j] < S && batch[j] < B) seen[(size_t)grp[j] * B + batch[j]] = 1;
    uint32_t npb = 0;
    for (size_t e = 0; e < seen.size(); ++e)
        if (seen[e]) {
            pb_group[npb] = (uint32_t)(e / B);
            pb_batch[npb] = (uint32_t)(e % B);
            id[e] = npb++;
        }
    std::vector<float> sum((size_t)npb * K, 0.0f), cnt(npb, 0.0f);
    for (uint64_t j = 0; j < N; ++j) {
        cell_to_pb[j] = UINT32_MAX;
        if (grp[j] >= S || batch[j] >= B) continue;
        const uint32_t p = id[(size_t)grp[j] * B + batch[j]];
        cell_to_pb[j] = p;
        const float w = mult ? mult[j] : 1.0f;
        for (int k = 0; k < K; ++k) sum[(size_t)p * K + k] += proj[(size_t)j * K + k] * w;
        cnt[p] += w;
    }
    for (uint32_t p = 0; p < npb; ++p) {
        // blocks with count <= 0 are filtered by the reference (:170); multiplicities are > 0 so none are
        const float inv = 1.0f / cnt[p];
        for (int k = 0; k < K; ++k) centroids[(size_t)p * K + k] = sum[(size_t)p * K + k] * inv;
        pb_count[p] = cnt[p];
    }
    return npb;
}

/* ---- pb_samples.rs:323-459 knn_distinct_pbsamples_in_batch / bbknn_match_one_pbsamp (pooled) ---------
 * out slot (p, b, r): r-th nearest distinct foreign pb-sample of batch b for pb-sample p's centroid,
 * b ascending, own batch left empty.  The adaptive query_k loop is restated literally. */
extern "C" void orc_pb_match(const float* proj, int K, uint64_t N, const uint32_t* batch, uint32_t B,
                             const uint32_t* cell_to_pb, const float* centroids, const uint32_t* pb_batch, uint32_t npb,
                             int knn, uint32_t* out_pb, float* out_dist, int nthreads) {
    std::vector<std::vector<uint32_t>> members(B);
    for (uint64_t j = 0; j < N; ++j)
        if (batch[j] < B) members[batch[j]].push_back((uint32_t)j);
    const size_t T = (size_t)B * knn;
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
    (void)nthreads;
#endif
#pragma omp parallel num_threads(nthreads > 0 ? nthreads : 1)
    {
        std::vector<std::pair<float, uint32_t>> sc;
        std::vector<std::pair<float, uint32_t>> best;  // (distance, pb) in first-occurrence order
#pragma omp for schedule(dynamic, 4)
        for (int64_t p = 0; p < (int64_t)npb; ++p) {
            uint32_t* op = out_pb + (size_t)p * T;
            float* od = out_dist + (size_t)p * T;
            for (size_t t = 0; t < T; ++t) {
                op[t] = UINT32_MAX;
                od[t] = INF;
            }
            for (uint32_t b = 0; b < B; ++b) {
                if (b == pb_batch[p]) continue;
                const size_t n = members[b].size();
                if (n == 0 || knn == 0) continue;
                scan_sorted(proj, K, members[b], centroids + (size_t)p * K, sc);
                size_t query_k = std::min<size_t>((size_t)knn * 4 + 1, n);
                while (true) {
                    best.clear();
                    for (size_t r = 0; r < query_k; ++r) {
                        const uint32_t other = cell_to_pb[sc[r].second];
                        if (other == UINT32_MAX || other == (uint32_t)p) continue;
                        const float d = std::sqrt(sc[r].first);
                        bool found = false;
                        for (auto& e : best)
                            if (e.second == other) {
                                if (d < e.first) e.first = d;
                                found = true;
                                break;
                            }
                        if (!found) best.push_back({d, other});
                    }
                    if (best.size() >= (size_t)knn || query_k >= n) break;
                    query_k = std::min<size_t>(query_k * 4, n);
                }
                // sort_by(partial_cmp) is stable; ties keep first-occurrence order
                std::stable_sort(best.begin(), best.end(),
                                 [](const std::pair<float, uint32_t>& a, const std::pair<float, uint32_t>& c) { return a.first < c.first; });
                for (size_t r = 0; r < best.size() && r < (size_t)knn; ++r) {
                    op[(size_t)b * knn + r] = best[r].second;
                    od[(size_t)b * knn + r] = best[r].first;
                }
            }
        }
    }
}

/* ---- stats.rs:698-784 collect_matched_stat_coarse -------------------------------------------------
 * gene_sums: D x npb dense (zero = gene absent from the pb-sample's sparse list).  The reference adds
 * pb-samples under a mutex in arbitrary order; the oracle fixes ascending pb-sample. */
extern "C" void orc_collect_matched_stat_coarse(const float* gene_sums, uint64_t D, uint32_t npb, const float* pb_count,
                                                const uint32_t* pb_to_group, uint32_t S, const uint32_t* mpb,
                                                const float* mdist, uint32_t T, float* imputed_ds, float* residual_ds) {
    std::memset(imputed_ds, 0, sizeof(float) * (size_t)D * S);
    std::memset(residual_ds, 0, sizeof(float) * (size_t)D * S);
    std::vector<float> yhat(D), w(T);
    std::vector<uint8_t> present(D);
    for (uint32_t p = 0; p < npb; ++p) {
        const uint32_t s = pb_to_group[p];
        const float sc_count = pb_count[p];
        if (sc_count < 1.0f || s >= S) continue;
        const uint32_t* mi = mpb + (size_t)p * T;
        const float* md = mdist + (size_t)p * T;
        bool any = false;
        float max_neg = -INF;
        for (uint32_t t = 0; t < T; ++t)
            if (mi[t] != UINT32_MAX) {
                any = true;
                max_neg = std::max(max_neg, -md[t]);
            }
        if (!any) continue;
        float wsum = 0.0f;
        for (uint32_t t = 0; t < T; ++t) {
            w[t] = 0.0f;
            if (mi[t] != UINT32_MAX) {
                w[t] = std::exp(-md[t] - max_neg);
                wsum += w[t];
            }
        }
        if (wsum > 0.0f)
            for (uint32_t t = 0; t < T; ++t) w[t] /= wsum;
        std::fill(yhat.begin(), yhat.end(), 0.0f);
        std::fill(present.begin(), present.end(), 0);
        for (uint32_t t = 0; t < T; ++t) {
            if (mi[t] == UINT32_MAX) continue;
            const float mc = pb_count[mi[t]];
            if (mc < 1.0f) continue;
            const float inv = 1.0f / mc;
            const float* gs = gene_sums + (size_t)mi[t] * D;
            for (uint64_t g = 0; g < D; ++g)
                if (gs[g] != 0.0f) {
                    present[g] = 1;
                    yhat[g] += w[t] * gs[g] * inv;
                }
        }
        float* imp = imputed_ds + (size_t)s * D;
        float* res = residual_ds + (size_t)s * D;
        const float* own = gene_sums + (size_t)p * D;
        for (uint64_t g = 0; g < D; ++g) {
            if (present[g]) imp[g] += sc_count * yhat[g];
            if (own[g] != 0.0f && present[g] && yhat[g] > 0.0f) res[g] += own[g] / yhat[g];
        }
    }
}

/* ---- refine.rs:741-769 compute_fine_to_coarse_mapping -----------------------------------------------
 * group_code[f] = binary code shared by the cells of fine group f. */
extern "C" uint32_t orc_fine_to_coarse(const uint64_t* group_code, uint32_t nfine, int coarse_dim, uint32_t* f2c) {
    const uint64_t mask = coarse_dim >= 64 ? ~0ull : ((1ull << coarse_dim) - 1ull);
    std::vector<uint64_t> uniq(nfine);
    for (uint32_t f = 0; f < nfine; ++f) uniq[f] = group_code[f] & mask;
    std::sort(uniq.begin(), uniq.end());
    uniq.erase(std::unique(uniq.begin(), uniq.end()), uniq.end());
    for (uint32_t f = 0; f < nfine; ++f)
        f2c[f] = (uint32_t)(std::lower_bound(uniq.begin(), uniq.end(), group_code[f] & mask) - uniq.begin());
    return (uint32_t)uniq.size();
}
