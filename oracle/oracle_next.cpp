/*
 * oracle_next.cpp — CPU restatement of the steps either side of the hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle.h).  Nothing under legume-rs_b200/ may include or link this.
 *
 *  - per-gene running statistics of a sparse block: SparseRunningStatistics::add_sparse_column / add_csc
 *    (matrix-util/src/sparse_stat.rs:64-108), driven by streaming_sparse_running_stats
 *    (data-beans-alg/src/sparse_streaming.rs:23-60).  Pinned by the reference's known-answer tests
 *    sparse_stat.rs:671-728 (tests/test_oracle_next.py).
 *  - Nystrom re-projection of every cell: nystrom_proj_visitor (senna/src/svd/fit.rs:433-466) with
 *    normalize_columns_inplace (matrix-util/src/dmatrix_util.rs:766-784), adjust_by_division_of_selected_inplace
 *    (:178-205), log1p, CSC scale_columns_inplace (:786-830) and the transposed product with the basis.
 *    PARITY UNPINNED: the reference holds no golden vectors for it; checked against a float64 restatement.
 */
#include <cmath>
#include <cstdint>
#include <vector>

#include "oracle.h"

extern "C" {

/* sparse_stat.rs:64-80: for every stored FINITE value: npos += (v > 0), s1 += v, s2 += v*v, in column order, f32 */
void orc_row_stats(const uint64_t* indptr, const uint64_t* indices, const float* data, uint64_t nrows, uint64_t ncols,
                   float* npos, float* s1, float* s2) {
    for (uint64_t g = 0; g < nrows; ++g) npos[g] = s1[g] = s2[g] = 0.0f;
    for (uint64_t j = 0; j < ncols; ++j)
        for (uint64_t t = indptr[j]; t < indptr[j + 1]; ++t) {
            const float v = data[t];
            if (!std::isfinite(v)) continue;
            const uint64_t g = indices[t];
            if (v > 0.0f) npos[g] += 1.0f;
            s1[g] += v;
            s2[g] += v * v;
        }
}

/* sparse_stat.rs:412-431: mean = s1 / n, variance = s2 / n - mean^2, std = sqrt(variance); n = columns seen
 * (1e-8 when none, :16-23) */
void orc_row_stats_moments(const float* s1, const float* s2, uint64_t nrows, uint64_t ncols_processed, float* mean,
                           float* variance, float* sd) {
    const float n = ncols_processed > 0 ? (float)ncols_processed : 1e-8f;
    for (uint64_t g = 0; g < nrows; ++g) {
        const float mu = s1[g] / n;
        mean[g] = mu;
        variance[g] = s2[g] / n - mu * mu;
        sd[g] = std::sqrt(variance[g]);
    }
}

/* nystrom_proj_visitor (senna/src/svd/fit.rs:433-466) for every column, serial f32 folds as the reference makes them.
 * basis_dk: D x K column-major; delta_dp: D x P column-major or NULL; pb_of_cell: N (ignored without delta);
 * out: K x N column-major. */
void orc_nystrom_project(const uint64_t* indptr, const uint64_t* indices, const float* data, uint64_t nrows,
                         uint64_t ncols, const float* basis_dk, int K, const float* delta_dp, const uint32_t* pb_of_cell,
                         uint32_t P, float column_sum_norm, float* out_kn) {
    std::vector<float> x;
    for (uint64_t j = 0; j < ncols; ++j) {
        const uint64_t lo = indptr[j], hi = indptr[j + 1];
        const size_t n = (size_t)(hi - lo);
        x.assign(data + lo, data + hi);
        /* normalize_columns_inplace (dmatrix_util.rs:766-784) */
        float denom = 0.0f;
        for (size_t k = 0; k < n; ++k) denom += x[k] * x[k];
        denom = std::fmax(std::sqrt(denom), 1e-8f);
        for (size_t k = 0; k < n; ++k) x[k] /= denom;
        /* x_dn *= column_sum_norm (fit.rs:447) */
        for (size_t k = 0; k < n; ++k) x[k] *= column_sum_norm;
        /* adjust_by_division_of_selected_inplace -> adjust_by_poisson_ratio (dmatrix_util.rs:226-258) */
        if (delta_dp && pb_of_cell[j] < P) {
            const float* d = delta_dp + (size_t)pb_of_cell[j] * nrows;
            float dsum = 0.0f, xsum = 0.0f;
            for (size_t k = 0; k < n; ++k) dsum += d[indices[lo + k]];
            for (size_t k = 0; k < n; ++k) xsum += x[k];
            const float scale = dsum > 0.0f ? xsum / dsum : 1.0f;
            for (size_t k = 0; k < n; ++k) {
                const float dk = d[indices[lo + k]];
                if (dk > 0.0f) x[k] /= dk * scale;
            }
        }
        /* log1p_inplace (:629-633) */
        for (size_t k = 0; k < n; ++k) x[k] = std::log1p(x[k]);
        /* CSC scale_columns_inplace (:791-824): moments over the stored entries only */
        float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f;
        for (size_t k = 0; k < n; ++k) {
            s0 += 1.0f;
            s1 += x[k];
            s2 += x[k] * x[k];
        }
        const float mu = s1 / std::fmax(s0, 1.0f);
        const float sig = std::sqrt(s2 / std::fmax(s0, 1.0f) - mu * mu);
        if (sig > 0.0f)
            for (size_t k = 0; k < n; ++k) x[k] = (x[k] - mu) / sig;
        else
            for (size_t k = 0; k < n; ++k) x[k] -= mu;
        /* chunk = (x_dn^T * basis_dk)^T (fit.rs:457): row j of x^T against every basis column, entries ascending */
        for (int kk = 0; kk < K; ++kk) {
            const float* b = basis_dk + (size_t)kk * nrows;
            float acc = 0.0f;
            for (size_t k = 0; k < n; ++k) acc += x[k] * b[indices[lo + k]];
            out_kn[(size_t)j * K + kk] = acc;
        }
    }
}

}  // extern "C"
