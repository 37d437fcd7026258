/*
 * oracle_bench.cpp — the CPU BASELINE leg of bench.py: the reference's arithmetic (oracle.cpp) in the reference's own
 * execution STRUCTURE, on all host threads.  TEST / MEASUREMENT INFRASTRUCTURE ONLY (oracle.h).
 *
 * The Rust workspace cannot be built in this image, so this is a port ("kind": "port"), arranged the way the
 * reference runs the path so that the timing means something:
 *   projection   visit_columns_by_block (data-beans/src/sparse_data_visitors.rs:60-88): rayon over column blocks of
 *                default_block_size(D) = clamp(1e6 / D, 100, 10 000) cells (matrix-util/src/utils.rs:86-94); per block
 *                read_columns_csc repacks the columns into per-column (row, value) buckets and assembles a block CSC
 *                (sparse_io_vector/read.rs:195-281), then ln_1p / normalise / ascending-row axpy into a block-local
 *                K x n chunk (random_projection.rs:169-194) which is copied into the shared output under the Mutex (:196)
 *   collapse     visit_columns_by_group (:93-124): rayon over groups; each job repacks its group's columns, then takes
 *                the GLOBAL Mutex and holds it across the whole accumulate loop (collapse_data/stats.rs:117-132).
 *                locked = 0 is the lock-free variant (every group owns its own output column), reported beside it.
 *   posterior    optimize (collapse_data/stats.rs:378-512): par_iter over gene blocks, optimize_block each
 * Results are bit-identical to oracle.cpp's serial functions (checked in tests/test_oracle_pinned.py).
 */
#include <algorithm>
#include <cmath>
#include <cstring>
#include <mutex>
#include <utility>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "oracle.h"

namespace {
struct BlockCsc {
    std::vector<size_t> offsets;
    std::vector<size_t> rows;
    std::vector<float> vals;
};
// read_columns_csc fast path (read.rs:205-219 buckets, :246-281 assembly); single backend, identity row map
void repack(const uint64_t* indptr, const uint64_t* indices, const float* data, const uint64_t* cells, size_t ncell,
            uint64_t first_cell, BlockCsc* out) {
    std::vector<std::vector<std::pair<uint32_t, float>>> buckets(ncell);
    for (size_t c = 0; c < ncell; ++c) {
        const uint64_t j = cells ? cells[c] : first_cell + c;
        const uint64_t s = indptr[j], e = indptr[j + 1];
        auto& b = buckets[c];
        b.reserve(e - s);
        for (uint64_t t = s; t < e; ++t) b.emplace_back((uint32_t)indices[t], data[t]);
    }
    size_t total = 0;
    for (auto& b : buckets) total += b.size();
    out->offsets.clear();
    out->rows.clear();
    out->vals.clear();
    out->offsets.reserve(ncell + 1);
    out->rows.reserve(total);
    out->vals.reserve(total);
    out->offsets.push_back(0);
    for (auto& b : buckets) {
        bool sorted = true;
        for (size_t i = 1; i < b.size(); ++i) sorted &= b[i - 1].first < b[i].first;
        if (!sorted) std::stable_sort(b.begin(), b.end(), [](auto& x, auto& y) { return x.first < y.first; });
        for (auto& rv : b) {
            out->rows.push_back(rv.first);
            out->vals.push_back(rv.second);
        }
        out->offsets.push_back(out->rows.size());
    }
}

__attribute__((target_clones("avx2", "default"))) void axpy_k(float a, const float* __restrict__ x, float* __restrict__ y, int K) {
    for (int k = 0; k < K; ++k) {
        const float p = a * x[k];  // product rounded, then the sum (-ffp-contract=off)
        y[k] = p + y[k];
    }
}
}  // namespace

extern "C" uint64_t orc_default_block_size(uint64_t num_features) {
    const uint64_t lo = 100, hi = 10000, target = 100 * 10000;
    if (num_features == 0) return lo;
    return std::min(std::max(target / num_features, lo), hi);
}

extern "C" void orc_bench_project_blocks(const uint64_t* indptr, const uint64_t* indices, const float* data, uint64_t D,
                                         uint64_t ncols, const float* basis_kd, int K, uint64_t block, int nthreads,
                                         float* proj_kn) {
    if (block == 0) block = orc_default_block_size(D);
    const int64_t njobs = (int64_t)((ncols + block - 1) / block);
    std::mutex out_lock;
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#endif
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads > 0 ? nthreads : 1)
    for (int64_t job = 0; job < njobs; ++job) {
        const uint64_t lb = (uint64_t)job * block, ub = std::min(ncols, lb + block);
        BlockCsc m;
        repack(indptr, indices, data, nullptr, ub - lb, lb, &m);
        for (auto& v : m.vals) v = log1pf(v);                                    // :181-183
        for (size_t c = 0; c + 1 < m.offsets.size(); ++c) {                      // normalize_columns_inplace
            float denom = 0.0f;
            for (size_t t = m.offsets[c]; t < m.offsets[c + 1]; ++t) denom += m.vals[t] * m.vals[t];
            denom = std::max(std::sqrt(denom), (float)1e-8);
            for (size_t t = m.offsets[c]; t < m.offsets[c + 1]; ++t) m.vals[t] /= denom;
        }
        std::vector<float> chunk((size_t)K * (ub - lb), 0.0f);
        for (size_t c = 0; c + 1 < m.offsets.size(); ++c) {                      // :188-194
            float* y = chunk.data() + c * K;
            for (size_t t = m.offsets[c]; t < m.offsets[c + 1]; ++t) axpy_k(m.vals[t], basis_kd + m.rows[t] * (size_t)K, y, K);
        }
        std::lock_guard<std::mutex> g(out_lock);                                  // :196
        std::memcpy(proj_kn + lb * (size_t)K, chunk.data(), chunk.size() * sizeof(float));
    }
}

extern "C" void orc_bench_collapse_groups(const uint64_t* indptr, const uint64_t* indices, const float* data, uint64_t D,
                                          uint64_t ncols, const uint32_t* group_of_cell, uint32_t S, int locked, int nthreads,
                                          float* sum_ds, float* size_s) {
    std::memset(sum_ds, 0, sizeof(float) * (size_t)D * S);
    std::memset(size_s, 0, sizeof(float) * S);
    // take_grouped_columns: cells ascending inside a group (groups.rs:26-33)
    std::vector<std::vector<uint64_t>> cols(S);
    for (uint64_t j = 0; j < ncols; ++j)
        if (group_of_cell[j] < S) cols[group_of_cell[j]].push_back(j);
    std::mutex stat_lock;
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#endif
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads > 0 ? nthreads : 1)
    for (int64_t s = 0; s < (int64_t)S; ++s) {
        BlockCsc m;
        repack(indptr, indices, data, cols[s].data(), cols[s].size(), 0, &m);
        auto accumulate = [&]() {
            float* col = sum_ds + (size_t)s * D;
            for (size_t c = 0; c + 1 < m.offsets.size(); ++c) {
                for (size_t t = m.offsets[c]; t < m.offsets[c + 1]; ++t) col[m.rows[t]] += m.vals[t] * 1.0f;
                size_s[s] += 1.0f;
            }
        };
        if (locked) {
            std::lock_guard<std::mutex> g(stat_lock);  // stats.rs:119: held across the whole loop
            accumulate();
        } else {
            accumulate();
        }
    }
}

/* optimize (stats.rs:462-478): gene blocks in parallel; the B <= 1 arm of optimize_block on each */
extern "C" void orc_bench_optimize_single_mt(const float* sum_ds, const float* size_s, uint64_t D, uint32_t S, float a0, float b0,
                                             int target, int nthreads, float* mean, float* sd, float* log_mean, float* log_sd) {
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#endif
    // the arithmetic is element-wise, so a gene block is a strided set of elements of every group's column: cut the
    // D x S plane into (group, gene range) tiles and hand each to orc_gamma_calibrate-style arithmetic
    const uint64_t gblock = 2048;
    const int64_t nb = (int64_t)((D + gblock - 1) / gblock);
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads > 0 ? nthreads : 1) collapse(2)
    for (int64_t blk = 0; blk < nb; ++blk)
        for (int64_t s = 0; s < (int64_t)S; ++s) {
            const uint64_t g0 = (uint64_t)blk * gblock, g1 = std::min(D, g0 + gblock);
            const size_t o = (size_t)s * D + g0;
            orc_optimize_single(sum_ds + o, size_s + s, g1 - g0, 1, a0, b0, target, mean + o, sd ? sd + o : nullptr,
                                log_mean ? log_mean + o : nullptr, log_sd ? log_sd + o : nullptr);
        }
}
