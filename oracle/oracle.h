/*
 * oracle.h — C API of the CPU oracle for the legume-rs hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under legume-rs_b200/ may include, link or
 * dlopen this.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may call it, and only as the checker.
 *
 * The reference (causalpathlab/legume-rs v0.3.2, Rust) cannot be compiled in
 * this image (no cargo/rustc, no vendored crates), so this is a RESTATEMENT of
 * its arithmetic in C++.  Every function cites the reference file:line it
 * follows.  Layouts follow nalgebra: all dense matrices are column-major f32.
 *
 * Pinning status:
 *   - trigamma / log_sd           pinned by matrix-param/src/dmatrix_gamma_tests.rs:9-32
 *   - posterior mean (a0+Σy)/(b0+n) pinned by data-beans-alg/tests/weighted_columns.rs:76-121
 *   - gene-blocked == whole fit    pinned by collapse_data/stats_tests.rs:33-97 (property)
 *   - exact kNN == brute force     pinned by matrix-util/src/knn/tests.rs:74-150 (property)
 *   - pad/level bookkeeping        pinned by collapse_data/refine.rs doc examples
 *   - projection, codes, digamma:  PARITY UNPINNED — the reference's tests hold no
 *     golden vectors for them and the third-party arithmetic
 *     (nalgebra 0.34.2 QR/SVD, special 0.13.1 digamma) is not on disk; those are
 *     restated from their published algorithms.
 */
#ifndef LEGUME_ORACLE_H
#define LEGUME_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ---- stage 1: projection (random_projection.rs:169-199, 341-415) ---------- */
/* raw K×N projection: log1p -> L2 normalise -> ascending-row axpy, no centring */
void orc_project_raw(const uint64_t* indptr, const uint64_t* indices, const float* data,
                     uint64_t ncols, const float* basis_kd, int K, float* proj_kn, int nthreads);
/* batch centring + per-cell standardise + clamp + re-standardise, in place.
 * batch may be NULL (no centring). */
void orc_project_finish(float* proj_kn, int K, uint64_t ncols, const uint32_t* batch, uint32_t nbatch);

/* ---- stage 2: binary codes (random_projection.rs:535-564, dmatrix_rsvd.rs) - */
/* returns 0 on success.  scratch outputs may be NULL.
 * out_q (K×kk), out_b (kk×N), out_u (kk×kk, f32), out_sigma (kk), out_mean (kk). */
int orc_binary_codes(const float* proj_kn, int K, uint64_t ncols, int kk, uint64_t* codes,
                     float* out_q, float* out_u, float* out_sigma, float* out_mean);
/* the INDEPENDENT restatement (oracle_svd.cpp): the reference's own route — f32 Householder bidiagonalisation +
 * implicit-shift QR SVD of B, then scale_columns_inplace with f32 left folds and [V > 0].  Defines each bit's
 * partition of the cells up to complement (singular-vector signs are free).  out_v: N x kk column-major or NULL. */
int orc_binary_codes_svd(const float* proj_kn, int K, uint64_t ncols, int kk, uint64_t* codes, float* out_v,
                         float* out_sigma);
/* Householder QR of a K×r block, thin Q (K×r) — nalgebra qr().q() restated */
void orc_householder_q(const float* a_kr, int K, int r, float* q_kr);
/* cyclic Jacobi on a symmetric n×n f64 matrix: eigenvalues descending */
void orc_jacobi_eig(const double* g, int n, double* evals, double* evecs);

/* ---- stage 3: group ids (sparse_io_vector/groups.rs:13-37) ---------------- */
/* lexicographic rank of code.to_string(); returns number of groups */
uint32_t orc_assign_groups(const uint64_t* codes, uint64_t n, uint32_t* group_of_cell);
/* refine.rs:21-35 + groups.rs: zero-padded labels => numeric order */
uint32_t orc_assign_groups_padded(const uint64_t* labels, uint64_t n, uint64_t k, uint32_t* group_of_cell);
/* refine.rs:718-734 */
int orc_level_sort_dims(int sort_dim, int num_levels, int* out_dims);

/* ---- stage 4: collapse (collapse_data/stats.rs:110-164) ------------------- */
void orc_collapse_basic(const uint64_t* indptr, const uint64_t* indices, const float* data,
                        uint64_t nrows, uint64_t ncols, const uint32_t* group_of_cell,
                        const float* mult /*or NULL*/, uint32_t S, float* sum_ds, float* size_s);
void orc_collapse_batch(const uint64_t* indptr, const uint64_t* indices, const float* data,
                        uint64_t nrows, uint64_t ncols, const uint32_t* group_of_cell,
                        const uint32_t* batch_of_cell, const float* mult, uint32_t S, uint32_t B,
                        float* sum_db, float* n_bs);
/* stats.rs:790-833 */
void orc_merge_stat(const float* fine_ds, uint64_t nrows, uint32_t nfine, const uint32_t* fine_to_coarse,
                    uint32_t ncoarse, float* coarse_ds);

/* ---- stage 5: Poisson-Gamma posterior (dmatrix_gamma.rs, stats.rs:206-368) - */
float orc_digamma(float x);   /* special 0.13.1 (AS 103), f32 arithmetic */
float orc_trigamma(float x);  /* special 0.13.1 (AS 121), f32 arithmetic */
/* target: 0 = All, 1 = MeanOnly, 2 = MeanAndLogMean.  Output planes may be NULL. */
void orc_gamma_calibrate(const float* num, const float* den, uint64_t n, float a0, float b0, int target,
                         float* mean, float* sd, float* log_mean, float* log_sd);
/* single-batch optimize_block (B<=1 arm): denom = size_s broadcast, optional sparsify */
void orc_optimize_single(const float* sum_ds, const float* size_s, uint64_t D, uint32_t S, float a0, float b0,
                         int target, float* mean, float* sd, float* log_mean, float* log_sd);
/* the two arms with panel observability attached (stats.rs:176-204, 299-322): size_ds / mask_db may be NULL */
void orc_optimize_single_obs(const float* sum_ds, const float* size_s, const float* size_ds, uint64_t D, uint32_t S, float a0,
                             float b0, int target, float* mean, float* sd, float* log_mean, float* log_sd);
void orc_optimize_batched_obs(const float* obs_ds, const float* imp_ds, const float* res_ds, const float* size_s,
                              const float* size_ds, const float* obs_db, const float* n_bs, const float* mask_db, uint64_t D,
                              uint32_t S, uint32_t B, float a0, float b0, int num_iter, int target, float* mu_obs, float* mu_adj,
                              float* mu_res, float* gamma, float* delta, float* mu_adj_log_mean);
/* batched optimize_block (B>1 arm), means only + optional log planes of mu_adj.
 * outs: each D×S (delta D×B); any may be NULL. */
void orc_optimize_batched(const float* obs_ds, const float* imp_ds, const float* res_ds, const float* size_s,
                          const float* obs_db, const float* n_bs, uint64_t D, uint32_t S, uint32_t B,
                          float a0, float b0, int num_iter, int target,
                          float* mu_obs, float* mu_adj, float* mu_res, float* gamma, float* delta,
                          float* mu_adj_log_mean);

/* ---- stage 6: exact kNN (knn/metric.rs:19-45, exact.rs:36-55, mod.rs:249-299) */
float orc_l2_sq(const float* a, const float* b, int d);
/* ref d×nr, qry d×nq column-major; exclude[q] = reference index to drop or UINT32_MAX.
 * out_idx/out_dist are k×nq (nearest first, true Euclidean), padded with UINT32_MAX / inf. */
void orc_knn_topk(const float* ref, uint64_t nr, const float* qry, uint64_t nq, int d, int k,
                  const uint32_t* exclude, uint32_t* out_idx, float* out_dist, int nthreads);


/* ---- stage 7: cross-batch neighbourhood adjustment (oracle_adjust.cpp) ------
 * per-cell path: batch.rs:182-234, matched.rs:173-260, collapse_data/stats.rs:26-108 */
/* order: B x B, row b = all batches by distance from batch b's centroid (itself first);
 * out_centroids (B x K... stored K-contiguous per batch) may be NULL */
void orc_batch_proximity(const float* proj_kn, int K, uint64_t ncols, const uint32_t* batch_of_cell, uint32_t B,
                         uint32_t* order, float* out_centroids);
/* target_order: B x nt (row = source batch) or NULL (targets 0..nt-1); slots whose target is the
 * source batch (or >= B) stay empty.  out_idx/out_dist: (nt*knn) x ncols, global cell indices. */
void orc_knn_match_batches(const float* proj_kn, int K, uint64_t ncols, const uint32_t* batch_of_cell, uint32_t B,
                           int knn, const uint32_t* target_order, uint32_t nt, uint32_t* out_idx, float* out_dist,
                           int nthreads);
/* imputed_sum_ds / residual_sum_ds (D x S), overwritten */
void orc_collect_matched_stat(const uint64_t* indptr, const uint64_t* indices, const float* data, uint64_t nrows,
                              uint64_t ncols, const uint32_t* group_of_cell, uint32_t S, const uint32_t* matched_idx,
                              const float* matched_dist, uint32_t T, float* imputed_ds, float* residual_ds);
/* pb-sample path: collapse_data/pb_samples.rs:94-459, stats.rs:698-784 */
uint32_t orc_pb_layout(const float* proj_kn, int K, uint64_t ncols, const uint32_t* group_of_cell, uint32_t S,
                       const uint32_t* batch_of_cell, uint32_t B, const float* mult, uint32_t* cell_to_pb,
                       uint32_t* pb_group, uint32_t* pb_batch, float* pb_count, float* centroids);
void orc_pb_match(const float* proj_kn, int K, uint64_t ncols, const uint32_t* batch_of_cell, uint32_t B,
                  const uint32_t* cell_to_pb, const float* centroids, const uint32_t* pb_batch, uint32_t npb, int knn,
                  uint32_t* out_pb, float* out_dist, int nthreads);
void orc_collect_matched_stat_coarse(const float* gene_sums_dp, uint64_t nrows, uint32_t npb, const float* pb_count,
                                     const uint32_t* pb_to_group, uint32_t S, const uint32_t* matched_pb,
                                     const float* matched_dist, uint32_t T, float* imputed_ds, float* residual_ds);
/* refine.rs:741-769; returns the number of coarse groups */
uint32_t orc_fine_to_coarse(const uint64_t* group_code, uint32_t nfine, int coarse_dim, uint32_t* fine_to_coarse);

/* ---- BBKNN + DC-Poisson refinement of the pb-sample partition (oracle_refine.cpp; refine_multilevel.rs:170-298) ----
 * gene_sums: E x M dense (an entity's row contiguous); bbknn: CSR of matched entities; initial / offsets: num_levels x E
 * (finest first; offsets may be NULL); out_levels: num_levels x E; out_k: num_levels.  Returns the accepted moves. */
uint64_t orc_refine_assignments(const float* gene_sums, uint32_t E, uint64_t M, const uint32_t* bb_ptr, const uint32_t* bb,
                                int num_levels, const uint32_t* initial, const uint32_t* offsets, int num_gibbs, int num_greedy,
                                int fisher, uint64_t seed, double stagnation, uint32_t* out_levels, uint32_t* out_k);

/* ---- synthetic counts (data-beans-sim/src/core.rs:155-203 restated with a
 *      counter-based RNG so CPU and GPU produce identical matrices) ---------- */
/* table entry e = ((k*B + b)*D + g): lam[e], p0[e]=exp(-lam[e]) precomputed by the caller,
 * npiece[e] >= 1.  Returns nnz; if indices==NULL only counts (fills indptr). */
uint64_t orc_sim_poisson_csc(uint64_t seed, uint64_t D, uint64_t col_lo, uint64_t col_hi,
                             const uint8_t* topic_of_cell, const uint8_t* batch_of_cell, uint32_t ntopic,
                             uint32_t nbatch, const float* lam, const float* p0, const uint8_t* npiece,
                             uint64_t* indptr, uint64_t* indices, float* data);

/* ---- the steps either side of the path (oracle_next.cpp) ---------------------- */
/* SparseRunningStatistics::add_csc (matrix-util/src/sparse_stat.rs:64-108), f32, column order */
void orc_row_stats(const uint64_t* indptr, const uint64_t* indices, const float* data, uint64_t nrows, uint64_t ncols,
                   float* npos, float* s1, float* s2);
/* mean / variance / std (sparse_stat.rs:412-431) */
void orc_row_stats_moments(const float* s1, const float* s2, uint64_t nrows, uint64_t ncols_processed, float* mean,
                           float* variance, float* sd);
/* nystrom_proj_visitor (senna/src/svd/fit.rs:433-466) */
void orc_nystrom_project(const uint64_t* indptr, const uint64_t* indices, const float* data, uint64_t nrows,
                         uint64_t ncols, const float* basis_dk, int K, const float* delta_dp, const uint32_t* pb_of_cell,
                         uint32_t P, float column_sum_norm, float* out_kn);

/* ---- the CPU BASELINE leg (oracle_bench.cpp): the same arithmetic in the reference's execution structure ---------- */
uint64_t orc_default_block_size(uint64_t num_features);   /* matrix-util/src/utils.rs:86-94 */
/* visit_columns_by_block + project_columns_visitor: block = 0 selects default_block_size(D) */
void orc_bench_project_blocks(const uint64_t* indptr, const uint64_t* indices, const float* data, uint64_t D,
                              uint64_t ncols, const float* basis_kd, int K, uint64_t block, int nthreads, float* proj_kn);
/* visit_columns_by_group + collect_basic_stat_visitor; locked = 1 holds the global Mutex across the accumulate loop
 * (stats.rs:119), locked = 0 is the lock-free variant */
void orc_bench_collapse_groups(const uint64_t* indptr, const uint64_t* indices, const float* data, uint64_t D,
                               uint64_t ncols, const uint32_t* group_of_cell, uint32_t S, int locked, int nthreads,
                               float* sum_ds, float* size_s);
/* optimize: par_iter over gene blocks (stats.rs:462-478), B <= 1 arm */
void orc_bench_optimize_single_mt(const float* sum_ds, const float* size_s, uint64_t D, uint32_t S, float a0, float b0,
                                  int target, int nthreads, float* mean, float* sd, float* log_mean, float* log_sd);

#ifdef __cplusplus
}
#endif
#endif
