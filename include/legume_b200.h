/*
 * legume_b200.h — C ABI of liblegume_b200.so: the B200-native replacement for the
 * data-parallel hot path of causalpathlab/legume-rs (v0.3.2).
 *
 * The reference has no FFI of its own; its boundary for this path is the set of Rust
 * trait methods listed in SURVEY.md §8(b).  Each entry point below names the reference
 * interface it replaces (file:line relative to the legume-rs tree).  INTEGRATION.md shows
 * the `legume-b200-sys` Rust binding a maintainer would add.
 *
 * Conventions (all follow the reference's own):
 *   - every call returns an int status (LG_OK == 0); lg_last_error(ctx) gives the message.
 *     Nothing aborts or throws across the boundary (the reference returns anyhow::Result).
 *   - dense matrices are column-major f32 (nalgebra DMatrix).  "K x N" means each cell's
 *     K-vector is contiguous.
 *   - the caller owns every buffer passed in or out; inputs are borrowed for the call only.
 *   - EVERY data pointer may be host memory or device memory of ctx's GPU; the library
 *     detects which (cudaPointerGetAttributes) and stages host buffers itself.  Device
 *     pointers are used in place, on ctx's stream, with no synchronisation on return;
 *     when any argument of a call is a host pointer the call synchronises the stream
 *     before returning so host outputs are valid.
 *   - one in-flight call per ctx; one ctx per device; callable from any thread.
 *   - there is no CPU fallback: with no usable CUDA device lg_ctx_create fails.
 */
#ifndef LEGUME_B200_H
#define LEGUME_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LG_OK 0
#define LG_ERR_INVALID 1   /* bad argument (shape mismatch, null pointer, out-of-range index) */
#define LG_ERR_CUDA 2      /* a CUDA runtime call failed; message holds cudaGetErrorString */
#define LG_ERR_NOMEM 3     /* device or host allocation failed */
#define LG_ERR_INTERNAL 4

#define LG_BLOCK_CELLS 1024 /* granularity of the order-fixed reductions over cells */

/* matrix_param::traits::CalibrateTarget (matrix-param/src/traits.rs:31-39) */
#define LG_TARGET_ALL 0
#define LG_TARGET_MEAN_ONLY 1
#define LG_TARGET_MEAN_AND_LOG_MEAN 2

typedef struct lg_ctx lg_ctx;
typedef struct lg_csc lg_csc; /* device-resident CSC block: indptr u64, row index u32, value f32 */

/* ---- context -------------------------------------------------------------------------- */
int lg_ctx_create(int device, lg_ctx** out);
int lg_ctx_destroy(lg_ctx* ctx);
const char* lg_last_error(const lg_ctx* ctx);
/* run on an externally owned cudaStream_t (e.g. torch's current stream); NULL = the legacy
 * default stream.  A fresh ctx runs on its own non-blocking stream. */
int lg_ctx_set_stream(lg_ctx* ctx, void* cuda_stream);
int lg_ctx_sync(lg_ctx* ctx);
/* number of kernels this library launched through ctx since creation */
uint64_t lg_ctx_launch_count(const lg_ctx* ctx);
/* bytes lg_csc_upload has put on the host->device link through ctx since creation (host arrays are narrowed /
 * packed before they travel, so this is less than the size of the arrays handed in) */
uint64_t lg_ctx_h2d_bytes(const lg_ctx* ctx);
/* how many calls a tensor-core path (K1 projection, K7 kNN) declined for a reason of capability — K above 53 or more
 * than 131 072 genes or a non-finite basis for the projection; k above 12, d above 126 for the kNN filter — and handed to
 * the (several times slower) CUDA-core kernel, and the last reason.  Each distinct reason is also printed once on
 * stderr unless LG_QUIET=1.  Results are the same either way. */
uint64_t lg_ctx_fallback_count(const lg_ctx* ctx);
const char* lg_ctx_last_fallback(const lg_ctx* ctx);
/* how many collapses summed the 1-bit pattern the projection of the same lg_hotpath_run_sharded call left behind instead of
 * streaming the CSC arrays a second time (LG_COLLAPSE_PATTERN=0 turns that off; blocks with more than 32 768 genes, counts
 * that are not whole numbers below 32 768 or a projection outside the tensor path take the CSC kernel; same sums either way) */
uint64_t lg_ctx_pattern_collapse_count(const lg_ctx* ctx);
/* lg_ctx_time_stages(ctx, 1): lg_hotpath_run_sharded brackets its six stages (projection, codes, groups, collapse,
 * all-reduce, posterior) with events and synchronises at its end; lg_hotpath_last_stage_ms copies the last call's six device
 * times in milliseconds (LG_ERR_INVALID when there is none).  Diagnostic: off by default. */
void lg_ctx_time_stages(lg_ctx* ctx, int on);
int lg_hotpath_last_stage_ms(const lg_ctx* ctx, float* out6);
const char* lg_version(void);

/* ---- data feed --------------------------------------------------------------------------
 * replaces SparseIo::csc_column_arrays() -> (&[u64] indptr, &[u64] indices, &[f32] data)
 * (data-beans/src/sparse_io/traits.rs:98-100; zarr impl sparse_backend/zarr.rs:982-994) and the
 * per-backend row remap of SparseIoVec::read_columns_csc (sparse_io_vector/read.rs:202-219).
 * Uploads columns [col_lo, col_hi) of the host arrays, narrowing row indices to u32 after a
 * range check; row_remap (length = nrows_backend, or NULL) maps a backend row to a row of the
 * shared feature axis, UINT32_MAX = drop the entry.  indptr has (ncols_total + 1) entries. */
int lg_csc_upload(lg_ctx* ctx, const uint64_t* indptr, const uint64_t* indices, const float* data,
                  uint64_t nrows, uint64_t col_lo, uint64_t col_hi, const uint32_t* row_remap,
                  lg_csc** out);
/* the same with the remap's length stated (nrows_backend): a backend row outside it is LG_ERR_INVALID instead of
 * the caller's promise.  With a remap the block is made CANONICAL the way read_columns_csc does (read.rs:246-281):
 * entries whose row maps to UINT32_MAX are dropped; a column that is no longer strictly ascending is stably sorted
 * by row and equal rows are folded by summing in that order (so the block's nnz may be smaller than the input's).
 * Without a remap the arrays must already be canonical CSC (rows strictly ascending inside a column) — checked
 * on the device, LG_ERR_INVALID otherwise: the projection bitmap, the running statistics and the matched-column
 * kernels rely on it. */
int lg_csc_upload_remap(lg_ctx* ctx, const uint64_t* indptr, const uint64_t* indices, const float* data,
                        uint64_t nrows, uint64_t col_lo, uint64_t col_hi, const uint32_t* row_remap,
                        uint64_t nrows_backend, lg_csc** out);
/* columns of several device blocks side by side over one feature axis (SparseIoVec::push of several backends,
 * data-beans/src/sparse_io_vector/mod.rs); the parts stay valid and owned by the caller */
int lg_csc_concat(lg_ctx* ctx, const lg_csc* const* parts, uint32_t nparts, lg_csc** out);
/* wrap arrays that already live on the device (no copy; the caller keeps them alive).  Canonical form is checked
 * lazily, by the first call that relies on it. */
int lg_csc_wrap_device(lg_ctx* ctx, const uint64_t* d_indptr, const uint32_t* d_indices,
                       const float* d_values, uint64_t nrows, uint64_t ncols, uint64_t nnz,
                       lg_csc** out);
/* on != 0: the block gets buffers of its own (about 3.9 KB per cell + 2 bytes per non-zero) into which every projection of it
 * (lg_project / lg_project_raw on the tensor path) leaves the 1-bit sparsity pattern and the list of counts != 1 that its scan
 * builds anyway; lg_collapse_basic / lg_collapse_batch with unit multiplicities then sum that pattern instead of streaming the
 * arrays a second (third, ...) time — the multilevel collapse sums the same block once per level.  Same sums, bit for bit.
 * No effect for blocks the pattern kernel cannot take (more than 32 768 genes); a block whose counts are not whole numbers below
 * 32 768, or with more than half of a cell's entries != 1, keeps the CSC kernel.  on == 2: only if the buffers fit four times
 * into the free device memory and LG_KEEP_PATTERN is not 0 (what the SparseIoVec mirrors ask for).  on == 0 releases the buffers
 * (lg_csc_free does too).  lg_hotpath_run_sharded does this on its own for the length of one call. */
int lg_csc_keep_pattern(lg_ctx* ctx, lg_csc* m, int on);
int lg_csc_free(lg_ctx* ctx, lg_csc* m);
int lg_csc_shape(const lg_csc* m, uint64_t* nrows, uint64_t* ncols, uint64_t* nnz);
/* raw device pointers of a block (for torch interop / tests) */
int lg_csc_device_arrays(const lg_csc* m, const uint64_t** d_indptr, const uint32_t** d_indices,
                         const float** d_values);
/* copy a block back to host arrays in the reference's u64/u64/f32 form */
int lg_csc_download(lg_ctx* ctx, const lg_csc* m, uint64_t* indptr, uint64_t* indices, float* data);

/* ---- stage 1: random projection -----------------------------------------------------------
 * replaces RandProjOps::project_columns_with_batch_correction_seeded and
 * project_columns_weighted_seeded (data-beans-alg/src/random_projection.rs:341-495).  The basis
 * is an INPUT (identical-projection-matrix contract): basis_kd is K x D column-major, i.e. the
 * reference's `basis_dk.transpose()` (:360); for the weighted variant the caller zeroes/scales
 * rows exactly as :438-444 do.  batch_of_cell: u32[ncols] in [0, nbatch) or NULL; a label outside that
 * range is LG_ERR_INVALID.
 * out_proj: K x ncols. */
int lg_project(lg_ctx* ctx, const lg_csc* m, const float* basis_kd, int K,
               const uint32_t* batch_of_cell, uint32_t nbatch, float* out_proj);
/* staged form, for cell-sharded runs (device pointers only):
 *   raw        project_columns_visitor                         (:169-199)
 *   partials   per-1024-cell-block f64 sums of proj per (batch, dim) and cell counts:
 *              out[blk][b*(K+1) + k], k == K holds the count   (:378-388)
 *   finalize   sums block partials in block order -> out[M] (f64); shards all-gather their
 *              partials in rank order first, so the result does not depend on the GPU count
 *   centre_scale  subtract batch means, per-cell standardise, track global (min,max) (:399-401)
 *   clamp_rescale clamp to [-4,4] and standardise again         (:402-407) */
int lg_project_raw(lg_ctx* ctx, const lg_csc* m, const float* basis_kd, int K, float* out_proj);
int lg_proj_batch_partials(lg_ctx* ctx, const float* d_proj, int K, uint64_t ncols,
                           const uint32_t* d_batch, uint32_t nbatch, double* d_partials);
int lg_block_partials_finalize(lg_ctx* ctx, const double* d_partials, uint64_t nblocks, uint32_t M,
                               double* d_out);
int lg_proj_centre_scale(lg_ctx* ctx, float* d_proj, int K, uint64_t ncols, const uint32_t* d_batch,
                         uint32_t nbatch, const double* d_batch_sums, float* d_minmax);
int lg_proj_clamp_rescale(lg_ctx* ctx, float* d_proj, int K, uint64_t ncols);
/* the same, decided ON THE DEVICE: d_minmax holds the global (min, max); a no-op when both lie inside [-4, 4] */
int lg_proj_clamp_rescale_if(lg_ctx* ctx, float* d_proj, int K, uint64_t ncols, const float* d_minmax);
/* EXACT-ORDER mode of the same stage: the reference's arithmetic operation by operation (ascending row, x divided by the
 * norm first, product and sum rounded separately — random_projection.rs:181-194, dmatrix_util.rs:770-778; batch means
 * as f32 left folds over the batch's cells in ascending order — random_projection.rs:380-387), so that for count data
 * (whole numbers below 65536, whose ln_1p comes from a libm-built table) the projection is BIT-IDENTICAL to the CPU
 * path and so are the codes, groups and sums derived from it.  Several times slower than lg_project (CUDA cores,
 * one basis row gathered per non-zero); lg_project stays the throughput path with a 1e-5 contract.
 *   lg_project_exact        composite, same arguments as lg_project
 *   lg_project_raw_exact    project_columns_visitor only
 *   lg_proj_batch_fold      continues the per-(batch, dim) f32 folds sum[nbatch*K] / cnt[nbatch] over this block's cells
 *                           (device pointers, zero them first); cell shards call it one after the other in rank order,
 *                           handing sum / cnt on, so the folds are those of one GPU
 *   lg_proj_centre_scale_exact   subtract -(sum / (f32)cnt), per-cell standardise, track (min, max) */
int lg_project_exact(lg_ctx* ctx, const lg_csc* m, const float* basis_kd, int K,
                     const uint32_t* batch_of_cell, uint32_t nbatch, float* out_proj);
int lg_project_raw_exact(lg_ctx* ctx, const lg_csc* m, const float* basis_kd, int K, float* out_proj);
int lg_proj_batch_fold(lg_ctx* ctx, const float* d_proj, int K, uint64_t ncols, const uint32_t* d_batch,
                       uint32_t nbatch, float* d_sum, uint64_t* d_cnt);
int lg_proj_centre_scale_exact(lg_ctx* ctx, float* d_proj, int K, uint64_t ncols, const uint32_t* d_batch,
                               uint32_t nbatch, const float* d_fold_sum, const uint64_t* d_fold_cnt, float* d_minmax);

/* ---- stage 2: binary codes -----------------------------------------------------------------
 * replaces binary_sort_columns (random_projection.rs:535-564) = rsvd (matrix-util/src/
 * dmatrix_rsvd.rs:85-180) + per-dimension standardise + sign bits.  codes: u64[ncols] < 2^kk. */
int lg_binary_codes(lg_ctx* ctx, const float* proj_kn, int K, uint64_t ncols, int kk, uint64_t* out_codes);
/* staged form (device pointers; basis and factor also take host pointers):
 *   basis      one-warp kernel: Q (K x kk) = first kk columns of qr(X[:, 0..r]).q(), r = min(kk+5, N) <= 21;
 *              first_cols_kr = the K-vectors of the first r cells (the head of the projection itself)
 *   gram       B = Q^T X (kk x ncols) and block partials of the upper triangle of B B^T
 *              (M = kk(kk+1)/2, entry (a,b), a<=b, at a*kk - a(a-1)/2 + (b-a))
 *   factor     one-warp kernel: cyclic Jacobi (f64) on the Gram sums -> U (kk x kk f32), sigma (kk), sign-fixed
 *   vproj      V = B^T U / sigma (kk x ncols) and block partials of its column sums (M = kk)
 *   means      mean[k] = (f32)(sums[k] / ncols_total)
 *   pack       warp-ballot sign packer: bit k of code_j = [V[k,j] > mean_k]
 * No stage reads anything back to the host: the whole of K3 is queued on the stream. */
int lg_codes_basis(lg_ctx* ctx, const float* first_cols_kr, int K, int r, int kk, float* out_q);
int lg_codes_gram(lg_ctx* ctx, const float* d_proj, int K, uint64_t ncols, const float* d_q, int kk,
                  float* d_b, double* d_partials);
int lg_codes_factor(lg_ctx* ctx, const double* gram_sums, const float* q, int K, int kk, float* out_u,
                    float* out_sigma);
int lg_codes_vproj(lg_ctx* ctx, const float* d_b, int kk, uint64_t ncols, const float* d_u,
                   const float* d_sigma, float* d_v, double* d_partials);
int lg_codes_means(lg_ctx* ctx, const double* d_sums, int kk, uint64_t ncols_total, float* d_mean);
int lg_codes_pack(lg_ctx* ctx, const float* d_v, int kk, uint64_t ncols, const float* d_mean,
                  uint64_t* d_codes);

/* ---- stage 3: group ids --------------------------------------------------------------------
 * replaces SparseIoVec::assign_groups (data-beans/src/sparse_io_vector/groups.rs:13-37): groups
 * ordered by the byte-wise order of code.to_string().  padded != 0 selects the refine path's
 * zero-padded labels (collapse_data/refine.rs:21-35, 393-399) = numeric order. */
int lg_assign_groups(lg_ctx* ctx, const uint64_t* codes, uint64_t ncols, int kk, int padded,
                     uint32_t* out_group_of_cell, uint32_t* out_num_groups);
/* staged: presence flags (u32[2^kk], OR-reducible across shards), host LUT, device map */
int lg_code_presence(lg_ctx* ctx, const uint64_t* d_codes, uint64_t ncols, int kk, uint32_t* d_present);
int lg_group_lut(lg_ctx* ctx, const uint32_t* present, int kk, int padded, uint32_t* out_lut,
                 uint32_t* out_num_groups);
int lg_codes_to_groups(lg_ctx* ctx, const uint64_t* d_codes, uint64_t ncols, int kk, const uint32_t* d_lut,
                       uint32_t* d_group_of_cell);

/* ---- stage 4: collapse ---------------------------------------------------------------------
 * replaces CollapsingOps::collect_basic_stat / collect_batch_stat
 * (data-beans-alg/src/collapse_data/mod.rs:477-483, stats.rs:110-164).
 *   sum_ds[g,s] += y*w, size_s[s] += w;  sum_db[g,b] += y*w, n_bs[b,s] += w.
 * mult = column multiplicity (sparse_io_vector/batch.rs:331-336) or NULL (all ones).
 * Outputs are OVERWRITTEN (zeroed first).  group ids >= S are skipped.
 * Sums of integer-valued counts are exact and order-independent below 2^24. */
int lg_collapse_basic(lg_ctx* ctx, const lg_csc* m, const uint32_t* group_of_cell, const float* mult,
                      uint32_t S, float* out_sum_ds, float* out_size_s);
int lg_collapse_batch(lg_ctx* ctx, const lg_csc* m, const uint32_t* group_of_cell,
                      const uint32_t* batch_of_cell, const float* mult, uint32_t S, uint32_t B,
                      float* out_sum_db, float* out_n_bs);
/* merge_stat (stats.rs:790-833): coarse[:, f2c[f]] += fine[:, f] */
int lg_merge_stat(lg_ctx* ctx, const float* fine_ds, uint64_t nrows, uint32_t nfine,
                  const uint32_t* fine_to_coarse, uint32_t ncoarse, float* out_coarse_ds);

/* ---- stage 5: Poisson-Gamma posterior --------------------------------------------------------
 * replaces GammaMatrix::update_stat + calibrate_with (matrix-param/src/dmatrix_gamma.rs:64-123,
 * traits.rs:61-77) and optimize/optimize_block (collapse_data/stats.rs:206-512).
 * a = a0 + num, b = b0 + den; mean = a/b; sd = sqrt(a)/b; log_mean = digamma(a) - ln(b);
 * log_sd = sqrt(trigamma(a)).  Output planes not selected by target, or NULL, are not written. */
int lg_gamma_calibrate(lg_ctx* ctx, const float* num, const float* den, uint64_t n, float a0, float b0,
                       int target, float* mean, float* sd, float* log_mean, float* log_sd);
/* optimize_block, B <= 1 arm (stats.rs:351-368): den[g,s] = size_s[s]; MeanOnly sparsifies */
int lg_optimize_single(lg_ctx* ctx, const float* sum_ds, const float* size_s, uint64_t D, uint32_t S,
                       float a0, float b0, int target, float* mean, float* sd, float* log_mean,
                       float* log_sd);
/* optimize_block, B > 1 arm (stats.rs:219-350): num_iter sweeps kept on-device.
 * Outputs are posterior means (delta is D x B); mu_adj_log_mean optional (target All / MeanAndLogMean). */
int lg_optimize_batched(lg_ctx* ctx, const float* obs_ds, const float* imp_ds, const float* res_ds,
                        const float* size_s, const float* obs_db, const float* n_bs, uint64_t D,
                        uint32_t S, uint32_t B, float a0, float b0, int num_iter, int target,
                        float* mu_obs, float* mu_adj, float* mu_res, float* gamma, float* delta,
                        float* mu_adj_log_mean);

/* optimize_block with panel observability attached (stats.rs:176-204, 299-322): size_ds (D x S, optional) replaces the
 * per-sample size in every denominator (add_effective_size / scale_by_effective_size), obs_mask_db (D x B of 0 / 1,
 * optional) multiplies both sides of the delta ratio.  With both NULL these are lg_optimize_single / _batched. */
int lg_optimize_single_obs(lg_ctx* ctx, const float* sum_ds, const float* size_s, const float* size_ds, uint64_t D,
                           uint32_t S, float a0, float b0, int target, float* mean, float* sd, float* log_mean,
                           float* log_sd);
int lg_optimize_batched_obs(lg_ctx* ctx, const float* obs_ds, const float* imp_ds, const float* res_ds,
                            const float* size_s, const float* size_ds, const float* obs_db, const float* n_bs,
                            const float* obs_mask_db, uint64_t D, uint32_t S, uint32_t B, float a0, float b0,
                            int num_iter, int target, float* mu_obs, float* mu_adj, float* mu_res, float* gamma,
                            float* delta, float* mu_adj_log_mean);
/* attach_observability (collapse_data/mod.rs:221-301): coverage is nsrc x D (1 = the backend measures the gene),
 * source_of_cell the backend every column came from.  size_ds[g, s] = multiplicity-weighted mass of sample s from the
 * backends that cover g; mask_db[g, b] = 1 iff some backend used by batch b covers g (optional; *out_mask_has_zero
 * tells whether any entry is 0 — the reference keeps the mask only then). */
int lg_attach_observability(lg_ctx* ctx, const uint8_t* coverage, uint32_t nsrc, const uint32_t* source_of_cell,
                            const uint32_t* group_of_cell, const uint32_t* batch_of_cell, const float* mult,
                            uint64_t ncols, uint64_t D, uint32_t S, uint32_t B, float* out_size_ds,
                            float* out_mask_db, int* out_mask_has_zero);

/* ---- stage 6: exact kNN ----------------------------------------------------------------------
 * replaces ColumnDict::search_by_query_data / match_by_query_name_against / search_others with the
 * exact backend (matrix-util/src/knn/mod.rs:152-299, exact.rs:36-55, metric.rs:19-45).
 * ref: d x nr, qry: d x nq (column-major).  exclude: u32[nq] reference index to drop per query
 * (UINT32_MAX = none) or NULL.  out_idx/out_dist: k x nq, nearest first, true Euclidean distance;
 * unused slots hold UINT32_MAX / +inf.  Ranking is by the reference's squared-distance arithmetic;
 * equal distances are ordered by lower index. */
int lg_knn_topk(lg_ctx* ctx, const float* ref, uint64_t nr, const float* qry, uint64_t nq, int d, int k,
                const uint32_t* exclude, uint32_t* out_idx, float* out_dist);

/* the same search leaving SQUARED distances (what the ranking is made on): the per-shard half of a search
 * whose reference cells are sharded over GPUs */
int lg_knn_topk_sq(lg_ctx* ctx, const float* ref, uint64_t nr, const float* qry, uint64_t nq, int d, int k,
                   const uint32_t* exclude, uint32_t* out_idx, float* out_sq);
/* top-k merge across reference-cell shards: shard_idx / shard_sq are nshard x nq x k lists from
 * lg_knn_topk_sq (local indices); shard s holds the reference cells shard_offset[s] ... of the global numbering.
 * Merged by (squared distance, lower global index), exactly the order of a single search over all cells;
 * exclude (u32[nq], global index, or NULL) is applied here, so shards must be searched with k+1 when it is used. */
int lg_knn_merge_topk(lg_ctx* ctx, const uint32_t* shard_idx, const float* shard_sq, uint32_t nshard, uint64_t nq, int k,
                      const uint64_t* shard_offset, const uint32_t* exclude, uint32_t* out_idx, float* out_dist);

/* ---- stage 7: cross-batch neighbourhood adjustment ---------------------------------------------
 * (a) per-cell path = CollapsingOps::collapse_columns with B > 1 (collapse_data/mod.rs:384-475):
 *     batch dictionaries + sort_batch_proximity (data-beans/src/sparse_io_vector/batch.rs:100-234),
 *     read_neighbouring_columns_csc / read_matched_columns_csc (sparse_io_vector/matched.rs:173-474),
 *     collect_matched_stat_visitor (collapse_data/stats.rs:26-108).
 * (b) pb-sample path = collapse_columns_multilevel_vec with B >= 2 (collapse_data/mod.rs:867-1050):
 *     build_pb_sample_layout / per_batch_sc_neighbors (collapse_data/pb_samples.rs:94-459),
 *     collect_matched_stat_coarse (collapse_data/stats.rs:698-784),
 *     compute_fine_to_coarse_mapping (collapse_data/refine.rs:741-769).                          */

/* sort_batch_proximity: out_order is B x B (row b = every batch by distance from batch b's centroid,
 * b itself first); out_centroids (K x B, may be NULL) = DMatrix::column_mean of each batch's cells */
int lg_batch_proximity(lg_ctx* ctx, const float* proj_kn, int K, uint64_t ncols, const uint32_t* batch_of_cell,
                       uint32_t B, uint32_t* out_order, float* out_centroids);
/* neighbouring_columns_triplets, kNN part: for source cell j of batch s and slot i < nt the target batch
 * is target_order[s*nt + i] (NULL: nt = B, targets 0..B-1).  Slots whose target is s itself or >= B stay
 * empty (skip_same_batch).  out_idx / out_dist: (nt*knn) x ncols; entry i*knn + r = r-th nearest cell of the
 * target batch as a GLOBAL cell index / Euclidean distance; UINT32_MAX / +inf when absent.  The
 * reference's knn_batches argument only sizes a Vec (matched.rs:201) and is not part of the result. */
int lg_knn_match_batches(lg_ctx* ctx, const float* proj_kn, int K, uint64_t ncols, const uint32_t* batch_of_cell,
                         uint32_t B, int knn, const uint32_t* target_order, uint32_t nt, uint32_t* out_idx,
                         float* out_dist);
/* collect_matched_stat_visitor over every group: W = softmax(-d) per source cell, y_hat = Y_matched W,
 * y1 <- y1 / (y_hat * sum(y1)/sum(y_hat)) where y_hat > 0; imputed[:, s] += y_hat, residual[:, s] += y1.
 * matched_idx / matched_dist: T x ncols as written by lg_knn_match_batches.  Outputs (D x S) are overwritten.
 * Sums are accumulated in a fixed order (cells ascending inside a group): bit-identical run to run. */
int lg_collect_matched_stat(lg_ctx* ctx, const lg_csc* m, const uint32_t* group_of_cell, uint32_t S,
                            const uint32_t* matched_idx, const float* matched_dist, uint32_t T,
                            float* out_imputed_ds, float* out_residual_ds);

/* build_pb_sample_layout: pb-sample = non-empty (group, batch) block, numbered group-major with batches
 * ascending inside a group.  Outputs: cell_to_pb u32[ncols]; pb_group / pb_batch / pb_count: capacity S*B;
 * centroids K x (S*B) (mean of the block's K-vectors, summed in ascending cell order); *out_num_pb.
 * mult = column multiplicity or NULL.  anchor / bulk batches are not supported. */
int lg_pb_layout(lg_ctx* ctx, const float* proj_kn, int K, uint64_t ncols, const uint32_t* group_of_cell, uint32_t S,
                 const uint32_t* batch_of_cell, uint32_t B, const float* mult, uint32_t* out_cell_to_pb,
                 uint32_t* out_pb_group, uint32_t* out_pb_batch, float* out_pb_count, float* out_centroids,
                 uint32_t* out_num_pb);
/* per_batch_sc_neighbors (pooled matching): for pb-sample p and batch b != pb_batch[p], the knn nearest
 * DISTINCT foreign pb-samples of batch b (distance of a pb-sample = its closest cell to p's centroid, which is
 * what the adaptive 4k+1, x4 search of pb_samples.rs:337-363 converges to).  out: (B*knn) x npb. */
int lg_pb_match(lg_ctx* ctx, const float* proj_kn, int K, uint64_t ncols, const uint32_t* batch_of_cell, uint32_t B,
                const uint32_t* cell_to_pb, const float* centroids, const uint32_t* pb_batch, uint32_t npb, int knn,
                uint32_t* out_matched_pb, float* out_matched_dist);
/* collect_matched_stat_coarse.  gene_sums: D x npb dense (= lg_collapse_basic with cell_to_pb as the label).
 * pb_to_group: layout.pb_group or a refined assignment.  Outputs (D x S) are overwritten. */
int lg_collect_matched_stat_coarse(lg_ctx* ctx, const float* gene_sums, uint64_t D, uint32_t npb, const float* pb_count,
                                   const uint32_t* pb_to_group, uint32_t S, const uint32_t* matched_pb,
                                   const float* matched_dist, uint32_t T, float* out_imputed_ds,
                                   float* out_residual_ds);
/* staged form of the pb-sample arm for cell-sharded runs (device pointers unless noted; exchanges are the caller's):
 *   pair_presence    flags[g*B + b] = 1 where a cell of group g and batch b exists      -> all-reduce(max)
 *   pb_ids           HOST math: ids of the present (group, batch) blocks, group-major, batches ascending
 *   cells_to_pb      cell -> pb-sample id
 *   centroid_fold    continues the serial folds sum (K x npb) / count (npb) over this shard's cells; shards run it
 *                    one after the other in rank order, handing sum / count on, so the centroids are those of one GPU
 *   centroid_finish  centroid = sum * (1 / count)
 *   min_keys         keys[(q - q0)*npb + p] = min over this shard's cells of pb-sample p of (l2_sq << 32 | global cell)
 *                    for the query centroids q0 .. q0+nq                                 -> all-reduce(min) as int64
 *   topk_keys        per query and batch the knn smallest keys -> (pb-sample, distance) lists               */
int lg_pair_presence(lg_ctx* ctx, const uint32_t* d_group, const uint32_t* d_batch, uint64_t ncols, uint32_t S, uint32_t B,
                     uint32_t* d_present);
int lg_pb_ids(lg_ctx* ctx, const uint32_t* present, uint32_t S, uint32_t B, uint32_t* out_id, uint32_t* out_pb_group,
              uint32_t* out_pb_batch, uint32_t* out_num_pb);
int lg_cells_to_pb(lg_ctx* ctx, const uint32_t* d_group, const uint32_t* d_batch, uint64_t ncols, uint32_t S, uint32_t B,
                   const uint32_t* d_id, uint32_t* d_cell_to_pb);
int lg_pb_centroid_fold(lg_ctx* ctx, const float* d_proj, int K, uint64_t ncols, const uint32_t* d_cell_to_pb, uint32_t npb,
                        const float* d_mult, float* d_sum, float* d_count);
int lg_pb_centroid_finish(lg_ctx* ctx, const float* d_sum, const float* d_count, uint32_t npb, int K, float* d_centroids);
int lg_pb_min_keys(lg_ctx* ctx, const float* d_proj, int K, uint64_t ncols, const uint32_t* d_cell_to_pb, uint64_t cell_offset,
                   const float* d_centroids, const uint32_t* d_pb_batch, uint32_t npb, uint32_t q0, uint32_t nq, uint64_t* d_keys);
int lg_pb_topk_keys(lg_ctx* ctx, const uint64_t* d_keys, uint32_t npb, uint32_t q0, uint32_t nq, uint32_t B,
                    const uint32_t* pb_batch, int knn, uint32_t* d_out_pb, float* d_out_dist);
/* compute_fine_to_coarse_mapping: coarse code = fine code & (2^coarse_dim - 1), ids by sorted unique code */
int lg_fine_to_coarse(lg_ctx* ctx, const uint64_t* codes, const uint32_t* group_of_cell, uint64_t ncols, uint32_t nfine,
                      int coarse_dim, uint32_t* out_fine_to_coarse, uint32_t* out_num_coarse);

/* ---- synthetic counts (benchmark input; data-beans-sim/src/core.rs:155-203) --------------------
 * y[g,j] ~ Poisson(lam[(topic_j*nbatch + batch_j)*D + g]) summed over npiece pieces, kept if > 0.5.
 * topic/batch arrays cover [col_lo, col_hi).  Counter-based RNG: identical to the oracle's twin. */
int lg_sim_poisson_csc(lg_ctx* ctx, uint64_t seed, uint64_t D, uint64_t col_lo, uint64_t col_hi,
                       const uint8_t* topic_of_cell, const uint8_t* batch_of_cell, uint32_t ntopic,
                       uint32_t nbatch, const float* lam, const float* p0, const uint8_t* npiece,
                       lg_csc** out);

/* ---- the nnz streams either side of the path ---------------------------------
 * lg_row_stats replaces SparseRunningStatistics::add_csc over a block (matrix-util/src/sparse_stat.rs:64-108) as
 * driven by streaming_sparse_running_stats (data-beans-alg/src/sparse_streaming.rs:23-60; callers hvg.rs:389,
 * gene_weighting.rs:113): per gene, over the FINITE stored values v: npos = #(v > 0), s1 = sum v, s2 = sum v*v.
 * Outputs are f64 (D each, overwritten): for count data they are exact integers, so per-shard results add up to
 * the same totals on any number of GPUs; the caller narrows to the reference's f32 after merging (merge = add,
 * sparse_stat.rs:183-196).  ncols_processed is the block's column count. */
int lg_row_stats(lg_ctx* ctx, const lg_csc* m, double* out_npos, double* out_s1, double* out_s2);
/* lg_nystrom_project replaces nystrom_proj_visitor over a block (senna/src/svd/fit.rs:433-466):
 *   x = y / max(||y||_2, 1e-8) * column_sum_norm;  x /= delta[row, pb(j)] * (sum x / sum delta) where delta > 0
 *   (adjust_by_poisson_ratio, matrix-util/src/dmatrix_util.rs:226-244);  z = ln(1 + x), standardised over the cell's
 *   stored entries (:791-824);  out[:, j] = sum_i z_i basis[i, :].
 * basis_dk: D x K column-major (the reference's DMatrix); delta_dp: D x P column-major or NULL (then pb_of_cell is
 * ignored); pb_of_cell: u32[N], values >= P leave the cell unadjusted; out: K x N column-major. */
int lg_nystrom_project(lg_ctx* ctx, const lg_csc* m, const float* basis_dk, int K, const float* delta_dp,
                       const uint32_t* pb_of_cell, uint32_t P, float column_sum_norm, float* out_proj_kn);

/* ---- BBKNN + DC-Poisson refinement of the pb-sample partition --------------------------
 * What `MultilevelParams::refine = Some(..)` runs between the hash partition and the collapse when there are two or more
 * batches (collapse_data/refine.rs:264-345 -> refine_multilevel.rs:170-298 -> dc_poisson.rs:778-915).  The entities are the
 * pb-samples, their profiles the dense pb-sample x gene sums the path already holds (npb x D, an entity's row contiguous;
 * a stored entry is a value > 0, the filter of Profiles::from_gene_sums, dc_poisson.rs:136-160).  The label bookkeeping
 * around the levels (compact_labels, project_to_refinement, sibling / candidate sets; dc_poisson.rs:493-633,
 * refine_multilevel.rs:85-112, 315-345) is host code in the mirrors; these entry points are the numeric part.
 *   lg_dcp_fisher_weights  Profiles::nb_fisher_weights (dc_poisson.rs:230-295; nb_dispersion.rs:58-151): per-gene serial f32
 *                          folds over the entities on the device, the D-long trend fit on the host.
 *   lg_dcp_profiles        weight_by_vec (:197-213) in place + every entity's size factor (serial f32 fold).
 *   lg_dcp_refine_level    refine_with_candidates_guarded (:778-915) for one level with RefineParams::parallel = true (Jacobi
 *                          sweeps) and no move guard: num_gibbs Gumbel-max sweeps (per-entity SmallRng streams from
 *                          jacobi_base_seed, early exit after three sweeps below `stagnation` * npb moves), then num_greedy
 *                          arg-max sweeps (exit on a sweep without moves).  Scores are accumulated in the reference's order
 *                          (one f64 accumulator per (entity, group) over ascending genes).  Candidates: CSR, groups ascending
 *                          inside an entity.  labels: in / out, values < k.  Pointers may be host or device memory. */
int lg_dcp_fisher_weights(lg_ctx* ctx, const float* profiles, uint64_t D, uint32_t npb, float* out_w);
int lg_dcp_profiles(lg_ctx* ctx, float* profiles, uint64_t D, uint32_t npb, const float* weights, float* out_size_factor);
int lg_dcp_refine_level(lg_ctx* ctx, const float* profiles, const float* size_factor, uint64_t D, uint32_t npb,
                        const uint32_t* cand_ptr, const uint32_t* cand, uint32_t k, int num_gibbs, int num_greedy,
                        uint64_t jacobi_base_seed, double stagnation, uint32_t* labels, uint64_t* out_moves);

/* ---- column-block ingest: the reference's zarr backend as the feed of the path ---------
 * lg_zarr_* replaces, for a matrix written by the reference's zarr backend (data-beans/src/sparse_backend/zarr.rs: a Zarr V3
 * FilesystemStore directory holding /by_column/{indptr u64, indices u64, data f32} as 1-D arrays in ~1 MiB chunks,
 * bytes(little) + zstd level 5, root attributes nrow / ncol / nnz; zarr.rs:31-64, 285-310, 515-523;
 * utilities/io_helpers.rs:105-115):
 *   SparseMtxData::open + num_rows / num_columns / num_non_zeros    zarr.rs:640-690, 515-523   -> lg_zarr_open, lg_zarr_shape
 *   preload_columns + csc_column_arrays over a column range            zarr.rs:573-587, 982-994   -> lg_zarr_read_columns_host
 *   read_columns_csc (the block handed to the visitors)                sparse_io_vector/read.rs:172-285 -> lg_zarr_read_columns
 * Chunks are inflated on the host cores (libzstd bound at run time with dlopen; LG_INGEST_THREADS, default all cores up to
 * 32) and the block goes through lg_csc_upload.  A path ending in `.zip` is opened as an archive of such a
 * directory (zarr_io.rs:30-85: entries under `<stem>/`, `<stem>.zarr/` or bare; stored or deflated; ZIP64) and read in place.  The hdf5 twin and
 * the /by_row copy are not read.
 * lg_zarr_open reports through `err` (it has no context yet); the other calls through lg_zarr_last_error. */
typedef struct lg_zarr lg_zarr;
int lg_zarr_open(const char* path, lg_zarr** out, char* err, size_t err_len);
void lg_zarr_close(lg_zarr* z);
const char* lg_zarr_last_error(const lg_zarr* z);
int lg_zarr_shape(const lg_zarr* z, uint64_t* nrows, uint64_t* ncols, uint64_t* nnz);
/* entries of columns [col_lo, col_hi) are [*first, *last) of the indices / data arrays */
int lg_zarr_column_extent(lg_zarr* z, uint64_t col_lo, uint64_t col_hi, uint64_t* first, uint64_t* last);
/* host arrays of the range: indptr (col_hi - col_lo + 1 entries, rebased to 0), indices and data (*last - *first each) */
int lg_zarr_read_columns_host(lg_zarr* z, uint64_t col_lo, uint64_t col_hi, uint64_t* indptr, uint64_t* indices, float* data);
/* the range as a device-resident block (rows = the store's nrow) */
int lg_zarr_read_columns(lg_ctx* ctx, lg_zarr* z, uint64_t col_lo, uint64_t col_hi, lg_csc** out);

/* ---- multi-GPU: cells sharded over the GPUs of one box, one process (or thread) per GPU ------------------------
 * (section 8b row 4 `lg_allreduce_stats`).  The reference is a single process; what these entry
 * points replace is the rayon reduction over column blocks inside project_columns / collect_basic_stat
 * (random_projection.rs:341-415, collapse_data/stats.rs:110-164) once the blocks live on different GPUs.
 * NCCL is bound at run time (dlopen libnccl.so.2).  Bootstrap: rank 0 calls lg_comm_unique_id, the host ships the
 * 128 bytes to every rank by its own means, every rank calls lg_comm_init on its context.  world == 1 needs no id and
 * makes every exchange a no-op. */
#define LG_COMM_ID_BYTES 128
int lg_comm_unique_id(lg_ctx* ctx, void* out_id_128_bytes);
int lg_comm_init(lg_ctx* ctx, const void* id_128_bytes, int rank, int world);
int lg_comm_info(lg_ctx* ctx, int* rank, int* world);
int lg_comm_destroy(lg_ctx* ctx);
/* in-place all-reduce(sum) over the ranks of the collapse statistics (device pointers, any may be NULL):
 * sum_ds D x S, size_s S, sum_db D x B, n_bs B x S.  Exact for count data in any order (sums below 2^24). */
int lg_allreduce_stats(lg_ctx* ctx, float* d_sum_ds, float* d_size_s, float* d_sum_db, float* d_n_bs, uint64_t D,
                       uint32_t S, uint32_t B);
/* The single-batch arm of the whole path on this rank's shard `m` (every shard but the last a multiple of 1024 cells,
 * rank 0 holding at least kk + 5): projection + batch centring -> binary codes -> groups -> collapse + all-reduce ->
 * posterior.  All pointers are device memory: d_proj K x ncols, d_codes / d_group per cell, d_sum_ds (and the optional
 * posterior planes) with room for D x 2^kk, d_size_s for 2^kk; *out_num_groups the number of groups found.  Results
 * are bit-identical for any number of ranks (order-sensitive sums travel as per-block partials in global order). */
int lg_hotpath_run_sharded(lg_ctx* ctx, const lg_csc* m, const float* d_basis_kd, int K, const uint32_t* d_batch,
                           uint32_t nbatch, int kk, int target, float* d_proj, uint64_t* d_codes, uint32_t* d_group,
                           uint32_t* out_num_groups, float* d_sum_ds, float* d_size_s, float* d_mean, float* d_sd,
                           float* d_log_mean, float* d_log_sd);

#ifdef __cplusplus
}
#endif
#endif /* LEGUME_B200_H */
