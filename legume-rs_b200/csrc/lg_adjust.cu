// lg_adjust.cu — stage 7: cross-batch neighbourhood adjustment (SURVEY.md §8a rows a14–a17).
//
//   per-cell path (CollapsingOps::collapse_columns, B > 1; collapse_data/mod.rs:384-475)
//     sort_batch_proximity            data-beans/src/sparse_io_vector/batch.rs:182-234
//     neighbouring_columns_triplets   data-beans/src/sparse_io_vector/matched.rs:173-260
//     collect_matched_stat_visitor    data-beans-alg/src/collapse_data/stats.rs:26-108
//   pb-sample path (collapse_columns_multilevel_vec, B >= 2; collapse_data/mod.rs:867-1050)
//     build_pb_sample_layout          collapse_data/pb_samples.rs:94-219
//     per_batch_sc_neighbors          collapse_data/pb_samples.rs:323-459
//     collect_matched_stat_coarse     collapse_data/stats.rs:698-784
//     compute_fine_to_coarse_mapping  collapse_data/refine.rs:741-769
//
// Everything that decides an INDEX (centroids, distances, rankings) uses the reference's exact f32
// arithmetic in the reference's order, so neighbour sets are bit-exact.  The weighted sums are
// accumulated in a fixed order (no floating-point atomics), so results are bit-identical run to run;
// they differ from the reference only through expf (last-bit) and stay inside the 1e-5 contract.
#include <cub/cub.cuh>

#include <algorithm>
#include <cmath>
#include <numeric>

#include <cstdlib>

#include "lg_common.cuh"
#include "lg_umma.cuh"

int lg_knn_topk_device(lg_ctx* ctx, const float* d_ref, uint64_t nr, const float* d_qry, uint64_t nq, int d, int k,
                       const uint32_t* d_ex, uint32_t* d_idx, float* d_dist, int squared);

namespace {

constexpr uint32_t NONE = 0xFFFFFFFFu;

// ------------------------------------------------------------------------------------------------
// helpers shared by the host wrappers
// ------------------------------------------------------------------------------------------------
__global__ void k_iota(uint32_t* p, uint64_t n) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = (uint32_t)i;
}
__global__ void k_fill_u32(uint32_t* p, uint64_t n, uint32_t v) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
__global__ void k_fill_f32(float* p, uint64_t n, float v) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
__global__ void k_label_counts(const uint32_t* __restrict__ label, uint64_t n, uint32_t L, unsigned int* __restrict__ counts) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && label[i] < L) atomicAdd(&counts[label[i]], 1u);
}

// stable sort of cell ids by a u32 label: cells stay ascending inside a label (labels >= L sort last)
int sort_cells_by_label(lg_ctx* ctx, LgStage& st, const uint32_t* d_label, uint64_t N, uint32_t** d_lab_sorted,
                        uint32_t** d_cell_sorted) {
    uint32_t* d_in;
    LG_TRY(st.scratch(N, &d_in));
    LG_TRY(st.scratch(N, d_lab_sorted));
    LG_TRY(st.scratch(N, d_cell_sorted));
    if (N == 0) return LG_OK;
    LG_LAUNCH(ctx, k_iota, (unsigned)((N + 255) / 256), 256, 0, d_in, N);
    size_t tmp_bytes = 0;
    LG_CUDA(ctx, cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_label, *d_lab_sorted, d_in, *d_cell_sorted, (int)N, 0, 32,
                                                 ctx->stream));
    char* d_tmp;
    LG_TRY(st.scratch(tmp_bytes, &d_tmp));
    LG_CUDA(ctx, cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, d_label, *d_lab_sorted, d_in, *d_cell_sorted, (int)N, 0, 32,
                                                 ctx->stream));
    ctx->launches += 4;
    return LG_OK;
}

// per-label counts on the host (one synchronisation)
int label_counts_host(lg_ctx* ctx, LgStage& st, const uint32_t* d_label, uint64_t N, uint32_t L, std::vector<uint32_t>& counts) {
    counts.assign(L, 0);
    if (L == 0 || N == 0) return LG_OK;
    unsigned int* d_counts;
    LG_TRY(st.scratch(L, &d_counts));
    LG_CUDA(ctx, cudaMemsetAsync(d_counts, 0, sizeof(unsigned int) * L, ctx->stream));
    LG_LAUNCH(ctx, k_label_counts, (unsigned)((N + 255) / 256), 256, 0, d_label, N, L, d_counts);
    LG_CUDA(ctx, cudaMemcpyAsync(counts.data(), d_counts, sizeof(uint32_t) * L, cudaMemcpyDeviceToHost, ctx->stream));
    LG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return LG_OK;
}

// small caller array (host or device) -> host vector
template <typename T>
int to_host(lg_ctx* ctx, const T* p, size_t n, std::vector<T>& out) {
    out.resize(n);
    if (n == 0) return LG_OK;
    if (lg_is_device_ptr(p)) {
        LG_CUDA(ctx, cudaMemcpyAsync(out.data(), p, sizeof(T) * n, cudaMemcpyDeviceToHost, ctx->stream));
        LG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    } else {
        std::copy(p, p + n, out.begin());
    }
    return LG_OK;
}
template <typename T>
int upload(lg_ctx* ctx, LgStage& st, const std::vector<T>& h, T** d) {
    LG_TRY(st.scratch(h.size(), d));
    if (!h.empty()) {
        // pageable source: the copy is staged by the runtime before the call returns
        LG_CUDA(ctx, cudaMemcpyAsync(*d, h.data(), sizeof(T) * h.size(), cudaMemcpyHostToDevice, ctx->stream));
        LG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return LG_OK;
}

// knn/metric.rs:19-45 — exact f32 arithmetic of the reference's distance
__device__ __forceinline__ float adj_l2_sq(const float* __restrict__ r, const float* __restrict__ q, int d) {
    float acc[16];
#pragma unroll
    for (int l = 0; l < 16; ++l) acc[l] = 0.0f;
    int c = 0;
    for (; c + 16 <= d; c += 16) {
#pragma unroll
        for (int l = 0; l < 16; ++l) {
            const float df = __fsub_rn(r[c + l], q[c + l]);
            acc[l] = __fadd_rn(acc[l], __fmul_rn(df, df));
        }
    }
    float sum = 0.0f;
#pragma unroll
    for (int l = 0; l < 16; ++l) sum = __fadd_rn(sum, acc[l]);
    for (; c < d; ++c) {
        const float df = __fsub_rn(r[c], q[c]);
        sum = __fadd_rn(sum, __fmul_rn(df, df));
    }
    return sum;
}

// ------------------------------------------------------------------------------------------------
// segment means in the reference's order: one block per segment, one thread per dimension, cells walked
// in ascending order with the loads unrolled ahead of the (serial) adds
//   MODE 0: nalgebra column_mean     acc = (1/n) * x + acc                 (batch.rs:196-205)
//   MODE 1: pb-sample centroid       acc += x * w ; then acc * (1/count)   (pb_samples.rs:141-160)
//   MODE 2: the same fold CONTINUED from running sums / counts (cell shards hand the fold on in rank order,
//           so the result is the single serial fold whatever the GPU count); no division
// ------------------------------------------------------------------------------------------------
template <int MODE>
__global__ void k_segment_mean(const float* __restrict__ proj, int K, const uint32_t* __restrict__ cell_sorted,
                               const uint64_t* __restrict__ seg_off, const float* __restrict__ mult, float* __restrict__ out_mean,
                               float* __restrict__ out_count) {
    const uint32_t s = blockIdx.x;
    const int k = threadIdx.x;
    const uint64_t lo = seg_off[s], hi = seg_off[s + 1];
    float cnt = (MODE == 2) ? out_count[s] : 0.0f;
    if (MODE >= 1) {
        // the count is a serial f32 sum of the multiplicities (every thread computes the same value)
        for (uint64_t i = lo; i < hi; ++i) cnt = __fadd_rn(cnt, mult ? mult[cell_sorted[i]] : 1.0f);
    }
    if (MODE == 2) __syncthreads();  // every thread has read the running count before thread 0 replaces it
    if (k >= K) return;
    const float denom = (MODE == 0) ? __fdiv_rn(1.0f, (float)(double)(hi - lo)) : 0.0f;
    float acc = (MODE == 2) ? out_mean[(size_t)s * K + k] : 0.0f;
    uint64_t i = lo;
    for (; i + 8 <= hi; i += 8) {
        float x[8], w[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const uint32_t c = cell_sorted[i + u];
            x[u] = proj[(size_t)c * K + k];
            w[u] = (MODE >= 1 && mult) ? mult[c] : 1.0f;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u)
            acc = (MODE == 0) ? __fadd_rn(__fmul_rn(denom, x[u]), acc) : __fadd_rn(acc, __fmul_rn(x[u], w[u]));
    }
    for (; i < hi; ++i) {
        const uint32_t c = cell_sorted[i];
        const float x = proj[(size_t)c * K + k];
        const float w = (MODE >= 1 && mult) ? mult[c] : 1.0f;
        acc = (MODE == 0) ? __fadd_rn(__fmul_rn(denom, x), acc) : __fadd_rn(acc, __fmul_rn(x, w));
    }
    if (MODE == 1) {
        const float inv = __fdiv_rn(1.0f, cnt);
        acc = __fmul_rn(acc, inv);
    }
    if (MODE >= 1 && k == 0 && out_count) out_count[s] = cnt;
    out_mean[(size_t)s * K + k] = acc;
}

__global__ void k_centroid_finish(const float* __restrict__ sum, const float* __restrict__ count, uint64_t n, int K,
                                  float* __restrict__ cen) {
    const uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n * (uint64_t)K) return;
    cen[e] = __fmul_rn(sum[e], __fdiv_rn(1.0f, count[e / K]));  // pb_samples.rs:171-175
}

__global__ void k_gather_rows(const float* __restrict__ src, int K, const uint32_t* __restrict__ rows, uint64_t n,
                              float* __restrict__ dst) {
    const uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n * (uint64_t)K) return;
    const uint64_t r = e / K;
    dst[e] = src[(size_t)rows[r] * K + (e % K)];
}

// kNN results of one (target batch, query range) -> global indices in the caller's slot layout
__global__ void k_scatter_matches(const uint32_t* __restrict__ knn_idx, const float* __restrict__ knn_dist, uint64_t nq, int knn,
                                  uint64_t q0, uint64_t ref0, const uint32_t* __restrict__ cell_sorted,
                                  const uint32_t* __restrict__ batch_sorted, const uint32_t* __restrict__ slot_of, uint32_t B,
                                  uint32_t b, uint32_t T, uint32_t* __restrict__ out_idx, float* __restrict__ out_dist) {
    const uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nq * (uint64_t)knn) return;
    const uint64_t q = e / knn;
    const int r = (int)(e % knn);
    const uint32_t s = batch_sorted[q0 + q];
    if (s >= B) return;
    const uint32_t slot = slot_of[(size_t)s * B + b];
    if (slot == NONE) return;
    const uint32_t c = cell_sorted[q0 + q];
    const uint32_t local = knn_idx[e];
    const size_t o = (size_t)c * T + (size_t)slot * knn + r;
    out_idx[o] = local == NONE ? NONE : cell_sorted[ref0 + local];
    out_dist[o] = knn_dist[e];
}

// ------------------------------------------------------------------------------------------------
// collect_matched_stat_visitor
// ------------------------------------------------------------------------------------------------
constexpr int MS_NW = 8;           // worker warps: each owns one gene sub-range of the CTA's range
constexpr int MS_THREADS = (MS_NW + 1) * 32;  // + one warp that prepares the next cell's descriptors
constexpr int MS_PF = 6;           // matched columns in flight per warp (cp.async ring: 2 x 16 B per lane per stage)
constexpr size_t MS_RING_BYTES = (size_t)MS_NW * MS_PF * 2 * 32 * 16;
constexpr int MS_SEG = 128;        // source cells per work item
constexpr int MS_MAXR = 16;        // gene ranges
constexpr int MS_PASS = 32 * 4;    // nnz per warp and pass

// per cell: sum of the column (one warp per cell)
__global__ void k_cell_colsum(const uint64_t* __restrict__ indptr, const float* __restrict__ values, uint64_t ncols,
                              float* __restrict__ colsum) {
    const int lane = threadIdx.x & 31;
    const uint64_t j = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (j >= ncols) return;
    const uint64_t lo = indptr[j], hi = indptr[j + 1];
    float s = 0.0f;
    for (uint64_t t = lo + lane; t < hi; t += 32) s += values[t];
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (lane == 0) colsum[j] = s;
}

// split[j][b] = number of entries of column j with row < b * wsub, b = 0 .. nb (rows are sorted: a binary search)
__global__ void k_cell_splits(const uint64_t* __restrict__ indptr, const uint32_t* __restrict__ indices, uint64_t ncols, uint32_t wsub,
                              uint32_t nb, uint32_t* __restrict__ split) {
    const uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= ncols * (uint64_t)(nb + 1)) return;
    const uint64_t j = e / (nb + 1);
    const uint32_t b = (uint32_t)(e % (nb + 1));
    const uint64_t lo = indptr[j], hi = indptr[j + 1];
    const uint64_t bound = (uint64_t)b * wsub;
    uint64_t a = lo, z = hi;
    while (a < z) {
        const uint64_t mid = (a + z) >> 1;
        if ((uint64_t)indices[mid] < bound) a = mid + 1;
        else z = mid;
    }
    split[e] = (uint32_t)(a - lo);
}
__global__ void k_max_part(const uint32_t* __restrict__ split, uint64_t ncols, uint32_t nb, unsigned int* __restrict__ max_part) {
    const uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= ncols * (uint64_t)nb) return;
    const uint64_t j = e / nb;
    const uint32_t b = (uint32_t)(e % nb);
    const unsigned int n = split[j * (nb + 1) + b + 1] - split[j * (nb + 1) + b];
    unsigned int m = n;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, off));
    if ((threadIdx.x & 31) == 0) atomicMax(max_part, m);
}

struct MatchW {
    uint32_t m;  // matched cell (>= ncols: empty slot)
    float w;     // softmax weight
};

// per source cell (one warp): softmax(-d) over its matched columns (dmatrix_util.rs:649-671: the MIN logit is
// subtracted) and the division scale sum(y1) / sum(y_hat) (dmatrix_util.rs:145-176)
__global__ void k_match_weights(const float* __restrict__ colsum, uint64_t ncols, const uint32_t* __restrict__ midx,
                                const float* __restrict__ mdist, uint32_t T, MatchW* __restrict__ mw, float* __restrict__ scale) {
    const int lane = threadIdx.x & 31;
    const uint64_t j = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (j >= ncols) return;
    const uint32_t* mi = midx + j * T;
    const float* md = mdist + j * T;
    float lmin = INFINITY;
    for (uint32_t t = lane; t < T; t += 32)
        if (mi[t] < ncols) lmin = fminf(lmin, -md[t]);
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) lmin = fminf(lmin, __shfl_xor_sync(0xffffffffu, lmin, off));
    float denom = 0.0f;
    for (uint32_t t = lane; t < T; t += 32)
        if (mi[t] < ncols) denom += expf(-md[t] - lmin);
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) denom += __shfl_xor_sync(0xffffffffu, denom, off);
    float dsum = 0.0f;
    for (uint32_t t = lane; t < T; t += 32) {
        const uint32_t m = mi[t];
        const bool ok = m < ncols;
        MatchW o;
        o.m = ok ? m : NONE;
        o.w = ok ? __fdiv_rn(expf(-md[t] - lmin), denom) : 0.0f;
        if (ok) dsum += o.w * colsum[m];
        mw[j * T + t] = o;
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) dsum += __shfl_xor_sync(0xffffffffu, dsum, off);
    if (lane == 0) scale[j] = dsum > 0.0f ? __fdiv_rn(colsum[j], dsum) : 1.0f;
}

__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src, bool on) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst_smem);
    const int bytes = on ? 16 : 0;  // 0 source bytes = zero fill, nothing is read
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// what one worker warp needs for one matched column (or for the cell's own column): its sub-range of the column
struct SubSeg {
    unsigned long long lo;  // first nnz inside this warp's gene sub-range
    uint32_t n;             // nnz inside it
    float w;                // softmax weight (own column: the division scale)
};

// grid = (segments, gene ranges).  The CTA walks its source cells in ascending order.  The gene range is cut into
// MS_NW sub-ranges, one per worker warp: rows are sorted inside a column, so a warp's share of any column is one
// contiguous piece, and because no other warp ever touches its genes the warp streams the matched columns in order
// with plain shared-memory read-modify-writes — fixed accumulation order, no CTA barrier, no atomics.  Each warp
// keeps MS_PF columns in flight through its own cp.async ring; a ninth warp resolves the NEXT cell's pieces
// (indptr + split look-ups) into a double-buffered table while the workers are busy.
__global__ void __launch_bounds__(MS_THREADS, 1) k_matched_stat(
    const uint64_t* __restrict__ indptr, const uint32_t* __restrict__ indices, const float* __restrict__ values, uint64_t nnz_total,
    const uint32_t* __restrict__ split, const uint32_t* __restrict__ cell_sorted, const uint32_t* __restrict__ seg_first,
    const uint32_t* __restrict__ seg_len, const uint32_t* __restrict__ seg_group, const uint8_t* __restrict__ seg_single,
    const MatchW* __restrict__ mw, const float* __restrict__ scale, uint64_t ncols, uint32_t T, int R, uint64_t D, uint32_t W,
    uint32_t wsub, uint32_t pmax, float* __restrict__ scratch, float* __restrict__ out_imp, float* __restrict__ out_res) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint4* ring = reinterpret_cast<uint4*>(smem_raw);                               // [MS_NW][MS_PF][2][32]
    SubSeg* tab = reinterpret_cast<SubSeg*>(smem_raw + MS_RING_BYTES);              // [2][T + 1][MS_NW]; entry T = own column
    float* imp = reinterpret_cast<float*>(tab + (size_t)2 * (T + 1) * MS_NW);
    float* res = imp + W;
    float* yhat = res + W;                                                          // [MS_NW][pmax]
    float* pat_v = yhat + (size_t)MS_NW * pmax;
    uint32_t* pat_g = reinterpret_cast<uint32_t*>(pat_v + (size_t)MS_NW * pmax);
    unsigned short* slot = reinterpret_cast<unsigned short*>(pat_g + (size_t)MS_NW * pmax);
    __shared__ uint64_t bar_full[2], bar_empty[2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t seg = blockIdx.x;
    const int r = blockIdx.y;
    const uint32_t g0 = (uint32_t)r * W;
    const uint32_t Wr = g0 < D ? (uint32_t)min((uint64_t)W, D - g0) : 0u;
    const uint32_t nb = (uint32_t)R * MS_NW;
    for (uint32_t g = tid; g < W; g += MS_THREADS) {
        imp[g] = 0.0f;
        res[g] = 0.0f;
        slot[g] = 0xFFFF;
    }
    if (tid == 0) {
        for (int b = 0; b < 2; ++b) {
            umma::mbar_init(&bar_full[b], 1);
            umma::mbar_init(&bar_empty[b], MS_NW);
        }
        umma::fence_barrier_init();
    }
    __syncthreads();
    const uint32_t p0 = seg_first[seg], np_cells = seg_len[seg];

    if (warp == MS_NW) {
        // ===== descriptor warp: one cell ahead of the workers =====
        for (uint32_t p = 0; p < np_cells; ++p) {
            const uint32_t buf = p & 1;
            if (p >= 2) umma::mbar_wait(&bar_empty[buf], ((p >> 1) - 1) & 1);
            const uint32_t j = cell_sorted[p0 + p];
            SubSeg* tb = tab + (size_t)buf * (T + 1) * MS_NW;
            for (uint32_t t = lane; t <= T; t += 32) {
                uint32_t m;
                float w;
                if (t < T) {
                    const MatchW x = mw[(uint64_t)j * T + t];
                    m = x.m;
                    w = x.w;
                } else {
                    m = j;
                    w = scale[j];
                }
                const bool ok = m < ncols;
                const unsigned long long base = ok ? indptr[m] : 0ull;
                const uint32_t* sp = split + (uint64_t)(ok ? m : 0) * (nb + 1) + (uint32_t)r * MS_NW;
                uint32_t prev = ok ? sp[0] : 0u;
#pragma unroll
                for (int w8 = 0; w8 < MS_NW; ++w8) {
                    const uint32_t nxt = ok ? sp[w8 + 1] : 0u;
                    SubSeg s;
                    s.lo = base + prev;
                    s.n = nxt - prev;
                    s.w = w;
                    tb[(size_t)t * MS_NW + w8] = s;
                    prev = nxt;
                }
            }
            __syncwarp();
            if (lane == 0) umma::mbar_arrive(&bar_full[buf]);
        }
    } else {
        // ===== worker warp: owns genes [g0 + warp * wsub, g0 + (warp + 1) * wsub) =====
        uint4* myring = ring + (size_t)warp * MS_PF * 2 * 32;
        float* my_yhat = yhat + (size_t)warp * pmax;
        float* my_pv = pat_v + (size_t)warp * pmax;
        uint32_t* my_pg = pat_g + (size_t)warp * pmax;
        auto fetch = [&](unsigned long long c, uint4& g, float4& v) {  // synchronous: long pieces and the array tail
            if (c + 4 <= nnz_total) {
                g = __ldg(reinterpret_cast<const uint4*>(indices + c));
                v = __ldg(reinterpret_cast<const float4*>(values + c));
            } else {
                uint32_t gg[4] = {0, 0, 0, 0};
                float vv[4] = {0.f, 0.f, 0.f, 0.f};
                for (int u = 0; u < 4; ++u)
                    if (c + u < nnz_total) {
                        gg[u] = indices[c + u];
                        vv[u] = values[c + u];
                    }
                g = make_uint4(gg[0], gg[1], gg[2], gg[3]);
                v = make_float4(vv[0], vv[1], vv[2], vv[3]);
            }
        };
        auto apply4 = [&](unsigned long long c, unsigned long long lo, uint32_t n, float w, const uint4& g, const float4& v) {
            const uint32_t e0 = (uint32_t)(c - lo);  // position of the group's first entry in the piece (wraps below 0: masked)
            const uint32_t gi[4] = {g.x - g0, g.y - g0, g.z - g0, g.w - g0};
            const float vi[4] = {v.x, v.y, v.z, v.w};
            bool ok[4];
            float a[4], term[4];
            unsigned short sl[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                ok[u] = (e0 + (uint32_t)u) < n;
                a[u] = ok[u] ? imp[gi[u]] : 0.0f;  // rows are distinct inside a column: the four updates never alias
                sl[u] = ok[u] ? slot[gi[u]] : (unsigned short)0xFFFF;
                term[u] = __fmul_rn(w, vi[u]);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (ok[u]) imp[gi[u]] = __fadd_rn(a[u], term[u]);
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (sl[u] != 0xFFFF) my_yhat[sl[u]] = __fadd_rn(my_yhat[sl[u]], term[u]);
        };
        for (uint32_t p = 0; p < np_cells; ++p) {
            const uint32_t buf = p & 1;
            umma::mbar_wait(&bar_full[buf], (p >> 1) & 1);
            const SubSeg* tb = tab + (size_t)buf * (T + 1) * MS_NW + warp;  // stride MS_NW between columns
            auto issue = [&](uint32_t t, int s) {
                if (t < T) {
                    const SubSeg d = tb[(size_t)t * MS_NW];
                    const unsigned long long c = (d.lo & ~3ull) + 4ull * lane;
                    const bool on = (c < d.lo + d.n) && (c + 4 <= nnz_total);
                    const unsigned long long cs = on ? c : 0ull;
                    cp_async16(&myring[(s * 2 + 0) * 32 + lane], indices + cs, on);
                    cp_async16(&myring[(s * 2 + 1) * 32 + lane], values + cs, on);
                }
                cp_async_commit();
            };
#pragma unroll
            for (int s = 0; s < MS_PF; ++s) issue(s, s);
            // this warp's share of the source cell's own column: pattern, values, slot map
            const SubSeg own = tb[(size_t)T * MS_NW];
            for (uint32_t i = lane; i < own.n; i += 32) {
                const uint32_t g = indices[own.lo + i] - g0;
                my_pg[i] = g;
                my_pv[i] = values[own.lo + i];
                my_yhat[i] = 0.0f;
                slot[g] = (unsigned short)i;
            }
            __syncwarp();
            for (uint32_t t0 = 0; t0 < T; t0 += MS_PF) {
#pragma unroll
                for (int s = 0; s < MS_PF; ++s) {
                    const uint32_t t = t0 + s;
                    cp_async_wait<MS_PF - 1>();  // this lane's copies for column t have landed
                    if (t < T) {
                        const SubSeg d = tb[(size_t)t * MS_NW];
                        if (d.n) {  // warp-uniform
                            const unsigned long long c0 = (d.lo & ~3ull) + 4ull * lane;
                            if (c0 < d.lo + d.n) {
                                uint4 g = myring[(s * 2 + 0) * 32 + lane];
                                const uint4 vb = myring[(s * 2 + 1) * 32 + lane];
                                float4 v = make_float4(__uint_as_float(vb.x), __uint_as_float(vb.y), __uint_as_float(vb.z), __uint_as_float(vb.w));
                                if (c0 + 4 > nnz_total) fetch(c0, g, v);
                                apply4(c0, d.lo, d.n, d.w, g, v);
                            }
                            for (unsigned long long c = c0 + MS_PASS; c < d.lo + d.n; c += MS_PASS) {  // pieces longer than one pass
                                uint4 g;
                                float4 v;
                                fetch(c, g, v);
                                apply4(c, d.lo, d.n, d.w, g, v);
                            }
                            __syncwarp();  // the next column may touch the same genes from other lanes
                        }
                    }
                    issue(t + MS_PF, s);
                }
            }
            cp_async_wait<0>();
            __syncwarp();
            const float sc = own.w;
            for (uint32_t i = lane; i < own.n; i += 32) {
                const float d = my_yhat[i];
                float x = my_pv[i];
                if (d > 0.0f) x = __fdiv_rn(x, __fmul_rn(d, sc));
                const uint32_t g = my_pg[i];
                res[g] = __fadd_rn(res[g], x);
                slot[g] = 0xFFFF;
            }
            __syncwarp();
            if (lane == 0) umma::mbar_arrive(&bar_empty[buf]);
        }
    }
    __syncthreads();
    float *dst_imp, *dst_res;
    if (seg_single[seg]) {
        dst_imp = out_imp + (size_t)seg_group[seg] * D + g0;
        dst_res = out_res + (size_t)seg_group[seg] * D + g0;
    } else {
        dst_imp = scratch + ((size_t)seg * 2) * D + g0;
        dst_res = scratch + ((size_t)seg * 2 + 1) * D + g0;
    }
    for (uint32_t g = tid; g < Wr; g += MS_THREADS) {
        dst_imp[g] = imp[g];
        dst_res[g] = res[g];
    }
}

// groups that span several segments: partial sums added in segment order
__global__ void k_segment_reduce(const float* __restrict__ scratch, const uint32_t* __restrict__ group_seg0, uint32_t S, uint64_t D,
                                 float* __restrict__ out_imp, float* __restrict__ out_res) {
    const uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= D * S) return;
    const uint32_t s = (uint32_t)(e / D);
    const uint64_t g = e % D;
    const uint32_t a = group_seg0[s], b = group_seg0[s + 1];
    if (b - a < 2) return;
    float x = 0.0f, y = 0.0f;
    for (uint32_t seg = a; seg < b; ++seg) {
        x = __fadd_rn(x, scratch[((size_t)seg * 2) * D + g]);
        y = __fadd_rn(y, scratch[((size_t)seg * 2 + 1) * D + g]);
    }
    out_imp[e] = x;
    out_res[e] = y;
}

// ------------------------------------------------------------------------------------------------
// pb-sample path
// ------------------------------------------------------------------------------------------------
__global__ void k_pair_presence(const uint32_t* __restrict__ grp, const uint32_t* __restrict__ batch, uint64_t n, uint32_t S,
                                uint32_t B, uint32_t* __restrict__ present) {
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n && grp[j] < S && batch[j] < B) present[(size_t)grp[j] * B + batch[j]] = 1u;
}
__global__ void k_cell_to_pb(const uint32_t* __restrict__ grp, const uint32_t* __restrict__ batch, uint64_t n, uint32_t S, uint32_t B,
                             const uint32_t* __restrict__ id, uint32_t* __restrict__ c2p) {
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) c2p[j] = (grp[j] < S && batch[j] < B) ? id[(size_t)grp[j] * B + batch[j]] : NONE;
}

constexpr int PM_Q = 256;     // queries (centroids) per CTA
constexpr int PM_CELLS = 64;  // cells of the target pb-sample per shared-memory tile
constexpr int PM_U = 4;       // cells scored per pass over a query's coordinates (register tile)

// key[q - q0][p] = min over the cells c of pb-sample p of (l2_sq(centroid_q, c) << 32 | c); ~0 when p is in
// q's own batch (which covers p == q) or empty.  The minimum key is the first cell of p met when the batch's
// cells are walked in (distance, index) order, i.e. what the reference's growing search reports for p.
// One thread per query; PM_U cells share every load of the query's coordinates, each with the reference's own
// 16 lane accumulators (knn/metric.rs:19-45), so the distances are bit-exact and the loop is FP-issue bound.
__global__ void __launch_bounds__(PM_Q) k_pb_min_dist(const float* __restrict__ proj, int K, const uint32_t* __restrict__ cell_sorted,
                                                      const uint64_t* __restrict__ pb_off, const float* __restrict__ centroids,
                                                      const uint32_t* __restrict__ pb_batch, uint32_t npb, uint32_t q0, uint32_t nq,
                                                      uint32_t cell_offset, unsigned long long* __restrict__ keys) {
    extern __shared__ float pm_smem[];
    const int ds = K | 1;
    float* qs = pm_smem;                        // PM_Q x ds
    float* cs = pm_smem + (size_t)PM_Q * ds;    // PM_CELLS x K
    __shared__ uint32_t cell_id[PM_CELLS];
    const uint32_t p = blockIdx.x;
    const uint32_t ql = blockIdx.y * PM_Q + threadIdx.x;
    const bool live = ql < nq;
    const uint32_t q = q0 + ql;
    const uint32_t nq_here = min((uint32_t)PM_Q, nq - blockIdx.y * PM_Q);
    for (uint32_t e = threadIdx.x; e < nq_here * (uint32_t)K; e += PM_Q)
        qs[(e / K) * ds + (e % K)] = centroids[(size_t)(q0 + blockIdx.y * PM_Q) * K + e];
    const uint32_t pbatch = pb_batch[p];
    const bool want = live && pb_batch[q] != pbatch;
    unsigned long long best = ~0ull;
    const uint64_t lo = pb_off[p], hi = pb_off[p + 1];
    const float* myq = qs + (size_t)threadIdx.x * ds;
    for (uint64_t base = lo; base < hi; base += PM_CELLS) {
        const int nt = (int)min((uint64_t)PM_CELLS, hi - base);
        __syncthreads();
        if ((int)threadIdx.x < nt) cell_id[threadIdx.x] = cell_sorted[base + threadIdx.x];
        __syncthreads();
        for (int e = threadIdx.x; e < nt * K; e += PM_Q) cs[e] = proj[(size_t)cell_id[e / K] * K + (e % K)];
        __syncthreads();
        if (!want) continue;
        int t = 0;
        for (; t + PM_U <= nt; t += PM_U) {
            float acc[PM_U][16];
#pragma unroll
            for (int u = 0; u < PM_U; ++u)
#pragma unroll
                for (int l = 0; l < 16; ++l) acc[u][l] = 0.0f;
            int c = 0;
            for (; c + 16 <= K; c += 16) {
#pragma unroll
                for (int l = 0; l < 16; ++l) {
                    const float qv = myq[c + l];
#pragma unroll
                    for (int u = 0; u < PM_U; ++u) {
                        const float df = __fsub_rn(cs[(size_t)(t + u) * K + c + l], qv);
                        acc[u][l] = __fadd_rn(acc[u][l], __fmul_rn(df, df));
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < PM_U; ++u) {
                float sum = 0.0f;
#pragma unroll
                for (int l = 0; l < 16; ++l) sum = __fadd_rn(sum, acc[u][l]);
                for (int cc = c; cc < K; ++cc) {
                    const float df = __fsub_rn(cs[(size_t)(t + u) * K + cc], myq[cc]);
                    sum = __fadd_rn(sum, __fmul_rn(df, df));
                }
                const unsigned long long key = ((unsigned long long)__float_as_uint(sum) << 32) | (cell_id[t + u] + cell_offset);
                best = key < best ? key : best;
            }
        }
        for (; t < nt; ++t) {
            const float d2 = adj_l2_sq(cs + (size_t)t * K, myq, K);
            const unsigned long long key = ((unsigned long long)__float_as_uint(d2) << 32) | (cell_id[t] + cell_offset);
            best = key < best ? key : best;
        }
    }
    if (live) keys[(size_t)ql * npb + p] = best;
}

// The same kernel with the query's first 16 * NCH coordinates in REGISTERS and the cell tile read with 128-bit broadcast
// loads: the shared-memory form above issues one LDS per (dim, cell) next to three FP operations and is bound by the
// one-wavefront-per-clock LSU pipe, not by FP issue.  Here an LDS.128 serves four dims of a cell (1 LDS per 12 FP
// operations).  Accumulation order is unchanged: 16 lane accumulators per cell over the full 16-chunks, folded left to
// right, then the tail dims one by one (knn/metric.rs:19-45) — bit-identical keys.
template <int NCH>
__global__ void __launch_bounds__(PM_Q, 2) k_pb_min_dist_reg(const float* __restrict__ proj, int K, const uint32_t* __restrict__ cell_sorted,
                                                             const uint64_t* __restrict__ pb_off, const float* __restrict__ centroids,
                                                             const uint32_t* __restrict__ pb_batch, uint32_t npb, uint32_t q0,
                                                             uint32_t nq, uint32_t cell_offset, unsigned long long* __restrict__ keys) {
    extern __shared__ __align__(16) float pm_smem[];
    const int ds = K | 1, KP = (K + 3) & ~3;
    float* cs = pm_smem;                              // PM_CELLS x KP (16-byte aligned rows)
    float* qs = pm_smem + (size_t)PM_CELLS * KP;      // PM_Q x ds
    __shared__ uint32_t cell_id[PM_CELLS];
    const uint32_t p = blockIdx.x;
    const uint32_t ql = blockIdx.y * PM_Q + threadIdx.x;
    const bool live = ql < nq;
    const uint32_t q = q0 + ql;
    const uint32_t nq_here = min((uint32_t)PM_Q, nq - blockIdx.y * PM_Q);
    for (uint32_t e = threadIdx.x; e < nq_here * (uint32_t)K; e += PM_Q)
        qs[(e / K) * ds + (e % K)] = centroids[(size_t)(q0 + blockIdx.y * PM_Q) * K + e];
    __syncthreads();
    const float* myq = qs + (size_t)threadIdx.x * ds;
    float qr[16 * NCH];
#pragma unroll
    for (int i = 0; i < 16 * NCH; ++i) qr[i] = (live || threadIdx.x < nq_here) ? myq[i] : 0.0f;
    const uint32_t pbatch = pb_batch[p];
    const bool want = live && pb_batch[q] != pbatch;
    unsigned long long best = ~0ull;
    const uint64_t lo = pb_off[p], hi = pb_off[p + 1];
    for (uint64_t base = lo; base < hi; base += PM_CELLS) {
        const int nt = (int)min((uint64_t)PM_CELLS, hi - base);
        __syncthreads();
        if ((int)threadIdx.x < nt) cell_id[threadIdx.x] = cell_sorted[base + threadIdx.x];
        __syncthreads();
        {   // tile fill without a division per element: (cell, dim) advanced incrementally
            int cc = (int)threadIdx.x / K, dd = (int)threadIdx.x % K;
            const int dc = PM_Q / K, dr = PM_Q % K;
            while (cc < nt) {
                cs[cc * KP + dd] = proj[(size_t)cell_id[cc] * K + dd];
                cc += dc;
                dd += dr;
                if (dd >= K) {
                    dd -= K;
                    ++cc;
                }
            }
        }
        __syncthreads();
        if (!want) continue;
        int t = 0;
        for (; t + PM_U <= nt; t += PM_U) {
            float acc[PM_U][16];  // first chunk peeled: 0 + x == x exactly, so the accumulators start as the first squares
#pragma unroll
            for (int ch = 0; ch < NCH; ++ch) {
#pragma unroll
                for (int l4 = 0; l4 < 4; ++l4) {
#pragma unroll
                    for (int u = 0; u < PM_U; ++u) {
                        const float4 cv = *reinterpret_cast<const float4*>(cs + (size_t)(t + u) * KP + 16 * ch + 4 * l4);
                        const float d0 = __fsub_rn(cv.x, qr[16 * ch + 4 * l4 + 0]);
                        const float d1 = __fsub_rn(cv.y, qr[16 * ch + 4 * l4 + 1]);
                        const float d2 = __fsub_rn(cv.z, qr[16 * ch + 4 * l4 + 2]);
                        const float d3 = __fsub_rn(cv.w, qr[16 * ch + 4 * l4 + 3]);
                        if (ch == 0) {
                            acc[u][4 * l4 + 0] = __fmul_rn(d0, d0);
                            acc[u][4 * l4 + 1] = __fmul_rn(d1, d1);
                            acc[u][4 * l4 + 2] = __fmul_rn(d2, d2);
                            acc[u][4 * l4 + 3] = __fmul_rn(d3, d3);
                        } else {
                            acc[u][4 * l4 + 0] = __fadd_rn(acc[u][4 * l4 + 0], __fmul_rn(d0, d0));
                            acc[u][4 * l4 + 1] = __fadd_rn(acc[u][4 * l4 + 1], __fmul_rn(d1, d1));
                            acc[u][4 * l4 + 2] = __fadd_rn(acc[u][4 * l4 + 2], __fmul_rn(d2, d2));
                            acc[u][4 * l4 + 3] = __fadd_rn(acc[u][4 * l4 + 3], __fmul_rn(d3, d3));
                        }
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < PM_U; ++u) {
                float sum = 0.0f;
#pragma unroll
                for (int l = 0; l < 16; ++l) sum = __fadd_rn(sum, acc[u][l]);
                for (int cc = 16 * NCH; cc < K; ++cc) {
                    const float df = __fsub_rn(cs[(size_t)(t + u) * KP + cc], myq[cc]);
                    sum = __fadd_rn(sum, __fmul_rn(df, df));
                }
                const unsigned long long key = ((unsigned long long)__float_as_uint(sum) << 32) | (cell_id[t + u] + cell_offset);
                best = key < best ? key : best;
            }
        }
        for (; t < nt; ++t) {
            const float d2 = adj_l2_sq(cs + (size_t)t * KP, myq, K);
            const unsigned long long key = ((unsigned long long)__float_as_uint(d2) << 32) | (cell_id[t] + cell_offset);
            best = key < best ? key : best;
        }
    }
    if (live) keys[(size_t)ql * npb + p] = best;
}

// launches the register form when the full 16-chunks of K fit it (K < 80), else the shared-memory form
static int launch_pb_min_dist(lg_ctx* ctx, dim3 grid, const float* d_proj, int K, const uint32_t* d_cell, const uint64_t* d_off,
                              const float* d_cen, const uint32_t* d_pbb, uint32_t npb, uint32_t q0, uint32_t nq, uint32_t cell_offset,
                              unsigned long long* d_keys) {
    const int nch = K / 16;
    const size_t smem_reg = ((size_t)PM_CELLS * ((K + 3) & ~3) + (size_t)PM_Q * (K | 1)) * sizeof(float);
    const char* force = getenv("LG_PB_SMEM_FORM");
    if (nch >= 1 && nch <= 4 && smem_reg <= ctx->smem_optin && !(force && force[0] == '1')) {
#define LG_PBMD(N)                                                                                                               \
    do {                                                                                                                         \
        LG_CUDA(ctx, cudaFuncSetAttribute(k_pb_min_dist_reg<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_reg));    \
        k_pb_min_dist_reg<N><<<grid, PM_Q, smem_reg, ctx->stream>>>(d_proj, K, d_cell, d_off, d_cen, d_pbb, npb, q0, nq,        \
                                                                    cell_offset, d_keys);                                       \
    } while (0)
        if (nch == 1) LG_PBMD(1);
        else if (nch == 2) LG_PBMD(2);
        else if (nch == 3) LG_PBMD(3);
        else LG_PBMD(4);
#undef LG_PBMD
    } else {
        const size_t smem = ((size_t)PM_Q * (K | 1) + (size_t)PM_CELLS * K) * sizeof(float);
        LG_CUDA(ctx, cudaFuncSetAttribute(k_pb_min_dist, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_pb_min_dist<<<grid, PM_Q, smem, ctx->stream>>>(d_proj, K, d_cell, d_off, d_cen, d_pbb, npb, q0, nq, cell_offset, d_keys);
    }
    ctx->launches++;
    LG_CUDA(ctx, cudaGetLastError());
    return LG_OK;
}

// one warp per (query, target batch): the knn smallest keys among that batch's pb-samples, ascending
__global__ void k_pb_topk(const unsigned long long* __restrict__ keys, uint32_t npb, uint32_t q0, uint32_t nq, uint32_t B,
                          const uint32_t* __restrict__ batch_pb, const uint32_t* __restrict__ batch_off, int knn,
                          uint32_t* __restrict__ out_pb, float* __restrict__ out_dist) {
    const int lane = threadIdx.x & 31;
    const uint64_t wid = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (wid >= (uint64_t)nq * B) return;
    const uint32_t ql = (uint32_t)(wid / B), b = (uint32_t)(wid % B);
    const unsigned long long* row = keys + (size_t)ql * npb;
    const uint32_t lo = batch_off[b], hi = batch_off[b + 1];
    unsigned long long last = 0;
    bool first = true;
    const size_t o = ((size_t)(q0 + ql) * B + b) * knn;
    for (int r = 0; r < knn; ++r) {
        unsigned long long best = ~0ull;
        uint32_t arg = NONE;
        for (uint32_t i = lo + lane; i < hi; i += 32) {
            const uint32_t p = batch_pb[i];
            const unsigned long long k = row[p];
            if ((first || k > last) && k < best) {
                best = k;
                arg = p;
            }
        }
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) {
            const unsigned long long ob = __shfl_xor_sync(0xffffffffu, best, off);
            const uint32_t oa = __shfl_xor_sync(0xffffffffu, arg, off);
            if (ob < best) {
                best = ob;
                arg = oa;
            }
        }
        if (lane == 0) {
            const bool empty = best == ~0ull;
            out_pb[o + r] = empty ? NONE : arg;
            out_dist[o + r] = empty ? INFINITY : __fsqrt_rn(__uint_as_float((uint32_t)(best >> 32)));
        }
        if (best == ~0ull) {
            // nothing left: the remaining slots stay empty
            if (lane == 0)
                for (int rr = r + 1; rr < knn; ++rr) {
                    out_pb[o + rr] = NONE;
                    out_dist[o + rr] = INFINITY;
                }
            break;
        }
        last = best;
        first = false;
    }
}

// stats.rs:715-731: softmax weights (MAX of -d subtracted), one warp per pb-sample
__global__ void k_pb_weights(const uint32_t* __restrict__ mpb, const float* __restrict__ mdist, uint32_t npb, uint32_t T,
                             float* __restrict__ w) {
    const int lane = threadIdx.x & 31;
    const uint64_t p = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (p >= npb) return;
    const uint32_t* mi = mpb + p * T;
    const float* md = mdist + p * T;
    float mx = -INFINITY;
    for (uint32_t t = lane; t < T; t += 32)
        if (mi[t] < npb) mx = fmaxf(mx, -md[t]);
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
    float sum = 0.0f;
    for (uint32_t t = lane; t < T; t += 32)
        if (mi[t] < npb) sum += expf(-md[t] - mx);
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
    for (uint32_t t = lane; t < T; t += 32) {
        float x = 0.0f;
        if (mi[t] < npb) {
            x = expf(-md[t] - mx);
            if (sum > 0.0f) x = __fdiv_rn(x, sum);
        }
        w[p * T + t] = x;
    }
}

// stats.rs:733-779: one thread per (gene, group); the group's pb-samples are folded in ascending order
__global__ void k_coarse_stat(const float* __restrict__ gene_sums, uint64_t D, uint32_t npb, const float* __restrict__ pb_count,
                              const uint32_t* __restrict__ group_pb, const uint32_t* __restrict__ group_off, uint32_t S,
                              const uint32_t* __restrict__ mpb, const float* __restrict__ w, uint32_t T, float* __restrict__ out_imp,
                              float* __restrict__ out_res) {
    const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t s = blockIdx.y;
    if (g >= D) return;
    float imp = 0.0f, res = 0.0f;
    for (uint32_t i = group_off[s]; i < group_off[s + 1]; ++i) {
        const uint32_t p = group_pb[i];
        const float cnt = pb_count[p];
        if (cnt < 1.0f) continue;
        float yhat = 0.0f;
        bool present = false, any = false;
        for (uint32_t t = 0; t < T; ++t) {
            const uint32_t m = mpb[(size_t)p * T + t];
            if (m >= npb) continue;
            any = true;
            const float mc = pb_count[m];
            if (mc < 1.0f) continue;
            const float gs = gene_sums[(size_t)m * D + g];
            if (gs != 0.0f) {
                present = true;
                yhat = __fadd_rn(yhat, __fmul_rn(__fmul_rn(w[(size_t)p * T + t], gs), __fdiv_rn(1.0f, mc)));
            }
        }
        if (!any || !present) continue;
        imp = __fadd_rn(imp, __fmul_rn(cnt, yhat));
        const float own = gene_sums[(size_t)p * D + g];
        if (own != 0.0f && yhat > 0.0f) res = __fadd_rn(res, __fdiv_rn(own, yhat));
    }
    out_imp[(size_t)s * D + g] = imp;
    out_res[(size_t)s * D + g] = res;
}

__global__ void k_group_code(const uint64_t* __restrict__ codes, const uint32_t* __restrict__ grp, uint64_t n, uint32_t nfine,
                             unsigned long long* __restrict__ group_code) {
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n && grp[j] < nfine) group_code[grp[j]] = codes[j];  // every cell of a group carries the same code
}

}  // namespace

// ====================================================================================================
extern "C" int lg_batch_proximity(lg_ctx* ctx, const float* proj_kn, int K, uint64_t ncols, const uint32_t* batch_of_cell,
                                  uint32_t B, uint32_t* out_order, float* out_centroids) {
    if (!ctx) return LG_ERR_INVALID;
    LG_REQUIRE(ctx, proj_kn && batch_of_cell && out_order, "lg_batch_proximity: null argument");
    LG_REQUIRE(ctx, K >= 1 && K <= 1024 && B >= 1 && B <= 65535 && ncols < 0xFFFFFFFFull, "lg_batch_proximity: bad shape");
    cudaSetDevice(ctx->device);
    LgStage st(ctx);
    const float* d_proj;
    const uint32_t* d_batch;
    LG_TRY(st.in(proj_kn, (size_t)ncols * K, &d_proj));
    LG_TRY(st.in(batch_of_cell, (size_t)ncols, &d_batch));
    std::vector<uint32_t> counts;
    LG_TRY(label_counts_host(ctx, st, d_batch, ncols, B, counts));
    std::vector<uint64_t> off(B + 1, 0);
    for (uint32_t b = 0; b < B; ++b) off[b + 1] = off[b] + counts[b];
    uint32_t *d_lab, *d_cell;
    LG_TRY(sort_cells_by_label(ctx, st, d_batch, ncols, &d_lab, &d_cell));
    uint64_t* d_off;
    LG_TRY(upload(ctx, st, off, &d_off));
    float* d_cen;
    LG_TRY(st.scratch((size_t)B * K, &d_cen));
    LG_LAUNCH(ctx, k_segment_mean<0>, B, ((K + 31) / 32) * 32, 0, d_proj, K, d_cell, d_off, (const float*)nullptr, d_cen,
              (float*)nullptr);
    std::vector<float> cen((size_t)B * K);
    LG_CUDA(ctx, cudaMemcpyAsync(cen.data(), d_cen, sizeof(float) * cen.size(), cudaMemcpyDeviceToHost, ctx->stream));
    LG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    // B x B exact scan on the host (B is tiny): search_by_query_name(b, nbatches, false)
    std::vector<uint32_t> order((size_t)B * B);
    std::vector<std::pair<float, uint32_t>> sc(B);
    for (uint32_t b = 0; b < B; ++b) {
        for (uint32_t o = 0; o < B; ++o) sc[o] = {lgh_l2_sq(&cen[(size_t)o * K], &cen[(size_t)b * K], K), o};
        std::sort(sc.begin(), sc.end());
        for (uint32_t o = 0; o < B; ++o) order[(size_t)b * B + o] = sc[o].second;
    }
    auto put = [&](void* dst, const void* src, size_t bytes) -> int {
        LG_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, ctx->stream));
        LG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        return LG_OK;
    };
    LG_TRY(put(out_order, order.data(), sizeof(uint32_t) * order.size()));
    if (out_centroids) LG_TRY(put(out_centroids, cen.data(), sizeof(float) * cen.size()));
    return st.finish();
}

extern "C" int lg_knn_match_batches(lg_ctx* ctx, const float* proj_kn, int K, uint64_t ncols, const uint32_t* batch_of_cell,
                                    uint32_t B, int knn, const uint32_t* target_order, uint32_t nt, uint32_t* out_idx,
                                    float* out_dist) {
    if (!ctx) return LG_ERR_INVALID;
    LG_REQUIRE(ctx, proj_kn && batch_of_cell && out_idx && out_dist, "lg_knn_match_batches: null argument");
    LG_REQUIRE(ctx, K >= 1 && K <= 256 && knn >= 1 && knn <= 1024 && B >= 1 && B <= 65535 && ncols < 0xFFFFFFFFull,
               "lg_knn_match_batches: bad shape");
    if (!target_order) nt = B;
    LG_REQUIRE(ctx, nt >= 1, "lg_knn_match_batches: no target slots");
    cudaSetDevice(ctx->device);
    LgStage st(ctx);
    const uint64_t N = ncols;
    const uint32_t T = nt * (uint32_t)knn;
    const float* d_proj;
    const uint32_t* d_batch;
    uint32_t* d_oidx;
    float* d_odist;
    LG_TRY(st.in(proj_kn, (size_t)N * K, &d_proj));
    LG_TRY(st.in(batch_of_cell, (size_t)N, &d_batch));
    LG_TRY(st.out(out_idx, (size_t)N * T, &d_oidx));
    LG_TRY(st.out(out_dist, (size_t)N * T, &d_odist));
    if (N == 0) return st.finish();
    // slot_of[s][b]: which output slot of a source cell of batch s holds target batch b
    std::vector<uint32_t> slot_of((size_t)B * B, NONE);
    {
        std::vector<uint32_t> order;
        if (target_order) LG_TRY(to_host(ctx, target_order, (size_t)B * nt, order));
        for (uint32_t s = 0; s < B; ++s)
            for (uint32_t i = 0; i < nt; ++i) {
                const uint32_t b = target_order ? order[(size_t)s * nt + i] : i;
                if (b < B && b != s && slot_of[(size_t)s * B + b] == NONE) slot_of[(size_t)s * B + b] = i;
            }
    }
    std::vector<uint32_t> counts;
    LG_TRY(label_counts_host(ctx, st, d_batch, N, B, counts));
    std::vector<uint64_t> off(B + 1, 0);
    for (uint32_t b = 0; b < B; ++b) off[b + 1] = off[b] + counts[b];
    const uint64_t nlive = off[B];  // cells with a valid batch sort first
    uint32_t *d_lab, *d_cell, *d_slot;
    LG_TRY(sort_cells_by_label(ctx, st, d_batch, N, &d_lab, &d_cell));
    LG_TRY(upload(ctx, st, slot_of, &d_slot));
    float* d_sorted;
    LG_TRY(st.scratch((size_t)N * K, &d_sorted));
    LG_LAUNCH(ctx, k_gather_rows, (unsigned)(((uint64_t)N * K + 255) / 256), 256, 0, d_proj, K, d_cell, N, d_sorted);
    LG_LAUNCH(ctx, k_fill_u32, (unsigned)(((uint64_t)N * T + 255) / 256), 256, 0, d_oidx, (uint64_t)N * T, NONE);
    LG_LAUNCH(ctx, k_fill_f32, (unsigned)(((uint64_t)N * T + 255) / 256), 256, 0, d_odist, (uint64_t)N * T, INFINITY);
    uint32_t* d_kidx;
    float* d_kdist;
    LG_TRY(st.scratch((size_t)N * knn, &d_kidx));
    LG_TRY(st.scratch((size_t)N * knn, &d_kdist));
    for (uint32_t b = 0; b < B; ++b) {
        const uint64_t nb = counts[b];
        if (nb == 0) continue;
        bool wanted = false;
        for (uint32_t s = 0; s < B && !wanted; ++s) wanted = slot_of[(size_t)s * B + b] != NONE && counts[s] > 0;
        if (!wanted) continue;
        // queries: every cell outside batch b = the two ranges around it in batch-sorted order
        const uint64_t ranges[2][2] = {{0, off[b]}, {off[b + 1], nlive}};
        for (int h = 0; h < 2; ++h) {
            const uint64_t q0 = ranges[h][0], nq = ranges[h][1] - ranges[h][0];
            if (nq == 0) continue;
            LG_TRY(lg_knn_topk_device(ctx, d_sorted + off[b] * K, nb, d_sorted + q0 * K, nq, K, knn, nullptr, d_kidx, d_kdist, 0));
            LG_LAUNCH(ctx, k_scatter_matches, (unsigned)((nq * knn + 255) / 256), 256, 0, d_kidx, d_kdist, nq, knn, q0, off[b], d_cell,
                      d_lab, d_slot, B, b, T, d_oidx, d_odist);
        }
    }
    return st.finish();
}

extern "C" int lg_collect_matched_stat(lg_ctx* ctx, const lg_csc* m, const uint32_t* group_of_cell, uint32_t S,
                                       const uint32_t* matched_idx, const float* matched_dist, uint32_t T, float* out_imputed_ds,
                                       float* out_residual_ds) {
    if (!ctx) return LG_ERR_INVALID;
    LG_REQUIRE(ctx, m && group_of_cell && matched_idx && matched_dist && out_imputed_ds && out_residual_ds,
               "lg_collect_matched_stat: null argument");
    LG_REQUIRE(ctx, T >= 1 && T <= 1024 && S >= 1, "lg_collect_matched_stat: T must be in [1, 1024] and S >= 1");
    cudaSetDevice(ctx->device);
    LG_TRY(lg_csc_require_canonical(ctx, m, "lg_collect_matched_stat"));
    LgStage st(ctx);
    const uint64_t N = m->ncols, D = m->nrows;
    LG_REQUIRE(ctx, N < 0xFFFFFFFFull, "lg_collect_matched_stat: more than 2^32-1 cells in one block");
    LG_REQUIRE(ctx, ((uintptr_t)m->indices & 15) == 0 && ((uintptr_t)m->values & 15) == 0,
               "lg_collect_matched_stat: the CSC index / value arrays must be 16-byte aligned");
    const uint32_t* d_group;
    const uint32_t* d_midx;
    const float* d_mdist;
    float *d_imp, *d_res;
    LG_TRY(st.in(group_of_cell, (size_t)N, &d_group));
    LG_TRY(st.in(matched_idx, (size_t)N * T, &d_midx));
    LG_TRY(st.in(matched_dist, (size_t)N * T, &d_mdist));
    LG_TRY(st.out(out_imputed_ds, (size_t)D * S, &d_imp));
    LG_TRY(st.out(out_residual_ds, (size_t)D * S, &d_res));
    LG_CUDA(ctx, cudaMemsetAsync(d_imp, 0, sizeof(float) * (size_t)D * S, ctx->stream));
    LG_CUDA(ctx, cudaMemsetAsync(d_res, 0, sizeof(float) * (size_t)D * S, ctx->stream));
    if (N == 0 || D == 0) return st.finish();

    // gene ranges: per range two f32 accumulators + a u16 slot map (10 B per gene) next to the cp.async rings, the
    // double-buffered piece table and the per-warp pattern buffers (12 B per own nnz of a sub-range)
    const size_t fixed = MS_RING_BYTES + (size_t)2 * (T + 1) * MS_NW * sizeof(SubSeg) + 2048;
    LG_REQUIRE(ctx, fixed + 4096 < ctx->smem_optin, "lg_collect_matched_stat: T too large for shared memory");
    const size_t budget = ctx->smem_optin - fixed;
    int R = 1;
    uint32_t W = 0, wsub = 0, pmax = 0;
    float* d_colsum;
    uint32_t* d_split = nullptr;
    unsigned int* d_maxpart;
    LG_TRY(st.scratch(N, &d_colsum));
    LG_TRY(st.scratch(1, &d_maxpart));
    LG_LAUNCH(ctx, k_cell_colsum, (unsigned)((N * 32 + 255) / 256), 256, 0, m->indptr, m->values, N, d_colsum);
    for (;; ++R) {
        LG_REQUIRE(ctx, R <= MS_MAXR, "lg_collect_matched_stat: too many genes for the shared-memory accumulators");
        wsub = (uint32_t)((D + (uint64_t)R * MS_NW - 1) / ((uint64_t)R * MS_NW));
        wsub = (wsub + 7) & ~7u;
        W = wsub * MS_NW;
        if ((size_t)W * 10 + 64 > budget) continue;
        const uint32_t nb = (uint32_t)R * MS_NW;
        LG_TRY(st.scratch((size_t)N * (nb + 1), &d_split));
        LG_CUDA(ctx, cudaMemsetAsync(d_maxpart, 0, sizeof(unsigned int), ctx->stream));
        LG_LAUNCH(ctx, k_cell_splits, (unsigned)((N * (nb + 1) + 255) / 256), 256, 0, m->indptr, m->indices, N, wsub, nb, d_split);
        LG_LAUNCH(ctx, k_max_part, (unsigned)((N * nb + 255) / 256), 256, 0, d_split, N, nb, d_maxpart);
        unsigned int* h = static_cast<unsigned int*>(ctx->pinned);
        LG_CUDA(ctx, cudaMemcpyAsync(h, d_maxpart, sizeof(unsigned int), cudaMemcpyDeviceToHost, ctx->stream));
        LG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        pmax = (*h + 3) & ~3u;
        if (pmax < 4) pmax = 4;
        if ((size_t)W * 10 + (size_t)MS_NW * pmax * 12 + 64 <= budget) break;
    }
    LG_REQUIRE(ctx, pmax < 65535, "lg_collect_matched_stat: a column holds more than 65534 entries in one gene sub-range");
    const size_t smem = MS_RING_BYTES + (size_t)2 * (T + 1) * MS_NW * sizeof(SubSeg) + (size_t)W * 10 + (size_t)MS_NW * pmax * 12 + 64;

    // segments of at most MS_SEG group-sorted cells, never crossing a group boundary
    std::vector<uint32_t> counts;
    LG_TRY(label_counts_host(ctx, st, d_group, N, S, counts));
    std::vector<uint32_t> seg_first, seg_len, seg_group, group_seg0(S + 1, 0);
    std::vector<uint8_t> seg_single;
    {
        uint64_t pos = 0;
        for (uint32_t s = 0; s < S; ++s) {
            group_seg0[s] = (uint32_t)seg_first.size();
            const uint32_t nseg = (counts[s] + MS_SEG - 1) / MS_SEG;
            for (uint32_t i = 0; i < nseg; ++i) {
                seg_first.push_back((uint32_t)(pos + (uint64_t)i * MS_SEG));
                seg_len.push_back(std::min<uint32_t>(MS_SEG, counts[s] - i * MS_SEG));
                seg_group.push_back(s);
                seg_single.push_back(nseg == 1);
            }
            pos += counts[s];
        }
        group_seg0[S] = (uint32_t)seg_first.size();
    }
    const uint32_t nseg = (uint32_t)seg_first.size();
    if (nseg == 0) return st.finish();
    uint32_t *d_lab, *d_cell, *d_seg_first, *d_seg_len, *d_seg_group, *d_group_seg0;
    uint8_t* d_seg_single;
    LG_TRY(sort_cells_by_label(ctx, st, d_group, N, &d_lab, &d_cell));
    LG_TRY(upload(ctx, st, seg_first, &d_seg_first));
    LG_TRY(upload(ctx, st, seg_len, &d_seg_len));
    LG_TRY(upload(ctx, st, seg_group, &d_seg_group));
    LG_TRY(upload(ctx, st, seg_single, &d_seg_single));
    LG_TRY(upload(ctx, st, group_seg0, &d_group_seg0));
    MatchW* d_mw;
    float *d_scale, *d_scratch;
    LG_TRY(st.scratch((size_t)N * T, &d_mw));
    LG_TRY(st.scratch(N, &d_scale));
    const bool multi = std::any_of(seg_single.begin(), seg_single.end(), [](uint8_t x) { return x == 0; });
    LG_TRY(st.scratch(multi ? (size_t)nseg * 2 * D : 1, &d_scratch));
    LG_LAUNCH(ctx, k_match_weights, (unsigned)((N * 32 + 255) / 256), 256, 0, d_colsum, N, d_midx, d_mdist, T, d_mw, d_scale);
    LG_CUDA(ctx, cudaFuncSetAttribute(k_matched_stat, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    {
        dim3 grid(nseg, (unsigned)R);
        k_matched_stat<<<grid, MS_THREADS, smem, ctx->stream>>>(m->indptr, m->indices, m->values, m->nnz, d_split, d_cell, d_seg_first, d_seg_len,
                                                               d_seg_group, d_seg_single, d_mw, d_scale, N, T, R, D, W, wsub, pmax, d_scratch,
                                                               d_imp, d_res);
        ctx->launches++;
        LG_CUDA(ctx, cudaGetLastError());
    }
    if (multi) LG_LAUNCH(ctx, k_segment_reduce, (unsigned)((D * S + 255) / 256), 256, 0, d_scratch, d_group_seg0, S, D, d_imp, d_res);
    return st.finish();
}

extern "C" int lg_pb_layout(lg_ctx* ctx, const float* proj_kn, int K, uint64_t ncols, const uint32_t* group_of_cell, uint32_t S,
                            const uint32_t* batch_of_cell, uint32_t B, const float* mult, uint32_t* out_cell_to_pb,
                            uint32_t* out_pb_group, uint32_t* out_pb_batch, float* out_pb_count, float* out_centroids,
                            uint32_t* out_num_pb) {
    if (!ctx) return LG_ERR_INVALID;
    LG_REQUIRE(ctx, proj_kn && group_of_cell && batch_of_cell && out_cell_to_pb && out_pb_group && out_pb_batch && out_pb_count &&
                        out_centroids && out_num_pb,
               "lg_pb_layout: null argument");
    LG_REQUIRE(ctx, K >= 1 && K <= 1024 && S >= 1 && B >= 1 && (uint64_t)S * B < 0x7FFFFFFFull && ncols < 0xFFFFFFFFull,
               "lg_pb_layout: bad shape");
    cudaSetDevice(ctx->device);
    LgStage st(ctx);
    const uint64_t N = ncols;
    const size_t cap = (size_t)S * B;
    const float *d_proj, *d_mult;
    const uint32_t *d_group, *d_batch;
    uint32_t *d_c2p, *d_pg, *d_pb;
    float *d_cnt, *d_cen;
    LG_TRY(st.in(proj_kn, (size_t)N * K, &d_proj));
    LG_TRY(st.in(group_of_cell, (size_t)N, &d_group));
    LG_TRY(st.in(batch_of_cell, (size_t)N, &d_batch));
    LG_TRY(st.in(mult, (size_t)N, &d_mult));
    LG_TRY(st.out(out_cell_to_pb, (size_t)N, &d_c2p));
    LG_TRY(st.out(out_pb_group, cap, &d_pg));
    LG_TRY(st.out(out_pb_batch, cap, &d_pb));
    LG_TRY(st.out(out_pb_count, cap, &d_cnt));
    LG_TRY(st.out(out_centroids, cap * K, &d_cen));
    uint32_t* d_present;
    LG_TRY(st.scratch(cap, &d_present));
    LG_CUDA(ctx, cudaMemsetAsync(d_present, 0, sizeof(uint32_t) * cap, ctx->stream));
    LG_CUDA(ctx, cudaMemsetAsync(d_pg, 0, sizeof(uint32_t) * cap, ctx->stream));
    LG_CUDA(ctx, cudaMemsetAsync(d_pb, 0, sizeof(uint32_t) * cap, ctx->stream));
    LG_CUDA(ctx, cudaMemsetAsync(d_cnt, 0, sizeof(float) * cap, ctx->stream));
    LG_CUDA(ctx, cudaMemsetAsync(d_cen, 0, sizeof(float) * cap * K, ctx->stream));
    if (N) LG_LAUNCH(ctx, k_pair_presence, (unsigned)((N + 255) / 256), 256, 0, d_group, d_batch, N, S, B, d_present);
    std::vector<uint32_t> present(cap);
    LG_CUDA(ctx, cudaMemcpyAsync(present.data(), d_present, sizeof(uint32_t) * cap, cudaMemcpyDeviceToHost, ctx->stream));
    LG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    std::vector<uint32_t> id(cap, NONE), pg, pb;
    for (size_t e = 0; e < cap; ++e)
        if (present[e]) {
            id[e] = (uint32_t)pg.size();
            pg.push_back((uint32_t)(e / B));
            pb.push_back((uint32_t)(e % B));
        }
    const uint32_t npb = (uint32_t)pg.size();
    *out_num_pb = npb;
    if (npb == 0) return lg_fail(ctx, LG_ERR_INVALID, "lg_pb_layout: no pb-samples built");  // pb_samples.rs:179-181
    uint32_t* d_id;
    LG_TRY(upload(ctx, st, id, &d_id));
    LG_CUDA(ctx, cudaMemcpyAsync(d_pg, pg.data(), sizeof(uint32_t) * npb, cudaMemcpyHostToDevice, ctx->stream));
    LG_CUDA(ctx, cudaMemcpyAsync(d_pb, pb.data(), sizeof(uint32_t) * npb, cudaMemcpyHostToDevice, ctx->stream));
    LG_LAUNCH(ctx, k_cell_to_pb, (unsigned)((N + 255) / 256), 256, 0, d_group, d_batch, N, S, B, d_id, d_c2p);
    std::vector<uint32_t> counts;
    LG_TRY(label_counts_host(ctx, st, d_c2p, N, npb, counts));  // also orders the two small copies above
    std::vector<uint64_t> off(npb + 1, 0);
    for (uint32_t p = 0; p < npb; ++p) off[p + 1] = off[p] + counts[p];
    uint32_t *d_lab, *d_cell;
    uint64_t* d_off;
    LG_TRY(sort_cells_by_label(ctx, st, d_c2p, N, &d_lab, &d_cell));
    LG_TRY(upload(ctx, st, off, &d_off));
    LG_LAUNCH(ctx, k_segment_mean<1>, npb, ((K + 31) / 32) * 32, 0, d_proj, K, d_cell, d_off, d_mult, d_cen, d_cnt);
    return st.finish();
}

extern "C" int lg_pb_match(lg_ctx* ctx, const float* proj_kn, int K, uint64_t ncols, const uint32_t* batch_of_cell, uint32_t B,
                           const uint32_t* cell_to_pb, const float* centroids, const uint32_t* pb_batch, uint32_t npb, int knn,
                           uint32_t* out_matched_pb, float* out_matched_dist) {
    if (!ctx) return LG_ERR_INVALID;
    LG_REQUIRE(ctx, proj_kn && batch_of_cell && cell_to_pb && centroids && pb_batch && out_matched_pb && out_matched_dist,
               "lg_pb_match: null argument");
    LG_REQUIRE(ctx, K >= 1 && K <= 256 && knn >= 1 && B >= 1 && npb >= 1 && ncols < 0xFFFFFFFFull, "lg_pb_match: bad shape");
    cudaSetDevice(ctx->device);
    LgStage st(ctx);
    const uint64_t N = ncols;
    const uint32_t T = B * (uint32_t)knn;
    const float *d_proj, *d_cen;
    const uint32_t *d_c2p, *d_pbb;
    uint32_t* d_mpb;
    float* d_md;
    (void)batch_of_cell;  // a cell's batch is its pb-sample's batch
    LG_TRY(st.in(proj_kn, (size_t)N * K, &d_proj));
    LG_TRY(st.in(cell_to_pb, (size_t)N, &d_c2p));
    LG_TRY(st.in(centroids, (size_t)npb * K, &d_cen));
    LG_TRY(st.in(pb_batch, (size_t)npb, &d_pbb));
    LG_TRY(st.out(out_matched_pb, (size_t)npb * T, &d_mpb));
    LG_TRY(st.out(out_matched_dist, (size_t)npb * T, &d_md));
    std::vector<uint32_t> pbb;
    LG_TRY(to_host(ctx, pb_batch, npb, pbb));
    std::vector<uint32_t> batch_off(B + 1, 0), batch_pb(npb);
    for (uint32_t p = 0; p < npb; ++p) {
        LG_REQUIRE(ctx, pbb[p] < B, "lg_pb_match: pb_batch out of range");
        batch_off[pbb[p] + 1]++;
    }
    for (uint32_t b = 0; b < B; ++b) batch_off[b + 1] += batch_off[b];
    {
        std::vector<uint32_t> cur(batch_off.begin(), batch_off.end() - 1);
        for (uint32_t p = 0; p < npb; ++p) batch_pb[cur[pbb[p]]++] = p;
    }
    std::vector<uint32_t> counts;
    LG_TRY(label_counts_host(ctx, st, d_c2p, N, npb, counts));
    std::vector<uint64_t> off(npb + 1, 0);
    for (uint32_t p = 0; p < npb; ++p) off[p + 1] = off[p] + counts[p];
    uint32_t *d_lab, *d_cell, *d_batch_pb, *d_batch_off;
    uint64_t* d_off;
    LG_TRY(sort_cells_by_label(ctx, st, d_c2p, N, &d_lab, &d_cell));
    LG_TRY(upload(ctx, st, off, &d_off));
    LG_TRY(upload(ctx, st, batch_pb, &d_batch_pb));
    LG_TRY(upload(ctx, st, batch_off, &d_batch_off));
    // query chunks keep the key matrix below ~2 GiB
    uint32_t QC = (uint32_t)std::min<uint64_t>(npb, std::max<uint64_t>(PM_Q, ((uint64_t)1 << 28) / npb));
    QC = ((QC + PM_Q - 1) / PM_Q) * PM_Q;
    unsigned long long* d_keys;
    LG_TRY(st.scratch((size_t)QC * npb, &d_keys));
    const size_t smem = ((size_t)PM_Q * (K | 1) + (size_t)PM_CELLS * K) * sizeof(float);
    LG_REQUIRE(ctx, smem <= ctx->smem_optin, "lg_pb_match: K too large for shared memory");
    LG_CUDA(ctx, cudaFuncSetAttribute(k_pb_min_dist, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (uint32_t q0 = 0; q0 < npb; q0 += QC) {
        const uint32_t nq = std::min(QC, npb - q0);
        dim3 grid(npb, (nq + PM_Q - 1) / PM_Q);
        LG_TRY(launch_pb_min_dist(ctx, grid, d_proj, K, d_cell, d_off, d_cen, d_pbb, npb, q0, nq, 0u, d_keys));
        LG_LAUNCH(ctx, k_pb_topk, (unsigned)(((uint64_t)nq * B * 32 + 255) / 256), 256, 0, d_keys, npb, q0, nq, B, d_batch_pb, d_batch_off,
                  knn, d_mpb, d_md);
    }
    return st.finish();
}

extern "C" int lg_collect_matched_stat_coarse(lg_ctx* ctx, const float* gene_sums, uint64_t D, uint32_t npb, const float* pb_count,
                                              const uint32_t* pb_to_group, uint32_t S, const uint32_t* matched_pb,
                                              const float* matched_dist, uint32_t T, float* out_imputed_ds, float* out_residual_ds) {
    if (!ctx) return LG_ERR_INVALID;
    LG_REQUIRE(ctx, gene_sums && pb_count && pb_to_group && matched_pb && matched_dist && out_imputed_ds && out_residual_ds,
               "lg_collect_matched_stat_coarse: null argument");
    LG_REQUIRE(ctx, npb >= 1 && S >= 1 && S <= 65535 && T >= 1, "lg_collect_matched_stat_coarse: bad shape");
    cudaSetDevice(ctx->device);
    LgStage st(ctx);
    const float *d_gs, *d_cnt, *d_md;
    const uint32_t* d_mpb;
    float *d_imp, *d_res;
    LG_TRY(st.in(gene_sums, (size_t)D * npb, &d_gs));
    LG_TRY(st.in(pb_count, (size_t)npb, &d_cnt));
    LG_TRY(st.in(matched_pb, (size_t)npb * T, &d_mpb));
    LG_TRY(st.in(matched_dist, (size_t)npb * T, &d_md));
    LG_TRY(st.out(out_imputed_ds, (size_t)D * S, &d_imp));
    LG_TRY(st.out(out_residual_ds, (size_t)D * S, &d_res));
    std::vector<uint32_t> p2g;
    LG_TRY(to_host(ctx, pb_to_group, npb, p2g));
    std::vector<uint32_t> group_off(S + 1, 0), group_pb(npb);
    uint32_t kept = 0;
    for (uint32_t p = 0; p < npb; ++p)
        if (p2g[p] < S) {
            group_off[p2g[p] + 1]++;
            ++kept;
        }
    for (uint32_t s = 0; s < S; ++s) group_off[s + 1] += group_off[s];
    {
        std::vector<uint32_t> cur(group_off.begin(), group_off.end() - 1);
        for (uint32_t p = 0; p < npb; ++p)
            if (p2g[p] < S) group_pb[cur[p2g[p]]++] = p;
    }
    (void)kept;
    uint32_t *d_group_pb, *d_group_off;
    float* d_w;
    LG_TRY(upload(ctx, st, group_pb, &d_group_pb));
    LG_TRY(upload(ctx, st, group_off, &d_group_off));
    LG_TRY(st.scratch((size_t)npb * T, &d_w));
    LG_LAUNCH(ctx, k_pb_weights, (unsigned)(((uint64_t)npb * 32 + 255) / 256), 256, 0, d_mpb, d_md, npb, T, d_w);
    if (D) {
        dim3 grid((unsigned)((D + 255) / 256), S);
        k_coarse_stat<<<grid, 256, 0, ctx->stream>>>(d_gs, D, npb, d_cnt, d_group_pb, d_group_off, S, d_mpb, d_w, T, d_imp, d_res);
        ctx->launches++;
        LG_CUDA(ctx, cudaGetLastError());
    }
    return st.finish();
}

extern "C" int lg_fine_to_coarse(lg_ctx* ctx, const uint64_t* codes, const uint32_t* group_of_cell, uint64_t ncols, uint32_t nfine,
                                 int coarse_dim, uint32_t* out_fine_to_coarse, uint32_t* out_num_coarse) {
    if (!ctx) return LG_ERR_INVALID;
    LG_REQUIRE(ctx, codes && group_of_cell && out_fine_to_coarse && out_num_coarse, "lg_fine_to_coarse: null argument");
    LG_REQUIRE(ctx, coarse_dim >= 0 && nfine >= 1, "lg_fine_to_coarse: bad shape");
    cudaSetDevice(ctx->device);
    LgStage st(ctx);
    const uint64_t* d_codes;
    const uint32_t* d_group;
    LG_TRY(st.in(codes, (size_t)ncols, &d_codes));
    LG_TRY(st.in(group_of_cell, (size_t)ncols, &d_group));
    unsigned long long* d_gc;
    LG_TRY(st.scratch(nfine, &d_gc));
    LG_CUDA(ctx, cudaMemsetAsync(d_gc, 0xFF, sizeof(unsigned long long) * nfine, ctx->stream));
    if (ncols) LG_LAUNCH(ctx, k_group_code, (unsigned)((ncols + 255) / 256), 256, 0, d_codes, d_group, ncols, nfine, d_gc);
    std::vector<unsigned long long> gc(nfine);
    LG_CUDA(ctx, cudaMemcpyAsync(gc.data(), d_gc, sizeof(unsigned long long) * nfine, cudaMemcpyDeviceToHost, ctx->stream));
    LG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const unsigned long long mask = coarse_dim >= 64 ? ~0ull : ((1ull << coarse_dim) - 1ull);
    std::vector<unsigned long long> uniq(nfine);
    for (uint32_t f = 0; f < nfine; ++f) {
        LG_REQUIRE(ctx, gc[f] != ~0ull, "lg_fine_to_coarse: a fine group has no cells");
        uniq[f] = gc[f] & mask;
    }
    std::sort(uniq.begin(), uniq.end());
    uniq.erase(std::unique(uniq.begin(), uniq.end()), uniq.end());
    std::vector<uint32_t> f2c(nfine);
    for (uint32_t f = 0; f < nfine; ++f) f2c[f] = (uint32_t)(std::lower_bound(uniq.begin(), uniq.end(), gc[f] & mask) - uniq.begin());
    *out_num_coarse = (uint32_t)uniq.size();
    LG_CUDA(ctx, cudaMemcpyAsync(out_fine_to_coarse, f2c.data(), sizeof(uint32_t) * nfine, cudaMemcpyDefault, ctx->stream));
    LG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return st.finish();
}

// ---- staged form of the pb-sample arm for cell-sharded runs (device pointers; the exchanges are the caller's) ----------
extern "C" int lg_pair_presence(lg_ctx* ctx, const uint32_t* d_group, const uint32_t* d_batch, uint64_t ncols, uint32_t S, uint32_t B,
                                uint32_t* d_present) {
    if (!ctx) return LG_ERR_INVALID;
    LG_REQUIRE(ctx, d_group && d_batch && d_present && S >= 1 && B >= 1, "lg_pair_presence: bad argument");
    cudaSetDevice(ctx->device);
    LG_CUDA(ctx, cudaMemsetAsync(d_present, 0, sizeof(uint32_t) * (size_t)S * B, ctx->stream));
    if (ncols) LG_LAUNCH(ctx, k_pair_presence, (unsigned)((ncols + 255) / 256), 256, 0, d_group, d_batch, ncols, S, B, d_present);
    return LG_OK;
}

// host math: pb-sample ids from the (all-reduced) presence flags, group-major with batches ascending
extern "C" int lg_pb_ids(lg_ctx* ctx, const uint32_t* present, uint32_t S, uint32_t B, uint32_t* out_id, uint32_t* out_pb_group,
                         uint32_t* out_pb_batch, uint32_t* out_num_pb) {
    if (!ctx) return LG_ERR_INVALID;
    LG_REQUIRE(ctx, present && out_id && out_pb_group && out_pb_batch && out_num_pb, "lg_pb_ids: null argument");
    LG_REQUIRE(ctx, !lg_is_device_ptr(present) && !lg_is_device_ptr(out_id), "lg_pb_ids is host math: pass host arrays");
    uint32_t n = 0;
    for (size_t e = 0; e < (size_t)S * B; ++e) {
        out_id[e] = NONE;
        if (present[e]) {
            out_id[e] = n;
            out_pb_group[n] = (uint32_t)(e / B);
            out_pb_batch[n] = (uint32_t)(e % B);
            ++n;
        }
    }
    *out_num_pb = n;
    return LG_OK;
}

extern "C" int lg_cells_to_pb(lg_ctx* ctx, const uint32_t* d_group, const uint32_t* d_batch, uint64_t ncols, uint32_t S, uint32_t B,
                              const uint32_t* d_id, uint32_t* d_cell_to_pb) {
    if (!ctx) return LG_ERR_INVALID;
    LG_REQUIRE(ctx, d_group && d_batch && d_id && d_cell_to_pb, "lg_cells_to_pb: null argument");
    cudaSetDevice(ctx->device);
    if (ncols) LG_LAUNCH(ctx, k_cell_to_pb, (unsigned)((ncols + 255) / 256), 256, 0, d_group, d_batch, ncols, S, B, d_id, d_cell_to_pb);
    return LG_OK;
}

// continues the serial centroid folds (sum of K-vectors x multiplicity, sum of multiplicities) over this shard's cells
extern "C" int lg_pb_centroid_fold(lg_ctx* ctx, const float* d_proj, int K, uint64_t ncols, const uint32_t* d_cell_to_pb, uint32_t npb,
                                   const float* d_mult, float* d_sum, float* d_count) {
    if (!ctx) return LG_ERR_INVALID;
    LG_REQUIRE(ctx, d_proj && d_cell_to_pb && d_sum && d_count && K >= 1 && K <= 1024 && npb >= 1 && ncols < 0xFFFFFFFFull,
               "lg_pb_centroid_fold: bad argument");
    cudaSetDevice(ctx->device);
    LgStage st(ctx);
    std::vector<uint32_t> counts;
    LG_TRY(label_counts_host(ctx, st, d_cell_to_pb, ncols, npb, counts));
    std::vector<uint64_t> off(npb + 1, 0);
    for (uint32_t p = 0; p < npb; ++p) off[p + 1] = off[p] + counts[p];
    uint32_t *d_lab, *d_cell;
    uint64_t* d_off;
    LG_TRY(sort_cells_by_label(ctx, st, d_cell_to_pb, ncols, &d_lab, &d_cell));
    LG_TRY(upload(ctx, st, off, &d_off));
    LG_LAUNCH(ctx, k_segment_mean<2>, npb, ((K + 31) / 32) * 32, 0, d_proj, K, d_cell, d_off, d_mult, d_sum, d_count);
    return st.finish();
}

extern "C" int lg_pb_centroid_finish(lg_ctx* ctx, const float* d_sum, const float* d_count, uint32_t npb, int K, float* d_centroids) {
    if (!ctx) return LG_ERR_INVALID;
    LG_REQUIRE(ctx, d_sum && d_count && d_centroids && K >= 1, "lg_pb_centroid_finish: bad argument");
    cudaSetDevice(ctx->device);
    if (npb) LG_LAUNCH(ctx, k_centroid_finish, (unsigned)(((uint64_t)npb * K + 255) / 256), 256, 0, d_sum, d_count, (uint64_t)npb, K, d_centroids);
    return LG_OK;
}

// keys[q - q0][p] over THIS shard's cells, cell indices offset to the global numbering (min-reducible across shards)
extern "C" int lg_pb_min_keys(lg_ctx* ctx, const float* d_proj, int K, uint64_t ncols, const uint32_t* d_cell_to_pb, uint64_t cell_offset,
                              const float* d_centroids, const uint32_t* d_pb_batch, uint32_t npb, uint32_t q0, uint32_t nq,
                              uint64_t* d_keys) {
    if (!ctx) return LG_ERR_INVALID;
    LG_REQUIRE(ctx, d_proj && d_cell_to_pb && d_centroids && d_pb_batch && d_keys, "lg_pb_min_keys: null argument");
    LG_REQUIRE(ctx, K >= 1 && K <= 256 && npb >= 1 && q0 + (uint64_t)nq <= npb && cell_offset + ncols < 0xFFFFFFFFull,
               "lg_pb_min_keys: bad shape");
    cudaSetDevice(ctx->device);
    LgStage st(ctx);
    std::vector<uint32_t> counts;
    LG_TRY(label_counts_host(ctx, st, d_cell_to_pb, ncols, npb, counts));
    std::vector<uint64_t> off(npb + 1, 0);
    for (uint32_t p = 0; p < npb; ++p) off[p + 1] = off[p] + counts[p];
    uint32_t *d_lab, *d_cell;
    uint64_t* d_off;
    LG_TRY(sort_cells_by_label(ctx, st, d_cell_to_pb, ncols, &d_lab, &d_cell));
    LG_TRY(upload(ctx, st, off, &d_off));
    const size_t smem = ((size_t)PM_Q * (K | 1) + (size_t)PM_CELLS * K) * sizeof(float);
    LG_REQUIRE(ctx, smem <= ctx->smem_optin, "lg_pb_min_keys: K too large for shared memory");
    LG_CUDA(ctx, cudaFuncSetAttribute(k_pb_min_dist, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (nq) {
        dim3 grid(npb, (nq + PM_Q - 1) / PM_Q);
        LG_TRY(launch_pb_min_dist(ctx, grid, d_proj, K, d_cell, d_off, d_centroids, d_pb_batch, npb, q0, nq, (uint32_t)cell_offset,
                                  reinterpret_cast<unsigned long long*>(d_keys)));
    }
    return st.finish();
}

extern "C" int lg_pb_topk_keys(lg_ctx* ctx, const uint64_t* d_keys, uint32_t npb, uint32_t q0, uint32_t nq, uint32_t B,
                               const uint32_t* pb_batch, int knn, uint32_t* d_out_pb, float* d_out_dist) {
    if (!ctx) return LG_ERR_INVALID;
    LG_REQUIRE(ctx, d_keys && pb_batch && d_out_pb && d_out_dist && knn >= 1 && B >= 1, "lg_pb_topk_keys: bad argument");
    cudaSetDevice(ctx->device);
    LgStage st(ctx);
    std::vector<uint32_t> pbb;
    LG_TRY(to_host(ctx, pb_batch, npb, pbb));
    std::vector<uint32_t> batch_off(B + 1, 0), batch_pb(npb);
    for (uint32_t p = 0; p < npb; ++p) {
        LG_REQUIRE(ctx, pbb[p] < B, "lg_pb_topk_keys: pb_batch out of range");
        batch_off[pbb[p] + 1]++;
    }
    for (uint32_t b = 0; b < B; ++b) batch_off[b + 1] += batch_off[b];
    {
        std::vector<uint32_t> cur(batch_off.begin(), batch_off.end() - 1);
        for (uint32_t p = 0; p < npb; ++p) batch_pb[cur[pbb[p]]++] = p;
    }
    uint32_t *d_batch_pb, *d_batch_off;
    LG_TRY(upload(ctx, st, batch_pb, &d_batch_pb));
    LG_TRY(upload(ctx, st, batch_off, &d_batch_off));
    if (nq)
        LG_LAUNCH(ctx, k_pb_topk, (unsigned)(((uint64_t)nq * B * 32 + 255) / 256), 256, 0, reinterpret_cast<const unsigned long long*>(d_keys),
                  npb, q0, nq, B, d_batch_pb, d_batch_off, knn, d_out_pb, d_out_dist);
    return st.finish();
}
