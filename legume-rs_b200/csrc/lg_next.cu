// lg_next.cu — the nnz streams either side of the hot path:
//   K10  per-gene running statistics   SparseRunningStatistics::add_csc  matrix-util/src/sparse_stat.rs:64-108,
//                                       streaming_sparse_running_stats    data-beans-alg/src/sparse_streaming.rs:23-60
//   K11  Nystrom re-projection          nystrom_proj_visitor              senna/src/svd/fit.rs:433-466
#include <cstdlib>

#include "lg_common.cuh"

namespace {
// 1 if every stored value is a non-negative whole number below 2^20 (same test as the collapse uses)
__global__ void k_rs_all_integral(const float* __restrict__ v, uint64_t n, int* __restrict__ not_integral) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    bool bad = false;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float x = __ldg(v + i);
        bad |= !(x >= 0.0f && x < 1048576.0f && x == truncf(x));
    }
    if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(not_integral, 1);
}
}  // namespace

// ---------------------------------------------------------------------------------------------
// K10: one pass over the nnz stream for (npos, s1, s2) per gene.
// Count data (every value a whole number in [0, 2^20), checked once per block and cached): a CTA keeps two
// shared-memory accumulators over the gene axis and every non-zero costs ONE native ATOMS.ADD on the first,
//   acc[g] += (1 << 21) | y               npos in the top 11 bits (<= 1024 cells between folds), s1 in the low 21
// and counts above one a second on a 16-bit field (two genes per word) holding y (y - 1), since
//   s2 = s1 + sum over y >= 2 of y (y - 1).
// Every RS_FOLD_CELLS cells the CTA folds both accumulators into its private slab in global memory with plain vector
// read-modify-writes; a last kernel adds the slabs in a fixed order.  Counts too large for the fields (y >= 2048
// resp. y >= 8, a ~1e-5 share) go to the slab directly with atomics.  Integer arithmetic throughout: exact,
// order-free, identical for any grid or GPU count.  Anything else (fractional / negative / non-finite values) takes
// the f64-atomic path below.
// ---------------------------------------------------------------------------------------------
constexpr int RS_THREADS = 1024;
constexpr int RS_CHUNK = 256;            // cells per work item (dynamic claim)
constexpr int RS_FOLD_CELLS = 1024;      // cells between folds: npos <= 1024 needs 11 bits
constexpr uint32_t RS_SHIFT = 21;        // s1 field: 1024 cells * (RS_YMAX - 1) < 2^21
constexpr uint32_t RS_YMAX = 2048;
constexpr uint32_t RS_XMAX = 8;          // y (y - 1) <= 42 for y < 8: 1024 cells * 42 < 2^16

template <bool VEC>
__global__ void __launch_bounds__(RS_THREADS, 1) k_row_stats_int(const uint64_t* __restrict__ indptr,
                                                                 const uint32_t* __restrict__ indices,
                                                                 const float* __restrict__ values, uint64_t nnz_total,
                                                                 uint64_t ncells, uint64_t D, uint32_t g0, uint32_t W,
                                                                 uint32_t* __restrict__ slab_npos,
                                                                 unsigned long long* __restrict__ slab_s1,
                                                                 unsigned long long* __restrict__ slab_extra,
                                                                 unsigned long long* __restrict__ next_chunk) {
    extern __shared__ __align__(16) uint32_t acc[];  // W packed (npos | s1), then ceil(W / 2) words of paired 16-bit extras
    __shared__ unsigned long long s_chunk;
    const uint32_t W2 = (W + 1) >> 1;
    uint32_t* ext = acc + ((W + 1) & ~1u);
    const uint64_t nchunks = (ncells + RS_CHUNK - 1) / RS_CHUNK;
    uint32_t* my_npos = slab_npos + (size_t)blockIdx.x * D + g0;
    unsigned long long* my_s1 = slab_s1 + (size_t)blockIdx.x * D + g0;
    unsigned long long* my_extra = slab_extra + (size_t)blockIdx.x * D + g0;
    for (uint32_t g = threadIdx.x; g < ((W + 1) & ~1u) + W2; g += RS_THREADS) acc[g] = 0;
    auto one = [&](uint32_t g, float v) {
        const uint32_t ge = g - g0;
        if (ge >= W) return;
        const uint32_t y = (uint32_t)v;
        if (y - 1u < RS_YMAX - 1u) {  // 1 <= y < RS_YMAX
            atomicAdd(&acc[ge], (1u << RS_SHIFT) | y);
        } else if (y) {  // too large for the packed field (a stored zero adds nothing)
            atomicAdd(&my_npos[ge], 1u);
            atomicAdd(&my_s1[ge], (unsigned long long)y);
        }
        if (y >= 2) {
            if (y < RS_XMAX) atomicAdd(&ext[ge >> 1], (y * (y - 1)) << ((ge & 1u) * 16));
            else atomicAdd(&my_extra[ge], (unsigned long long)y * (y - 1));
        }
    };
    auto fold = [&]() {  // shared accumulators -> this CTA's slab (plain read-modify-writes: nobody else owns it)
        __syncthreads();
        for (uint32_t w = threadIdx.x; w < W2; w += RS_THREADS) {
            const uint32_t a0 = acc[2 * w], a1 = (2 * w + 1 < W) ? acc[2 * w + 1] : 0u, e = ext[w];
            if (a0) {
                my_npos[2 * w] += a0 >> RS_SHIFT;
                my_s1[2 * w] += a0 & ((1u << RS_SHIFT) - 1u);
                acc[2 * w] = 0;
            }
            if (a1) {
                my_npos[2 * w + 1] += a1 >> RS_SHIFT;
                my_s1[2 * w + 1] += a1 & ((1u << RS_SHIFT) - 1u);
                acc[2 * w + 1] = 0;
            }
            if (e) {
                if (e & 0xffffu) my_extra[2 * w] += e & 0xffffu;
                if (e >> 16) my_extra[2 * w + 1] += e >> 16;
                ext[w] = 0;
            }
        }
        __syncthreads();
    };
    uint32_t pending = 0;  // cells accumulated since the last fold
    while (true) {
        __syncthreads();
        if (threadIdx.x == 0) s_chunk = atomicAdd(next_chunk, 1ull);
        __syncthreads();
        const uint64_t chunk = s_chunk;
        if (chunk >= nchunks) break;
        if (pending + RS_CHUNK > RS_FOLD_CELLS) {
            fold();
            pending = 0;
        }
        pending += RS_CHUNK;
        const uint64_t c0 = chunk * RS_CHUNK;
        const uint64_t c1 = (c0 + RS_CHUNK) < ncells ? (c0 + RS_CHUNK) : ncells;
        // the chunk's cells are consecutive columns: one contiguous nnz range, walked by all warps together
        const uint64_t lo = indptr[c0], hi = indptr[c1];
        if constexpr (VEC) {
            constexpr int VU = 4;
            const uint64_t hi4 = nnz_total & ~3ull;
            for (uint64_t c = (lo & ~3ull) + 4ull * threadIdx.x; c < hi; c += 4ull * RS_THREADS * VU) {
                uint4 gq[VU];
                float4 vq[VU];
#pragma unroll
                for (int u = 0; u < VU; ++u) {
                    const uint64_t cu = c + 4ull * RS_THREADS * u;
                    gq[u] = make_uint4(0, 0, 0, 0);
                    vq[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (cu < hi) {
                        if (cu < hi4) {
                            gq[u] = __ldg(reinterpret_cast<const uint4*>(indices + cu));
                            vq[u] = __ldg(reinterpret_cast<const float4*>(values + cu));
                        } else {  // the last, partial group of the whole array
                            uint32_t gg[4] = {0, 0, 0, 0};
                            float vv[4] = {0.f, 0.f, 0.f, 0.f};
                            for (int e = 0; e < 4; ++e)
                                if (cu + e < nnz_total) {
                                    gg[e] = indices[cu + e];
                                    vv[e] = values[cu + e];
                                }
                            gq[u] = make_uint4(gg[0], gg[1], gg[2], gg[3]);
                            vq[u] = make_float4(vv[0], vv[1], vv[2], vv[3]);
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < VU; ++u) {
                    const uint64_t cu = c + 4ull * RS_THREADS * u;
                    const uint32_t ge[4] = {gq[u].x, gq[u].y, gq[u].z, gq[u].w};
                    const float ve[4] = {vq[u].x, vq[u].y, vq[u].z, vq[u].w};
                    if (cu >= lo && cu + 3 < hi) {  // interior group: no per-entry range test
#pragma unroll
                        for (int e = 0; e < 4; ++e) one(ge[e], ve[e]);
                    } else {
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            if (cu + e >= lo && cu + e < hi) one(ge[e], ve[e]);
                    }
                }
            }
        } else {
            for (uint64_t t = lo + threadIdx.x; t < hi; t += RS_THREADS) one(__ldg(indices + t), __ldg(values + t));
        }
    }
    fold();
}

__global__ void k_row_stats_finish(const uint32_t* __restrict__ slab_npos, const unsigned long long* __restrict__ slab_s1,
                                   const unsigned long long* __restrict__ slab_extra, uint64_t D, uint32_t nslab,
                                   double* __restrict__ npos, double* __restrict__ s1, double* __restrict__ s2) {
    const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= D) return;
    unsigned long long a = 0, b = 0, x = 0;
    for (uint32_t s = 0; s < nslab; ++s) {
        a += slab_npos[(size_t)s * D + g];
        b += slab_s1[(size_t)s * D + g];
        x += slab_extra[(size_t)s * D + g];
    }
    npos[g] = (double)a;
    s1[g] = (double)b;
    s2[g] = (double)(b + x);
}

constexpr int RS_REPL = 8;  // replicas of the f64 accumulators (spreads same-address atomics)
// general values: f64 atomics on replicated accumulators (sums agree to ~1e-16 relative run to run; non-finite skipped)
__global__ void __launch_bounds__(256) k_row_stats_f64(const uint32_t* __restrict__ indices, const float* __restrict__ values,
                                                       uint64_t nnz, uint64_t D, double* __restrict__ rep) {
    double* mine = rep + (size_t)(blockIdx.x % RS_REPL) * 3 * D;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < nnz; t += stride) {
        const float v = __ldg(values + t);
        if (!isfinite(v)) continue;
        const uint32_t g = __ldg(indices + t);
        if (v > 0.0f) atomicAdd(&mine[g], 1.0);
        atomicAdd(&mine[D + g], (double)v);
        atomicAdd(&mine[2 * D + g], (double)v * (double)v);
    }
}
__global__ void k_row_stats_f64_finish(const double* __restrict__ rep, uint64_t D, double* __restrict__ npos,
                                       double* __restrict__ s1, double* __restrict__ s2) {
    const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= D) return;
    double a = 0, b = 0, c = 0;
    for (int r = 0; r < RS_REPL; ++r) {
        a += rep[(size_t)r * 3 * D + g];
        b += rep[(size_t)r * 3 * D + D + g];
        c += rep[(size_t)r * 3 * D + 2 * D + g];
    }
    npos[g] = a;
    s1[g] = b;
    s2[g] = c;
}

extern "C" int lg_row_stats(lg_ctx* ctx, const lg_csc* m, double* out_npos, double* out_s1, double* out_s2) {
    if (!ctx) return LG_ERR_INVALID;
    LG_REQUIRE(ctx, m && out_npos && out_s1 && out_s2, "lg_row_stats: null argument");
    cudaSetDevice(ctx->device);
    LG_TRY(lg_csc_require_canonical(ctx, m, "lg_row_stats"));
    const uint64_t D = m->nrows, N = m->ncols;
    LgStage st(ctx);
    double *d_npos, *d_s1, *d_s2;
    LG_TRY(st.out(out_npos, D, &d_npos));
    LG_TRY(st.out(out_s1, D, &d_s1));
    LG_TRY(st.out(out_s2, D, &d_s2));
    if (D == 0) return st.finish();
    if (N == 0 || m->nnz == 0) {
        LG_CUDA(ctx, cudaMemsetAsync(d_npos, 0, D * sizeof(double), ctx->stream));
        LG_CUDA(ctx, cudaMemsetAsync(d_s1, 0, D * sizeof(double), ctx->stream));
        LG_CUDA(ctx, cudaMemsetAsync(d_s2, 0, D * sizeof(double), ctx->stream));
        return st.finish();
    }
    if (m->int_valued < 0) {
        int* d_flag;
        LG_TRY(st.scratch(1, &d_flag));
        LG_CUDA(ctx, cudaMemsetAsync(d_flag, 0, sizeof(int), ctx->stream));
        LG_LAUNCH(ctx, k_rs_all_integral, ctx->num_sms * 8, 256, 0, m->values, m->nnz, d_flag);
        int* h_flag = static_cast<int*>(ctx->pinned);
        LG_CUDA(ctx, cudaMemcpyAsync(h_flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        LG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        m->int_valued = (*h_flag == 0) ? 1 : 0;
    }
    if (m->int_valued == 1) {
        const uint32_t nslab = (uint32_t)ctx->num_sms;
        uint32_t* d_slab_npos;
        unsigned long long *d_slab_s1, *d_slab_extra, *d_next;
        LG_TRY(st.scratch((size_t)nslab * D, &d_slab_npos));
        LG_TRY(st.scratch((size_t)nslab * D, &d_slab_s1));
        LG_TRY(st.scratch((size_t)nslab * D, &d_slab_extra));
        LG_TRY(st.scratch(1, &d_next));
        LG_CUDA(ctx, cudaMemsetAsync(d_slab_npos, 0, (size_t)nslab * D * 4, ctx->stream));
        LG_CUDA(ctx, cudaMemsetAsync(d_slab_s1, 0, (size_t)nslab * D * 8, ctx->stream));
        LG_CUDA(ctx, cudaMemsetAsync(d_slab_extra, 0, (size_t)nslab * D * 8, ctx->stream));
        // 6 bytes of shared memory per gene: one pass when the gene axis fits (D <= ~38k), else gene windows
        const uint32_t Wmax = (uint32_t)(((ctx->smem_optin - 1024) / 6) & ~1ull);
        const bool vec = (((uintptr_t)m->indices | (uintptr_t)m->values) & 15) == 0;
        for (uint64_t g0 = 0; g0 < D; g0 += Wmax) {
            const uint32_t W = (uint32_t)((D - g0) < Wmax ? (D - g0) : Wmax);
            const size_t smem = ((size_t)((W + 1) & ~1u) + ((W + 1) >> 1)) * sizeof(uint32_t);
            LG_CUDA(ctx, cudaMemsetAsync(d_next, 0, sizeof(unsigned long long), ctx->stream));
            if (vec) {
                LG_CUDA(ctx, cudaFuncSetAttribute(k_row_stats_int<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                LG_LAUNCH(ctx, k_row_stats_int<true>, nslab, RS_THREADS, smem, m->indptr, m->indices, m->values, m->nnz, N, D,
                          (uint32_t)g0, W, d_slab_npos, d_slab_s1, d_slab_extra, d_next);
            } else {
                LG_CUDA(ctx, cudaFuncSetAttribute(k_row_stats_int<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                LG_LAUNCH(ctx, k_row_stats_int<false>, nslab, RS_THREADS, smem, m->indptr, m->indices, m->values, m->nnz, N, D,
                          (uint32_t)g0, W, d_slab_npos, d_slab_s1, d_slab_extra, d_next);
            }
        }
        LG_LAUNCH(ctx, k_row_stats_finish, (unsigned)((D + 255) / 256), 256, 0, d_slab_npos, d_slab_s1, d_slab_extra, D, nslab,
                  d_npos, d_s1, d_s2);
    } else {
        double* d_rep;
        LG_TRY(st.scratch((size_t)RS_REPL * 3 * D, &d_rep));
        LG_CUDA(ctx, cudaMemsetAsync(d_rep, 0, (size_t)RS_REPL * 3 * D * sizeof(double), ctx->stream));
        LG_LAUNCH(ctx, k_row_stats_f64, ctx->num_sms * 8, 256, 0, m->indices, m->values, m->nnz, D, d_rep);
        LG_LAUNCH(ctx, k_row_stats_f64_finish, (unsigned)((D + 255) / 256), 256, 0, d_rep, D, d_npos, d_s1, d_s2);
    }
    return st.finish();
}

// ---------------------------------------------------------------------------------------------
// K11: Nystrom re-projection, one warp per cell (senna/src/svd/fit.rs:433-466).  Per cell j with stored counts y:
//   x = y / max(||y||, 1e-8) * c                               normalize_columns_inplace, *= column_sum_norm
//   d = delta[row, pb(j)];  x /= d * (sum x / sum d)  where d > 0    adjust_by_poisson_ratio (dmatrix_util.rs:226-244)
//   z = ln(1 + x);  z = (z - mean z) / sd z  over the stored entries (CSC scale_columns_inplace :791-824; z - mean when sd = 0)
//   out[:, j] = sum_i z_i basis[i, :]
// The reference's per-cell scalars are ill-conditioned in f32: sd^2 = s2/n - mean^2 cancels about three digits (z ~ 5.5,
// sd ~ 0.2), so ONE ulp of difference in a log1p (CUDA's vs glibc's) already moves every output of a cell by ~1e-4, and
// the reference's own serial f32 folds sit up to ~5e-4 from exact arithmetic.  Bit-chasing that is meaningless, so the
// per-cell sums are accumulated in f64 here (lane partials + a butterfly): the result is within 1e-5 of the float64
// restatement of the reference's formulas, which is what the tests hold it to (and within the reference's own error of
// the f32 oracle).  Sweeps re-read the cell's entries from L1/L2: A (sum y^2, sum y, sum d), C (sum z, sum z^2),
// D gather-FMA of the basis rows (two entries at a time, one per half-warp, as float2).
// ---------------------------------------------------------------------------------------------
constexpr int NY_WARPS = 8;

template <bool HAS_DELTA>
__global__ void __launch_bounds__(NY_WARPS * 32) k_nystrom(const uint64_t* __restrict__ indptr, const uint32_t* __restrict__ indices,
                                                          const float* __restrict__ values, uint64_t ncols, uint64_t D,
                                                          const float* __restrict__ basis_kd, int K,
                                                          const float* __restrict__ delta_dp, const uint32_t* __restrict__ pb_of_cell,
                                                          uint32_t P, float csn, float* __restrict__ out) {
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const uint64_t warp0 = (uint64_t)blockIdx.x * NY_WARPS + wib;
    const uint64_t nwarps = (uint64_t)gridDim.x * NY_WARPS;
    const int half = lane >> 4, l = lane & 15;
    for (uint64_t j = warp0; j < ncols; j += nwarps) {
        const uint64_t lo = indptr[j], hi = indptr[j + 1];
        const float* dcol = nullptr;
        if (HAS_DELTA) {
            const uint32_t pb = pb_of_cell[j];
            dcol = pb < P ? delta_dp + (size_t)pb * D : nullptr;  // an unassigned cell is left unadjusted
        }
        // sweep A: sum y^2, sum y, sum d
        double sq = 0.0, sy = 0.0, sd = 0.0;
        for (uint64_t t = lo + lane; t < hi; t += 32) {
            const double y = (double)__ldg(values + t);
            sq = fma(y, y, sq);
            sy += y;
            if (HAS_DELTA && dcol) sd += (double)__ldg(dcol + __ldg(indices + t));
        }
        sq = lg_butterfly32(sq);
        sy = lg_butterfly32(sy);
        if (HAS_DELTA) sd = lg_butterfly32(sd);
        const float denom = fmaxf((float)sqrt(sq), 1e-8f);
        // sum x = sum (y / denom) * c; the ratio xsum / dsum is well conditioned
        const float ratio = (HAS_DELTA && dcol && sd > 0.0) ? (float)(sy / (double)denom * (double)csn / sd) : 1.0f;
        auto zval = [&](uint64_t t, uint32_t g) {
            float x = __fmul_rn(__fdiv_rn(__ldg(values + t), denom), csn);
            if (HAS_DELTA && dcol) {
                const float d = __ldg(dcol + g);
                if (d > 0.0f) x = __fdiv_rn(x, __fmul_rn(d, ratio));
            }
            return log1pf(x);
        };
        // sweep C: moments of z over the stored entries
        double s1 = 0.0, s2 = 0.0;
        for (uint64_t t = lo + lane; t < hi; t += 32) {
            const double z = (double)zval(t, (HAS_DELTA && dcol) ? __ldg(indices + t) : 0u);
            s1 += z;
            s2 = fma(z, z, s2);
        }
        s1 = lg_butterfly32(s1);
        s2 = lg_butterfly32(s2);
        const double nn = (double)(hi - lo) > 1.0 ? (double)(hi - lo) : 1.0;
        const double mud = s1 / nn;
        const double var = s2 / nn - mud * mud;
        // a constant column (variance zero up to f64 rounding) is only centred (dmatrix_util.rs:815-819)
        const double inv_sig = var > 1e-12 * (s2 / nn) ? 1.0 / sqrt(var) : 1.0;
        // sweep D: out = sum_i w_i basis[i, :].  Each batch of 32 entries is summed in f32 and folded into an f64 total:
        // a plain f32 chain over ~1500 terms of size O(1) drifts by ~4e-5 of the result (the reference's serial f32 sum
        // does too), which would be the largest error left in the path
        double tot[4] = {0.0, 0.0, 0.0, 0.0};
        const bool pairs = (K & 1) == 0 && K <= 64;
        for (uint64_t base = lo; base < hi; base += 32) {
            const uint64_t t = base + lane;
            uint32_t g = 0;
            float w = 0.0f;
            if (t < hi) {
                g = __ldg(indices + t);
                const float z = zval(t, g);
                w = (float)(((double)z - mud) * inv_sig);  // z - mean cancels ~2 digits: centred in f64
            }
            const int cnt = (hi - base) < 32 ? (int)(hi - base) : 32;
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
            if (pairs) {
                for (int e0 = 0; e0 < cnt; e0 += 8) {
                    float2 b0[4], b1[4];
                    float ws[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int src = e0 + 2 * e + half;  // lanes beyond cnt carry w = 0 and gene 0
                        const uint32_t ge = __shfl_sync(0xffffffffu, g, src);
                        ws[e] = __shfl_sync(0xffffffffu, w, src);
                        const float2* brow = reinterpret_cast<const float2*>(basis_kd + (size_t)ge * K);
                        b0[e] = (2 * l < K) ? __ldg(brow + l) : make_float2(0.f, 0.f);
                        b1[e] = (2 * (l + 16) < K) ? __ldg(brow + l + 16) : make_float2(0.f, 0.f);
                    }
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        acc[0] = fmaf(ws[e], b0[e].x, acc[0]);
                        acc[1] = fmaf(ws[e], b0[e].y, acc[1]);
                        acc[2] = fmaf(ws[e], b1[e].x, acc[2]);
                        acc[3] = fmaf(ws[e], b1[e].y, acc[3]);
                    }
                }
            } else {  // lanes as dims, up to 128 dims
                for (int e = 0; e < cnt; ++e) {
                    const uint32_t ge = __shfl_sync(0xffffffffu, g, e);
                    const float we = __shfl_sync(0xffffffffu, w, e);
                    const float* brow = basis_kd + (size_t)ge * K;
#pragma unroll
                    for (int a = 0; a < 4; ++a)
                        if (lane + 32 * a < K) acc[a] = fmaf(we, __ldg(brow + lane + 32 * a), acc[a]);
                }
            }
#pragma unroll
            for (int a = 0; a < 4; ++a) tot[a] += (double)acc[a];
        }
        float* o = out + (size_t)j * K;
        if (pairs) {
#pragma unroll
            for (int a = 0; a < 4; ++a) tot[a] += __shfl_xor_sync(0xffffffffu, tot[a], 16);
            if (lane < 16) {
                if (2 * l < K) {
                    o[2 * l] = (float)tot[0];
                    o[2 * l + 1] = (float)tot[1];
                }
                if (2 * (l + 16) < K) {
                    o[2 * l + 32] = (float)tot[2];
                    o[2 * l + 33] = (float)tot[3];
                }
            }
        } else {
#pragma unroll
            for (int a = 0; a < 4; ++a)
                if (lane + 32 * a < K) o[lane + 32 * a] = (float)tot[a];
        }
    }
}

int lg_project_raw_umma(lg_ctx* ctx, const lg_csc* m, const float* d_basis, int K, float* d_out, int* used, int mode, float csn);  // lg_project_umma.cu

// basis_dk (D x K column-major, the reference's DMatrix) -> rows of K contiguous dims
__global__ void k_transpose_dk(const float* __restrict__ src_dk, uint64_t D, int K, float* __restrict__ dst_kd) {
    const uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= D * (uint64_t)K) return;
    const uint64_t g = e / K;
    const int k = (int)(e % K);
    dst_kd[e] = src_dk[(size_t)k * D + g];
}

extern "C" int lg_nystrom_project(lg_ctx* ctx, const lg_csc* m, const float* basis_dk, int K, const float* delta_dp,
                                  const uint32_t* pb_of_cell, uint32_t P, float column_sum_norm, float* out_proj_kn) {
    if (!ctx) return LG_ERR_INVALID;
    LG_REQUIRE(ctx, m && basis_dk && out_proj_kn, "lg_nystrom_project: null argument");
    LG_REQUIRE(ctx, K >= 1 && K <= 128, "lg_nystrom_project: K must be in [1, 128]");
    LG_REQUIRE(ctx, !delta_dp || (pb_of_cell && P >= 1), "lg_nystrom_project: delta needs the pseudobulk of every cell");
    cudaSetDevice(ctx->device);
    LG_TRY(lg_csc_require_canonical(ctx, m, "lg_nystrom_project"));
    const uint64_t D = m->nrows, N = m->ncols;
    LgStage st(ctx);
    const float *d_basis, *d_delta;
    const uint32_t* d_pb;
    float *d_out, *d_bt;
    LG_TRY(st.in(basis_dk, (size_t)D * K, &d_basis));
    LG_TRY(st.in(delta_dp, delta_dp ? (size_t)D * P : 0, &d_delta));
    LG_TRY(st.in(pb_of_cell, delta_dp ? (size_t)N : 0, &d_pb));
    LG_TRY(st.out(out_proj_kn, (size_t)K * N, &d_out));
    if (N == 0 || D == 0) return st.finish();
    LG_TRY(st.scratch((size_t)D * K, &d_bt));
    LG_LAUNCH(ctx, k_transpose_dk, (unsigned)((D * K + 255) / 256), 256, 0, d_basis, D, K, d_bt);
    // without a batch divisor every count of one maps to the same value, so the tensor path of K1 applies: the 0/1
    // pattern against the quantised basis on tcgen05, the counts above one on CUDA cores (LG_K11_CUDA_CORES=1: A/B runs)
    const char* force = getenv("LG_K11_CUDA_CORES");
    if (!d_delta && !(force && force[0] == '1')) {
        int used = 0;
        LG_TRY(lg_project_raw_umma(ctx, m, d_bt, K, d_out, &used, 1, column_sum_norm));
        if (used) return st.finish();
    }
    uint64_t blocks = (N + NY_WARPS - 1) / NY_WARPS;
    const uint64_t cap = (uint64_t)ctx->num_sms * 32;
    if (blocks > cap) blocks = cap;
    if (d_delta)
        LG_LAUNCH(ctx, k_nystrom<true>, (unsigned)blocks, NY_WARPS * 32, 0, m->indptr, m->indices, m->values, N, D, d_bt, K, d_delta, d_pb, P,
                  column_sum_norm, d_out);
    else
        LG_LAUNCH(ctx, k_nystrom<false>, (unsigned)blocks, NY_WARPS * 32, 0, m->indptr, m->indices, m->values, N, D, d_bt, K,
                  (const float*)nullptr, (const uint32_t*)nullptr, 0u, column_sum_norm, d_out);
    return st.finish();
}
