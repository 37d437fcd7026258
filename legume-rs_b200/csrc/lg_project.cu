// lg_project.cu — stage 1: random projection of the sparse gene x cell matrix.
//   K1  project_columns_visitor       data-beans-alg/src/random_projection.rs:169-199
//   K2  batch centring / standardise / clamp                         :378-407
#include <cmath>
#include <cstdlib>

#include "lg_common.cuh"

// ---------------------------------------------------------------------------------------------
// K1 (baseline CUDA-core form): one warp per cell.  Lanes load 32 nnz at a time with coalesced
// loads, log1p them, then each nnz's (row, x) pair is broadcast and every lane accumulates its
// NACC dims of the K-vector.  The 1/||x|| factor is applied once at the end (the reference
// divides every x first; the difference is rounding only, inside the 1e-5 contract).
// ---------------------------------------------------------------------------------------------
template <int NACC>
__global__ void __launch_bounds__(256) k_project_raw_warp(const uint64_t* __restrict__ indptr,
                                                          const uint32_t* __restrict__ indices,
                                                          const float* __restrict__ values, uint64_t ncols,
                                                          const float* __restrict__ basis_kd, int K,
                                                          float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const uint64_t warp0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t j = warp0; j < ncols; j += nwarps) {
        const uint64_t lo = indptr[j], hi = indptr[j + 1];
        float acc[NACC];
#pragma unroll
        for (int a = 0; a < NACC; ++a) acc[a] = 0.0f;
        float nsq = 0.0f;
        for (uint64_t base = lo; base < hi; base += 32) {
            const uint64_t t = base + lane;
            uint32_t idx = 0;
            float x = 0.0f;
            if (t < hi) {
                idx = __ldg(indices + t);
                x = log1pf(__ldg(values + t));
            }
            nsq = fmaf(x, x, nsq);
            const int n = (hi - base) < 32 ? (int)(hi - base) : 32;
            for (int s = 0; s < n; ++s) {
                const uint32_t i = __shfl_sync(0xffffffffu, idx, s);
                const float xs = __shfl_sync(0xffffffffu, x, s);
                const float* row = basis_kd + (size_t)i * K;
#pragma unroll
                for (int a = 0; a < NACC; ++a) {
                    const int k = lane + 32 * a;
                    if (k < K) acc[a] = fmaf(xs, __ldg(row + k), acc[a]);
                }
            }
        }
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) nsq += __shfl_xor_sync(0xffffffffu, nsq, off);
        const float denom = fmaxf(sqrtf(nsq), 1e-8f);
#pragma unroll
        for (int a = 0; a < NACC; ++a) {
            const int k = lane + 32 * a;
            if (k < K) out[(size_t)j * K + k] = acc[a] / denom;
        }
    }
}


// ---------------------------------------------------------------------------------------------
// K1, EXACT-ORDER form (LG_PROJECT_EXACT): the reference's arithmetic operation by operation, so that the
// projection of count data is bit-identical to the CPU path and the sign bits / groups derived from it are too:
//   x = ln_1p(y)                               random_projection.rs:181-183   (libm log1pf: a host-built table for whole
//                                                                              counts below 65536, so no second libm is involved)
//   denom = max(sqrt(fold(x*x)), 1e-8); x /= denom   dmatrix_util.rs:770-778   (sequential fold, product rounded, then the sum)
//   chunk[:, j] = x_i * B[:, i] + chunk[:, j]   random_projection.rs:188-194   (ascending row, product rounded, then the sum)
// ---------------------------------------------------------------------------------------------
constexpr int LG_LOG1P_TAB = 65536;

__device__ __forceinline__ float exact_log1p(float y, const float* __restrict__ tab) {
    const int yi = (int)y;
    if (y == (float)yi && yi >= 0 && yi < LG_LOG1P_TAB) return __ldg(tab + yi);
    return (float)log1p((double)y);  // correctly rounded from f64: what libm's log1pf returns in all but rare cases
}

// one thread per cell: the norm is ONE dependent chain per cell, so a warp runs 32 cells' chains side by side
__global__ void __launch_bounds__(256) k_cell_norm_exact(const uint64_t* __restrict__ indptr, const float* __restrict__ values,
                                                         uint64_t ncols, const float* __restrict__ tab, float* __restrict__ denom) {
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= ncols) return;
    const uint64_t lo = indptr[j], hi = indptr[j + 1];
    float s = 0.0f;
    for (uint64_t t = lo; t < hi; ++t) {
        const float x = exact_log1p(__ldg(values + t), tab);
        s = __fadd_rn(s, __fmul_rn(x, x));
    }
    denom[j] = fmaxf(__fsqrt_rn(s), 1e-8f);
}

template <int NACC>
__global__ void __launch_bounds__(256) k_project_raw_exact(const uint64_t* __restrict__ indptr, const uint32_t* __restrict__ indices,
                                                           const float* __restrict__ values, uint64_t ncols,
                                                           const float* __restrict__ basis_kd, int K, const float* __restrict__ tab,
                                                           const float* __restrict__ denom, float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const uint64_t warp0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t j = warp0; j < ncols; j += nwarps) {
        const uint64_t lo = indptr[j], hi = indptr[j + 1];
        const float dn = denom[j];
        float acc[NACC];
#pragma unroll
        for (int a = 0; a < NACC; ++a) acc[a] = 0.0f;
        for (uint64_t base = lo; base < hi; base += 32) {
            const uint64_t t = base + lane;
            uint32_t idx = 0;
            float x = 0.0f;
            if (t < hi) {
                idx = __ldg(indices + t);
                x = __fdiv_rn(exact_log1p(__ldg(values + t), tab), dn);
            }
            const int n = (hi - base) < 32 ? (int)(hi - base) : 32;
            for (int s = 0; s < n; ++s) {
                const uint32_t i = __shfl_sync(0xffffffffu, idx, s);
                const float xs = __shfl_sync(0xffffffffu, x, s);
                const float* row = basis_kd + (size_t)i * K;
#pragma unroll
                for (int a = 0; a < NACC; ++a) {
                    const int k = lane + 32 * a;
                    if (k < K) acc[a] = __fadd_rn(__fmul_rn(xs, __ldg(row + k)), acc[a]);
                }
            }
        }
#pragma unroll
        for (int a = 0; a < NACC; ++a) {
            const int k = lane + 32 * a;
            if (k < K) out[(size_t)j * K + k] = acc[a];
        }
    }
}

// the table is libm's: built on the host once per context
static int exact_log1p_table(lg_ctx* ctx, const float** tab) {
    if (!ctx->log1p_tab) {
        std::vector<float> h(LG_LOG1P_TAB);
        for (int i = 0; i < LG_LOG1P_TAB; ++i) h[i] = log1pf((float)i);
        LG_CUDA(ctx, cudaMalloc(&ctx->log1p_tab, LG_LOG1P_TAB * sizeof(float)));
        LG_CUDA(ctx, cudaMemcpy(ctx->log1p_tab, h.data(), LG_LOG1P_TAB * sizeof(float), cudaMemcpyHostToDevice));
    }
    *tab = ctx->log1p_tab;
    return LG_OK;
}

static int launch_project_raw_exact(lg_ctx* ctx, const lg_csc* m, const float* d_basis, int K, float* d_out) {
    if (m->ncols == 0) return LG_OK;
    LG_TRY(lg_csc_require_canonical(ctx, m, "lg_project_exact"));  // "ascending row" is the order of the sum
    const float* tab;
    LG_TRY(exact_log1p_table(ctx, &tab));
    LgStage st(ctx);
    float* d_denom;
    LG_TRY(st.scratch((size_t)m->ncols, &d_denom));
    LG_LAUNCH(ctx, k_cell_norm_exact, (unsigned)((m->ncols + 255) / 256), 256, 0, m->indptr, m->values, m->ncols, tab, d_denom);
    const int nacc = (K + 31) / 32;
    uint64_t blocks = (m->ncols + 7) / 8;
    const uint64_t cap = (uint64_t)ctx->num_sms * 32;
    if (blocks > cap) blocks = cap;
#define LG_EXACT_CASE(NA)                                                                                                  \
    case NA:                                                                                                               \
        LG_LAUNCH(ctx, k_project_raw_exact<NA>, (unsigned)blocks, 256, 0, m->indptr, m->indices, m->values, m->ncols, d_basis, \
                  K, tab, d_denom, d_out);                                                                                 \
        break
    switch (nacc) {
        LG_EXACT_CASE(1);
        LG_EXACT_CASE(2);
        LG_EXACT_CASE(3);
        LG_EXACT_CASE(4);
        default: return lg_fail(ctx, LG_ERR_INVALID, "lg_project: K must be in [1, 128]");
    }
#undef LG_EXACT_CASE
    return LG_OK;
}

int lg_project_raw_umma(lg_ctx* ctx, const lg_csc* m, const float* d_basis, int K, float* d_out, int* used, int mode, float csn);

static int launch_project_raw(lg_ctx* ctx, const lg_csc* m, const float* d_basis, int K, float* d_out) {
    if (m->ncols == 0) return LG_OK;
    LG_TRY(lg_csc_require_canonical(ctx, m, "lg_project"));  // the pattern bitmap needs unique rows
    // tensor path first (lg_project_umma.cu); LG_K1_CUDA_CORES=1 forces the warp-per-cell kernel for A/B runs
    const char* force = getenv("LG_K1_CUDA_CORES");
    if (!(force && force[0] == '1')) {
        int used = 0;
        LG_TRY(lg_project_raw_umma(ctx, m, d_basis, K, d_out, &used, 0, 0.0f));
        if (used) return LG_OK;
    }
    const int nacc = (K + 31) / 32;
    uint64_t warps = m->ncols;
    uint64_t blocks = (warps + 7) / 8;
    const uint64_t cap = (uint64_t)ctx->num_sms * 32;
    if (blocks > cap) blocks = cap;
    switch (nacc) {
        case 1: LG_LAUNCH(ctx, k_project_raw_warp<1>, (unsigned)blocks, 256, 0, m->indptr, m->indices, m->values, m->ncols, d_basis, K, d_out); break;
        case 2: LG_LAUNCH(ctx, k_project_raw_warp<2>, (unsigned)blocks, 256, 0, m->indptr, m->indices, m->values, m->ncols, d_basis, K, d_out); break;
        case 3: LG_LAUNCH(ctx, k_project_raw_warp<3>, (unsigned)blocks, 256, 0, m->indptr, m->indices, m->values, m->ncols, d_basis, K, d_out); break;
        case 4: LG_LAUNCH(ctx, k_project_raw_warp<4>, (unsigned)blocks, 256, 0, m->indptr, m->indices, m->values, m->ncols, d_basis, K, d_out); break;
        default: return lg_fail(ctx, LG_ERR_INVALID, "lg_project: K must be in [1, 128]");
    }
    return LG_OK;
}

extern "C" int lg_project_raw(lg_ctx* ctx, const lg_csc* m, const float* basis_kd, int K, float* out_proj) {
    if (!ctx) return LG_ERR_INVALID;
    LG_REQUIRE(ctx, m && basis_kd && out_proj, "lg_project_raw: null argument");
    LG_REQUIRE(ctx, K >= 1 && K <= 128, "lg_project_raw: K must be in [1, 128]");
    cudaSetDevice(ctx->device);
    LgStage st(ctx);
    const float* d_basis;
    float* d_out;
    LG_TRY(st.in(basis_kd, (size_t)K * m->nrows, &d_basis));
    LG_TRY(st.out(out_proj, (size_t)K * m->ncols, &d_out));
    LG_TRY(launch_project_raw(ctx, m, d_basis, K, d_out));
    return st.finish();
}

// ---------------------------------------------------------------------------------------------
// K2a: per-1024-cell-block f64 partial sums of proj per (batch, dim) + cell counts.
// One block = LG_BLOCK_CELLS threads = one cell each.  partials[blk][b*(K+1) + k].
// ---------------------------------------------------------------------------------------------
// STAGED: the block's 1024 x K tile (one contiguous run of the projection) is first brought into shared memory with
// 128-bit streams and a row stride of K | 1 words; the per-value reads of a thread's own row are then conflict-free
// shared-memory loads instead of 32-sector global ones (one per lane and value: 0.33 -> 0.1 ms at 1M cells).  Same sums,
// same tree.
template <bool STAGED>
__global__ void __launch_bounds__(1024) k_batch_partials(const float* __restrict__ proj, int K, uint64_t ncols,
                                                         const uint32_t* __restrict__ batch, uint32_t nbatch,
                                                         double* __restrict__ partials) {
    __shared__ double stage[LG_SUMS_BATCH * 32];
    extern __shared__ float ptile[];
    const uint64_t cell0 = (uint64_t)blockIdx.x * LG_BLOCK_CELLS;
    const uint64_t cell = cell0 + threadIdx.x;
    const bool live = cell < ncols;
    const uint32_t myb = live ? (batch ? batch[cell] : 0u) : 0xffffffffu;
    const float* row = proj + (size_t)cell * K;
    if constexpr (STAGED) {
        const int KS = K | 1;
        const int ncell = (ncols - cell0) < (uint64_t)LG_BLOCK_CELLS ? (int)(ncols - cell0) : LG_BLOCK_CELLS;
        const int total = ncell * K, nvec = total >> 2;
        const float* src = proj + cell0 * K;  // 16-byte aligned: the host checks the base, a block is 4096 K bytes
        const int step_r = (4 * LG_BLOCK_CELLS) / K, step_c = (4 * LG_BLOCK_CELLS) % K;
        int r = (4 * (int)threadIdx.x) / K, c = (4 * (int)threadIdx.x) % K;
        for (int v = threadIdx.x; v < nvec; v += LG_BLOCK_CELLS) {
            const float4 q = __ldcs(reinterpret_cast<const float4*>(src) + v);
            const float qq[4] = {q.x, q.y, q.z, q.w};
            int rr = r, cc = c;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                ptile[rr * KS + cc] = qq[i];
                if (++cc == K) {
                    cc = 0;
                    ++rr;
                }
            }
            r += step_r;
            c += step_c;
            if (c >= K) {
                c -= K;
                ++r;
            }
        }
        for (int e = (nvec << 2) + threadIdx.x; e < total; e += LG_BLOCK_CELLS) ptile[(e / K) * KS + (e % K)] = src[e];
        __syncthreads();
        row = ptile + (size_t)threadIdx.x * KS;
    }
    double* outp = partials + (size_t)blockIdx.x * nbatch * (K + 1);
    // nbatch * (K + 1) sums (value k of batch b, then the cell count), LG_SUMS_BATCH at a time: same tree as
    // lg_block_sum_1024, one barrier per batch instead of two per value
    const uint32_t M = nbatch * (uint32_t)(K + 1);
    for (uint32_t base = 0; base < M; base += LG_SUMS_BATCH) {
        const uint32_t cnt = min((uint32_t)LG_SUMS_BATCH, M - base);
        for (uint32_t i = 0; i < cnt; ++i) {
            const uint32_t e = base + i, b = e / (uint32_t)(K + 1);
            const int k = (int)(e % (uint32_t)(K + 1));
            double v = 0.0;
            if (live && myb == b) v = (k < K) ? (double)row[k] : 1.0;
            lg_block_sums_stage1(v, (int)i, stage);
        }
        __syncthreads();
        lg_block_sums_stage2(stage, (int)cnt, outp + base);
    }
}

extern "C" int lg_proj_batch_partials(lg_ctx* ctx, const float* d_proj, int K, uint64_t ncols, const uint32_t* d_batch,
                                      uint32_t nbatch, double* d_partials) {
    if (!ctx) return LG_ERR_INVALID;
    LG_REQUIRE(ctx, d_proj && d_partials && nbatch >= 1, "lg_proj_batch_partials: null argument");
    cudaSetDevice(ctx->device);
    const uint64_t nblk = (ncols + LG_BLOCK_CELLS - 1) / LG_BLOCK_CELLS;
    if (nblk == 0) return LG_OK;
    const size_t tile = (size_t)LG_BLOCK_CELLS * (K | 1) * sizeof(float);
    if (tile + sizeof(double) * LG_SUMS_BATCH * 32 + 1024 <= ctx->smem_optin && ((uintptr_t)d_proj & 15) == 0) {
        LG_CUDA(ctx, cudaFuncSetAttribute(k_batch_partials<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tile));
        LG_LAUNCH(ctx, k_batch_partials<true>, (unsigned)nblk, LG_BLOCK_CELLS, tile, d_proj, K, ncols, d_batch, nbatch, d_partials);
    } else {
        LG_LAUNCH(ctx, k_batch_partials<false>, (unsigned)nblk, LG_BLOCK_CELLS, 0, d_proj, K, ncols, d_batch, nbatch, d_partials);
    }
    return LG_OK;
}

// sum block partials in block order: out[m] = ((p[0][m] + p[1][m]) + p[2][m]) + ...
// The adds are one dependent chain in block order (that order is the contract: results do not depend on how the cells are
// sharded), the loads are not.  A CTA owns 32 consecutive values m; all of its threads stream tiles of PF_TB blocks x 32
// values into a shared-memory ring with cp.async (PF_ST tiles in flight), and one warp folds the rows of a landed tile, a
// lane per value.  The chain then runs at the latency of a shared-memory load + one f64 add per block instead of one L2
// round trip per 16 blocks: with the 9 768 blocks of eight 1.25M-cell shards 0.22 ms -> 0.05 ms per call, three calls per pass.
constexpr int PF_MT = 32, PF_TB = 32, PF_ST = 4, PF_THREADS = 256;
__global__ void __launch_bounds__(PF_THREADS) k_partials_finalize(const double* __restrict__ partials, uint64_t nblocks, uint32_t M,
                                                                  double* __restrict__ out) {
    __shared__ double tile[PF_ST][PF_TB][PF_MT];
    const uint32_t m0 = blockIdx.x * PF_MT;
    const uint32_t mt = min((uint32_t)PF_MT, M - m0);
    const uint64_t ntiles = (nblocks + PF_TB - 1) / PF_TB;
    auto issue = [&](uint64_t t) {
        if (t < ntiles) {
            double(*buf)[PF_MT] = tile[t % PF_ST];
            for (uint32_t i = threadIdx.x; i < PF_TB * PF_MT; i += PF_THREADS) {
                const uint32_t r = i / PF_MT, c = i % PF_MT;
                const uint64_t b = t * PF_TB + r;
                if (b < nblocks && c < mt) {
                    const uint32_t d = (uint32_t)__cvta_generic_to_shared(&buf[r][c]);
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(partials + b * M + m0 + c) : "memory");
                }
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");  // an empty group keeps the wait count uniform
    };
    for (int s = 0; s < PF_ST; ++s) issue((uint64_t)s);
    double acc = 0.0;
    for (uint64_t t = 0; t < ntiles; ++t) {
        asm volatile("cp.async.wait_group %0;" ::"n"(PF_ST - 1) : "memory");
        __syncthreads();
        if (threadIdx.x < mt) {
            const double(*buf)[PF_MT] = tile[t % PF_ST];
            const uint32_t rows = (uint32_t)min((uint64_t)PF_TB, nblocks - t * PF_TB);
            if (rows == PF_TB) {
#pragma unroll
                for (int r = 0; r < PF_TB; ++r) acc = acc + buf[r][threadIdx.x];
            } else {
                for (uint32_t r = 0; r < rows; ++r) acc = acc + buf[r][threadIdx.x];
            }
        }
        __syncthreads();
        issue(t + PF_ST);
    }
    if (threadIdx.x < mt) out[m0 + threadIdx.x] = acc;
}

extern "C" int lg_block_partials_finalize(lg_ctx* ctx, const double* d_partials, uint64_t nblocks, uint32_t M,
                                          double* d_out) {
    if (!ctx) return LG_ERR_INVALID;
    LG_REQUIRE(ctx, d_partials && d_out && M >= 1, "lg_block_partials_finalize: null argument");
    cudaSetDevice(ctx->device);
    LG_LAUNCH(ctx, k_partials_finalize, (M + PF_MT - 1) / PF_MT, PF_THREADS, 0, d_partials, nblocks, M, d_out);
    return LG_OK;
}

// ---------------------------------------------------------------------------------------------
// K2b/K2c: per-cell standardise (nalgebra scale_columns_inplace, dmatrix_util.rs:986-995) done by
// one thread per cell in the reference's own sequential order, on a smem tile so global traffic
// stays coalesced.  MODE 0: subtract batch mean first, track (min,max).  MODE 1: clamp first.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void atomic_max_f32(float* addr, float v) {
    if (v >= 0.0f) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
    else atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}
__device__ __forceinline__ void atomic_min_f32(float* addr, float v) {
    if (v >= 0.0f) atomicMin(reinterpret_cast<int*>(addr), __float_as_int(v));
    else atomicMax(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

constexpr int SCALE_CELLS = 128;

template <int MODE>
__global__ void __launch_bounds__(SCALE_CELLS) k_scale_cells(float* __restrict__ proj, int K, uint64_t ncols,
                                                             const uint32_t* __restrict__ batch, uint32_t nbatch,
                                                             const double* __restrict__ batch_sums,
                                                             const float* __restrict__ fold_sum,
                                                             const unsigned long long* __restrict__ fold_cnt,
                                                             float* __restrict__ minmax) {
    extern __shared__ float tile[];  // SCALE_CELLS rows of stride KS (odd -> conflict-free row walks)
    if constexpr (MODE == 1) {
        // gated form (lg_proj_clamp_rescale_if): `minmax` holds the GLOBAL (min, max); nothing to do inside [-4, 4]
        if (minmax && !(minmax[1] > 4.0f || minmax[0] < -4.0f)) return;
    }
    const int KS = K | 1;
    float* neg_mean = tile + SCALE_CELLS * KS;  // nbatch * K
    float lmin = INFINITY, lmax = -INFINITY;
    // a grid-stride walk over the 128-cell tiles: one tile per CTA in the usual launch; the gated clamp is launched with a
    // few CTAs per SM only, so that the common "nothing to clamp" case costs a handful of CTAs instead of one per tile
    const uint64_t ntiles = (ncols + SCALE_CELLS - 1) / SCALE_CELLS;
    for (uint64_t tile_i = blockIdx.x; tile_i < ntiles; tile_i += gridDim.x) {
    const uint64_t cell0 = tile_i * SCALE_CELLS;
    const int ncell = (ncols - cell0) < (uint64_t)SCALE_CELLS ? (int)(ncols - cell0) : SCALE_CELLS;
    const size_t total = (size_t)ncell * K;
    const float* src = proj + cell0 * K;
    // the tile is one contiguous run of ncell * K floats: 128-bit streams, (row, col) advanced without a division per element
    const bool vec_io = ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
    const int total_i = (int)total, nvec = vec_io ? (total_i >> 2) : 0;
    const int step_r = (4 * SCALE_CELLS) / K, step_c = (4 * SCALE_CELLS) % K;
    {
        int r = (4 * (int)threadIdx.x) / K, c = (4 * (int)threadIdx.x) % K;
        for (int v = threadIdx.x; v < nvec; v += SCALE_CELLS) {
            const float4 q = reinterpret_cast<const float4*>(src)[v];
            const float qq[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                int cc = c + i, rr = r;
                if (cc >= K) {
                    cc -= K;
                    ++rr;
                }
                tile[rr * KS + cc] = qq[i];
            }
            r += step_r;
            c += step_c;
            if (c >= K) {
                c -= K;
                ++r;
            }
        }
        for (int e = 4 * nvec + threadIdx.x; e < total_i; e += SCALE_CELLS) tile[(e / K) * KS + (e % K)] = src[e];
    }
    const bool centre = MODE == 0 && (batch_sums || fold_sum);
    if (MODE == 0 && batch_sums) {
        for (uint32_t e = threadIdx.x; e < nbatch * (uint32_t)K; e += SCALE_CELLS) {
            const uint32_t b = e / K, k = e % K;
            const double cnt = batch_sums[(size_t)b * (K + 1) + K];
            neg_mean[e] = cnt > 0.0 ? -(float)(batch_sums[(size_t)b * (K + 1) + k] / cnt) : 0.0f;
        }
    } else if (MODE == 0 && fold_sum) {
        // exact-order form: the reference's own f32 fold divided by the f32 cell count (random_projection.rs:380-387)
        for (uint32_t e = threadIdx.x; e < nbatch * (uint32_t)K; e += SCALE_CELLS) {
            const unsigned long long cnt = fold_cnt[e / K];
            neg_mean[e] = cnt ? -__fdiv_rn(fold_sum[e], (float)(double)cnt) : 0.0f;
        }
    }
    __syncthreads();
    if ((int)threadIdx.x < ncell) {
        float* x = tile + threadIdx.x * KS;
        if (centre) {
            // a label outside [0, nbatch) is rejected by the callers before any launch; the clamp keeps the read in bounds
            const uint32_t bb = batch ? batch[cell0 + threadIdx.x] : 0u;
            const float* nm = neg_mean + (size_t)(bb < nbatch ? bb : 0u) * K;
            for (int k = 0; k < K; ++k) x[k] = __fadd_rn(x[k], nm[k]);
        }
        if (MODE == 1)
            for (int k = 0; k < K; ++k) x[k] = fminf(fmaxf(x[k], -4.0f), 4.0f);
        const float nf = (float)K;
        float s = 0.0f;
        for (int k = 0; k < K; ++k) s = __fadd_rn(s, x[k]);
        const float mu = __fdiv_rn(s, nf);
        float v = 0.0f;
        for (int k = 0; k < K; ++k) {
            const float d = __fsub_rn(x[k], mu);
            v = __fadd_rn(v, __fmul_rn(d, d));
        }
        v = __fdiv_rn(v, nf);
        const float sig = __fsqrt_rn(v);
        const float nmu = -mu;
        for (int k = 0; k < K; ++k) {
            float y = __fadd_rn(x[k], nmu);
            if (sig > 0.0f) y = __fdiv_rn(y, sig);
            x[k] = y;
            lmin = fminf(lmin, y);
            lmax = fmaxf(lmax, y);
        }
    }
    __syncthreads();
    float* dst = proj + cell0 * K;
    {
        int r = (4 * (int)threadIdx.x) / K, c = (4 * (int)threadIdx.x) % K;
        for (int v = threadIdx.x; v < nvec; v += SCALE_CELLS) {
            float qq[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                int cc = c + i, rr = r;
                if (cc >= K) {
                    cc -= K;
                    ++rr;
                }
                qq[i] = tile[rr * KS + cc];
            }
            reinterpret_cast<float4*>(dst)[v] = make_float4(qq[0], qq[1], qq[2], qq[3]);
            r += step_r;
            c += step_c;
            if (c >= K) {
                c -= K;
                ++r;
            }
        }
        for (int e = 4 * nvec + threadIdx.x; e < total_i; e += SCALE_CELLS) dst[e] = tile[(e / K) * KS + (e % K)];
    }
    __syncthreads();  // the tile is overwritten by the next round
    }
    if (MODE == 0 && minmax) {
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) {
            lmin = fminf(lmin, __shfl_xor_sync(0xffffffffu, lmin, off));
            lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, off));
        }
        if ((threadIdx.x & 31) == 0) {
            if (lmin != INFINITY) atomic_min_f32(minmax, lmin);
            if (lmax != -INFINITY) atomic_max_f32(minmax + 1, lmax);
        }
    }
}

__global__ void k_init_minmax(float* mm) {
    mm[0] = INFINITY;
    mm[1] = -INFINITY;
}

static size_t scale_smem(int K, uint32_t nbatch) {
    return ((size_t)SCALE_CELLS * (K | 1) + (size_t)nbatch * K) * sizeof(float);
}

extern "C" int lg_proj_centre_scale(lg_ctx* ctx, float* d_proj, int K, uint64_t ncols, const uint32_t* d_batch,
                                    uint32_t nbatch, const double* d_batch_sums, float* d_minmax) {
    if (!ctx) return LG_ERR_INVALID;
    LG_REQUIRE(ctx, d_proj && K >= 1 && K <= 128, "lg_proj_centre_scale: bad argument");
    cudaSetDevice(ctx->device);
    if (d_minmax) LG_LAUNCH(ctx, k_init_minmax, 1, 1, 0, d_minmax);
    if (ncols == 0) return LG_OK;
    if (!d_batch_sums) nbatch = 0;
    const size_t smem = scale_smem(K, nbatch);
    LG_REQUIRE(ctx, smem <= ctx->smem_optin, "lg_proj_centre_scale: nbatch*K too large for shared memory");
    LG_CUDA(ctx, cudaFuncSetAttribute(k_scale_cells<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const uint64_t grid = (ncols + SCALE_CELLS - 1) / SCALE_CELLS;
    LG_LAUNCH(ctx, k_scale_cells<0>, (unsigned)grid, SCALE_CELLS, smem, d_proj, K, ncols, d_batch, nbatch,
              d_batch_sums, (const float*)nullptr, (const unsigned long long*)nullptr, d_minmax);
    return LG_OK;
}

// ---------------------------------------------------------------------------------------------
// K2, EXACT-ORDER form: the batch means are the reference's left folds over the batch's cells in ascending cell
// order (random_projection.rs:380-387; nalgebra column_mean).  One block per batch, one thread per dim: the adds of a
// (batch, dim) pair are ONE dependent f32 chain (that order is the contract), cells staged through shared memory so
// the global reads stay coalesced.  sum / cnt are carried IN and OUT: cell shards call it one after the other in rank
// order, handing the running folds on, so the result is the single-GPU fold for any GPU count.
// ---------------------------------------------------------------------------------------------
constexpr int FOLD_CELLS = 64;
__global__ void __launch_bounds__(128) k_batch_fold_exact(const float* __restrict__ proj, int K, uint64_t ncols,
                                                          const uint32_t* __restrict__ batch, float* __restrict__ sum,
                                                          unsigned long long* __restrict__ cnt) {
    extern __shared__ float tile[];  // FOLD_CELLS * K values, then FOLD_CELLS batch ids
    uint32_t* ids = reinterpret_cast<uint32_t*>(tile + FOLD_CELLS * K);
    const uint32_t b = blockIdx.x;
    const int k = threadIdx.x;
    float s = k < K ? sum[(size_t)b * K + k] : 0.0f;
    unsigned long long c = cnt[b];
    for (uint64_t c0 = 0; c0 < ncols; c0 += FOLD_CELLS) {
        const int nc = (ncols - c0) < (uint64_t)FOLD_CELLS ? (int)(ncols - c0) : FOLD_CELLS;
        for (int e = threadIdx.x; e < nc * K; e += blockDim.x) tile[e] = proj[c0 * K + e];
        for (int e = threadIdx.x; e < nc; e += blockDim.x) ids[e] = batch ? batch[c0 + e] : 0u;
        __syncthreads();
        if (k < K) {
            for (int i = 0; i < nc; ++i)
                if (ids[i] == b) {
                    s = __fadd_rn(s, tile[i * K + k]);
                    ++c;
                }
        }
        __syncthreads();
    }
    if (k < K) sum[(size_t)b * K + k] = s;
    if (k == 0) cnt[b] = c;
}

extern "C" int lg_proj_batch_fold(lg_ctx* ctx, const float* d_proj, int K, uint64_t ncols, const uint32_t* d_batch,
                                  uint32_t nbatch, float* d_sum, uint64_t* d_cnt) {
    if (!ctx) return LG_ERR_INVALID;
    LG_REQUIRE(ctx, d_proj && d_sum && d_cnt && nbatch >= 1 && K >= 1 && K <= 128, "lg_proj_batch_fold: bad argument");
    cudaSetDevice(ctx->device);
    if (ncols == 0) return LG_OK;
    const size_t smem = (size_t)FOLD_CELLS * (K + 1) * sizeof(float);
    LG_LAUNCH(ctx, k_batch_fold_exact, nbatch, 128, smem, d_proj, K, ncols, d_batch, d_sum,
              reinterpret_cast<unsigned long long*>(d_cnt));
    return LG_OK;
}

extern "C" int lg_proj_centre_scale_exact(lg_ctx* ctx, float* d_proj, int K, uint64_t ncols, const uint32_t* d_batch,
                                          uint32_t nbatch, const float* d_fold_sum, const uint64_t* d_fold_cnt,
                                          float* d_minmax) {
    if (!ctx) return LG_ERR_INVALID;
    LG_REQUIRE(ctx, d_proj && K >= 1 && K <= 128, "lg_proj_centre_scale_exact: bad argument");
    cudaSetDevice(ctx->device);
    if (d_minmax) LG_LAUNCH(ctx, k_init_minmax, 1, 1, 0, d_minmax);
    if (ncols == 0) return LG_OK;
    if (!d_fold_sum || !d_fold_cnt) nbatch = 0;
    const size_t smem = scale_smem(K, nbatch);
    LG_REQUIRE(ctx, smem <= ctx->smem_optin, "lg_proj_centre_scale_exact: nbatch*K too large for shared memory");
    LG_CUDA(ctx, cudaFuncSetAttribute(k_scale_cells<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const uint64_t grid = (ncols + SCALE_CELLS - 1) / SCALE_CELLS;
    LG_LAUNCH(ctx, k_scale_cells<0>, (unsigned)grid, SCALE_CELLS, smem, d_proj, K, ncols, d_batch, nbatch,
              (const double*)nullptr, nbatch ? d_fold_sum : nullptr,
              reinterpret_cast<const unsigned long long*>(nbatch ? d_fold_cnt : nullptr), d_minmax);
    return LG_OK;
}

extern "C" int lg_proj_clamp_rescale(lg_ctx* ctx, float* d_proj, int K, uint64_t ncols) {
    return lg_proj_clamp_rescale_if(ctx, d_proj, K, ncols, nullptr);
}

extern "C" int lg_proj_clamp_rescale_if(lg_ctx* ctx, float* d_proj, int K, uint64_t ncols, const float* d_minmax) {
    if (!ctx) return LG_ERR_INVALID;
    LG_REQUIRE(ctx, d_proj && K >= 1 && K <= 128, "lg_proj_clamp_rescale: bad argument");
    cudaSetDevice(ctx->device);
    if (ncols == 0) return LG_OK;
    const size_t smem = scale_smem(K, 0);
    LG_CUDA(ctx, cudaFuncSetAttribute(k_scale_cells<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    uint64_t grid = (ncols + SCALE_CELLS - 1) / SCALE_CELLS;
    if (d_minmax && grid > (uint64_t)ctx->num_sms * 8) grid = (uint64_t)ctx->num_sms * 8;  // gated: usually nothing to do
    LG_LAUNCH(ctx, k_scale_cells<1>, (unsigned)grid, SCALE_CELLS, smem, d_proj, K, ncols, (const uint32_t*)nullptr, 0u,
              (const double*)nullptr, (const float*)nullptr, (const unsigned long long*)nullptr, const_cast<float*>(d_minmax));
    return LG_OK;
}

// ---------------------------------------------------------------------------------------------
// composite: project_columns_with_batch_correction_seeded (random_projection.rs:341-415)
// ---------------------------------------------------------------------------------------------
// a batch label outside [0, nbatch) is an error (the reference indexes per-batch vectors with it and would panic)
__global__ void k_check_labels(const uint32_t* __restrict__ label, uint64_t n, uint32_t bound, int* __restrict__ flag) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    bool bad = false;
    for (; i < n; i += stride) bad |= label[i] >= bound;
    if (bad) atomicOr(flag, 1);
}
int lg_check_labels(lg_ctx* ctx, const uint32_t* d_label, uint64_t n, uint32_t bound, const char* what) {
    if (!d_label || n == 0) return LG_OK;
    LgStage st(ctx);
    int* d_flag;
    LG_TRY(st.scratch(1, &d_flag));
    LG_CUDA(ctx, cudaMemsetAsync(d_flag, 0, sizeof(int), ctx->stream));
    uint64_t blocks = (n + 1023) / 1024;
    if (blocks > 1184) blocks = 1184;
    LG_LAUNCH(ctx, k_check_labels, (unsigned)blocks, 256, 0, d_label, n, bound, d_flag);
    int* h_flag = static_cast<int*>(ctx->pinned);
    LG_CUDA(ctx, cudaMemcpyAsync(h_flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    LG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (*h_flag) return lg_fail(ctx, LG_ERR_INVALID, std::string(what) + ": label out of range");
    return LG_OK;
}

static int project_impl(lg_ctx* ctx, const lg_csc* m, const float* basis_kd, int K, const uint32_t* batch_of_cell,
                        uint32_t nbatch, float* out_proj, bool exact);

extern "C" int lg_project(lg_ctx* ctx, const lg_csc* m, const float* basis_kd, int K, const uint32_t* batch_of_cell,
                          uint32_t nbatch, float* out_proj) {
    return project_impl(ctx, m, basis_kd, K, batch_of_cell, nbatch, out_proj, false);
}
extern "C" int lg_project_exact(lg_ctx* ctx, const lg_csc* m, const float* basis_kd, int K, const uint32_t* batch_of_cell,
                                uint32_t nbatch, float* out_proj) {
    return project_impl(ctx, m, basis_kd, K, batch_of_cell, nbatch, out_proj, true);
}
extern "C" int lg_project_raw_exact(lg_ctx* ctx, const lg_csc* m, const float* basis_kd, int K, float* out_proj) {
    if (!ctx) return LG_ERR_INVALID;
    LG_REQUIRE(ctx, m && basis_kd && out_proj, "lg_project_raw_exact: null argument");
    LG_REQUIRE(ctx, K >= 1 && K <= 128, "lg_project_raw_exact: K must be in [1, 128]");
    cudaSetDevice(ctx->device);
    LgStage st(ctx);
    const float* d_basis;
    float* d_out;
    LG_TRY(st.in(basis_kd, (size_t)K * m->nrows, &d_basis));
    LG_TRY(st.out(out_proj, (size_t)K * m->ncols, &d_out));
    LG_TRY(launch_project_raw_exact(ctx, m, d_basis, K, d_out));
    return st.finish();
}

static int project_impl(lg_ctx* ctx, const lg_csc* m, const float* basis_kd, int K, const uint32_t* batch_of_cell,
                        uint32_t nbatch, float* out_proj, bool exact) {
    if (!ctx) return LG_ERR_INVALID;
    LG_REQUIRE(ctx, m && basis_kd && out_proj, "lg_project: null argument");
    LG_REQUIRE(ctx, K >= 1 && K <= 128, "lg_project: K must be in [1, 128]");
    LG_REQUIRE(ctx, !batch_of_cell || nbatch >= 1, "lg_project: nbatch must be >= 1 with batch labels");
    cudaSetDevice(ctx->device);
    LgStage st(ctx);
    const float* d_basis;
    const uint32_t* d_batch;
    float* d_out;
    LG_TRY(st.in(basis_kd, (size_t)K * m->nrows, &d_basis));
    LG_TRY(st.in(batch_of_cell, (size_t)m->ncols, &d_batch));
    LG_TRY(st.out(out_proj, (size_t)K * m->ncols, &d_out));
    if (d_batch) LG_TRY(lg_check_labels(ctx, d_batch, m->ncols, nbatch, "lg_project: batch_of_cell"));
    float* d_mm;
    LG_TRY(st.scratch(2, &d_mm));
    float* h_mm = static_cast<float*>(ctx->pinned);
    if (exact) {
        LG_TRY(launch_project_raw_exact(ctx, m, d_basis, K, d_out));
        float* d_fsum = nullptr;
        uint64_t* d_fcnt = nullptr;
        if (d_batch && m->ncols) {
            LG_TRY(st.scratch((size_t)nbatch * K, &d_fsum));
            LG_TRY(st.scratch((size_t)nbatch, &d_fcnt));
            LG_CUDA(ctx, cudaMemsetAsync(d_fsum, 0, (size_t)nbatch * K * sizeof(float), ctx->stream));
            LG_CUDA(ctx, cudaMemsetAsync(d_fcnt, 0, (size_t)nbatch * sizeof(uint64_t), ctx->stream));
            LG_TRY(lg_proj_batch_fold(ctx, d_out, K, m->ncols, d_batch, nbatch, d_fsum, d_fcnt));
        }
        LG_TRY(lg_proj_centre_scale_exact(ctx, d_out, K, m->ncols, d_batch, nbatch, d_fsum, d_fcnt, d_mm));
        LG_CUDA(ctx, cudaMemcpyAsync(h_mm, d_mm, 2 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
        LG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (m->ncols && (h_mm[1] > 4.0f || h_mm[0] < -4.0f)) LG_TRY(lg_proj_clamp_rescale(ctx, d_out, K, m->ncols));
        return st.finish();
    }
    LG_TRY(launch_project_raw(ctx, m, d_basis, K, d_out));
    const uint64_t nblk = (m->ncols + LG_BLOCK_CELLS - 1) / LG_BLOCK_CELLS;
    double* d_sums = nullptr;
    if (d_batch && nblk) {
        const uint32_t M = nbatch * (K + 1);
        double* d_part;
        LG_TRY(st.scratch((size_t)nblk * M, &d_part));
        LG_TRY(st.scratch((size_t)M, &d_sums));
        LG_TRY(lg_proj_batch_partials(ctx, d_out, K, m->ncols, d_batch, nbatch, d_part));
        LG_TRY(lg_block_partials_finalize(ctx, d_part, nblk, M, d_sums));
    }
    LG_TRY(lg_proj_centre_scale(ctx, d_out, K, m->ncols, d_batch, nbatch, d_sums, d_mm));
    // the clamp branch is a global decision (:401): read back (min, max)
    LG_CUDA(ctx, cudaMemcpyAsync(h_mm, d_mm, 2 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    LG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (m->ncols && (h_mm[1] > 4.0f || h_mm[0] < -4.0f)) LG_TRY(lg_proj_clamp_rescale(ctx, d_out, K, m->ncols));
    return st.finish();
}
