// lg_codes.cu — stage 2 (binary codes) and stage 3 (group ids).
//   binary_sort_columns   data-beans-alg/src/random_projection.rs:535-564
//   rsvd                  matrix-util/src/dmatrix_rsvd.rs:85-180   (as-written semantics)
//   assign_groups         data-beans/src/sparse_io_vector/groups.rs:13-37
#include <algorithm>
#include <cmath>
#include <string>

#include "lg_common.cuh"

constexpr int KK_MAX = 16;

// ---------------------------------------------------------------------------------------------
// K3a: B = Q^T X per cell (sequential fma over the K axis, as the reference's gemm does for
// k <= kc) and the per-1024-cell-block f64 partials of the upper triangle of B B^T.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k_codes_gram(const float* __restrict__ proj, int K, uint64_t ncols,
                                                     const float* __restrict__ q, int kk, float* __restrict__ bout,
                                                     double* __restrict__ partials) {
    extern __shared__ float qs[];  // K * kk
    for (int e = threadIdx.x; e < K * kk; e += blockDim.x) qs[e] = q[e];
    __syncthreads();
    const uint64_t cell = (uint64_t)blockIdx.x * LG_BLOCK_CELLS + threadIdx.x;
    const bool live = cell < ncols;
    float b[KK_MAX];
#pragma unroll
    for (int i = 0; i < KK_MAX; ++i) b[i] = 0.0f;
    if (live) {
        // every thread walks its own row (staging the block's rows through shared memory with coalesced loads was tried:
        // 209 KB of tiles leave one block per SM with serialised phases, 0.70 -> 0.82 ms for the stage)
        const float* x = proj + (size_t)cell * K;
        for (int k = 0; k < K; ++k) {
            const float xv = x[k];
#pragma unroll
            for (int i = 0; i < KK_MAX; ++i)
                if (i < kk) b[i] = fmaf(qs[i * K + k], xv, b[i]);
        }
#pragma unroll
        for (int i = 0; i < KK_MAX; ++i)
            if (i < kk) bout[(size_t)cell * kk + i] = b[i];
    }
    const int M = kk * (kk + 1) / 2;
    double* outp = partials + (size_t)blockIdx.x * M;
    // the kk (kk + 1) / 2 Gram entries through the batched block sum (same tree as lg_block_sum_1024, one barrier in all)
    __shared__ double stage[(KK_MAX * (KK_MAX + 1) / 2) * 32];
    int slot = 0;
#pragma unroll
    for (int a = 0; a < KK_MAX; ++a) {
#pragma unroll
        for (int c = 0; c < KK_MAX; ++c) {
            if (c >= a && a < kk && c < kk) {
                lg_block_sums_stage1(live ? (double)b[a] * (double)b[c] : 0.0, slot, stage);
                ++slot;
            }
        }
    }
    __syncthreads();
    lg_block_sums_stage2(stage, M, outp);
}

// K3b: V[k] = (sum_i U[i,k] * B[i]) / sigma_k per cell, plus block partials of the column sums.
__global__ void __launch_bounds__(1024) k_codes_vproj(const float* __restrict__ bin, int kk, uint64_t ncols,
                                                      const float* __restrict__ u, const float* __restrict__ sigma,
                                                      float* __restrict__ vout, double* __restrict__ partials) {
    __shared__ float us[KK_MAX * KK_MAX];
    __shared__ float sg[KK_MAX];
    for (int e = threadIdx.x; e < kk * kk; e += blockDim.x) us[e] = u[e];
    if ((int)threadIdx.x < kk) sg[threadIdx.x] = sigma[threadIdx.x];
    __syncthreads();
    const uint64_t cell = (uint64_t)blockIdx.x * LG_BLOCK_CELLS + threadIdx.x;
    const bool live = cell < ncols;
    float b[KK_MAX], v[KK_MAX];
#pragma unroll
    for (int i = 0; i < KK_MAX; ++i) {
        b[i] = 0.0f;
        v[i] = 0.0f;
    }
    if (live) {
#pragma unroll
        for (int i = 0; i < KK_MAX; ++i)
            if (i < kk) b[i] = bin[(size_t)cell * kk + i];
#pragma unroll
        for (int k = 0; k < KK_MAX; ++k) {
            if (k < kk) {
                float acc = 0.0f;
#pragma unroll
                for (int i = 0; i < KK_MAX; ++i)
                    if (i < kk) acc = fmaf(us[k * kk + i], b[i], acc);
                v[k] = sg[k] > 0.0f ? __fdiv_rn(acc, sg[k]) : 0.0f;
                vout[(size_t)cell * kk + k] = v[k];
            }
        }
    }
    double* outp = partials + (size_t)blockIdx.x * kk;
    __shared__ double stage[KK_MAX * 32];
#pragma unroll
    for (int k = 0; k < KK_MAX; ++k)
        if (k < kk) lg_block_sums_stage1(live ? (double)v[k] : 0.0, k, stage);
    __syncthreads();
    lg_block_sums_stage2(stage, kk, outp);
}

// K3c: warp-ballot sign packer.  V is cell-major (kk floats per cell); a warp takes 32 cells =
// 32*kk consecutive floats, reads them in kk fully coalesced rounds, ballots the predicate
// [V > mean] into kk 32-bit words, and lane c cuts its kk-bit window out of that bit stream.
__global__ void __launch_bounds__(256) k_codes_pack(const float* __restrict__ v, int kk, uint64_t ncols,
                                                    const float* __restrict__ mean, uint64_t* __restrict__ codes) {
    __shared__ float ms[KK_MAX];
    if ((int)threadIdx.x < kk) ms[threadIdx.x] = mean[threadIdx.x];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t cell0 = warp * 32;
    if (cell0 >= ncols) return;
    const uint64_t total = ncols * (uint64_t)kk;
    const uint64_t base = cell0 * (uint64_t)kk;
    const unsigned start = (unsigned)lane * kk;  // first bit of this lane's window
    const int wi = start >> 5, off = start & 31;
    unsigned lo = 0, hi = 0;
    for (int it = 0; it < kk; ++it) {
        const uint64_t e = base + (uint64_t)it * 32 + lane;
        bool pred = false;
        if (e < total) pred = v[e] > ms[(it * 32 + lane) % kk];
        const unsigned w = __ballot_sync(0xffffffffu, pred);
        if (it == wi) lo = w;
        if (it == wi + 1) hi = w;
    }
    const unsigned window = __funnelshift_r(lo, hi, off);
    const uint64_t cell = cell0 + lane;
    if (cell < ncols) codes[cell] = (uint64_t)(window & ((1u << kk) - 1u));
}

// ---- staged entry points --------------------------------------------------------------------
extern "C" int lg_codes_basis(lg_ctx* ctx, const float* first_cols_kr, int K, int r, int kk, float* out_q) {
    if (!ctx) return LG_ERR_INVALID;
    LG_REQUIRE(ctx, first_cols_kr && out_q && K >= 1 && r >= 1 && kk >= 1 && kk <= r && kk <= K,
               "lg_codes_basis: bad argument");
    LG_REQUIRE(ctx, !lg_is_device_ptr(first_cols_kr) && !lg_is_device_ptr(out_q), "lg_codes_basis takes host arrays");
    std::vector<float> qf((size_t)K * r);
    lgh_householder_q(first_cols_kr, K, r, qf.data());
    std::copy(qf.begin(), qf.begin() + (size_t)K * kk, out_q);
    return LG_OK;
}

extern "C" int lg_codes_gram(lg_ctx* ctx, const float* d_proj, int K, uint64_t ncols, const float* d_q, int kk,
                             float* d_b, double* d_partials) {
    if (!ctx) return LG_ERR_INVALID;
    LG_REQUIRE(ctx, d_proj && d_q && d_b && d_partials, "lg_codes_gram: null argument");
    LG_REQUIRE(ctx, kk >= 1 && kk <= KK_MAX && K >= 1, "lg_codes_gram: kk must be in [1, 16]");
    cudaSetDevice(ctx->device);
    const uint64_t nblk = (ncols + LG_BLOCK_CELLS - 1) / LG_BLOCK_CELLS;
    if (!nblk) return LG_OK;
    LG_LAUNCH(ctx, k_codes_gram, (unsigned)nblk, LG_BLOCK_CELLS, (size_t)K * kk * sizeof(float), d_proj, K, ncols, d_q,
              kk, d_b, d_partials);
    return LG_OK;
}

extern "C" int lg_codes_factor(lg_ctx* ctx, const double* gram_sums, const float* q, int K, int kk, float* out_u,
                               float* out_sigma) {
    if (!ctx) return LG_ERR_INVALID;
    LG_REQUIRE(ctx, gram_sums && q && out_u && out_sigma && kk >= 1 && kk <= KK_MAX, "lg_codes_factor: bad argument");
    LG_REQUIRE(ctx, !lg_is_device_ptr(gram_sums) && !lg_is_device_ptr(q) && !lg_is_device_ptr(out_u),
               "lg_codes_factor takes host arrays");
    std::vector<double> G((size_t)kk * kk), ev(kk), U((size_t)kk * kk);
    int slot = 0;
    for (int a = 0; a < kk; ++a)
        for (int b = a; b < kk; ++b) {
            G[(size_t)b * kk + a] = gram_sums[slot];
            G[(size_t)a * kk + b] = gram_sums[slot];
            ++slot;
        }
    lgh_jacobi_eig(G.data(), kk, ev.data(), U.data());
    for (int k = 0; k < kk; ++k) {
        // sign convention: largest-magnitude component of Q u_k (in R^K) positive, first index wins ties
        double best = 0.0, bestv = 0.0;
        for (int d = 0; d < K; ++d) {
            double s = 0.0;
            for (int i = 0; i < kk; ++i) s += (double)q[(size_t)i * K + d] * U[(size_t)k * kk + i];
            if (std::fabs(s) > best) {
                best = std::fabs(s);
                bestv = s;
            }
        }
        const double flip = bestv < 0.0 ? -1.0 : 1.0;
        for (int i = 0; i < kk; ++i) out_u[(size_t)k * kk + i] = (float)(flip * U[(size_t)k * kk + i]);
        out_sigma[k] = (float)std::sqrt(std::max(ev[k], 0.0));
    }
    return LG_OK;
}

extern "C" int lg_codes_vproj(lg_ctx* ctx, const float* d_b, int kk, uint64_t ncols, const float* d_u,
                              const float* d_sigma, float* d_v, double* d_partials) {
    if (!ctx) return LG_ERR_INVALID;
    LG_REQUIRE(ctx, d_b && d_u && d_sigma && d_v && d_partials, "lg_codes_vproj: null argument");
    LG_REQUIRE(ctx, kk >= 1 && kk <= KK_MAX, "lg_codes_vproj: kk must be in [1, 16]");
    cudaSetDevice(ctx->device);
    const uint64_t nblk = (ncols + LG_BLOCK_CELLS - 1) / LG_BLOCK_CELLS;
    if (!nblk) return LG_OK;
    LG_LAUNCH(ctx, k_codes_vproj, (unsigned)nblk, LG_BLOCK_CELLS, 0, d_b, kk, ncols, d_u, d_sigma, d_v, d_partials);
    return LG_OK;
}

extern "C" int lg_codes_pack(lg_ctx* ctx, const float* d_v, int kk, uint64_t ncols, const float* d_mean,
                             uint64_t* d_codes) {
    if (!ctx) return LG_ERR_INVALID;
    LG_REQUIRE(ctx, d_v && d_mean && d_codes, "lg_codes_pack: null argument");
    LG_REQUIRE(ctx, kk >= 1 && kk <= KK_MAX, "lg_codes_pack: kk must be in [1, 16]");
    cudaSetDevice(ctx->device);
    if (!ncols) return LG_OK;
    const uint64_t warps = (ncols + 31) / 32;
    LG_LAUNCH(ctx, k_codes_pack, (unsigned)((warps + 7) / 8), 256, 0, d_v, kk, ncols, d_mean, d_codes);
    return LG_OK;
}

__global__ void k_means_from_sums(const double* __restrict__ sums, int kk, uint64_t ncols, float* __restrict__ mean) {
    const int k = threadIdx.x;
    if (k < kk) mean[k] = (float)(sums[k] / (double)ncols);
}

// ---- composite: binary_sort_columns -------------------------------------------------------------
extern "C" int lg_binary_codes(lg_ctx* ctx, const float* proj_kn, int K, uint64_t ncols, int kk, uint64_t* out_codes) {
    if (!ctx) return LG_ERR_INVALID;
    LG_REQUIRE(ctx, proj_kn && out_codes, "lg_binary_codes: null argument");
    LG_REQUIRE(ctx, kk >= 1 && kk <= KK_MAX && kk <= K && (uint64_t)kk <= ncols,
               "lg_binary_codes: need 1 <= kk <= min(16, K, ncols)");
    cudaSetDevice(ctx->device);
    LgStage st(ctx);
    const float* d_proj;
    uint64_t* d_codes;
    LG_TRY(st.in(proj_kn, (size_t)K * ncols, &d_proj));
    LG_TRY(st.out(out_codes, (size_t)ncols, &d_codes));
    // dmatrix_rsvd.rs:145-153: rank = min(K, N); oversample 5 only when rank > kk
    int rank = (int)std::min<uint64_t>((uint64_t)K, ncols);
    int r = rank > kk ? kk + 5 : rank;
    if ((uint64_t)r > ncols) r = (int)ncols;
    std::vector<float> first((size_t)K * r), q((size_t)K * kk);
    LG_CUDA(ctx, cudaMemcpyAsync(first.data(), d_proj, first.size() * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    LG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    LG_TRY(lg_codes_basis(ctx, first.data(), K, r, kk, q.data()));

    const uint64_t nblk = (ncols + LG_BLOCK_CELLS - 1) / LG_BLOCK_CELLS;
    const int M = kk * (kk + 1) / 2;
    float *d_q, *d_b, *d_v, *d_u, *d_sig, *d_mean;
    double *d_part, *d_sums;
    LG_TRY(st.scratch((size_t)K * kk, &d_q));
    LG_TRY(st.scratch((size_t)kk * ncols, &d_b));
    LG_TRY(st.scratch((size_t)kk * ncols, &d_v));
    LG_TRY(st.scratch((size_t)kk * kk, &d_u));
    LG_TRY(st.scratch((size_t)kk, &d_sig));
    LG_TRY(st.scratch((size_t)kk, &d_mean));
    LG_TRY(st.scratch((size_t)nblk * M, &d_part));
    LG_TRY(st.scratch((size_t)M, &d_sums));
    LG_CUDA(ctx, cudaMemcpyAsync(d_q, q.data(), q.size() * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    LG_TRY(lg_codes_gram(ctx, d_proj, K, ncols, d_q, kk, d_b, d_part));
    LG_TRY(lg_block_partials_finalize(ctx, d_part, nblk, M, d_sums));
    std::vector<double> gram(M);
    LG_CUDA(ctx, cudaMemcpyAsync(gram.data(), d_sums, M * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    LG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    std::vector<float> u((size_t)kk * kk), sig(kk);
    LG_TRY(lg_codes_factor(ctx, gram.data(), q.data(), K, kk, u.data(), sig.data()));
    LG_CUDA(ctx, cudaMemcpyAsync(d_u, u.data(), u.size() * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    LG_CUDA(ctx, cudaMemcpyAsync(d_sig, sig.data(), sig.size() * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    LG_TRY(lg_codes_vproj(ctx, d_b, kk, ncols, d_u, d_sig, d_v, d_part));
    LG_TRY(lg_block_partials_finalize(ctx, d_part, nblk, kk, d_sums));
    LG_LAUNCH(ctx, k_means_from_sums, 1, 32, 0, d_sums, kk, ncols, d_mean);
    LG_TRY(lg_codes_pack(ctx, d_v, kk, ncols, d_mean, d_codes));
    // u/sig/q host vectors must outlive the async copies
    LG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return st.finish();
}

// ---------------------------------------------------------------------------------------------
// Stage 3: group ids.  Presence flags are written once per distinct code per warp
// (__match_any_sync leader), the lexicographic LUT is built on the host over <= 2^kk codes.
// ---------------------------------------------------------------------------------------------
__global__ void k_code_presence(const uint64_t* __restrict__ codes, uint64_t ncols, uint32_t ncodes,
                                uint32_t* __restrict__ present) {
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = j < ncols;
    const unsigned active = __ballot_sync(0xffffffffu, live);
    if (!live) return;
    const uint32_t c = (uint32_t)codes[j];
    const unsigned peers = __match_any_sync(active, c);
    if ((threadIdx.x & 31) == (unsigned)(__ffs(peers) - 1) && c < ncodes) present[c] = 1u;
}
__global__ void k_codes_to_groups(const uint64_t* __restrict__ codes, uint64_t ncols, uint32_t ncodes,
                                  const uint32_t* __restrict__ lut, uint32_t* __restrict__ group) {
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < ncols) {
        const uint64_t c = codes[j];
        group[j] = c < ncodes ? lut[c] : 0xffffffffu;
    }
}

extern "C" int lg_code_presence(lg_ctx* ctx, const uint64_t* d_codes, uint64_t ncols, int kk, uint32_t* d_present) {
    if (!ctx) return LG_ERR_INVALID;
    LG_REQUIRE(ctx, d_codes && d_present && kk >= 1 && kk <= 24, "lg_code_presence: bad argument");
    cudaSetDevice(ctx->device);
    LG_CUDA(ctx, cudaMemsetAsync(d_present, 0, sizeof(uint32_t) << kk, ctx->stream));
    if (!ncols) return LG_OK;
    LG_LAUNCH(ctx, k_code_presence, (unsigned)((ncols + 255) / 256), 256, 0, d_codes, ncols, 1u << kk, d_present);
    return LG_OK;
}

extern "C" int lg_group_lut(lg_ctx* ctx, const uint32_t* present, int kk, int padded, uint32_t* out_lut,
                            uint32_t* out_num_groups) {
    if (!ctx) return LG_ERR_INVALID;
    LG_REQUIRE(ctx, present && out_lut && kk >= 1 && kk <= 24, "lg_group_lut: bad argument");
    LG_REQUIRE(ctx, !lg_is_device_ptr(present) && !lg_is_device_ptr(out_lut), "lg_group_lut takes host arrays");
    const uint32_t ncodes = 1u << kk;
    std::vector<uint32_t> keys;
    for (uint32_t c = 0; c < ncodes; ++c) {
        out_lut[c] = 0xffffffffu;
        if (present[c]) keys.push_back(c);
    }
    if (!padded) {
        // groups.rs:20-24: sort by key.to_string(), byte-wise
        std::vector<std::string> names(ncodes);
        for (uint32_t c : keys) names[c] = std::to_string(c);
        std::sort(keys.begin(), keys.end(), [&](uint32_t a, uint32_t b) { return names[a] < names[b]; });
    }  // padded labels (refine.rs:21-35) sort numerically: keys are already ascending
    for (uint32_t g = 0; g < keys.size(); ++g) out_lut[keys[g]] = g;
    if (out_num_groups) *out_num_groups = (uint32_t)keys.size();
    return LG_OK;
}

extern "C" int lg_codes_to_groups(lg_ctx* ctx, const uint64_t* d_codes, uint64_t ncols, int kk, const uint32_t* d_lut,
                                  uint32_t* d_group) {
    if (!ctx) return LG_ERR_INVALID;
    LG_REQUIRE(ctx, d_codes && d_lut && d_group && kk >= 1 && kk <= 24, "lg_codes_to_groups: bad argument");
    cudaSetDevice(ctx->device);
    if (!ncols) return LG_OK;
    LG_LAUNCH(ctx, k_codes_to_groups, (unsigned)((ncols + 255) / 256), 256, 0, d_codes, ncols, 1u << kk, d_lut, d_group);
    return LG_OK;
}

extern "C" int lg_assign_groups(lg_ctx* ctx, const uint64_t* codes, uint64_t ncols, int kk, int padded,
                                uint32_t* out_group, uint32_t* out_num_groups) {
    if (!ctx) return LG_ERR_INVALID;
    LG_REQUIRE(ctx, codes && out_group && kk >= 1 && kk <= 24, "lg_assign_groups: bad argument");
    cudaSetDevice(ctx->device);
    LgStage st(ctx);
    const uint64_t* d_codes;
    uint32_t *d_group, *d_present, *d_lut;
    LG_TRY(st.in(codes, (size_t)ncols, &d_codes));
    LG_TRY(st.out(out_group, (size_t)ncols, &d_group));
    const uint32_t ncodes = 1u << kk;
    LG_TRY(st.scratch(ncodes, &d_present));
    LG_TRY(st.scratch(ncodes, &d_lut));
    LG_TRY(lg_code_presence(ctx, d_codes, ncols, kk, d_present));
    std::vector<uint32_t> present(ncodes), lut(ncodes);
    LG_CUDA(ctx, cudaMemcpyAsync(present.data(), d_present, ncodes * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    LG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    uint32_t ng = 0;
    LG_TRY(lg_group_lut(ctx, present.data(), kk, padded, lut.data(), &ng));
    LG_CUDA(ctx, cudaMemcpyAsync(d_lut, lut.data(), ncodes * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
    LG_TRY(lg_codes_to_groups(ctx, d_codes, ncols, kk, d_lut, d_group));
    LG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (out_num_groups) *out_num_groups = ng;
    return st.finish();
}
