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
        if ((K & 1) == 0 && ((uintptr_t)proj & 7) == 0) {
            // 64-bit loads of the row (the lanes of a warp sit 4 K bytes apart, so every load costs one wavefront per lane
            // whatever its width: half the loads, half the wavefronts); the fma order over k is unchanged
            const float2* x2 = reinterpret_cast<const float2*>(x);
            for (int k = 0; k < K; k += 2) {
                const float2 xv = x2[k >> 1];
#pragma unroll
                for (int i = 0; i < KK_MAX; ++i)
                    if (i < kk) b[i] = fmaf(qs[i * K + k], xv.x, b[i]);
#pragma unroll
                for (int i = 0; i < KK_MAX; ++i)
                    if (i < kk) b[i] = fmaf(qs[i * K + k + 1], xv.y, b[i]);
            }
        } else {
            for (int k = 0; k < K; ++k) {
                const float xv = x[k];
#pragma unroll
                for (int i = 0; i < KK_MAX; ++i)
                    if (i < kk) b[i] = fmaf(qs[i * K + k], xv, b[i]);
            }
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


// ---------------------------------------------------------------------------------------------
// K3 small factorisations ON THE DEVICE (one warp each): the K x r Householder QR whose first kk columns are the
// range basis Q (dmatrix_rsvd.rs:129-131, `.qr().q()`), and the kk x kk symmetric eigen-decomposition that stands in
// for SVD(B).  They used to run on the host between kernels (two stream synchronisations per call);
// here every reduction keeps a fixed sequential order — a lane owns a column (QR) or a row (Jacobi) and walks it in
// index order — and every operation is an explicit round-to-nearest intrinsic (no FMA contraction), so the result
// does not depend on the launch shape.
// ---------------------------------------------------------------------------------------------
constexpr int QR_RMAX = KK_MAX + 5;  // r = kk + 5 columns at most

__global__ void __launch_bounds__(32) k_codes_basis(const float* __restrict__ first_kr, int K, int r, int kk,
                                                    float* __restrict__ q_out) {
    extern __shared__ float qr_s[];  // A (K x r), then Q (K x dim), column-major
    const int lane = threadIdx.x;
    const int dim = K < r ? K : r;
    float* A = qr_s;
    float* Q = qr_s + (size_t)K * r;
    __shared__ float diag[QR_RMAX];
    for (int e = lane; e < K * r; e += 32) A[e] = first_kr[e];
    __syncwarp();
    for (int c = 0; c < dim; ++c) {
        float* col = A + (size_t)c * K;
        // reflection axis of A[c.., c]: every lane folds the same values in the same order (no broadcast needed)
        float sq = 0.0f;
        for (int i = c; i < K; ++i) sq = __fadd_rn(sq, __fmul_rn(col[i], col[i]));
        const float norm = __fsqrt_rn(sq);
        const float head = col[c];
        const float signed_norm = head < 0.0f ? -norm : norm;
        const float factor = __fmul_rn(__fadd_rn(sq, __fmul_rn(fabsf(head), norm)), 2.0f);
        __syncwarp();
        if (lane == 0) col[c] = __fadd_rn(head, signed_norm);
        __syncwarp();
        if (factor != 0.0f) {
            const float fs = __fsqrt_rn(factor);
            for (int i = c + lane; i < K; i += 32) col[i] = __fdiv_rn(col[i], fs);
            __syncwarp();
            float n2 = 0.0f;
            for (int i = c; i < K; ++i) n2 = __fadd_rn(n2, __fmul_rn(col[i], col[i]));
            const float nn = __fsqrt_rn(n2);
            __syncwarp();
            for (int i = c + lane; i < K; i += 32) col[i] = __fdiv_rn(col[i], nn);
            __syncwarp();
            const float dg = -signed_norm;
            if (lane == 0) diag[c] = dg;
            const float sign = dg > 0.0f ? 1.0f : (dg < 0.0f ? -1.0f : 0.0f);
            const float m_two = __fmul_rn(sign, -2.0f);
            const int j = c + 1 + lane;  // a lane reflects its own column
            if (j < r) {
                float* cj = A + (size_t)j * K;
                float dot = 0.0f;
                for (int i = c; i < K; ++i) dot = __fadd_rn(dot, __fmul_rn(col[i], cj[i]));
                const float f = __fmul_rn(dot, m_two);
                for (int i = c; i < K; ++i) cj[i] = __fadd_rn(__fmul_rn(f, col[i]), __fmul_rn(sign, cj[i]));
            }
        } else if (lane == 0) {
            diag[c] = signed_norm;
        }
        __syncwarp();
    }
    // QR::q(): identity, reflected by the stored axes from the last to the first
    for (int e = lane; e < K * dim; e += 32) Q[e] = (e % K) == (e / K) ? 1.0f : 0.0f;
    __syncwarp();
    for (int c = dim - 1; c >= 0; --c) {
        const float* col = A + (size_t)c * K;
        const float dg = diag[c];
        const float sign = dg > 0.0f ? 1.0f : (dg < 0.0f ? -1.0f : 0.0f);
        const float m_two = __fmul_rn(sign, -2.0f);
        const int j = c + lane;
        if (j < dim) {
            float* qj = Q + (size_t)j * K;
            float dot = 0.0f;
            for (int i = c; i < K; ++i) dot = __fadd_rn(dot, __fmul_rn(col[i], qj[i]));
            const float f = __fmul_rn(dot, m_two);
            for (int i = c; i < K; ++i) qj[i] = __fadd_rn(__fmul_rn(f, col[i]), __fmul_rn(sign, qj[i]));
        }
        __syncwarp();
    }
    const int keep = kk < dim ? kk : dim;
    for (int e = lane; e < K * kk; e += 32) q_out[e] = (e / K) < keep ? Q[e] : 0.0f;
}

// cyclic Jacobi on the kk x kk Gram matrix (f64), eigenvalues descending (stable), then the sign convention
// (largest-magnitude component of Q u_k positive, first index wins ties) and sigma = sqrt(max(lambda, 0))
__global__ void __launch_bounds__(32) k_codes_factor(const double* __restrict__ gram_sums, const float* __restrict__ q, int K,
                                                     int n, float* __restrict__ out_u, float* __restrict__ out_sigma) {
    __shared__ double a[KK_MAX * KK_MAX], v[KK_MAX * KK_MAX];
    __shared__ int order[KK_MAX];
    const int lane = threadIdx.x;
#define JA(i, j) a[(j) * n + (i)]
#define JV(i, j) v[(j) * n + (i)]
    if (lane == 0) {
        int slot = 0;
        for (int x = 0; x < n; ++x)
            for (int y = x; y < n; ++y) {
                JA(x, y) = gram_sums[slot];
                JA(y, x) = gram_sums[slot];
                ++slot;
            }
    }
    for (int e = lane; e < n * n; e += 32) v[e] = (e % n) == (e / n) ? 1.0 : 0.0;
    __syncwarp();
    for (int sweep = 0; sweep < 64; ++sweep) {
        double off = 0.0, dg = 0.0;
        for (int p = 0; p < n; ++p) {
            dg = __dadd_rn(dg, __dmul_rn(JA(p, p), JA(p, p)));
            for (int qq = p + 1; qq < n; ++qq) off = __dadd_rn(off, __dmul_rn(JA(p, qq), JA(p, qq)));
        }
        if (off <= 1e-60 || off <= __dmul_rn(1e-32, dg)) break;
        for (int p = 0; p < n - 1; ++p)
            for (int qq = p + 1; qq < n; ++qq) {
                const double apq = JA(p, qq);
                if (apq == 0.0) continue;  // warp-uniform: every lane reads the same element
                const double theta = __ddiv_rn(__dsub_rn(JA(qq, qq), JA(p, p)), __dmul_rn(2.0, apq));
                const double t = __ddiv_rn(theta >= 0.0 ? 1.0 : -1.0,
                                           __dadd_rn(fabs(theta), __dsqrt_rn(__dadd_rn(__dmul_rn(theta, theta), 1.0))));
                const double c = __ddiv_rn(1.0, __dsqrt_rn(__dadd_rn(__dmul_rn(t, t), 1.0)));
                const double s = __dmul_rn(t, c);
                __syncwarp();
                if (lane < n) {  // columns p, q: lane = row
                    const double akp = JA(lane, p), akq = JA(lane, qq);
                    JA(lane, p) = __dsub_rn(__dmul_rn(c, akp), __dmul_rn(s, akq));
                    JA(lane, qq) = __dadd_rn(__dmul_rn(s, akp), __dmul_rn(c, akq));
                }
                __syncwarp();
                if (lane < n) {  // rows p, q: lane = column
                    const double apk = JA(p, lane), aqk = JA(qq, lane);
                    JA(p, lane) = __dsub_rn(__dmul_rn(c, apk), __dmul_rn(s, aqk));
                    JA(qq, lane) = __dadd_rn(__dmul_rn(s, apk), __dmul_rn(c, aqk));
                    const double vkp = JV(lane, p), vkq = JV(lane, qq);
                    JV(lane, p) = __dsub_rn(__dmul_rn(c, vkp), __dmul_rn(s, vkq));
                    JV(lane, qq) = __dadd_rn(__dmul_rn(s, vkp), __dmul_rn(c, vkq));
                }
                __syncwarp();
            }
    }
    __syncwarp();
    if (lane == 0) {  // stable insertion sort, descending eigenvalue
        for (int i = 0; i < n; ++i) order[i] = i;
        for (int i = 1; i < n; ++i) {
            const int o = order[i];
            int j = i - 1;
            while (j >= 0 && JA(order[j], order[j]) < JA(o, o)) {
                order[j + 1] = order[j];
                --j;
            }
            order[j + 1] = o;
        }
    }
    __syncwarp();
    if (lane < n) {  // lane = singular vector k
        const int src = order[lane];
        double best = 0.0, bestv = 0.0;
        for (int d = 0; d < K; ++d) {
            double s = 0.0;
            for (int i = 0; i < n; ++i) s = __dadd_rn(s, __dmul_rn((double)q[(size_t)i * K + d], JV(i, src)));
            if (fabs(s) > best) {
                best = fabs(s);
                bestv = s;
            }
        }
        const double flip = bestv < 0.0 ? -1.0 : 1.0;
        for (int i = 0; i < n; ++i) out_u[(size_t)lane * n + i] = (float)__dmul_rn(flip, JV(i, src));
        const double ev = JA(src, src);
        out_sigma[lane] = (float)__dsqrt_rn(ev > 0.0 ? ev : 0.0);
    }
#undef JA
#undef JV
}

// ---- staged entry points --------------------------------------------------------------------
extern "C" int lg_codes_basis(lg_ctx* ctx, const float* first_cols_kr, int K, int r, int kk, float* out_q) {
    if (!ctx) return LG_ERR_INVALID;
    LG_REQUIRE(ctx, first_cols_kr && out_q && K >= 1 && K <= 128 && r >= 1 && r <= QR_RMAX && kk >= 1 && kk <= r && kk <= K,
               "lg_codes_basis: bad argument");
    cudaSetDevice(ctx->device);
    LgStage st(ctx);
    const float* d_first;
    float* d_q;
    LG_TRY(st.in(first_cols_kr, (size_t)K * r, &d_first));
    LG_TRY(st.out(out_q, (size_t)K * kk, &d_q));
    LG_LAUNCH(ctx, k_codes_basis, 1, 32, (size_t)2 * K * r * sizeof(float), d_first, K, r, kk, d_q);
    return st.finish();
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
    LG_REQUIRE(ctx, gram_sums && q && out_u && out_sigma && kk >= 1 && kk <= KK_MAX && K >= 1, "lg_codes_factor: bad argument");
    cudaSetDevice(ctx->device);
    LgStage st(ctx);
    const double* d_g;
    const float* d_q;
    float *d_u, *d_s;
    LG_TRY(st.in(gram_sums, (size_t)kk * (kk + 1) / 2, &d_g));
    LG_TRY(st.in(q, (size_t)K * kk, &d_q));
    LG_TRY(st.out(out_u, (size_t)kk * kk, &d_u));
    LG_TRY(st.out(out_sigma, (size_t)kk, &d_s));
    LG_LAUNCH(ctx, k_codes_factor, 1, 32, 0, d_g, d_q, K, kk, d_u, d_s);
    return st.finish();
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

extern "C" int lg_codes_means(lg_ctx* ctx, const double* d_sums, int kk, uint64_t ncols_total, float* d_mean) {
    if (!ctx) return LG_ERR_INVALID;
    LG_REQUIRE(ctx, d_sums && d_mean && kk >= 1 && kk <= KK_MAX && ncols_total >= 1, "lg_codes_means: bad argument");
    cudaSetDevice(ctx->device);
    LG_LAUNCH(ctx, k_means_from_sums, 1, 32, 0, d_sums, kk, ncols_total, d_mean);
    return LG_OK;
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
    // everything stays on the stream: the first r cells' K-vectors are the head of the projection itself
    LG_TRY(lg_codes_basis(ctx, d_proj, K, r, kk, d_q));
    LG_TRY(lg_codes_gram(ctx, d_proj, K, ncols, d_q, kk, d_b, d_part));
    LG_TRY(lg_block_partials_finalize(ctx, d_part, nblk, M, d_sums));
    LG_TRY(lg_codes_factor(ctx, d_sums, d_q, K, kk, d_u, d_sig));
    LG_TRY(lg_codes_vproj(ctx, d_b, kk, ncols, d_u, d_sig, d_v, d_part));
    LG_TRY(lg_block_partials_finalize(ctx, d_part, nblk, kk, d_sums));
    LG_LAUNCH(ctx, k_means_from_sums, 1, 32, 0, d_sums, kk, ncols, d_mean);
    LG_TRY(lg_codes_pack(ctx, d_v, kk, ncols, d_mean, d_codes));
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
