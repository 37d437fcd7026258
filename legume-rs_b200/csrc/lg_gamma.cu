// lg_gamma.cu — stage 5: Poisson-Gamma posterior update, fused elementwise.
//   GammaMatrix::update_stat + calibrate_with   matrix-param/src/dmatrix_gamma.rs:64-123, traits.rs:61-77
//   optimize_block                              data-beans-alg/src/collapse_data/stats.rs:206-368
// digamma / trigamma follow the `special` crate's algorithms (AS 103 / AS 121) in f32.
#include "lg_common.cuh"

//
// The recurrences sum reciprocals (up to 9 for digamma, 5 for trigamma) when the argument is small, which is
// the common case for count sums.  One IEEE division per term made the kernel compute-bound (0.26 ms for 30.7 M
// elements against 0.09 ms of HBM time), so the terms are accumulated as ONE fraction n/d (two FMA-pipe
// operations per term, no overflow: d < 1e7) and divided once; 1/z and ln z use the MUFU approximations
// (2 ulp), far inside the 1e-5 contract.  Whole-number sums below LG_GLUT additionally come from a per-CTA
// shared-memory table of the very same functions, so count data never walks the recurrences at all.
__device__ __forceinline__ float lg_digamma(float p) {
    const float C = 8.5f, S = 1e-5f, S3 = 8.333333333e-2f, S4 = 8.333333333e-3f, S5 = 3.968253968e-3f;
    const float EULER = 0.57721566490153286f;
    if (!(p > 0.0f)) return __int_as_float(0x7fc00000);
    if (p <= S) return -EULER - __fdiv_rn(1.0f, p);
    float value = 0.0f, z = p;
    if (z < C) {
        float n = 0.0f, d = 1.0f;  // sum_i 1/z_i = n / d
        do {
            n = fmaf(n, z, d);
            d *= z;
            z += 1.0f;
        } while (z < C);
        value = -__fdividef(n, d);
    }
    float r = __fdividef(1.0f, z);
    value += __logf(z) - 0.5f * r;
    r *= r;
    value -= r * (S3 - r * (S4 - r * S5));
    return value;
}
__device__ __forceinline__ float lg_trigamma(float x) {
    const float A = 1e-4f, Bc = 5.0f, B2 = 0.1666666667f, B4 = -0.03333333333f, B6 = 0.02380952381f,
                B8 = -0.03333333333f;
    if (!(x > 0.0f)) return __int_as_float(0x7fc00000);
    if (x <= A) return __fdiv_rn(1.0f, x * x);
    float value = 0.0f, z = x;
    if (z < Bc) {
        float n = 0.0f, d = 1.0f;  // sum_i 1/z_i^2 = n / d
        do {
            const float zz = z * z;
            n = fmaf(n, zz, d);
            d *= zz;
            z += 1.0f;
        } while (z < Bc);
        value = __fdividef(n, d);
    }
    const float rz = __fdividef(1.0f, z);
    const float y = rz * rz;
    value += 0.5f * y + (1.0f + y * (B2 + y * (B4 + y * (B6 + y * B8)))) * rz;
    return value;
}

constexpr int LG_GLUT = 1024;  // whole-number sums below this hit the shared-memory tables
struct GammaLut {
    float dig[LG_GLUT];   // digamma(a0 + k)
    float lsd[LG_GLUT];   // sqrt(trigamma(a0 + k))
    float sqa[LG_GLUT];   // sqrt(a0 + k)
};
__device__ __forceinline__ void gamma_lut_build(GammaLut& t, float a0, int target) {
    if (target == LG_TARGET_MEAN_ONLY) return;
    for (int k = threadIdx.x; k < LG_GLUT; k += blockDim.x) {
        const float a = a0 + (float)k;
        t.dig[k] = lg_digamma(a);
        if (target == LG_TARGET_ALL) {
            t.lsd[k] = __fsqrt_rn(lg_trigamma(a));
            t.sqa[k] = __fsqrt_rn(a);
        }
    }
}
// one element of calibrate_with (dmatrix_gamma.rs:96-123); logb = ln(b) is hoisted by the column kernel
template <bool LM, bool ALL>
__device__ __forceinline__ void gamma_element(const GammaLut& t, float x, float a0, float b, float logb, int sparsify,
                                              float& mean, float& sd, float& lm, float& lsd) {
    const float a = a0 + x;
    mean = (sparsify && x == 0.0f) ? 0.0f : __fdiv_rn(a, b);
    const int k = (int)x;
    if ((unsigned)k < (unsigned)LG_GLUT && x == (float)k) {
        if (LM) lm = t.dig[k] - logb;
        if (ALL) {
            sd = __fdiv_rn(t.sqa[k], b);
            lsd = t.lsd[k];
        }
    } else {
        if (LM) lm = lg_digamma(a) - logb;
        if (ALL) {
            sd = __fdiv_rn(__fsqrt_rn(a), b);
            lsd = __fsqrt_rn(lg_trigamma(a));
        }
    }
}

// den[e] is a full plane (GammaMatrix::update_stat with arbitrary B)
__global__ void __launch_bounds__(256) k_gamma_calibrate(const float* __restrict__ num, const float* __restrict__ den,
                                                         uint64_t n, float a0, float b0, int target, int sparsify,
                                                         float* __restrict__ mean, float* __restrict__ sd,
                                                         float* __restrict__ log_mean, float* __restrict__ log_sd) {
    __shared__ GammaLut t;
    gamma_lut_build(t, a0, target);
    __syncthreads();
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += stride) {
        const float x = num[e];
        const float b = b0 + (0.0f + den[e]);
        float m, s = 0.f, lm = 0.f, ls = 0.f;
        if (target == LG_TARGET_ALL) gamma_element<true, true>(t, x, a0, b, logf(b), sparsify, m, s, lm, ls);
        else if (target == LG_TARGET_MEAN_ONLY) gamma_element<false, false>(t, x, a0, b, 0.0f, sparsify, m, s, lm, ls);
        else gamma_element<true, false>(t, x, a0, b, logf(b), sparsify, m, s, lm, ls);
        if (mean) mean[e] = m;
        if (target != LG_TARGET_MEAN_ONLY && log_mean) log_mean[e] = lm;
        if (target == LG_TARGET_ALL) {
            if (sd) sd[e] = s;
            if (log_sd) log_sd[e] = ls;
        }
    }
}

// den = size_s[column] broadcast down each column of the D x S plane (optimize, B <= 1: stats.rs:330-368).
// blockIdx.y walks columns, blockIdx.x gene quads: no per-element division, ln(b) once per column, 128-bit streams.
template <int TARGET, bool VEC4>
__global__ void __launch_bounds__(256) k_gamma_columns(const float* __restrict__ num, const float* __restrict__ size_s,
                                                       uint64_t D, uint32_t S, float a0, float b0, int sparsify,
                                                       float* __restrict__ mean, float* __restrict__ sd,
                                                       float* __restrict__ log_mean, float* __restrict__ log_sd) {
    constexpr bool LM = TARGET != LG_TARGET_MEAN_ONLY, ALL = TARGET == LG_TARGET_ALL;
    __shared__ GammaLut t;
    gamma_lut_build(t, a0, TARGET);
    __syncthreads();
    for (uint32_t col = blockIdx.y; col < S; col += gridDim.y) {
        const float b = b0 + (0.0f + size_s[col]);
        const float logb = LM ? logf(b) : 0.0f;
        const uint64_t base = (uint64_t)col * D;
        if constexpr (VEC4) {
            const uint64_t nq = D >> 2;  // D % 4 == 0 and 16-byte aligned planes (checked by the host)
            const float4* src = reinterpret_cast<const float4*>(num + base);
            for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += (uint64_t)gridDim.x * blockDim.x) {
                const float4 x = __ldcs(src + q);
                float4 m, s, lm, ls;
                gamma_element<LM, ALL>(t, x.x, a0, b, logb, sparsify, m.x, s.x, lm.x, ls.x);
                gamma_element<LM, ALL>(t, x.y, a0, b, logb, sparsify, m.y, s.y, lm.y, ls.y);
                gamma_element<LM, ALL>(t, x.z, a0, b, logb, sparsify, m.z, s.z, lm.z, ls.z);
                gamma_element<LM, ALL>(t, x.w, a0, b, logb, sparsify, m.w, s.w, lm.w, ls.w);
                if (mean) __stcs(reinterpret_cast<float4*>(mean + base) + q, m);
                if (LM && log_mean) __stcs(reinterpret_cast<float4*>(log_mean + base) + q, lm);
                if (ALL && sd) __stcs(reinterpret_cast<float4*>(sd + base) + q, s);
                if (ALL && log_sd) __stcs(reinterpret_cast<float4*>(log_sd + base) + q, ls);
            }
        } else {
            for (uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g < D; g += (uint64_t)gridDim.x * blockDim.x) {
                float m, s, lm, ls;
                gamma_element<LM, ALL>(t, num[base + g], a0, b, logb, sparsify, m, s, lm, ls);
                if (mean) mean[base + g] = m;
                if (LM && log_mean) log_mean[base + g] = lm;
                if (ALL && sd) sd[base + g] = s;
                if (ALL && log_sd) log_sd[base + g] = ls;
            }
        }
    }
}

static unsigned ew_grid(lg_ctx* ctx, uint64_t n) {
    uint64_t g = (n + 255) / 256;
    const uint64_t cap = (uint64_t)ctx->num_sms * 16;
    return (unsigned)(g < cap ? (g ? g : 1) : cap);
}

extern "C" int lg_gamma_calibrate(lg_ctx* ctx, const float* num, const float* den, uint64_t n, float a0, float b0, int target,
                                  float* mean, float* sd, float* log_mean, float* log_sd) {
    if (!ctx) return LG_ERR_INVALID;
    LG_REQUIRE(ctx, num && den, "lg_gamma_calibrate: null argument");
    LG_REQUIRE(ctx, target >= 0 && target <= 2, "lg_gamma_calibrate: bad target");
    cudaSetDevice(ctx->device);
    LgStage st(ctx);
    const float *d_num, *d_den;
    float *d_mean, *d_sd, *d_lm, *d_ls;
    LG_TRY(st.in(num, n, &d_num));
    LG_TRY(st.in(den, n, &d_den));
    LG_TRY(st.out(mean, n, &d_mean));
    LG_TRY(st.out(sd, n, &d_sd));
    LG_TRY(st.out(log_mean, n, &d_lm));
    LG_TRY(st.out(log_sd, n, &d_ls));
    if (n) LG_LAUNCH(ctx, k_gamma_calibrate, ew_grid(ctx, n), 256, 0, d_num, d_den, n, a0, b0, target, 0, d_mean, d_sd, d_lm, d_ls);
    return st.finish();
}

extern "C" int lg_optimize_single(lg_ctx* ctx, const float* sum_ds, const float* size_s, uint64_t D, uint32_t S, float a0,
                                  float b0, int target, float* mean, float* sd, float* log_mean, float* log_sd) {
    return lg_optimize_single_obs(ctx, sum_ds, size_s, nullptr, D, S, a0, b0, target, mean, sd, log_mean, log_sd);
}

extern "C" int lg_optimize_single_obs(lg_ctx* ctx, const float* sum_ds, const float* size_s, const float* size_ds, uint64_t D,
                                      uint32_t S, float a0, float b0, int target, float* mean, float* sd, float* log_mean,
                                      float* log_sd) {
    if (!ctx) return LG_ERR_INVALID;
    LG_REQUIRE(ctx, sum_ds && (size_s || size_ds), "lg_optimize_single: null argument");
    LG_REQUIRE(ctx, target >= 0 && target <= 2, "lg_optimize_single: bad target");
    cudaSetDevice(ctx->device);
    const uint64_t n = D * (uint64_t)S;
    LgStage st(ctx);
    const float *d_num, *d_size;
    float *d_mean, *d_sd, *d_lm, *d_ls;
    const float* d_size_ds;
    LG_TRY(st.in(sum_ds, n, &d_num));
    LG_TRY(st.in(size_s, (size_t)S, &d_size));
    LG_TRY(st.in(size_ds, n, &d_size_ds));
    LG_TRY(st.out(mean, n, &d_mean));
    LG_TRY(st.out(sd, n, &d_sd));
    LG_TRY(st.out(log_mean, n, &d_lm));
    LG_TRY(st.out(log_sd, n, &d_ls));
    // MeanOnly drops the prior baseline where nothing was observed (stats.rs:357-359)
    const int sparsify = target == LG_TARGET_MEAN_ONLY;
    if (n && d_size_ds) {
        // per-(gene, sample) denominators (add_effective_size with size_ds attached, stats.rs:176-186)
        LG_LAUNCH(ctx, k_gamma_calibrate, ew_grid(ctx, n), 256, 0, d_num, d_size_ds, n, a0, b0, target, sparsify, d_mean, d_sd,
                  d_lm, d_ls);
    } else if (n) {
        const uintptr_t align = (uintptr_t)d_num | (uintptr_t)d_mean | (uintptr_t)d_sd | (uintptr_t)d_lm | (uintptr_t)d_ls;
        const bool vec = (D % 4 == 0) && (align & 15) == 0;
        const uint64_t per_col = vec ? D / 4 : D;
        uint64_t gx = (per_col + 255) / 256;
        if (gx > 64) gx = 64;
        uint64_t gy = ((uint64_t)ctx->num_sms * 8 + gx - 1) / gx;  // ~8 CTAs per SM in total
        if (gy > S) gy = S;
        if (gy > 65535) gy = 65535;
        const dim3 grid((unsigned)gx, (unsigned)gy);
#define LG_GAMMA_COLS(T, V)                                                                                              \
    LG_LAUNCH(ctx, (k_gamma_columns<T, V>), grid, 256, 0, d_num, d_size, D, S, a0, b0, sparsify, d_mean, d_sd, d_lm, d_ls)
        if (target == LG_TARGET_ALL) {
            if (vec) LG_GAMMA_COLS(LG_TARGET_ALL, true); else LG_GAMMA_COLS(LG_TARGET_ALL, false);
        } else if (target == LG_TARGET_MEAN_ONLY) {
            if (vec) LG_GAMMA_COLS(LG_TARGET_MEAN_ONLY, true); else LG_GAMMA_COLS(LG_TARGET_MEAN_ONLY, false);
        } else {
            if (vec) LG_GAMMA_COLS(LG_TARGET_MEAN_AND_LOG_MEAN, true); else LG_GAMMA_COLS(LG_TARGET_MEAN_AND_LOG_MEAN, false);
        }
#undef LG_GAMMA_COLS
    }
    return st.finish();
}

// ---------------------------------------------------------------------------------------------
// B > 1 arm: the whole coordinate-descent loop for one (gene, sample) element stays in registers.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_optimize_sweeps(const float* __restrict__ obs, const float* __restrict__ imp,
                                                         const float* __restrict__ res, const float* __restrict__ size_s,
                                                         const float* __restrict__ size_ds, uint64_t n, uint64_t D, float a0,
                                                         float b0, int num_iter, int target,
                                                         float* __restrict__ mu_obs, float* __restrict__ mu_adj,
                                                         float* __restrict__ mu_res, float* __restrict__ gamma,
                                                         float* __restrict__ mu_adj_lm) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += stride) {
        const float o = obs[e], im = imp[e], r = res[e];
        const float sz = size_ds ? size_ds[e] : size_s[e / D];
        const float m_res = __fdiv_rn(a0 + r, b0 + (0.0f + sz));
        float m_gam = 0.0f, m_adj = 0.0f, a_adj = a0, b_adj = b0;
        for (int it = 0; it < num_iter; ++it) {
            float denom = __fmul_rn(__fadd_rn(m_res, m_gam), sz);
            a_adj = a0 + __fadd_rn(o, im);
            b_adj = b0 + denom;
            m_adj = __fdiv_rn(a_adj, b_adj);
            denom = __fmul_rn(m_adj, sz);
            m_gam = __fdiv_rn(a0 + im, b0 + denom);
        }
        const bool sp = target == LG_TARGET_MEAN_ONLY;
        if (mu_obs) mu_obs[e] = (sp && o == 0.0f) ? 0.0f : __fdiv_rn(a0 + o, b0 + (0.0f + sz));
        if (mu_adj) mu_adj[e] = (sp && __fadd_rn(o, im) == 0.0f) ? 0.0f : m_adj;
        if (mu_res) mu_res[e] = (sp && r == 0.0f) ? 0.0f : m_res;
        if (gamma) gamma[e] = (sp && im == 0.0f) ? 0.0f : m_gam;
        if (mu_adj_lm && target != LG_TARGET_MEAN_ONLY) mu_adj_lm[e] = lg_digamma(a_adj) - logf(b_adj);
    }
}

// delta[g,b] = (a0 + obs_db[g,b]) / (b0 + sum_s mu_adj[g,s] * n_bs[b,s])   (stats.rs:296-324)
// one thread per (gene, batch); the sum runs over s in order with fma, genes are the fast axis.
// mu_adj here must be the un-sparsified mean, so it is recomputed by the caller into scratch.
// obs_mask_db (optional, D x B of 0 / 1) multiplies BOTH sides of the ratio so that a masked entry lands on the prior
// exactly (stats.rs:299-322)
__global__ void __launch_bounds__(256) k_optimize_delta(const float* __restrict__ mu_adj, const float* __restrict__ n_bs,
                                                        const float* __restrict__ obs_db, const float* __restrict__ mask_db,
                                                        uint64_t D, uint32_t S, uint32_t B, float a0, float b0,
                                                        float* __restrict__ delta) {
    const uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= D * B) return;
    const uint64_t g = e % D;
    const uint32_t b = (uint32_t)(e / D);
    float acc = 0.0f;
    for (uint32_t s = 0; s < S; ++s) acc = fmaf(mu_adj[(size_t)s * D + g], n_bs[(size_t)s * B + b], acc);
    float num = obs_db[e];
    if (mask_db) {
        acc = __fmul_rn(acc, mask_db[e]);
        num = __fmul_rn(num, mask_db[e]);
    }
    delta[e] = __fdiv_rn(a0 + num, b0 + acc);
}

extern "C" int lg_optimize_batched(lg_ctx* ctx, const float* obs_ds, const float* imp_ds, const float* res_ds,
                                   const float* size_s, const float* obs_db, const float* n_bs, uint64_t D, uint32_t S,
                                   uint32_t B, float a0, float b0, int num_iter, int target, float* mu_obs, float* mu_adj,
                                   float* mu_res, float* gamma, float* delta, float* mu_adj_log_mean) {
    return lg_optimize_batched_obs(ctx, obs_ds, imp_ds, res_ds, size_s, nullptr, obs_db, n_bs, nullptr, D, S, B, a0, b0, num_iter,
                                   target, mu_obs, mu_adj, mu_res, gamma, delta, mu_adj_log_mean);
}

extern "C" int lg_optimize_batched_obs(lg_ctx* ctx, const float* obs_ds, const float* imp_ds, const float* res_ds,
                                       const float* size_s, const float* size_ds, const float* obs_db, const float* n_bs,
                                       const float* obs_mask_db, uint64_t D, uint32_t S, uint32_t B, float a0, float b0,
                                       int num_iter, int target, float* mu_obs, float* mu_adj, float* mu_res, float* gamma,
                                       float* delta, float* mu_adj_log_mean) {
    if (!ctx) return LG_ERR_INVALID;
    LG_REQUIRE(ctx, obs_ds && imp_ds && res_ds && (size_s || size_ds), "lg_optimize_batched: null statistic");
    LG_REQUIRE(ctx, !delta || (obs_db && n_bs), "lg_optimize_batched: delta needs obs_db and n_bs");
    LG_REQUIRE(ctx, target >= 0 && target <= 2 && num_iter >= 0, "lg_optimize_batched: bad target / num_iter");
    cudaSetDevice(ctx->device);
    const uint64_t n = D * (uint64_t)S;
    LgStage st(ctx);
    const float *d_obs, *d_imp, *d_res, *d_size, *d_obs_db, *d_nbs, *d_size_ds, *d_mask;
    float *d_mo, *d_ma, *d_mr, *d_g, *d_delta, *d_lm;
    LG_TRY(st.in(size_ds, n, &d_size_ds));
    LG_TRY(st.in(obs_mask_db, (size_t)D * B, &d_mask));
    LG_TRY(st.in(obs_ds, n, &d_obs));
    LG_TRY(st.in(imp_ds, n, &d_imp));
    LG_TRY(st.in(res_ds, n, &d_res));
    LG_TRY(st.in(size_s, (size_t)S, &d_size));
    LG_TRY(st.in(obs_db, (size_t)D * B, &d_obs_db));
    LG_TRY(st.in(n_bs, (size_t)B * S, &d_nbs));
    LG_TRY(st.out(mu_obs, n, &d_mo));
    LG_TRY(st.out(mu_adj, n, &d_ma));
    LG_TRY(st.out(mu_res, n, &d_mr));
    LG_TRY(st.out(gamma, n, &d_g));
    LG_TRY(st.out(delta, (size_t)D * B, &d_delta));
    LG_TRY(st.out(mu_adj_log_mean, n, &d_lm));
    if (n == 0) return st.finish();
    // delta reads the dense mu_adj mean; under MeanOnly the caller's plane is sparsified, so use scratch
    float* d_ma_dense = d_ma;
    if (d_delta && (target == LG_TARGET_MEAN_ONLY || !d_ma)) {
        LG_TRY(st.scratch(n, &d_ma_dense));
        LG_LAUNCH(ctx, k_optimize_sweeps, ew_grid(ctx, n), 256, 0, d_obs, d_imp, d_res, d_size, d_size_ds, n, D, a0, b0, num_iter,
                  LG_TARGET_ALL, (float*)nullptr, d_ma_dense, (float*)nullptr, (float*)nullptr, (float*)nullptr);
    }
    LG_LAUNCH(ctx, k_optimize_sweeps, ew_grid(ctx, n), 256, 0, d_obs, d_imp, d_res, d_size, d_size_ds, n, D, a0, b0, num_iter,
              target, d_mo, d_ma, d_mr, d_g, d_lm);
    if (d_delta) {
        const uint64_t tot = D * (uint64_t)B;
        LG_LAUNCH(ctx, k_optimize_delta, (unsigned)((tot + 255) / 256), 256, 0, d_ma_dense, d_nbs, d_obs_db, d_mask, D, S, B, a0, b0,
                  d_delta);
    }
    return st.finish();
}

// ---------------------------------------------------------------------------------------------
// attach_observability (collapse_data/mod.rs:221-301): per-(gene, sample) effective sizes and the (gene, batch) mask
// of delta when the backends of a SparseIoVec cover different rows.
// ---------------------------------------------------------------------------------------------
__global__ void k_obs_counts(const uint32_t* __restrict__ source, const uint32_t* __restrict__ group,
                             const uint32_t* __restrict__ batch, const float* __restrict__ mult, uint64_t ncols, uint32_t nsrc,
                             uint32_t S, uint32_t B, float* __restrict__ count_ss, unsigned int* __restrict__ used_sb) {
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= ncols) return;
    const uint32_t src = source[j];
    if (src >= nsrc) return;
    const uint32_t s = group[j];
    if (s < S) atomicAdd(&count_ss[(size_t)s * nsrc + src], mult ? mult[j] : 1.0f);
    if (batch && batch[j] < B) used_sb[(size_t)src * B + batch[j]] = 1u;
}
// size_ds[g, s] = sum over the sources that cover g, in source order, of the mass of sample s that came from the source
__global__ void __launch_bounds__(256) k_obs_size(const uint8_t* __restrict__ cover, const float* __restrict__ count_ss,
                                                  uint64_t D, uint32_t S, uint32_t nsrc, float* __restrict__ size_ds) {
    const uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= D * S) return;
    const uint64_t g = e % D;
    const uint32_t s = (uint32_t)(e / D);
    float acc = 0.0f;
    for (uint32_t src = 0; src < nsrc; ++src)
        if (cover[(size_t)src * D + g]) acc = __fadd_rn(acc, count_ss[(size_t)s * nsrc + src]);
    size_ds[e] = acc;
}
__global__ void __launch_bounds__(256) k_obs_mask(const uint8_t* __restrict__ cover, const unsigned int* __restrict__ used_sb,
                                                  uint64_t D, uint32_t B, uint32_t nsrc, float* __restrict__ mask_db,
                                                  int* __restrict__ any_zero) {
    const uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= D * B) return;
    const uint64_t g = e % D;
    const uint32_t b = (uint32_t)(e / D);
    float m = 0.0f;
    for (uint32_t src = 0; src < nsrc; ++src)
        if (used_sb[(size_t)src * B + b] && cover[(size_t)src * D + g]) m = 1.0f;
    mask_db[e] = m;
    if (m == 0.0f) *any_zero = 1;
}

extern "C" int lg_attach_observability(lg_ctx* ctx, const uint8_t* coverage, uint32_t nsrc, const uint32_t* source_of_cell,
                                       const uint32_t* group_of_cell, const uint32_t* batch_of_cell, const float* mult,
                                       uint64_t ncols, uint64_t D, uint32_t S, uint32_t B, float* out_size_ds,
                                       float* out_mask_db, int* out_mask_has_zero) {
    if (!ctx) return LG_ERR_INVALID;
    LG_REQUIRE(ctx, coverage && source_of_cell && group_of_cell && out_size_ds && nsrc >= 1, "lg_attach_observability: null argument");
    LG_REQUIRE(ctx, !out_mask_db || (batch_of_cell && B >= 1 && out_mask_has_zero), "lg_attach_observability: the mask needs batches");
    cudaSetDevice(ctx->device);
    LgStage st(ctx);
    const uint8_t* d_cov;
    const uint32_t *d_src, *d_grp, *d_bat;
    const float* d_mult;
    float *d_size, *d_mask, *d_count;
    unsigned int* d_used;
    int* d_zero;
    LG_TRY(st.in(coverage, (size_t)nsrc * D, &d_cov));
    LG_TRY(st.in(source_of_cell, (size_t)ncols, &d_src));
    LG_TRY(st.in(group_of_cell, (size_t)ncols, &d_grp));
    LG_TRY(st.in(batch_of_cell, (size_t)ncols, &d_bat));
    LG_TRY(st.in(mult, (size_t)ncols, &d_mult));
    LG_TRY(st.out(out_size_ds, (size_t)D * S, &d_size));
    LG_TRY(st.out(out_mask_db, (size_t)D * B, &d_mask));
    LG_TRY(st.scratch((size_t)S * nsrc, &d_count));
    LG_TRY(st.scratch((size_t)nsrc * (B ? B : 1), &d_used));
    LG_TRY(st.scratch(1, &d_zero));
    LG_CUDA(ctx, cudaMemsetAsync(d_count, 0, sizeof(float) * (size_t)S * nsrc, ctx->stream));
    LG_CUDA(ctx, cudaMemsetAsync(d_used, 0, sizeof(unsigned int) * (size_t)nsrc * (B ? B : 1), ctx->stream));
    LG_CUDA(ctx, cudaMemsetAsync(d_zero, 0, sizeof(int), ctx->stream));
    if (ncols)
        LG_LAUNCH(ctx, k_obs_counts, (unsigned)((ncols + 255) / 256), 256, 0, d_src, d_grp, d_mask ? d_bat : nullptr, d_mult, ncols, nsrc,
                  S, B, d_count, d_used);
    const uint64_t tot = D * (uint64_t)S;
    if (tot) LG_LAUNCH(ctx, k_obs_size, (unsigned)((tot + 255) / 256), 256, 0, d_cov, d_count, D, S, nsrc, d_size);
    if (d_mask) {
        const uint64_t tb = D * (uint64_t)B;
        if (tb) LG_LAUNCH(ctx, k_obs_mask, (unsigned)((tb + 255) / 256), 256, 0, d_cov, d_used, D, B, nsrc, d_mask, d_zero);
        int* h = static_cast<int*>(ctx->pinned);
        LG_CUDA(ctx, cudaMemcpyAsync(h, d_zero, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        LG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        *out_mask_has_zero = *h;
    }
    return st.finish();
}
