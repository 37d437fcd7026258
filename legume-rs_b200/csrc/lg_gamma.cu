// lg_gamma.cu — stage 5: Poisson-Gamma posterior update, fused elementwise.
//   GammaMatrix::update_stat + calibrate_with   matrix-param/src/dmatrix_gamma.rs:64-123, traits.rs:61-77
//   optimize_block                              data-beans-alg/src/collapse_data/stats.rs:206-368
// digamma / trigamma follow the `special` crate's algorithms (AS 103 / AS 121) in f32.
#include "lg_common.cuh"

__device__ __forceinline__ float lg_digamma(float p) {
    const float C = 8.5f, S = 1e-5f, S3 = 8.333333333e-2f, S4 = 8.333333333e-3f, S5 = 3.968253968e-3f;
    const float EULER = 0.57721566490153286f;
    if (!(p > 0.0f)) return __int_as_float(0x7fc00000);
    if (p <= S) return -EULER - __fdiv_rn(1.0f, p);
    float value = 0.0f, z = p;
    while (z < C) {
        value -= __fdiv_rn(1.0f, z);
        z += 1.0f;
    }
    float r = __fdiv_rn(1.0f, z);
    value += logf(z) - 0.5f * r;
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
    while (z < Bc) {
        value += __fdiv_rn(1.0f, z * z);
        z += 1.0f;
    }
    const float y = __fdiv_rn(1.0f, z * z);
    value += 0.5f * y + __fdiv_rn(1.0f + y * (B2 + y * (B4 + y * (B6 + y * B8))), z);
    return value;
}

// den_mode 0: den[e] is a full plane; 1: den = size_s[e / D] broadcast down each column
template <int DEN_MODE>
__global__ void __launch_bounds__(256) k_gamma_calibrate(const float* __restrict__ num, const float* __restrict__ den,
                                                         uint64_t n, uint64_t D, float a0, float b0, int target,
                                                         int sparsify, float* __restrict__ mean, float* __restrict__ sd,
                                                         float* __restrict__ log_mean, float* __restrict__ log_sd) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += stride) {
        const float x = num[e];
        const float dn = DEN_MODE == 0 ? den[e] : (0.0f + den[e / D]);
        const float a = a0 + x, b = b0 + dn;
        if (mean) mean[e] = (sparsify && x == 0.0f) ? 0.0f : __fdiv_rn(a, b);
        if (target != LG_TARGET_MEAN_ONLY && log_mean) log_mean[e] = lg_digamma(a) - logf(b);
        if (target == LG_TARGET_ALL) {
            if (sd) sd[e] = __fdiv_rn(__fsqrt_rn(a), b);
            if (log_sd) log_sd[e] = __fsqrt_rn(lg_trigamma(a));
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
    if (n) LG_LAUNCH(ctx, k_gamma_calibrate<0>, ew_grid(ctx, n), 256, 0, d_num, d_den, n, (uint64_t)1, a0, b0, target, 0, d_mean, d_sd, d_lm, d_ls);
    return st.finish();
}

extern "C" int lg_optimize_single(lg_ctx* ctx, const float* sum_ds, const float* size_s, uint64_t D, uint32_t S, float a0,
                                  float b0, int target, float* mean, float* sd, float* log_mean, float* log_sd) {
    if (!ctx) return LG_ERR_INVALID;
    LG_REQUIRE(ctx, sum_ds && size_s, "lg_optimize_single: null argument");
    LG_REQUIRE(ctx, target >= 0 && target <= 2, "lg_optimize_single: bad target");
    cudaSetDevice(ctx->device);
    const uint64_t n = D * (uint64_t)S;
    LgStage st(ctx);
    const float *d_num, *d_size;
    float *d_mean, *d_sd, *d_lm, *d_ls;
    LG_TRY(st.in(sum_ds, n, &d_num));
    LG_TRY(st.in(size_s, (size_t)S, &d_size));
    LG_TRY(st.out(mean, n, &d_mean));
    LG_TRY(st.out(sd, n, &d_sd));
    LG_TRY(st.out(log_mean, n, &d_lm));
    LG_TRY(st.out(log_sd, n, &d_ls));
    // MeanOnly drops the prior baseline where nothing was observed (stats.rs:357-359)
    const int sparsify = target == LG_TARGET_MEAN_ONLY;
    if (n) LG_LAUNCH(ctx, k_gamma_calibrate<1>, ew_grid(ctx, n), 256, 0, d_num, d_size, n, D, a0, b0, target, sparsify, d_mean, d_sd, d_lm, d_ls);
    return st.finish();
}

// ---------------------------------------------------------------------------------------------
// B > 1 arm: the whole coordinate-descent loop for one (gene, sample) element stays in registers.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_optimize_sweeps(const float* __restrict__ obs, const float* __restrict__ imp,
                                                         const float* __restrict__ res, const float* __restrict__ size_s,
                                                         uint64_t n, uint64_t D, float a0, float b0, int num_iter, int target,
                                                         float* __restrict__ mu_obs, float* __restrict__ mu_adj,
                                                         float* __restrict__ mu_res, float* __restrict__ gamma,
                                                         float* __restrict__ mu_adj_lm) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += stride) {
        const float o = obs[e], im = imp[e], r = res[e];
        const float sz = size_s[e / D];
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
__global__ void __launch_bounds__(256) k_optimize_delta(const float* __restrict__ mu_adj, const float* __restrict__ n_bs,
                                                        const float* __restrict__ obs_db, uint64_t D, uint32_t S, uint32_t B,
                                                        float a0, float b0, float* __restrict__ delta) {
    const uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= D * B) return;
    const uint64_t g = e % D;
    const uint32_t b = (uint32_t)(e / D);
    float acc = 0.0f;
    for (uint32_t s = 0; s < S; ++s) acc = fmaf(mu_adj[(size_t)s * D + g], n_bs[(size_t)s * B + b], acc);
    delta[e] = __fdiv_rn(a0 + obs_db[e], b0 + acc);
}

extern "C" int lg_optimize_batched(lg_ctx* ctx, const float* obs_ds, const float* imp_ds, const float* res_ds,
                                   const float* size_s, const float* obs_db, const float* n_bs, uint64_t D, uint32_t S,
                                   uint32_t B, float a0, float b0, int num_iter, int target, float* mu_obs, float* mu_adj,
                                   float* mu_res, float* gamma, float* delta, float* mu_adj_log_mean) {
    if (!ctx) return LG_ERR_INVALID;
    LG_REQUIRE(ctx, obs_ds && imp_ds && res_ds && size_s, "lg_optimize_batched: null statistic");
    LG_REQUIRE(ctx, !delta || (obs_db && n_bs), "lg_optimize_batched: delta needs obs_db and n_bs");
    LG_REQUIRE(ctx, target >= 0 && target <= 2 && num_iter >= 0, "lg_optimize_batched: bad target / num_iter");
    cudaSetDevice(ctx->device);
    const uint64_t n = D * (uint64_t)S;
    LgStage st(ctx);
    const float *d_obs, *d_imp, *d_res, *d_size, *d_obs_db, *d_nbs;
    float *d_mo, *d_ma, *d_mr, *d_g, *d_delta, *d_lm;
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
        LG_LAUNCH(ctx, k_optimize_sweeps, ew_grid(ctx, n), 256, 0, d_obs, d_imp, d_res, d_size, n, D, a0, b0, num_iter,
                  LG_TARGET_ALL, (float*)nullptr, d_ma_dense, (float*)nullptr, (float*)nullptr, (float*)nullptr);
    }
    LG_LAUNCH(ctx, k_optimize_sweeps, ew_grid(ctx, n), 256, 0, d_obs, d_imp, d_res, d_size, n, D, a0, b0, num_iter, target,
              d_mo, d_ma, d_mr, d_g, d_lm);
    if (d_delta) {
        const uint64_t tot = D * (uint64_t)B;
        LG_LAUNCH(ctx, k_optimize_delta, (unsigned)((tot + 255) / 256), 256, 0, d_ma_dense, d_nbs, d_obs_db, D, S, B, a0, b0, d_delta);
    }
    return st.finish();
}
