// lg_refine.cu — BBKNN + DC-Poisson refinement of the pb-sample partition: the sweeps of
// data-beans-alg/src/dc_poisson.rs on the device.
//
// What the reference does (refine_multilevel.rs:170-298 -> dc_poisson.rs:778-915, RefineParams::parallel = true, the
// default): per level, top-down, every pb-sample ("entity") scores the groups of its candidate set under a Poisson plug-in
//     s(e, k) = size_factor[e] * log_size_offset[k] + sum_{g : y_eg > 0} y_eg * log_gene[k, g]          (dc_poisson.rs:405-431)
// against a frozen snapshot of the statistics (Jacobi sweep, :733-776), picks a group (Gumbel-max sample in the Gibbs sweeps,
// arg-max in the greedy ones), and the proposals are then applied one entity after the other (:661-686, delta_move :352-377).
//
// Layout here: the profiles are the DENSE entity x feature matrix the path already holds (the pb-sample gene sums of
// lg_collapse_basic, weighted in place); a stored entry is a value > 0, exactly the filter of Profiles::from_gene_sums
// (:136-160).  gene_sum f64[k][M] and log_gene f32[k][M] live in HBM (k <= 2^kk groups: at most 368 MB at 30 000 genes).
//   k_dcp_scores   one thread per (entity, candidate) pair walks the entity's row in ascending feature order with ONE f64
//                  accumulator — the reference's order, so the scores are the reference's bit for bit (the products of two
//                  f32 are exact in f64); both rows stream as 128-bit loads.
//   k_dcp_apply    one thread per feature applies the accepted moves in entity order to its column of gene_sum (the order
//                  matters per (group, feature) cell only), coalesced over features.
//   k_dcp_logs     log_gene = (f32) ln(gene_sum + 1e-9).
// The picks (a few thousand Gumbel draws per sweep from per-entity xoshiro256++ streams), the k-long size sums and the sweep
// control (stagnation / no-move exits) stay on the host between the kernels.  The label bookkeeping around the levels
// (compact_labels, project_to_refinement, sibling and candidate sets) is host code in the mirrors.
#include <chrono>
#include <cmath>
#include <cstring>
#include <limits>
#include <thread>
#include <type_traits>

#include "lg_common.cuh"

namespace {

constexpr double DCP_LOG_EPS = 1e-9;  // dc_poisson.rs:33

// rand 0.10 SmallRng on 64-bit targets: xoshiro256++ seeded through SplitMix64
struct SmallRng {
    uint64_t s[4];
    explicit SmallRng(uint64_t seed) {
        for (int i = 0; i < 4; ++i) {
            seed += 0x9E3779B97F4A7C15ull;
            uint64_t z = seed;
            z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
            z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
            s[i] = z ^ (z >> 31);
        }
    }
    static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
    uint64_t next() {
        const uint64_t out = rotl(s[0] + s[3], 23) + s[0], t = s[1] << 17;
        s[2] ^= s[0];
        s[3] ^= s[1];
        s[1] ^= s[2];
        s[0] ^= s[3];
        s[2] ^= t;
        s[3] = rotl(s[3], 45);
        return out;
    }
    double range(double low, double high) {  // UniformFloat<f64>::sample_single
        double scale = high - low;
        for (;;) {
            const uint64_t bits = (next() >> 12) | 0x3FF0000000000000ull;
            double v;
            memcpy(&v, &bits, 8);
            const double res = (v - 1.0) * scale + low;
            if (res < high) return res;
            scale = std::nextafter(scale, 0.0);
        }
    }
};

// per feature, over the entities in ascending order: s1 = sum v, s2 = sum v * v of the stored entries (f32 folds with separate
// multiply and add: SparseRunningStatistics::add_sparse_column, matrix-util/src/sparse_stat.rs:66-79)
__global__ void k_dcp_colstats(const float* __restrict__ P, uint32_t E, uint64_t M, float* __restrict__ s1, float* __restrict__ s2) {
    const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= M) return;
    float a = 0.0f, b = 0.0f;
    for (uint32_t e = 0; e < E; ++e) {
        const float v = P[(size_t)e * M + g];
        if (v > 0.0f && isfinite(v)) {
            a = __fadd_rn(a, v);
            b = __fadd_rn(b, __fmul_rn(v, v));
        }
    }
    s1[g] = a;
    s2[g] = b;
}

// weight_by_vec (dc_poisson.rs:197-213): stored entries times the feature weight; anything else becomes a clean zero
__global__ void k_dcp_weight(float* __restrict__ P, uint64_t total, uint64_t M, const float* __restrict__ w) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const float v = P[i];
    P[i] = v > 0.0f ? (w ? __fmul_rn(v, w[i % M]) : v) : 0.0f;
}

// size factor of an entity: the serial f32 fold of its stored entries in ascending feature order (:150, :204-210)
__global__ void k_dcp_size_factor(const float* __restrict__ P, uint32_t E, uint64_t M, float* __restrict__ sf) {
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    const float* row = P + (size_t)e * M;
    float s = 0.0f;
    for (uint64_t g = 0; g < M; ++g) {
        const float v = row[g];
        if (v > 0.0f) s = __fadd_rn(s, v);
    }
    sf[e] = s;
}

// DcPoissonStats::from_profiles (:318-330): gene_sum[z_e][g] += v over the entities in ascending order
__global__ void k_dcp_init_sums(const float* __restrict__ P, uint32_t E, uint64_t M, const uint32_t* __restrict__ label,
                                double* __restrict__ gs) {
    const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= M) return;
    for (uint32_t e = 0; e < E; ++e) {
        const float v = P[(size_t)e * M + g];
        if (v > 0.0f) gs[(size_t)label[e] * M + g] += (double)v;
    }
}

__global__ void k_dcp_logs(const double* __restrict__ gs, uint64_t total, float* __restrict__ lg) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < total) lg[i] = (float)log(gs[i] + DCP_LOG_EPS);
}

// one thread per (entity, candidate group) pair
template <bool VEC>
__global__ void __launch_bounds__(128) k_dcp_scores(const float* __restrict__ P, uint64_t M, const float* __restrict__ sf,
                                                    const float* __restrict__ lg, const float* __restrict__ lso,
                                                    const uint32_t* __restrict__ pair_e, const uint32_t* __restrict__ pair_k, uint64_t npairs,
                                                    double* __restrict__ out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npairs) return;
    const uint32_t e = pair_e[i], k = pair_k[i];
    const float* row = P + (size_t)e * M;
    const float* lrow = lg + (size_t)k * M;
    double acc = (double)sf[e] * (double)lso[k];
    if (VEC) {
        const float4* r4 = reinterpret_cast<const float4*>(row);
        const float4* l4 = reinterpret_cast<const float4*>(lrow);
        for (uint64_t q = 0; q < M / 4; ++q) {
            const float4 v = r4[q], l = l4[q];
            // the product of two f32 is exact in f64, so each fma is the reference's `acc += v as f64 * l as f64`
            if (v.x > 0.0f) acc = fma((double)v.x, (double)l.x, acc);
            if (v.y > 0.0f) acc = fma((double)v.y, (double)l.y, acc);
            if (v.z > 0.0f) acc = fma((double)v.z, (double)l.z, acc);
            if (v.w > 0.0f) acc = fma((double)v.w, (double)l.w, acc);
        }
    } else {
        for (uint64_t g = 0; g < M; ++g) {
            const float v = row[g];
            if (v > 0.0f) acc = fma((double)v, (double)lrow[g], acc);
        }
    }
    out[i] = acc;
}

// The same scores with the operands staged through shared memory.  In the one-thread-per-pair form every pair streams its own
// copy of a log_gene row (4 * M bytes per pair and sweep out of L2: 14 GB per sweep at 120 000 pairs x 30 000 genes, which is
// what bounded it).  Here a CTA owns up to DCP_PAIRS consecutive pairs — consecutive entities, whose candidate groups overlap —
// and walks the feature axis in tiles of DCP_GT: all threads copy the tile of the CTA's DISTINCT log_gene rows and of its
// entities' profile rows into shared memory (coalesced 512-byte row pieces), then thread p continues ITS pair's single f64
// accumulator over the tile, ascending.  The order of every pair's additions is unchanged, so the scores are bit-identical to
// k_dcp_scores; the traffic drops to (rows + entities) * 4 * M bytes per CTA.  Row stride DCP_GT + 1: lanes that read different
// rows at the same column hit different banks, lanes on the same row broadcast.
constexpr int DCP_PAIRS = 512, DCP_ROWS = 128, DCP_ENTS = 64, DCP_GT = 128, DCP_LD = DCP_GT + 1;
constexpr size_t DCP_SMEM = (size_t)(DCP_ROWS + DCP_ENTS) * DCP_LD * sizeof(float);
__global__ void __launch_bounds__(DCP_PAIRS, 2) k_dcp_scores_tiled(const float* __restrict__ P, uint64_t M, const float* __restrict__ sf,
                                                                  const float* __restrict__ lg, const float* __restrict__ lso,
                                                                  const uint32_t* __restrict__ pair_e, const uint32_t* __restrict__ pair_k,
                                                                  const uint32_t* __restrict__ cta_pair, const uint32_t* __restrict__ cta_row,
                                                                  const uint32_t* __restrict__ rows, const uint32_t* __restrict__ cta_ent,
                                                                  const uint32_t* __restrict__ ents, const uint16_t* __restrict__ pair_slot,
                                                                  const uint16_t* __restrict__ pair_el, double* __restrict__ out) {
    extern __shared__ float dcp_sm[];  // (nr + ne) rows of DCP_LD floats: the log_gene rows first, then the profile rows
    __shared__ const float* base[DCP_ROWS + DCP_ENTS];
    const uint32_t p0 = cta_pair[blockIdx.x], np = cta_pair[blockIdx.x + 1] - p0;
    const uint32_t r0 = cta_row[blockIdx.x], nr = cta_row[blockIdx.x + 1] - r0;
    const uint32_t e0 = cta_ent[blockIdx.x], ne = cta_ent[blockIdx.x + 1] - e0;
    if (threadIdx.x < nr + ne)
        base[threadIdx.x] = threadIdx.x < nr ? lg + (size_t)rows[r0 + threadIdx.x] * M : P + (size_t)ents[e0 + threadIdx.x - nr] * M;
    const bool live = threadIdx.x < np;
    double acc = 0.0;
    const float *my_l = dcp_sm, *my_p = dcp_sm;
    if (live) {
        const uint32_t i = p0 + threadIdx.x;
        acc = (double)sf[pair_e[i]] * (double)lso[pair_k[i]];
        my_l = dcp_sm + (size_t)pair_slot[i] * DCP_LD;
        my_p = dcp_sm + (size_t)(nr + pair_el[i]) * DCP_LD;
    }
    const uint32_t total = (nr + ne) * DCP_GT;
    constexpr int U = 8;  // row pieces in flight per thread: the copy is latency-bound otherwise (one L2 round trip per element)
    for (uint64_t g0 = 0; g0 < M; g0 += DCP_GT) {
        __syncthreads();  // the previous tile has been consumed (and `base` is visible on the first round)
        for (uint32_t idx = threadIdx.x; idx < total; idx += DCP_PAIRS * U) {
            float v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const uint32_t i = idx + u * DCP_PAIRS;
                v[u] = 0.0f;
                if (i < total) {
                    const uint64_t g = g0 + (i % DCP_GT);
                    if (g < M) v[u] = base[i / DCP_GT][g];
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const uint32_t i = idx + u * DCP_PAIRS;
                if (i < total) dcp_sm[(i / DCP_GT) * DCP_LD + (i % DCP_GT)] = v[u];
            }
        }
        __syncthreads();
        if (live) {
#pragma unroll 8
            for (int j = 0; j < DCP_GT; ++j) {
                const float v = my_p[j];
                if (v > 0.0f) acc = fma((double)v, (double)my_l[j], acc);  // columns past M hold 0
            }
        }
    }
    if (live) out[p0 + threadIdx.x] = acc;
}

// delta_move (:352-377) for a list of accepted moves in entity order; one thread per feature
__global__ void k_dcp_apply(const float* __restrict__ P, uint64_t M, const uint32_t* __restrict__ mv, uint32_t nmoves, double* __restrict__ gs) {
    const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= M) return;
    for (uint32_t i = 0; i < nmoves; ++i) {
        const uint32_t e = mv[3 * i], from = mv[3 * i + 1], to = mv[3 * i + 2];
        const float v = P[(size_t)e * M + g];
        if (v > 0.0f) {
            gs[(size_t)from * M + g] -= (double)v;
            gs[(size_t)to * M + g] += (double)v;
        }
    }
}

// nb_dispersion.rs:58-151 on the host: D-long vectors, f64 weighted log-linear fit
struct Trend {
    float a = -std::numeric_limits<float>::infinity(), b = 0.0f;
    float phi_at(float mu) const {
        if (!std::isfinite(mu) || mu <= 0.0f) return 0.0f;
        const float phi = std::exp(a + b * std::log(mu));
        return std::fmin(std::fmax(phi, 0.0f), 100.0f);
    }
};
Trend fit_trend(const std::vector<float>& means, const std::vector<float>& vars) {
    std::vector<double> x, y, w;
    for (size_t i = 0; i < means.size(); ++i) {
        const float mu = means[i], var = vars[i];
        if (!std::isfinite(mu) || !std::isfinite(var) || mu < 1e-4f) continue;
        const double phi_hat = (double)((var - mu) / (mu * mu));
        if (phi_hat <= 0.0) continue;
        x.push_back(std::log((double)mu));
        y.push_back(std::log(phi_hat));
        w.push_back((double)mu);
    }
    Trend t;
    if (x.size() < 2) return t;
    double ws = 0.0, xm = 0.0, ym = 0.0;
    for (double v : w) ws += v;
    for (size_t i = 0; i < x.size(); ++i) xm += x[i] * w[i];
    for (size_t i = 0; i < x.size(); ++i) ym += y[i] * w[i];
    xm /= ws;
    ym /= ws;
    double sxx = 0.0, sxy = 0.0;
    for (size_t i = 0; i < x.size(); ++i) {
        const double dx = x[i] - xm;
        sxx += w[i] * dx * dx;
        sxy += w[i] * dx * (y[i] - ym);
    }
    if (sxx <= 0.0) {
        t.a = (float)ym;
        return t;
    }
    const double b = sxy / sxx;
    t.a = (float)(ym - b * xm);
    t.b = (float)b;
    return t;
}

}  // namespace

// Profiles::nb_fisher_weights (dc_poisson.rs:230-295): w_g = 1 / (1 + pi_g * mean size * phi(mean_g)), phi the NB dispersion
// trend fitted over the features.  profiles: npb x D (an entity's row contiguous), out_w: D, host or device.
extern "C" int lg_dcp_fisher_weights(lg_ctx* ctx, const float* profiles, uint64_t D, uint32_t npb, float* out_w) {
    if (!ctx) return LG_ERR_INVALID;
    LG_REQUIRE(ctx, profiles && out_w && D >= 1, "lg_dcp_fisher_weights: null argument");
    cudaSetDevice(ctx->device);
    LgStage st(ctx);
    const float* d_p;
    float *d_s1, *d_s2;
    LG_TRY(st.in(profiles, (size_t)npb * D, &d_p));
    LG_TRY(st.scratch((size_t)D, &d_s1));
    LG_TRY(st.scratch((size_t)D, &d_s2));
    LG_LAUNCH(ctx, k_dcp_colstats, (unsigned)((D + 127) / 128), 128, 0, d_p, npb, D, d_s1, d_s2);
    std::vector<float> s1(D), s2(D), means(D), vars(D), w(D);
    LG_CUDA(ctx, cudaMemcpyAsync(s1.data(), d_s1, sizeof(float) * D, cudaMemcpyDeviceToHost, ctx->stream));
    LG_CUDA(ctx, cudaMemcpyAsync(s2.data(), d_s2, sizeof(float) * D, cudaMemcpyDeviceToHost, ctx->stream));
    LG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const float n = npb > 0 ? (float)npb : 1e-8f;  // safe_denom (sparse_stat.rs:17-24)
    for (uint64_t g = 0; g < D; ++g) {
        const float mu = s1[g] / n;
        means[g] = mu;
        vars[g] = s2[g] / n - mu * mu;
    }
    const Trend trend = fit_trend(means, vars);
    double total = 0.0;
    for (uint64_t g = 0; g < D; ++g) total += (double)s1[g];
    const float avg_s = npb > 0 ? (float)(total / (double)npb) : 1.0f;
    const float inv_total = total > 0.0 ? 1.0f / (float)total : 0.0f;
    for (uint64_t g = 0; g < D; ++g) w[g] = 1.0f / (1.0f + s1[g] * inv_total * avg_s * trend.phi_at(means[g]));
    LG_CUDA(ctx, cudaMemcpyAsync(out_w, w.data(), sizeof(float) * D, cudaMemcpyDefault, ctx->stream));
    LG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return LG_OK;
}

// Profiles::from_gene_sums + weight_by_vec (dc_poisson.rs:136-213): the stored entries (> 0) of `profiles` (npb x D, in place)
// times the feature weights (NULL: unweighted), and every entity's size factor (serial f32 fold of its stored entries).
extern "C" int lg_dcp_profiles(lg_ctx* ctx, float* profiles, uint64_t D, uint32_t npb, const float* weights, float* out_size_factor) {
    if (!ctx) return LG_ERR_INVALID;
    LG_REQUIRE(ctx, profiles && out_size_factor && D >= 1, "lg_dcp_profiles: null argument");
    cudaSetDevice(ctx->device);
    LgStage st(ctx);
    float *d_p, *d_sf;
    const float* d_w;
    LG_TRY(st.out(profiles, (size_t)npb * D, &d_p, true));
    LG_TRY(st.in(weights, (size_t)D, &d_w));
    LG_TRY(st.out(out_size_factor, (size_t)npb, &d_sf));
    const uint64_t total = (uint64_t)npb * D;
    if (total) {
        LG_LAUNCH(ctx, k_dcp_weight, (unsigned)((total + 255) / 256), 256, 0, d_p, total, D, d_w);
        LG_LAUNCH(ctx, k_dcp_size_factor, (npb + 63) / 64, 64, 0, d_p, npb, D, d_sf);
    }
    return st.finish();
}

// refine_with_candidates_guarded (dc_poisson.rs:778-915) with Jacobi sweeps and no move guard, for ONE level.
//   profiles npb x D (weighted; stored entry = value > 0), size_factor npb; candidates as CSR (cand_ptr npb + 1 entries, groups
//   ascending inside an entity, the entity's own group among them); labels npb in / out, values < k; jacobi_base_seed = the odd
//   u64 the caller drew from the refinement's SmallRng for this level (:824).  *out_moves: accepted moves over all sweeps.
extern "C" int lg_dcp_refine_level(lg_ctx* ctx, const float* profiles, const float* size_factor, uint64_t D, uint32_t npb,
                                   const uint32_t* cand_ptr, const uint32_t* cand, uint32_t k, int num_gibbs, int num_greedy,
                                   uint64_t jacobi_base_seed, double stagnation, uint32_t* labels, uint64_t* out_moves) {
    if (!ctx) return LG_ERR_INVALID;
    LG_REQUIRE(ctx, profiles && size_factor && cand_ptr && labels && D >= 1 && k >= 1, "lg_dcp_refine_level: null argument");
    LG_REQUIRE(ctx, num_gibbs >= 0 && num_greedy >= 0, "lg_dcp_refine_level: negative sweep count");
    cudaSetDevice(ctx->device);
    if (out_moves) *out_moves = 0;
    if (npb == 0) return LG_OK;
    LgStage st(ctx);
    const float *d_p, *d_sf_in;
    LG_TRY(st.in(profiles, (size_t)npb * D, &d_p));
    LG_TRY(st.in(size_factor, (size_t)npb, &d_sf_in));
    // host copies of the small arrays: candidate sets, labels, size factors
    std::vector<uint32_t> cp(npb + 1), mem(npb);
    std::vector<float> sf(npb);
    LG_CUDA(ctx, cudaMemcpyAsync(cp.data(), cand_ptr, sizeof(uint32_t) * (npb + 1), cudaMemcpyDefault, ctx->stream));
    LG_CUDA(ctx, cudaMemcpyAsync(mem.data(), labels, sizeof(uint32_t) * npb, cudaMemcpyDefault, ctx->stream));
    LG_CUDA(ctx, cudaMemcpyAsync(sf.data(), size_factor, sizeof(float) * npb, cudaMemcpyDefault, ctx->stream));
    LG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    LG_REQUIRE(ctx, cp[0] == 0, "lg_dcp_refine_level: cand_ptr must start at 0");
    for (uint32_t e = 0; e < npb; ++e) LG_REQUIRE(ctx, cp[e + 1] >= cp[e], "lg_dcp_refine_level: cand_ptr not monotone");
    const uint32_t ncand = cp[npb];
    std::vector<uint32_t> cd(ncand ? ncand : 1);
    if (ncand) {
        LG_REQUIRE(ctx, cand, "lg_dcp_refine_level: null candidates");
        LG_CUDA(ctx, cudaMemcpyAsync(cd.data(), cand, sizeof(uint32_t) * ncand, cudaMemcpyDefault, ctx->stream));
        LG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    for (uint32_t e = 0; e < npb; ++e) {
        LG_REQUIRE(ctx, mem[e] < k, "lg_dcp_refine_level: label out of range");
        for (uint32_t i = cp[e]; i < cp[e + 1]; ++i) LG_REQUIRE(ctx, cd[i] < k, "lg_dcp_refine_level: candidate group out of range");
    }
    // the (entity, group) pairs that are scored: entities with fewer than two candidates stay where they are (:752-755)
    std::vector<uint32_t> pe, pk, first(npb + 1, 0);
    for (uint32_t e = 0; e < npb; ++e) {
        first[e] = (uint32_t)pe.size();
        if (cp[e + 1] - cp[e] >= 2)
            for (uint32_t i = cp[e]; i < cp[e + 1]; ++i) {
                pe.push_back(e);
                pk.push_back(cd[i]);
            }
    }
    first[npb] = (uint32_t)pe.size();
    const uint64_t npairs = pe.size();
    if (npairs == 0 || (num_gibbs == 0 && num_greedy == 0)) return LG_OK;

    const uint64_t KM = (uint64_t)k * D;
    double *d_gs, *d_scores;
    float *d_lg, *d_lso;
    uint32_t *d_pe, *d_pk, *d_mv, *d_lab;
    LG_TRY(st.scratch((size_t)KM, &d_gs));
    LG_TRY(st.scratch((size_t)KM, &d_lg));
    LG_TRY(st.scratch((size_t)k, &d_lso));
    LG_TRY(st.scratch((size_t)npairs, &d_scores));
    LG_TRY(st.scratch((size_t)npairs, &d_pe));
    LG_TRY(st.scratch((size_t)npairs, &d_pk));
    LG_TRY(st.scratch((size_t)3 * npb, &d_mv));
    LG_TRY(st.scratch((size_t)npb, &d_lab));
    LG_CUDA(ctx, cudaMemcpyAsync(d_pe, pe.data(), sizeof(uint32_t) * npairs, cudaMemcpyHostToDevice, ctx->stream));
    LG_CUDA(ctx, cudaMemcpyAsync(d_pk, pk.data(), sizeof(uint32_t) * npairs, cudaMemcpyHostToDevice, ctx->stream));
    LG_CUDA(ctx, cudaMemcpyAsync(d_lab, mem.data(), sizeof(uint32_t) * npb, cudaMemcpyHostToDevice, ctx->stream));
    LG_CUDA(ctx, cudaMemsetAsync(d_gs, 0, sizeof(double) * KM, ctx->stream));
    LG_LAUNCH(ctx, k_dcp_init_sums, (unsigned)((D + 127) / 128), 128, 0, d_p, npb, D, d_lab, d_gs);
    LG_LAUNCH(ctx, k_dcp_logs, (unsigned)((KM + 255) / 256), 256, 0, d_gs, KM, d_lg);
    // size sums and their log offsets: k-long, on the host in f64 (:329, :338-342, :369-374)
    const double m_eps = (double)D * DCP_LOG_EPS;
    std::vector<double> size_sum(k, 0.0);
    for (uint32_t e = 0; e < npb; ++e) size_sum[mem[e]] += (double)sf[e];
    std::vector<float> lso(k);
    for (uint32_t c = 0; c < k; ++c) lso[c] = (float)(-std::log(size_sum[c] + m_eps));
    LG_CUDA(ctx, cudaMemcpyAsync(d_lso, lso.data(), sizeof(float) * k, cudaMemcpyHostToDevice, ctx->stream));

    const bool vec = (D % 4 == 0) && ((reinterpret_cast<uintptr_t>(d_p) & 15) == 0) && ((reinterpret_cast<uintptr_t>(d_lg) & 15) == 0);
    // plan of the tiled score kernel: consecutive entities go to one CTA while its pairs, its distinct candidate groups and its
    // entities fit (DCP_PAIRS / DCP_ROWS / DCP_ENTS); the candidate sets are fixed for the level, so the plan is built once.
    // An entity with more than DCP_ROWS candidates (a coarsest level beyond 2^7 groups) keeps the one-thread-per-pair kernel.
    bool tiled = !getenv("LG_DCP_UNTILED") && DCP_SMEM <= ctx->smem_optin;
    std::vector<uint32_t> cta_pair{0}, cta_row{0}, cta_ent{0}, rows_flat, ents_flat;
    std::vector<uint16_t> pslot(npairs), pel(npairs);
    if (tiled) {
        std::vector<uint32_t> slot_of(k, 0xFFFFFFFFu), cur_rows;
        uint32_t cur_pairs = 0, cur_ents = 0;
        auto close = [&]() {
            for (uint32_t r : cur_rows) slot_of[r] = 0xFFFFFFFFu;
            rows_flat.insert(rows_flat.end(), cur_rows.begin(), cur_rows.end());
            cta_pair.push_back(cta_pair.back() + cur_pairs);
            cta_row.push_back((uint32_t)rows_flat.size());
            cta_ent.push_back((uint32_t)ents_flat.size());
            cur_rows.clear();
            cur_pairs = cur_ents = 0;
        };
        for (uint32_t e = 0; e < npb && tiled; ++e) {
            const uint32_t a = first[e], b = first[e + 1];
            if (a == b) continue;
            if (b - a > (uint32_t)DCP_ROWS || b - a > (uint32_t)DCP_PAIRS) {
                tiled = false;
                break;
            }
            uint32_t fresh = 0;
            for (uint32_t i = a; i < b; ++i) fresh += slot_of[pk[i]] == 0xFFFFFFFFu;
            if (cur_pairs + (b - a) > (uint32_t)DCP_PAIRS || cur_rows.size() + fresh > (size_t)DCP_ROWS || cur_ents + 1 > (uint32_t)DCP_ENTS) close();
            for (uint32_t i = a; i < b; ++i) {
                if (slot_of[pk[i]] == 0xFFFFFFFFu) {
                    slot_of[pk[i]] = (uint32_t)cur_rows.size();
                    cur_rows.push_back(pk[i]);
                }
                pslot[i] = (uint16_t)slot_of[pk[i]];
                pel[i] = (uint16_t)cur_ents;
            }
            ents_flat.push_back(e);
            cur_pairs += b - a;
            ++cur_ents;
        }
        if (tiled && cur_pairs) close();
    }
    uint32_t *d_cta_pair = nullptr, *d_cta_row = nullptr, *d_cta_ent = nullptr, *d_rows = nullptr, *d_ents = nullptr;
    uint16_t *d_pslot = nullptr, *d_pel = nullptr;
    const uint32_t ncta = tiled ? (uint32_t)cta_pair.size() - 1 : 0;
    if (tiled) {
        auto up = [&](const auto& v, auto** d) -> int {
            using T = typename std::remove_reference<decltype(v)>::type::value_type;
            LG_TRY(st.scratch(v.size(), d));
            if (!v.empty()) LG_CUDA(ctx, cudaMemcpyAsync(*d, v.data(), sizeof(T) * v.size(), cudaMemcpyHostToDevice, ctx->stream));
            return LG_OK;
        };
        LG_TRY(up(cta_pair, &d_cta_pair));
        LG_TRY(up(cta_row, &d_cta_row));
        LG_TRY(up(cta_ent, &d_cta_ent));
        LG_TRY(up(rows_flat, &d_rows));
        LG_TRY(up(ents_flat, &d_ents));
        LG_TRY(up(pslot, &d_pslot));
        LG_TRY(up(pel, &d_pel));
        LG_CUDA(ctx, cudaFuncSetAttribute(k_dcp_scores_tiled, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DCP_SMEM));
        LG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // the plan vectors are host temporaries
    }
    std::vector<double> scores(npairs);
    std::vector<uint32_t> prop(npb), mv;
    uint64_t total_moves = 0;
    // LG_TRACE=1: where a level's time goes (scores kernel + read-back, the host picks, applying the moves), on stderr
    const char* tr = getenv("LG_TRACE");
    const bool trace = tr && tr[0] == '1';
    double t_score = 0.0, t_pick = 0.0, t_apply = 0.0;
    int nsweeps = 0;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
        return std::chrono::duration<double, std::milli>(b - a).count();
    };
    auto sweep = [&](bool gibbs, uint64_t sweep_seed, uint64_t* moved) -> int {
        const auto t0 = now();
        if (tiled)
            LG_LAUNCH(ctx, k_dcp_scores_tiled, ncta, DCP_PAIRS, DCP_SMEM, d_p, D, d_sf_in, d_lg, d_lso, d_pe, d_pk, d_cta_pair, d_cta_row, d_rows,
                      d_cta_ent, d_ents, d_pslot, d_pel, d_scores);
        else if (vec) LG_LAUNCH(ctx, k_dcp_scores<true>, (unsigned)((npairs + 127) / 128), 128, 0, d_p, D, d_sf_in, d_lg, d_lso, d_pe, d_pk, npairs, d_scores);
        else LG_LAUNCH(ctx, k_dcp_scores<false>, (unsigned)((npairs + 127) / 128), 128, 0, d_p, D, d_sf_in, d_lg, d_lso, d_pe, d_pk, npairs, d_scores);
        LG_CUDA(ctx, cudaMemcpyAsync(scores.data(), d_scores, sizeof(double) * npairs, cudaMemcpyDeviceToHost, ctx->stream));
        LG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        const auto t1 = now();
        // the picks are independent per entity (per-entity RNG streams): a few host threads share them
        auto pick_range = [&](uint32_t e_lo, uint32_t e_hi) {
        for (uint32_t e = e_lo; e < e_hi; ++e) {
            prop[e] = mem[e];
            const uint32_t a = first[e], b = first[e + 1];
            if (a == b) continue;
            if (gibbs) {  // sample_categorical_log (:451-471): one draw per finite slot, ascending group
                SmallRng rng(sweep_seed ^ ((uint64_t)e * 2654435761ull));
                double best = -std::numeric_limits<double>::infinity();
                uint32_t pick = pk[a];
                for (uint32_t i = a; i < b; ++i) {
                    if (!std::isfinite(scores[i])) continue;
                    const double u = rng.range(1e-12, 1.0);
                    const double key = scores[i] + (-std::log(-std::log(u)));
                    if (key > best) {
                        best = key;
                        pick = pk[i];
                    }
                }
                prop[e] = pick;
            } else {  // argmax_log_restricted (:473-488): the first of equal scores
                uint32_t bi = a;
                for (uint32_t i = a + 1; i < b; ++i)
                    if (scores[i] > scores[bi]) bi = i;
                prop[e] = pk[bi];
            }
        }
        };
        {
            const unsigned hc = std::thread::hardware_concurrency();
            const uint32_t nt = gibbs && npairs >= 16384 ? std::min<uint32_t>(hc ? hc : 4u, 8u) : 1u;
            std::vector<std::thread> pool;
            const uint32_t per = (npb + nt - 1) / nt;
            for (uint32_t t = 1; t < nt; ++t) pool.emplace_back(pick_range, std::min(npb, t * per), std::min(npb, (t + 1) * per));
            pick_range(0, std::min(npb, per));
            for (auto& th : pool) th.join();
        }
        // apply_proposals (:661-686) in entity order
        mv.clear();
        for (uint32_t e = 0; e < npb; ++e) {
            if (prop[e] == mem[e]) continue;
            mv.push_back(e);
            mv.push_back(mem[e]);
            mv.push_back(prop[e]);
            const double s = (double)sf[e];
            size_sum[mem[e]] -= s;
            size_sum[prop[e]] += s;
            lso[mem[e]] = (float)(-std::log(size_sum[mem[e]] + m_eps));
            lso[prop[e]] = (float)(-std::log(size_sum[prop[e]] + m_eps));
            mem[e] = prop[e];
        }
        *moved = mv.size() / 3;
        const auto t2 = now();
        if (*moved) {
            LG_CUDA(ctx, cudaMemcpyAsync(d_mv, mv.data(), sizeof(uint32_t) * mv.size(), cudaMemcpyHostToDevice, ctx->stream));
            LG_CUDA(ctx, cudaMemcpyAsync(d_lso, lso.data(), sizeof(float) * k, cudaMemcpyHostToDevice, ctx->stream));
            LG_LAUNCH(ctx, k_dcp_apply, (unsigned)((D + 127) / 128), 128, 0, d_p, D, d_mv, (uint32_t)*moved, d_gs);
            LG_LAUNCH(ctx, k_dcp_logs, (unsigned)((KM + 255) / 256), 256, 0, d_gs, KM, d_lg);
            LG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // mv / lso are reused by the next sweep
        }
        t_score += ms(t0, t1);
        t_pick += ms(t1, t2);
        t_apply += ms(t2, now());
        ++nsweeps;
        return LG_OK;
    };
    int low = 0;
    for (int s = 0; s < num_gibbs; ++s) {
        uint64_t moved = 0;
        LG_TRY(sweep(true, jacobi_base_seed * (uint64_t)(s + 1), &moved));
        total_moves += moved;
        if (stagnation > 0.0) {
            if ((double)moved < stagnation * (double)npb) {
                if (++low >= 3) break;
            } else {
                low = 0;
            }
        }
    }
    for (int s = 0; s < num_greedy; ++s) {
        uint64_t moved = 0;
        LG_TRY(sweep(false, 0, &moved));
        total_moves += moved;
        if (moved == 0) break;
    }
    LG_CUDA(ctx, cudaMemcpyAsync(labels, mem.data(), sizeof(uint32_t) * npb, cudaMemcpyDefault, ctx->stream));
    LG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (out_moves) *out_moves = total_moves;
    if (trace)
        fprintf(stderr, "[lg_dcp_refine_level] %u entities x %llu features, k = %u, %llu scored pairs, %d sweeps, %llu moves: scores %.1f ms, picks %.1f ms, apply %.1f ms\n",
                npb, (unsigned long long)D, k, (unsigned long long)npairs, nsweeps, (unsigned long long)total_moves, t_score, t_pick, t_apply);
    return LG_OK;
}
