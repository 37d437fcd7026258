/*
 * oracle_refine.cpp — CPU restatement of the BBKNN + DC-Poisson refinement of the pb-sample partition
 * (SURVEY.md section 8f rank 3).  TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Restated from the reference (all citations relative to /root/reference, causalpathlab/legume-rs v0.3.2):
 *   data-beans-alg/src/dc_poisson.rs
 *     :128-160  Profiles::from_gene_sums        :197-213  weight_by_vec          :230-295  nb_fisher_weights
 *     :307-350  DcPoissonStats::from_profiles   :352-377  delta_move             :413-431  compute_log_probs_restricted
 *     :451-471  sample_categorical_log          :473-488  argmax_log_restricted  :493-509  compact_labels
 *     :518-550  compute_sibling_sets            :599-633  intersect_with_siblings_fallback
 *     :661-686  apply_proposals                 :733-776  sweep_jacobi           :778-915  refine_with_candidates_guarded
 *               (params.parallel = true, the default, with NoGuard; the Gauss-Seidel arm is not restated)
 *   data-beans-alg/src/nb_dispersion.rs :58-151   DispersionTrend::fit / phi_at / fisher_weight
 *   matrix-util/src/sparse_stat.rs :66-79, 643-659  add_sparse_column, mean, variance
 *   data-beans-alg/src/refine_multilevel.rs
 *     :85-112   build_candidate_sets            :170-298  refine_assignments
 *     :315-345  project_to_refinement / child_offset_within_parent
 *
 * Pinned on the reference's own tests (tests/test_oracle_refine.py): compact_labels, the two sibling-set cases
 * (dc_poisson_tests.rs:93-130), the candidate fall-back, child_offset_within_parent and project_to_refinement
 * (refine_multilevel_tests.rs:4-56), delta moves == recompute and restricted == unrestricted scores
 * (dc_poisson_tests.rs:38-91), an empty block stays finite (:84-91).
 * Third-party arithmetic restated here, PARITY UNPINNED: rand 0.10.1 SmallRng (xoshiro256++ seeded through SplitMix64),
 * UniformFloat<f64>::sample_single (53-bit mantissa draw, value0_1 * scale + low), libm log.  The reference's Fisher
 * weights come out of a rayon fold whose f32 addition order depends on the thread count; here the rows are folded in
 * ascending entity order.
 */
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <map>
#include <unordered_map>
#include <utility>
#include <vector>

#include "oracle.h"

namespace {

constexpr double LOG_EPS = 1e-9;  // dc_poisson.rs:33

// ---- rand 0.10 SmallRng on 64-bit targets ------------------------------------------------------
struct SmallRng {
    uint64_t s[4];
    explicit SmallRng(uint64_t seed) {  // Xoshiro256PlusPlus::seed_from_u64: SplitMix64 fills the state
        for (int i = 0; i < 4; ++i) {
            seed += 0x9E3779B97F4A7C15ull;
            uint64_t z = seed;
            z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
            z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
            s[i] = z ^ (z >> 31);
        }
    }
    static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
    uint64_t next_u64() {
        const uint64_t out = rotl(s[0] + s[3], 23) + s[0];
        const uint64_t t = s[1] << 17;
        s[2] ^= s[0];
        s[3] ^= s[1];
        s[1] ^= s[2];
        s[0] ^= s[3];
        s[2] ^= t;
        s[3] = rotl(s[3], 45);
        return out;
    }
    // UniformFloat<f64>::sample_single(low, high): [1, 2) from the top 52 bits, minus one, times scale plus low
    double range_f64(double low, double high) {
        double scale = high - low;
        for (;;) {
            const uint64_t bits = (next_u64() >> 12) | 0x3FF0000000000000ull;
            double v12;
            memcpy(&v12, &bits, 8);
            const double v01 = v12 - 1.0;
            const double res = v01 * scale + low;
            if (res < high) return res;
            scale = std::nextafter(scale, 0.0);  // decrease_masked: one ulp down
        }
    }
};

// ---- Profiles (dc_poisson.rs:128-213) -------------------------------------------------------------
struct Profiles {
    std::vector<std::vector<std::pair<uint32_t, float>>> rows;
    std::vector<float> size_factor;
    size_t num_features = 0;
};

// from a dense entity x feature matrix (the product's gene_sums): entries > 0, ascending feature
Profiles from_gene_sums(const float* P, uint32_t E, uint64_t M) {
    Profiles p;
    p.num_features = M;
    p.rows.resize(E);
    p.size_factor.resize(E);
    for (uint32_t e = 0; e < E; ++e) {
        float sf = 0.0f;
        for (uint64_t g = 0; g < M; ++g) {
            const float v = P[(size_t)e * M + g];
            if (v > 0.0f) {
                p.rows[e].push_back({(uint32_t)g, v});
                sf += v;
            }
        }
        p.size_factor[e] = sf;
    }
    return p;
}

void weight_by_vec(Profiles& p, const float* w) {
    for (size_t e = 0; e < p.rows.size(); ++e) {
        float sf = 0.0f;
        for (auto& gv : p.rows[e]) {
            gv.second *= w[gv.first];
            sf += gv.second;
        }
        p.size_factor[e] = sf;
    }
}

// nb_dispersion.rs:58-151
struct Trend {
    float a = -std::numeric_limits<float>::infinity(), b = 0.0f;
    float phi_at(float mu) const {
        if (!std::isfinite(mu) || mu <= 0.0f) return 0.0f;
        const float log_phi = a + b * std::log(mu);
        const float phi = std::exp(log_phi);
        return std::min(std::max(phi, 0.0f), 100.0f);
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
    double w_sum = 0.0, xm = 0.0, ym = 0.0;
    for (double v : w) w_sum += v;
    for (size_t i = 0; i < x.size(); ++i) xm += x[i] * w[i];
    for (size_t i = 0; i < x.size(); ++i) ym += y[i] * w[i];
    xm /= w_sum;
    ym /= w_sum;
    double sxx = 0.0, sxy = 0.0;
    for (size_t i = 0; i < x.size(); ++i) {
        const double dx = x[i] - xm;
        sxx += w[i] * dx * dx;
        sxy += w[i] * dx * (y[i] - ym);
    }
    if (sxx <= 0.0) {
        t.a = (float)ym;
        t.b = 0.0f;
        return t;
    }
    const double b = sxy / sxx;
    t.a = (float)(ym - b * xm);
    t.b = (float)b;
    return t;
}

// dc_poisson.rs:230-295 (rows folded in ascending entity order)
std::vector<float> nb_fisher_weights(const Profiles& p) {
    const size_t M = p.num_features, E = p.rows.size();
    std::vector<float> s1(M, 0.0f), s2(M, 0.0f);
    for (const auto& row : p.rows)
        for (const auto& gv : row)
            if (std::isfinite(gv.second)) {
                s1[gv.first] += gv.second;
                s2[gv.first] += gv.second * gv.second;
            }
    const float n = (float)std::max<size_t>(E, 1);  // safe_denom(ncols_processed)
    std::vector<float> means(M), vars(M);
    for (size_t g = 0; g < M; ++g) {
        const float mu = s1[g] / n;
        means[g] = mu;
        vars[g] = s2[g] / n - mu * mu;
    }
    const Trend trend = fit_trend(means, vars);
    double total = 0.0;
    for (size_t g = 0; g < M; ++g) total += (double)s1[g];
    const float avg_s = E > 0 ? (float)(total / (double)E) : 1.0f;
    const float inv_total = total > 0.0 ? 1.0f / (float)total : 0.0f;
    std::vector<float> w(M);
    for (size_t g = 0; g < M; ++g) w[g] = 1.0f / (1.0f + s1[g] * inv_total * avg_s * trend.phi_at(means[g]));
    return w;
}

// ---- sufficient statistics (dc_poisson.rs:307-377) ----------------------------------------------
struct Stats {
    size_t k = 0, m = 0;
    std::vector<uint32_t> membership;
    std::vector<double> gene_sum, size_sum;
    std::vector<float> log_gene, log_size_offset;
    Stats(const Profiles& p, size_t k_, const uint32_t* mem) : k(k_), m(p.num_features) {
        membership.assign(mem, mem + p.rows.size());
        gene_sum.assign(k * m, 0.0);
        size_sum.assign(k, 0.0);
        for (size_t e = 0; e < p.rows.size(); ++e) {
            const size_t base = (size_t)mem[e] * m;
            for (const auto& gv : p.rows[e]) gene_sum[base + gv.first] += (double)gv.second;
            size_sum[mem[e]] += (double)p.size_factor[e];
        }
        log_gene.resize(k * m);
        for (size_t i = 0; i < k * m; ++i) log_gene[i] = (float)std::log(gene_sum[i] + LOG_EPS);
        const double m_eps = (double)m * LOG_EPS;
        log_size_offset.resize(k);
        for (size_t i = 0; i < k; ++i) log_size_offset[i] = (float)(-std::log(size_sum[i] + m_eps));
    }
    void delta_move(size_t e, size_t from, size_t to, const Profiles& p) {
        if (from == to) return;
        const double m_eps = (double)m * LOG_EPS;
        for (const auto& gv : p.rows[e]) {
            const size_t a = from * m + gv.first, b = to * m + gv.first;
            gene_sum[a] -= (double)gv.second;
            gene_sum[b] += (double)gv.second;
            log_gene[a] = (float)std::log(gene_sum[a] + LOG_EPS);
            log_gene[b] = (float)std::log(gene_sum[b] + LOG_EPS);
        }
        const double sf = (double)p.size_factor[e];
        size_sum[from] -= sf;
        size_sum[to] += sf;
        log_size_offset[from] = (float)(-std::log(size_sum[from] + m_eps));
        log_size_offset[to] = (float)(-std::log(size_sum[to] + m_eps));
        membership[e] = (uint32_t)to;
    }
};

double score(size_t e, size_t kk, const Stats& st, const Profiles& p) {  // dc_poisson.rs:424-429
    double acc = (double)p.size_factor[e] * (double)st.log_size_offset[kk];
    const size_t base = kk * st.m;
    for (const auto& gv : p.rows[e]) acc += (double)gv.second * (double)st.log_gene[base + gv.first];
    return acc;
}

// ---- label bookkeeping ------------------------------------------------------------------------------
template <typename Key>
size_t compact_labels(const std::vector<Key>& labels, std::vector<uint32_t>& out) {
    std::map<Key, uint32_t> seen;
    out.resize(labels.size());
    uint32_t next = 0;
    for (size_t i = 0; i < labels.size(); ++i) {
        auto it = seen.find(labels[i]);
        if (it == seen.end()) it = seen.emplace(labels[i], next++).first;
        out[i] = it->second;
    }
    return next;
}

// dc_poisson.rs:518-550: siblings[e] = sorted children of e's parent (all groups at the coarsest level)
std::vector<std::vector<uint32_t>> sibling_sets(const uint32_t* level, const uint32_t* parent, size_t E, size_t k) {
    std::vector<std::vector<uint32_t>> out(E);
    if (!parent) {
        std::vector<uint32_t> all(k);
        for (size_t i = 0; i < k; ++i) all[i] = (uint32_t)i;
        for (auto& s : out) s = all;
        return out;
    }
    std::unordered_map<uint32_t, std::vector<uint32_t>> kids;
    for (size_t e = 0; e < E; ++e) {
        auto& v = kids[parent[e]];
        if (std::find(v.begin(), v.end(), level[e]) == v.end()) v.push_back(level[e]);
    }
    for (auto& kv : kids) std::sort(kv.second.begin(), kv.second.end());
    for (size_t e = 0; e < E; ++e) out[e] = kids[parent[e]];
    return out;
}

// dc_poisson.rs:599-633
std::vector<uint32_t> intersect_fallback(const std::vector<uint32_t>& sib, const std::vector<uint32_t>& ngroups, uint32_t current) {
    if (sib.empty()) return {};
    if (sib.size() == 1) return sib;
    std::vector<uint32_t> inter;
    for (uint32_t g : sib)
        if (std::binary_search(ngroups.begin(), ngroups.end(), g)) inter.push_back(g);
    if (inter.empty()) return sib;
    if (std::find(inter.begin(), inter.end(), current) == inter.end()) {
        inter.push_back(current);
        std::sort(inter.begin(), inter.end());
    }
    return inter;
}

// refine_multilevel.rs:85-112
std::vector<std::vector<uint32_t>> candidate_sets(const std::vector<std::vector<uint32_t>>& sib, const uint32_t* bb_ptr, const uint32_t* bb,
                                                  const uint32_t* labels) {
    std::vector<std::vector<uint32_t>> out(sib.size());
    for (size_t e = 0; e < sib.size(); ++e) {
        std::vector<uint32_t> ng;
        for (uint32_t i = bb_ptr[e]; i < bb_ptr[e + 1]; ++i) ng.push_back(labels[bb[i]]);
        std::sort(ng.begin(), ng.end());
        ng.erase(std::unique(ng.begin(), ng.end()), ng.end());
        out[e] = intersect_fallback(sib[e], ng, labels[e]);
    }
    return out;
}

// dc_poisson.rs:778-915 with parallel = true and NoGuard
uint64_t refine_level(const Profiles& p, const std::vector<std::vector<uint32_t>>& cand, size_t k, int num_gibbs, int num_greedy,
                      uint64_t base_seed, double stagnation, uint32_t* labels) {
    const size_t E = p.rows.size();
    Stats st(p, k, labels);
    std::vector<uint32_t> prop(E);
    std::vector<double> lp;
    uint64_t total = 0;
    auto apply = [&]() {  // :661-686
        uint64_t moves = 0;
        for (size_t e = 0; e < E; ++e) {
            const uint32_t old = st.membership[e];
            if (prop[e] == old) continue;
            st.delta_move(e, old, prop[e], p);
            ++moves;
        }
        return moves;
    };
    int low = 0;
    for (int sweep = 0; sweep < num_gibbs; ++sweep) {
        const uint64_t sweep_seed = base_seed * (uint64_t)(sweep + 1);
        for (size_t e = 0; e < E; ++e) {
            if (cand[e].size() < 2) {
                prop[e] = st.membership[e];
                continue;
            }
            SmallRng rng(sweep_seed ^ ((uint64_t)e * 2654435761ull));
            double best = -std::numeric_limits<double>::infinity();
            uint32_t pick = cand[e][0];
            for (uint32_t c : cand[e]) {  // ascending labels == the order of the finite slots of log_probs (:451-471)
                const double l = score(e, c, st, p);
                if (!std::isfinite(l)) continue;
                const double u = rng.range_f64(1e-12, 1.0);
                const double key = l + (-std::log(-std::log(u)));
                if (key > best) {
                    best = key;
                    pick = c;
                }
            }
            prop[e] = pick;
        }
        const uint64_t moves = apply();
        total += moves;
        if (stagnation > 0.0) {
            if ((double)moves < stagnation * (double)E) {
                if (++low >= 3) break;
            } else {
                low = 0;
            }
        }
    }
    for (int sweep = 0; sweep < num_greedy; ++sweep) {
        for (size_t e = 0; e < E; ++e) {
            if (cand[e].size() < 2) {
                prop[e] = st.membership[e];
                continue;
            }
            uint32_t best = cand[e][0];
            double bv = score(e, best, st, p);
            for (size_t i = 1; i < cand[e].size(); ++i) {
                const double l = score(e, cand[e][i], st, p);
                if (l > bv) {
                    bv = l;
                    best = cand[e][i];
                }
            }
            prop[e] = best;
        }
        const uint64_t moves = apply();
        total += moves;
        if (moves == 0) break;
    }
    memcpy(labels, st.membership.data(), sizeof(uint32_t) * E);
    return total;
}

std::vector<std::vector<uint32_t>> to_sets(const uint32_t* ptr, const uint32_t* flat, size_t E) {
    std::vector<std::vector<uint32_t>> out(E);
    for (size_t e = 0; e < E; ++e) out[e].assign(flat + ptr[e], flat + ptr[e + 1]);
    return out;
}

}  // namespace

// ---- C interface (ctypes) ------------------------------------------------------------------------
extern "C" uint64_t orc_smallrng_u64(uint64_t seed, int skip) {
    SmallRng r(seed);
    for (int i = 0; i < skip; ++i) r.next_u64();
    return r.next_u64();
}
extern "C" double orc_smallrng_range_f64(uint64_t seed, int skip, double lo, double hi) {
    SmallRng r(seed);
    for (int i = 0; i < skip; ++i) r.range_f64(lo, hi);
    return r.range_f64(lo, hi);
}

extern "C" uint32_t orc_compact_labels(const uint64_t* labels, uint64_t n, uint32_t* out) {
    std::vector<uint64_t> v(labels, labels + n);
    std::vector<uint32_t> o;
    const size_t k = compact_labels(v, o);
    if (n) memcpy(out, o.data(), sizeof(uint32_t) * n);
    return (uint32_t)k;
}

// refine_multilevel.rs:315-320: dense labels of the (child, parent) pairs in first-appearance order
extern "C" uint32_t orc_project_to_refinement(const uint32_t* child, const uint32_t* parent, uint64_t n, uint32_t* out) {
    std::vector<std::pair<uint32_t, uint32_t>> v(n);
    for (uint64_t i = 0; i < n; ++i) v[i] = {child[i], parent[i]};
    std::vector<uint32_t> o;
    const size_t k = compact_labels(v, o);
    if (n) memcpy(out, o.data(), sizeof(uint32_t) * n);
    return (uint32_t)k;
}

// refine_multilevel.rs:333-345
extern "C" void orc_child_offset_within_parent(const uint32_t* child, const uint32_t* parent, uint64_t n, uint32_t* out) {
    std::unordered_map<uint32_t, std::unordered_map<uint32_t, uint32_t>> per;
    for (uint64_t i = 0; i < n; ++i) {
        auto& local = per[parent[i]];
        auto it = local.find(child[i]);
        if (it == local.end()) it = local.emplace(child[i], (uint32_t)local.size()).first;
        out[i] = it->second;
    }
}

// sibling sets / candidate sets as CSR (out_ptr E + 1 entries; returns the number of entries, written while they fit `cap`)
extern "C" uint64_t orc_sibling_sets(const uint32_t* level, const uint32_t* parent_or_null, uint64_t E, uint32_t k, uint32_t* out_ptr,
                                     uint32_t* out, uint64_t cap) {
    const auto s = sibling_sets(level, parent_or_null, E, k);
    uint64_t n = 0;
    for (uint64_t e = 0; e < E; ++e) {
        out_ptr[e] = (uint32_t)n;
        for (uint32_t g : s[e]) {
            if (n < cap) out[n] = g;
            ++n;
        }
    }
    out_ptr[E] = (uint32_t)n;
    return n;
}
extern "C" uint64_t orc_candidate_sets(const uint32_t* sib_ptr, const uint32_t* sib, const uint32_t* bb_ptr, const uint32_t* bb,
                                       const uint32_t* labels, uint64_t E, uint32_t* out_ptr, uint32_t* out, uint64_t cap) {
    const auto c = candidate_sets(to_sets(sib_ptr, sib, E), bb_ptr, bb, labels);
    uint64_t n = 0;
    for (uint64_t e = 0; e < E; ++e) {
        out_ptr[e] = (uint32_t)n;
        for (uint32_t g : c[e]) {
            if (n < cap) out[n] = g;
            ++n;
        }
    }
    out_ptr[E] = (uint32_t)n;
    return n;
}

extern "C" void orc_dcp_fisher_weights(const float* P, uint32_t E, uint64_t M, float* out_w) {
    const Profiles p = from_gene_sums(P, E, M);
    const auto w = nb_fisher_weights(p);
    memcpy(out_w, w.data(), sizeof(float) * M);
}

// weighted profile values (in place) and size factors of a dense entity x feature matrix
extern "C" void orc_dcp_profiles(float* P, uint32_t E, uint64_t M, const float* w_or_null, float* out_sf) {
    Profiles p = from_gene_sums(P, E, M);
    if (w_or_null) weight_by_vec(p, w_or_null);
    for (uint32_t e = 0; e < E; ++e) {
        for (const auto& gv : p.rows[e]) P[(size_t)e * M + gv.first] = gv.second;
        out_sf[e] = p.size_factor[e];
    }
}

// sufficient statistics after a list of moves (entity, to) applied in order: for the delta == recompute property
extern "C" void orc_dcp_stats(const float* P, uint32_t E, uint64_t M, uint32_t k, const uint32_t* labels, const uint32_t* move_e,
                              const uint32_t* move_to, uint64_t nmoves, double* out_gene_sum, float* out_log_gene, double* out_size_sum,
                              float* out_log_size_offset, uint32_t* out_membership) {
    const Profiles p = from_gene_sums(P, E, M);
    Stats st(p, k, labels);
    for (uint64_t i = 0; i < nmoves; ++i) st.delta_move(move_e[i], st.membership[move_e[i]], move_to[i], p);
    memcpy(out_gene_sum, st.gene_sum.data(), sizeof(double) * st.gene_sum.size());
    memcpy(out_log_gene, st.log_gene.data(), sizeof(float) * st.log_gene.size());
    memcpy(out_size_sum, st.size_sum.data(), sizeof(double) * k);
    memcpy(out_log_size_offset, st.log_size_offset.data(), sizeof(float) * k);
    memcpy(out_membership, st.membership.data(), sizeof(uint32_t) * E);
}

// scores of entity e against every block (the unrestricted form of the reference's tests)
extern "C" void orc_dcp_scores(const float* P, uint32_t E, uint64_t M, uint32_t k, const uint32_t* labels, uint32_t e, double* out_k) {
    const Profiles p = from_gene_sums(P, E, M);
    const Stats st(p, k, labels);
    for (uint32_t c = 0; c < k; ++c) out_k[c] = score(e, c, st, p);
}

// one level: P are the (already weighted) profile values, candidates as CSR; labels in / out; returns the accepted moves
extern "C" uint64_t orc_dcp_refine_level(const float* P, uint32_t E, uint64_t M, const uint32_t* cand_ptr, const uint32_t* cand, uint32_t k,
                                         int num_gibbs, int num_greedy, uint64_t jacobi_base_seed, double stagnation, uint32_t* labels) {
    const Profiles p = from_gene_sums(P, E, M);  // size factors: the serial f32 fold of the stored values, as weight_by_vec leaves them
    return refine_level(p, to_sets(cand_ptr, cand, E), k, num_gibbs, num_greedy, jacobi_base_seed, stagnation, labels);
}

// refine_assignments (refine_multilevel.rs:170-298).  gene_sums: E x M dense; bbknn: CSR of matched pb-samples;
// initial / offsets: num_levels x E (finest first; offsets may be NULL -> child_offset_within_parent); fisher != 0 applies the
// NB Fisher-information weights.  out_levels: num_levels x E, out_k: num_levels.  Returns the total number of moves.
extern "C" uint64_t orc_refine_assignments(const float* gene_sums, uint32_t E, uint64_t M, const uint32_t* bb_ptr, const uint32_t* bb,
                                           int num_levels, const uint32_t* initial, const uint32_t* offsets, int num_gibbs, int num_greedy,
                                           int fisher, uint64_t seed, double stagnation, uint32_t* out_levels, uint32_t* out_k) {
    std::vector<std::vector<uint32_t>> refined(num_levels);
    std::vector<size_t> ks(num_levels);
    for (int l = 0; l < num_levels; ++l) {
        std::vector<uint64_t> v(E);
        for (uint32_t e = 0; e < E; ++e) v[e] = initial[(size_t)l * E + e];
        ks[l] = compact_labels(v, refined[l]);
    }
    uint64_t total = 0;
    if (num_gibbs != 0 || num_greedy != 0) {
        Profiles p = from_gene_sums(gene_sums, E, M);
        if (fisher) {
            const auto w = nb_fisher_weights(p);
            weight_by_vec(p, w.data());
        }
        SmallRng rng(seed);
        for (int level = num_levels - 1; level >= 0; --level) {
            if (level + 1 < num_levels) {
                std::vector<uint32_t> off(E);
                if (offsets) memcpy(off.data(), offsets + (size_t)level * E, sizeof(uint32_t) * E);
                else orc_child_offset_within_parent(initial + (size_t)level * E, initial + (size_t)(level + 1) * E, E, off.data());
                refined[level].resize(E);
                ks[level] = orc_project_to_refinement(off.data(), refined[level + 1].data(), E, refined[level].data());
            }
            const size_t k = ks[level];
            const auto sib = sibling_sets(refined[level].data(), level + 1 < num_levels ? refined[level + 1].data() : nullptr, E, k);
            const auto cand = candidate_sets(sib, bb_ptr, bb, refined[level].data());
            const uint64_t base_seed = rng.next_u64() | 1ull;
            total += refine_level(p, cand, k, num_gibbs, num_greedy, base_seed, stagnation, refined[level].data());
            std::vector<uint64_t> v(refined[level].begin(), refined[level].end());
            ks[level] = compact_labels(v, refined[level]);
        }
    }
    for (int l = 0; l < num_levels; ++l) {
        memcpy(out_levels + (size_t)l * E, refined[l].data(), sizeof(uint32_t) * E);
        out_k[l] = (uint32_t)ks[l];
    }
    return total;
}
