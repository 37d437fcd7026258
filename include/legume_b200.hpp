// legume_b200.hpp — C++ host-side mirror of the legume-rs operator interface for the hot path, over the C ABI
// of liblegume_b200.so (include/legume_b200.h).
//
// The reference's host language is Rust; this image has no cargo/rustc, so — besides the Rust binding written out in
// INTEGRATION.md — the host layer a compiled caller links against is this header: same names, argument meaning and
// error behaviour as the reference's traits, one class per reference type:
//
//   legume::DMatrix          nalgebra::DMatrix<f32>               column-major, (nrows, ncols)
//   legume::SparseIoVec      data_beans::sparse_io_vector::SparseIoVec with its derived caches
//                            + RandProjOps      data-beans-alg/src/random_projection.rs:43-162
//                            + CollapsingOps    data-beans-alg/src/collapse_data/mod.rs:315-361
//                            + MultilevelCollapsingOps (un-refined path)   collapse_data/mod.rs:867-1050
//   legume::GammaMatrix      matrix-param/src/dmatrix_gamma.rs (TwoStatParam + Inference)
//   legume::CollapsedStat / CollapsedOut / optimize      collapse_data/stats.rs:378-582
//   legume::ColumnDict       matrix-util/src/knn/mod.rs:62-299 (exact backend)
//
// Every failure is a legume::Error (the reference returns anyhow::Error); nothing here computes on the CPU — without
// a CUDA device Context's constructor throws.  Header-only; link with -llegume_b200.
#ifndef LEGUME_B200_HPP
#define LEGUME_B200_HPP
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <map>
#include <memory>
#include <optional>
#include <set>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "legume_b200.h"

namespace legume {

constexpr uint64_t DEFAULT_PROJECTION_SEED = 0x50524F4A50524F4Aull;  // random_projection.rs:41
constexpr size_t DEFAULT_KNN = 10;                                   // collapse_data/mod.rs:27
constexpr size_t DEFAULT_OPT_ITER = 100;                             // collapse_data/mod.rs:28
constexpr size_t DEFAULT_NUM_LEVELS = 2;                             // collapse_data/stats.rs:688

enum class CalibrateTarget : int { All = LG_TARGET_ALL, MeanOnly = LG_TARGET_MEAN_ONLY, MeanAndLogMean = LG_TARGET_MEAN_AND_LOG_MEAN };

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error("legume_b200 error " + std::to_string(c) + ": " + m), code(c) {}
};

// nalgebra::DMatrix<f32>: column-major
struct DMatrix {
    size_t nrows = 0, ncols = 0;
    std::vector<float> data;
    DMatrix() = default;
    DMatrix(size_t r, size_t c, float fill = 0.0f) : nrows(r), ncols(c), data(r * c, fill) {}
    float& operator()(size_t i, size_t j) { return data[j * nrows + i]; }
    float operator()(size_t i, size_t j) const { return data[j * nrows + i]; }
    const float* column(size_t j) const { return data.data() + j * nrows; }
};

class Context {
   public:
    explicit Context(int device = 0) {
        const int rc = lg_ctx_create(device, &h_);
        if (rc != LG_OK) throw Error(rc, "lg_ctx_create failed: no usable CUDA device (there is no CPU fallback)");
    }
    ~Context() { lg_ctx_destroy(h_); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    lg_ctx* get() const { return h_; }
    void check(int rc) const {
        if (rc != LG_OK) throw Error(rc, lg_last_error(h_));
    }
    uint64_t launch_count() const { return lg_ctx_launch_count(h_); }

   private:
    lg_ctx* h_ = nullptr;
};

// rank of label.to_string() in byte-wise order (batch.rs:274-275, groups.rs:20-24)
template <typename T>
inline std::pair<std::vector<uint32_t>, std::vector<std::string>> rank_labels(const std::vector<T>& labels) {
    std::vector<std::string> strs;
    strs.reserve(labels.size());
    for (const auto& x : labels) {
        if constexpr (std::is_convertible_v<T, std::string>) strs.push_back(std::string(x));
        else strs.push_back(std::to_string(x));
    }
    std::vector<std::string> keys = strs;
    std::sort(keys.begin(), keys.end());
    keys.erase(std::unique(keys.begin(), keys.end()), keys.end());
    std::vector<uint32_t> idx(strs.size());
    for (size_t i = 0; i < strs.size(); ++i) idx[i] = (uint32_t)(std::lower_bound(keys.begin(), keys.end(), strs[i]) - keys.begin());
    return {idx, keys};
}

// collapse_data/refine.rs:718-734 (f32 arithmetic, round half away from zero)
inline std::vector<size_t> compute_level_sort_dims(size_t finest, size_t num_levels) {
    if (num_levels <= 1) return {finest};
    const size_t coarsest = std::min<size_t>(7, finest);
    std::vector<size_t> dims;
    for (size_t level = 0; level < num_levels; ++level) {
        const float t = (float)level / (float)(num_levels - 1);
        const float dim = (float)finest - t * (float)(finest - coarsest);
        const size_t d = (size_t)std::round(dim);
        if (dims.empty() || dims.back() != d) dims.push_back(d);
    }
    return dims;
}

// random_projection.rs:535-564
inline std::vector<uint64_t> binary_sort_columns(const Context& ctx, const DMatrix& proj_kn, size_t kk) {
    std::vector<uint64_t> codes(proj_kn.ncols);
    ctx.check(lg_binary_codes(ctx.get(), proj_kn.data.data(), (int)proj_kn.nrows, proj_kn.ncols, (int)kk, codes.data()));
    return codes;
}

struct RandColProjOut {
    DMatrix basis;  // D x K
    DMatrix proj;   // K x N
};

// collapse_data/stats.rs:546-582
struct CollapsedStat {
    DMatrix observed_sum_ds, imputed_sum_ds, residual_sum_ds;  // D x S
    std::vector<float> size_s;                                 // S
    DMatrix observed_sum_db;                                   // D x B
    DMatrix n_bs;                                              // B x S
    // panel observability (stats.rs:556-567): per-(gene, sample) effective sizes and the (gene, batch) mask of delta;
    // nullopt = fully observed, the historical code path
    std::optional<DMatrix> size_ds, obs_mask_db;
    CollapsedStat(size_t ngene, size_t nsample, size_t nbatch)
        : observed_sum_ds(ngene, nsample), imputed_sum_ds(ngene, nsample), residual_sum_ds(ngene, nsample), size_s(nsample, 0.0f),
          observed_sum_db(ngene, nbatch), n_bs(nbatch, nsample) {}
    size_t num_genes() const { return observed_sum_ds.nrows; }
    size_t num_samples() const { return observed_sum_ds.ncols; }
    size_t num_batches() const { return observed_sum_db.ncols; }
};

// matrix-param/src/dmatrix_gamma.rs — the posterior planes a calibration fills
struct GammaPosterior {
    DMatrix mean, sd, log_mean, log_sd;
    const DMatrix& posterior_mean() const { return mean; }
    const DMatrix& posterior_sd() const { return sd; }
    const DMatrix& posterior_log_mean() const { return log_mean; }
    const DMatrix& posterior_log_sd() const { return log_sd; }
};

// matrix-param/src/dmatrix_gamma.rs:11-123 (TwoStatParam + Inference)
class GammaMatrix {
   public:
    GammaMatrix(const Context& ctx, size_t nrows, size_t ncols, float a0, float b0)
        : ctx_(ctx), a0_(a0), b0_(b0), a_stat_(nrows, ncols, a0), b_stat_(nrows, ncols, b0), post_{DMatrix(nrows, ncols), {}, {}, {}} {}
    void update_stat(const DMatrix& a, const DMatrix& b) {
        reset_stat();
        add_stat(a, b);
    }
    void add_stat(const DMatrix& a, const DMatrix& b) {
        if (a.data.size() != a_stat_.data.size() || b.data.size() != b_stat_.data.size()) throw Error(LG_ERR_INVALID, "shape mismatch");
        for (size_t i = 0; i < a.data.size(); ++i) {
            a_stat_.data[i] += a.data[i];
            b_stat_.data[i] += b.data[i];
        }
    }
    void reset_stat() {
        std::fill(a_stat_.data.begin(), a_stat_.data.end(), a0_);
        std::fill(b_stat_.data.begin(), b_stat_.data.end(), b0_);
    }
    void calibrate() { calibrate_with(CalibrateTarget::All); }
    void calibrate_with(CalibrateTarget target) {
        const size_t r = a_stat_.nrows, c = a_stat_.ncols;
        post_.mean = DMatrix(r, c);
        float *sd = nullptr, *lm = nullptr, *ls = nullptr;
        if (target == CalibrateTarget::All) {
            post_.sd = DMatrix(r, c);
            post_.log_sd = DMatrix(r, c);
            sd = post_.sd.data.data();
            ls = post_.log_sd.data.data();
        }
        if (target != CalibrateTarget::MeanOnly) {
            post_.log_mean = DMatrix(r, c);
            lm = post_.log_mean.data.data();
        }
        // a_stat / b_stat already carry the hyper-parameters (fill + add, dmatrix_gamma.rs:64-75)
        ctx_.check(lg_gamma_calibrate(ctx_.get(), a_stat_.data.data(), b_stat_.data.data(), a_stat_.data.size(), 0.0f, 0.0f, (int)target,
                                      post_.mean.data.data(), sd, lm, ls));
    }
    const DMatrix& posterior_mean() const { return post_.mean; }
    const DMatrix& posterior_sd() const { return post_.sd; }
    const DMatrix& posterior_log_mean() const { return post_.log_mean; }
    const DMatrix& posterior_log_sd() const { return post_.log_sd; }
    size_t nrows() const { return a_stat_.nrows; }
    size_t ncols() const { return a_stat_.ncols; }

   private:
    const Context& ctx_;
    float a0_, b0_;
    DMatrix a_stat_, b_stat_;
    GammaPosterior post_;
};

// collapse_data/stats.rs:516-522
struct CollapsedOut {
    GammaPosterior mu_observed;
    std::optional<GammaPosterior> mu_adjusted, mu_residual, gamma, delta;
};

// collapse_data/stats.rs:378-512; the gene-block loop is numerically inert (:370-377) and disappears
inline CollapsedOut optimize(const Context& ctx, const CollapsedStat& stat, std::pair<float, float> hyper = {1.0f, 1.0f},
                             size_t num_iter = DEFAULT_OPT_ITER, CalibrateTarget target = CalibrateTarget::All) {
    const size_t D = stat.num_genes(), S = stat.num_samples(), B = stat.num_batches();
    CollapsedOut out;
    if (B <= 1) {
        GammaPosterior& p = out.mu_observed;
        p.mean = DMatrix(D, S);
        float *sd = nullptr, *lm = nullptr, *ls = nullptr;
        if (target == CalibrateTarget::All) {
            p.sd = DMatrix(D, S);
            p.log_sd = DMatrix(D, S);
            sd = p.sd.data.data();
            ls = p.log_sd.data.data();
        }
        if (target != CalibrateTarget::MeanOnly) {
            p.log_mean = DMatrix(D, S);
            lm = p.log_mean.data.data();
        }
        ctx.check(lg_optimize_single_obs(ctx.get(), stat.observed_sum_ds.data.data(), stat.size_s.data(),
                                         stat.size_ds ? stat.size_ds->data.data() : nullptr, D, (uint32_t)S, hyper.first, hyper.second,
                                         (int)target, p.mean.data.data(), sd, lm, ls));
        return out;
    }
    out.mu_observed.mean = DMatrix(D, S);
    out.mu_adjusted.emplace();
    out.mu_residual.emplace();
    out.gamma.emplace();
    out.delta.emplace();
    out.mu_adjusted->mean = DMatrix(D, S);
    out.mu_residual->mean = DMatrix(D, S);
    out.gamma->mean = DMatrix(D, S);
    out.delta->mean = DMatrix(D, B);
    float* lm = nullptr;
    if (target != CalibrateTarget::MeanOnly) {
        out.mu_adjusted->log_mean = DMatrix(D, S);
        lm = out.mu_adjusted->log_mean.data.data();
    }
    ctx.check(lg_optimize_batched_obs(ctx.get(), stat.observed_sum_ds.data.data(), stat.imputed_sum_ds.data.data(),
                                      stat.residual_sum_ds.data.data(), stat.size_s.data(), stat.size_ds ? stat.size_ds->data.data() : nullptr,
                                      stat.observed_sum_db.data.data(), stat.n_bs.data.data(),
                                      stat.obs_mask_db ? stat.obs_mask_db->data.data() : nullptr, D, (uint32_t)S, (uint32_t)B, hyper.first,
                                      hyper.second, (int)num_iter, (int)target, out.mu_observed.mean.data.data(),
                                      out.mu_adjusted->mean.data.data(), out.mu_residual->mean.data.data(), out.gamma->mean.data.data(),
                                      out.delta->mean.data.data(), lm));
    return out;
}

// dc_poisson.rs:71-117 RefineParams::default().  Served: Jacobi sweeps (parallel = true), raw profiles, the two feature weightings.
struct RefineParams {
    size_t num_gibbs = 20, num_greedy = 10;
    bool fisher_info_nb = true;  // FeatureWeighting::FisherInfoNb (false: FeatureWeighting::None)
    uint64_t seed = 42;
    double gibbs_stagnation = 0.005;
};

// collapse_data/mod.rs:64-130.  MultilevelParams::new sets refine = Some(default): with one batch refine_or_identity keeps the
// compacted hash partition (refine.rs:126-147), with two or more the BBKNN + DC-Poisson refinement runs (refine_assignments
// below).  refine = false is the legacy un-refined descent.
struct MultilevelParams {
    size_t knn_pb_samples = DEFAULT_KNN, num_levels = DEFAULT_NUM_LEVELS, sort_dim = 12, num_opt_iter = DEFAULT_OPT_ITER;
    bool refine = true;
    RefineParams refine_params;
    CalibrateTarget output_calibration = CalibrateTarget::All;
    explicit MultilevelParams(size_t proj_dim) : sort_dim(std::min<size_t>(proj_dim, 12)) {}
};

// MultilevelCollapseOut (collapse_data/mod.rs): the levels finest-first and every level's cell -> pb map
struct MultilevelCollapseOut {
    std::vector<CollapsedOut> levels;
    std::vector<std::vector<uint32_t>> cell_to_pb_per_level;
    std::vector<CollapsedStat> stats;
};

// dc_poisson.rs:493-509: labels -> 0..k in order of first appearance
inline std::pair<std::vector<uint32_t>, uint32_t> compact_labels(const std::vector<uint64_t>& labels) {
    std::map<uint64_t, uint32_t> lut;
    std::vector<uint32_t> out;
    out.reserve(labels.size());
    for (uint64_t g : labels) out.push_back(lut.emplace(g, (uint32_t)lut.size()).first->second);
    return {out, (uint32_t)lut.size()};
}
// refine.rs:43-62: the coarse label of the first pb-sample of every fine group
inline std::vector<uint32_t> fine_to_coarse_from_refined(const std::vector<uint32_t>& p2f, const std::vector<uint32_t>& p2c,
                                                         uint32_t num_fine) {
    std::vector<uint32_t> m(num_fine, 0xFFFFFFFFu);
    for (size_t p = 0; p < p2f.size(); ++p)
        if (m[p2f[p]] == 0xFFFFFFFFu) m[p2f[p]] = p2c[p];
    return m;
}

// ---- BBKNN + DC-Poisson refinement of the pb-sample partition (refine_multilevel.rs, dc_poisson.rs) ----------------------
// rand 0.10 SmallRng on 64-bit targets (xoshiro256++ seeded through SplitMix64): the refinement draws one u64 per level
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
    uint64_t next_u64() {
        auto rotl = [](uint64_t x, int k) { return (x << k) | (x >> (64 - k)); };
        const uint64_t out = rotl(s[0] + s[3], 23) + s[0], t = s[1] << 17;
        s[2] ^= s[0];
        s[3] ^= s[1];
        s[1] ^= s[2];
        s[0] ^= s[3];
        s[2] ^= t;
        s[3] = rotl(s[3], 45);
        return out;
    }
};
// refine_multilevel.rs:315-320: dense labels of the (child, parent) pairs in order of first appearance
inline std::pair<std::vector<uint32_t>, uint32_t> project_to_refinement(const std::vector<uint32_t>& child, const std::vector<uint32_t>& parent) {
    std::map<std::pair<uint32_t, uint32_t>, uint32_t> lut;
    std::vector<uint32_t> out;
    out.reserve(child.size());
    for (size_t i = 0; i < child.size(); ++i) out.push_back(lut.emplace(std::make_pair(child[i], parent[i]), (uint32_t)lut.size()).first->second);
    return {out, (uint32_t)lut.size()};
}
// refine_multilevel.rs:333-345
inline std::vector<uint32_t> child_offset_within_parent(const std::vector<uint32_t>& child, const std::vector<uint32_t>& parent) {
    std::map<uint32_t, std::map<uint32_t, uint32_t>> per;
    std::vector<uint32_t> out(child.size());
    for (size_t i = 0; i < child.size(); ++i) {
        auto& local = per[parent[i]];
        out[i] = local.emplace(child[i], (uint32_t)local.size()).first->second;
    }
    return out;
}
// dc_poisson.rs:518-550
inline std::vector<std::vector<uint32_t>> compute_sibling_sets(const std::vector<std::vector<uint32_t>>& refined, size_t level, uint32_t k) {
    const size_t n = refined[level].size();
    std::vector<std::vector<uint32_t>> out(n);
    if (level + 1 >= refined.size()) {
        std::vector<uint32_t> all(k);
        for (uint32_t i = 0; i < k; ++i) all[i] = i;
        for (auto& v : out) v = all;
        return out;
    }
    std::map<uint32_t, std::set<uint32_t>> kids;
    for (size_t e = 0; e < n; ++e) kids[refined[level + 1][e]].insert(refined[level][e]);
    for (size_t e = 0; e < n; ++e) {
        const auto& c = kids[refined[level + 1][e]];
        out[e].assign(c.begin(), c.end());
    }
    return out;
}
// dc_poisson.rs:599-633
inline std::vector<uint32_t> intersect_with_siblings_fallback(const std::vector<uint32_t>& siblings, const std::vector<uint32_t>& neighbor_groups,
                                                              uint32_t current) {
    if (siblings.size() <= 1) return siblings;
    std::vector<uint32_t> inter;
    for (uint32_t g : siblings)
        if (std::binary_search(neighbor_groups.begin(), neighbor_groups.end(), g)) inter.push_back(g);
    if (inter.empty()) return siblings;
    if (std::find(inter.begin(), inter.end(), current) == inter.end()) {
        inter.push_back(current);
        std::sort(inter.begin(), inter.end());
    }
    return inter;
}
// refine_multilevel.rs:85-112
inline std::vector<std::vector<uint32_t>> build_candidate_sets(const std::vector<std::vector<uint32_t>>& siblings,
                                                               const std::vector<std::vector<uint32_t>>& bbknn, const std::vector<uint32_t>& labels) {
    std::vector<std::vector<uint32_t>> out(siblings.size());
    for (size_t e = 0; e < siblings.size(); ++e) {
        std::vector<uint32_t> ng;
        for (uint32_t j : bbknn[e]) ng.push_back(labels[j]);
        std::sort(ng.begin(), ng.end());
        ng.erase(std::unique(ng.begin(), ng.end()), ng.end());
        out[e] = intersect_with_siblings_fallback(siblings[e], ng, labels[e]);
    }
    return out;
}
struct RefinedAssignment {  // refine_multilevel.rs:44-47
    std::vector<std::vector<uint32_t>> pbsamp_to_group;
    std::vector<uint32_t> num_groups_per_level;
    uint64_t moves = 0;
};
// refine_assignments (refine_multilevel.rs:170-298).  gene_sums: D x npb column-major (a pb-sample's genes contiguous);
// bbknn: the matched foreign pb-samples of every pb-sample; initial / offsets: per level (finest first) one entry per pb-sample,
// an empty offsets level falls back to child_offset_within_parent.
inline RefinedAssignment refine_assignments(const Context& ctx, const std::vector<float>& gene_sums, size_t D, uint32_t npb,
                                            const std::vector<std::vector<uint32_t>>& bbknn, const std::vector<std::vector<uint32_t>>& initial,
                                            const std::vector<std::vector<uint32_t>>& offsets, const RefineParams& params) {
    if (initial.empty()) throw Error(LG_ERR_INVALID, "no levels");
    const size_t L = initial.size();
    RefinedAssignment out;
    for (size_t l = 0; l < L; ++l) {
        if (initial[l].size() != npb)
            throw Error(LG_ERR_INVALID, "level " + std::to_string(l) + " has " + std::to_string(initial[l].size()) + " entries, expected " + std::to_string(npb));
        auto cl = compact_labels(std::vector<uint64_t>(initial[l].begin(), initial[l].end()));
        out.pbsamp_to_group.push_back(std::move(cl.first));
        out.num_groups_per_level.push_back(cl.second);
    }
    if (params.num_gibbs == 0 && params.num_greedy == 0) return out;  // :215-222
    std::vector<float> prof(gene_sums), w, sf(npb);
    if (params.fisher_info_nb) {
        w.resize(D);
        ctx.check(lg_dcp_fisher_weights(ctx.get(), prof.data(), D, npb, w.data()));
    }
    ctx.check(lg_dcp_profiles(ctx.get(), prof.data(), D, npb, w.empty() ? nullptr : w.data(), sf.data()));
    SmallRng rng(params.seed);
    auto& refined = out.pbsamp_to_group;
    for (size_t level = L; level-- > 0;) {
        if (level + 1 < L) {  // re-anchor in the REFINED parent by the child hash relative to its parent (:255-280)
            const std::vector<uint32_t> off = (level < offsets.size() && !offsets[level].empty()) ? offsets[level]
                                                                                                   : child_offset_within_parent(initial[level], initial[level + 1]);
            auto pr = project_to_refinement(off, refined[level + 1]);
            refined[level] = std::move(pr.first);
            out.num_groups_per_level[level] = pr.second;
        }
        const uint32_t k = out.num_groups_per_level[level];
        const auto cand = build_candidate_sets(compute_sibling_sets(refined, level, k), bbknn, refined[level]);
        std::vector<uint32_t> cptr(npb + 1, 0), cflat;
        for (uint32_t e = 0; e < npb; ++e) {
            cflat.insert(cflat.end(), cand[e].begin(), cand[e].end());
            cptr[e + 1] = (uint32_t)cflat.size();
        }
        const uint64_t base_seed = rng.next_u64() | 1ull;  // dc_poisson.rs:824
        uint64_t moves = 0;
        ctx.check(lg_dcp_refine_level(ctx.get(), prof.data(), sf.data(), D, npb, cptr.data(), cflat.data(), k, (int)params.num_gibbs,
                                      (int)params.num_greedy, base_seed, params.gibbs_stagnation, refined[level].data(), &moves));
        out.moves += moves;
        auto cl = compact_labels(std::vector<uint64_t>(refined[level].begin(), refined[level].end()));  // a sweep can empty a group (:292-295)
        refined[level] = std::move(cl.first);
        out.num_groups_per_level[level] = cl.second;
    }
    return out;
}

// data-beans/src/sparse_io_vector: one preloaded backend's columns on the device + the derived caches (mod.rs:70-85)
// matrix-util/src/sparse_stat.rs:33-198, 404-431 with T = f32.  The sufficient statistics are kept as f64 (exact whole
// numbers for count data: blocks and shards merge to the same totals in any order) and narrowed by the accessors.
class SparseRunningStatistics {
   public:
    explicit SparseRunningStatistics(size_t nrows) : npos_(nrows, 0.0), s1_(nrows, 0.0), s2_(nrows, 0.0) {}
    size_t nrows() const { return npos_.size(); }
    size_t ncols_processed() const { return ncols_; }
    // add_csc over a device-resident block (:97-108)
    void add_block(const Context& ctx, const lg_csc* block) {
        uint64_t D = 0, N = 0, nnz = 0;
        ctx.check(lg_csc_shape(block, &D, &N, &nnz));
        if (D != npos_.size()) throw Error(LG_ERR_INVALID, "SparseRunningStatistics: row count mismatch");
        std::vector<double> a(D), b(D), c(D);
        ctx.check(lg_row_stats(ctx.get(), block, a.data(), b.data(), c.data()));
        for (size_t g = 0; g < D; ++g) {
            npos_[g] += a[g];
            s1_[g] += b[g];
            s2_[g] += c[g];
        }
        ncols_ += N;
    }
    void merge(const SparseRunningStatistics& o) {  // :183-196
        for (size_t g = 0; g < npos_.size(); ++g) {
            npos_[g] += o.npos_[g];
            s1_[g] += o.s1_[g];
            s2_[g] += o.s2_[g];
        }
        ncols_ += o.ncols_;
    }
    std::vector<float> count_positives() const { return narrow(npos_); }
    std::vector<float> sum() const { return narrow(s1_); }
    std::vector<float> mean() const {
        std::vector<float> m = narrow(s1_);
        for (auto& x : m) x /= denom();
        return m;
    }
    std::vector<float> variance() const {  // s2 / n - mean^2 (:417-427)
        std::vector<float> v = narrow(s2_), m = mean();
        for (size_t g = 0; g < v.size(); ++g) v[g] = v[g] / denom() - m[g] * m[g];
        return v;
    }
    std::vector<float> std() const {
        std::vector<float> v = variance();
        for (auto& x : v) x = std::sqrt(x);
        return v;
    }

   private:
    float denom() const { return ncols_ > 0 ? (float)ncols_ : 1e-8f; }  // safe_denom (:16-23)
    static std::vector<float> narrow(const std::vector<double>& v) { return std::vector<float>(v.begin(), v.end()); }
    std::vector<double> npos_, s1_, s2_;
    size_t ncols_ = 0;
};

class SparseIoVec {
   public:
    // SparseIo::csc_column_arrays() -> (&[u64] indptr, &[u64] indices, &[f32] data)   (sparse_io/traits.rs:98-100)
    SparseIoVec(const Context& ctx, const std::vector<uint64_t>& indptr, const std::vector<uint64_t>& indices,
                const std::vector<float>& data, size_t nrows)
        : ctx_(ctx) {
        if (indptr.empty()) throw Error(LG_ERR_INVALID, "empty indptr");
        ctx_.check(lg_csc_upload(ctx_.get(), indptr.data(), indices.data(), data.data(), nrows, 0, indptr.size() - 1, nullptr, &csc_));
        ctx_.check(lg_csc_keep_pattern(ctx_.get(), csc_, 2));  // one projection, then one collapse per level: keep the pattern if memory allows
        uint64_t r, c, z;
        lg_csc_shape(csc_, &r, &c, &z);
        nrows_ = r;
        ncols_ = c;
    }
    // open_sparse_matrix(zarr) + SparseIoVec::push: the columns [col_lo, col_hi) of a store written by the reference's
    // zarr backend (data-beans/src/sparse_backend/zarr.rs), inflated on the host cores and fed through lg_csc_upload
    SparseIoVec(const Context& ctx, const std::string& zarr_file, uint64_t col_lo = 0, uint64_t col_hi = UINT64_MAX) : ctx_(ctx) {
        lg_zarr* z = nullptr;
        char err[512] = {0};
        int rc = lg_zarr_open(zarr_file.c_str(), &z, err, sizeof err);
        if (rc != LG_OK) throw Error(rc, err);
        uint64_t r = 0, c = 0, nz = 0;
        lg_zarr_shape(z, &r, &c, &nz);
        if (col_hi == UINT64_MAX) col_hi = c;
        rc = lg_zarr_read_columns(ctx_.get(), z, col_lo, col_hi, &csc_);
        lg_zarr_close(z);
        ctx_.check(rc);
        ctx_.check(lg_csc_keep_pattern(ctx_.get(), csc_, 2));
        lg_csc_shape(csc_, &r, &c, &nz);
        nrows_ = r;
        ncols_ = c;
    }
    ~SparseIoVec() { lg_csc_free(ctx_.get(), csc_); }
    SparseIoVec(const SparseIoVec&) = delete;
    SparseIoVec& operator=(const SparseIoVec&) = delete;
    size_t num_rows() const { return nrows_; }
    uint64_t refine_moves() const { return refine_moves_; }  // accepted DC-Poisson moves of the last multilevel collapse
    size_t num_columns() const { return ncols_; }
    const lg_csc* block() const { return csc_; }
    // lg_csc_keep_pattern: projections of this block leave their 1-bit pattern + list of counts != 1 behind, and the collapses
    // that follow (one per level of the multilevel scheme) sum those instead of streaming the arrays again; same sums
    void keep_pattern(bool on = true) { ctx_.check(lg_csc_keep_pattern(ctx_.get(), csc_, on ? 1 : 0)); }

    // ---- batch.rs:259-336 ----
    template <typename T>
    void register_batch_membership(const std::vector<T>& labels) {
        if (labels.size() != ncols_) throw Error(LG_ERR_INVALID, "batch membership length mismatches the number of columns");
        auto r = rank_labels(labels);
        col_to_batch_ = std::move(r.first);
        batch_names_ = std::move(r.second);
    }
    size_t num_batches() const { return batch_names_.size(); }
    const std::vector<uint32_t>& col_to_batch() const { return col_to_batch_; }
    void register_column_multiplicity(const std::vector<float>& w) {
        if (w.size() != ncols_) throw Error(LG_ERR_INVALID, "column multiplicity length mismatches the number of columns");
        for (float x : w)
            if (!(x > 0.0f)) throw Error(LG_ERR_INVALID, "column multiplicity must be strictly positive");
        multiplicity_ = w;
    }

    // ---- groups.rs:13-37 ----
    template <typename T>
    void assign_groups(const std::vector<T>& column_to_group) {
        if (column_to_group.size() != ncols_) throw Error(LG_ERR_INVALID, "group membership length mismatches the number of columns");
        auto r = rank_labels(column_to_group);
        col_to_group_ = std::move(r.first);
        num_groups_ = r.second.size();
    }
    size_t num_groups() const { return num_groups_; }
    const std::vector<uint32_t>& get_group_membership() const {
        if (col_to_group_.empty() && ncols_) throw Error(LG_ERR_INVALID, "groups were not assigned");
        return col_to_group_;
    }

    // ---- RandProjOps (random_projection.rs:341-527).  The basis is an input (identical-projection-matrix contract):
    //      basis_dk is D x K as the reference returns it; block_size is accepted and ignored. ----
    template <typename T>
    RandColProjOut project_columns_with_batch_correction(const DMatrix& basis_dk, std::optional<size_t> /*block_size*/,
                                                         const std::vector<T>* batch_membership) const {
        if (basis_dk.nrows != nrows_) throw Error(LG_ERR_INVALID, "basis must be D x K");
        const size_t K = basis_dk.ncols;
        DMatrix basis_kd(K, nrows_);  // basis_dk.transpose() (:360)
        for (size_t g = 0; g < nrows_; ++g)
            for (size_t k = 0; k < K; ++k) basis_kd(k, g) = basis_dk(g, k);
        std::vector<uint32_t> batch;
        uint32_t nb = 0;
        if (batch_membership && batch_membership->size() == ncols_) {  // else: warn and skip the centring (:389-395)
            auto r = rank_labels(*batch_membership);
            batch = std::move(r.first);
            nb = (uint32_t)r.second.size();
        }
        RandColProjOut out{basis_dk, DMatrix(K, ncols_)};
        ctx_.check(lg_project(ctx_.get(), csc_, basis_kd.data.data(), (int)K, nb ? batch.data() : nullptr, nb, out.proj.data.data()));
        return out;
    }
    RandColProjOut project_columns(const DMatrix& basis_dk, std::optional<size_t> block_size = std::nullopt) const {
        return project_columns_with_batch_correction<uint32_t>(basis_dk, block_size, nullptr);
    }
    template <typename T>
    RandColProjOut project_columns_weighted(DMatrix basis_dk, std::optional<size_t> block_size, const std::vector<T>* batch_membership,
                                            const std::vector<float>& row_weights) const {
        if (row_weights.size() != nrows_) throw Error(LG_ERR_INVALID, "row_weights length mismatch");
        for (size_t g = 0; g < nrows_; ++g) {  // random_projection.rs:438-444
            const float w = row_weights[g];
            for (size_t k = 0; k < basis_dk.ncols; ++k) {
                if (w <= 0.0f) basis_dk(g, k) = 0.0f;
                else if (std::fabs(w - 1.0f) > 1e-6f) basis_dk(g, k) *= w;
            }
        }
        return project_columns_with_batch_correction(basis_dk, block_size, batch_membership);
    }
    // random_projection.rs:506-527: returns max code + 1 and assigns the groups
    size_t partition_columns_to_groups(const DMatrix& proj_kn, std::optional<size_t> num_features = std::nullopt) {
        if (proj_kn.ncols != ncols_) throw Error(LG_ERR_INVALID, "number of columns mismatch");
        const size_t kk = std::min({proj_kn.nrows, num_features.value_or(proj_kn.nrows), ncols_});
        binary_codes_ = binary_sort_columns(ctx_, proj_kn, kk);
        sort_dim_ = kk;
        col_to_group_.assign(ncols_, 0);
        uint32_t ng = 0;
        ctx_.check(lg_assign_groups(ctx_.get(), binary_codes_.data(), ncols_, (int)kk, 0, col_to_group_.data(), &ng));
        num_groups_ = ng;
        return binary_codes_.empty() ? 0 : (size_t)*std::max_element(binary_codes_.begin(), binary_codes_.end()) + 1;
    }
    const std::vector<uint64_t>& binary_codes() const { return binary_codes_; }

    // ---- CollapsingOps (collapse_data/mod.rs:315-500) ----
    void collect_basic_stat(CollapsedStat& stat) const {
        ctx_.check(lg_collapse_basic(ctx_.get(), csc_, get_group_membership().data(), mult(), (uint32_t)stat.num_samples(),
                                     stat.observed_sum_ds.data.data(), stat.size_s.data()));
    }
    void collect_batch_stat(CollapsedStat& stat) const {
        if (col_to_batch_.empty()) throw Error(LG_ERR_INVALID, "batches were not registered");
        ctx_.check(lg_collapse_batch(ctx_.get(), csc_, get_group_membership().data(), col_to_batch_.data(), mult(),
                                     (uint32_t)stat.num_samples(), (uint32_t)stat.num_batches(), stat.observed_sum_db.data.data(),
                                     stat.n_bs.data.data()));
    }
    // ---- the nnz streams either side of the path ----
    // data-beans-alg/src/sparse_streaming.rs:23-60
    SparseRunningStatistics streaming_sparse_running_stats(std::optional<size_t> block_size = std::nullopt) const {
        (void)block_size;
        SparseRunningStatistics st(nrows_);
        st.add_block(ctx_, csc_);
        return st;
    }
    // nystrom_proj_visitor over every column (senna/src/svd/fit.rs:433-466); the pseudobulk of a cell is its group.
    // basis_dk: D x K; delta_dp: D x P or nullptr; returns K x N
    DMatrix nystrom_project(const DMatrix& basis_dk, const DMatrix* delta_dp = nullptr, float column_sum_norm = 1e4f) const {
        if (basis_dk.nrows != nrows_) throw Error(LG_ERR_INVALID, "nystrom_project: basis rows mismatch the number of genes");
        if (delta_dp && delta_dp->nrows != nrows_) throw Error(LG_ERR_INVALID, "nystrom_project: delta rows mismatch the number of genes");
        DMatrix out(basis_dk.ncols, ncols_);
        ctx_.check(lg_nystrom_project(ctx_.get(), csc_, basis_dk.data.data(), (int)basis_dk.ncols, delta_dp ? delta_dp->data.data() : nullptr,
                                      delta_dp ? get_group_membership().data() : nullptr, delta_dp ? (uint32_t)delta_dp->ncols : 0u,
                                      column_sum_norm, out.data.data()));
        return out;
    }
    // collapse_data/mod.rs:364-383 -> register_batches_dmatrix (batch.rs:46-234): the exact backend needs no index
    template <typename T>
    void build_hnsw_per_batch(const DMatrix& proj_kn, const std::vector<T>& batch_membership) {
        register_batch_membership(batch_membership);
        batch_proj_ = proj_kn;
        proximity_.clear();
        const uint32_t B = (uint32_t)num_batches();
        if (B > 2) {
            proximity_.resize((size_t)B * B);
            ctx_.check(lg_batch_proximity(ctx_.get(), proj_kn.data.data(), (int)proj_kn.nrows, ncols_, col_to_batch_.data(), B,
                                          proximity_.data(), nullptr));
        }
    }
    // collect_matched_stat_visitor over every group (stats.rs:26-108); knn_batches only sizes a Vec in the reference
    void collect_matched_stat(size_t /*knn_batches*/, size_t knn_cells, const std::vector<uint32_t>* reference_indices,
                              CollapsedStat& stat) const {
        if (batch_proj_.data.empty()) throw Error(LG_ERR_INVALID, "no knn lookup");
        const uint32_t B = (uint32_t)num_batches();
        std::vector<uint32_t> order;
        uint32_t nt = B;
        if (reference_indices) {
            nt = (uint32_t)reference_indices->size();
            for (uint32_t s = 0; s < B; ++s) order.insert(order.end(), reference_indices->begin(), reference_indices->end());
        } else if (!proximity_.empty()) {
            order = proximity_;
        }
        const size_t T = (size_t)nt * knn_cells;
        std::vector<uint32_t> midx(ncols_ * T);
        std::vector<float> mdist(ncols_ * T);
        ctx_.check(lg_knn_match_batches(ctx_.get(), batch_proj_.data.data(), (int)batch_proj_.nrows, ncols_, col_to_batch_.data(), B,
                                        (int)knn_cells, order.empty() ? nullptr : order.data(), nt, midx.data(), mdist.data()));
        ctx_.check(lg_collect_matched_stat(ctx_.get(), csc_, get_group_membership().data(), (uint32_t)stat.num_samples(), midx.data(),
                                           mdist.data(), (uint32_t)T, stat.imputed_sum_ds.data.data(), stat.residual_sum_ds.data.data()));
    }
    // collapse_data/mod.rs:384-475
    CollapsedOut collapse_columns(std::optional<size_t> knn_batches = std::nullopt, std::optional<size_t> knn_cells = std::nullopt,
                                  const std::vector<std::string>* reference_batch_names = nullptr,
                                  std::optional<size_t> num_opt_iter = std::nullopt, CollapsedStat* stat_out = nullptr) const {
        if (col_to_group_.empty()) throw Error(LG_ERR_INVALID, "The columns were not assigned before. Call `assign_columns_to_groups`");
        const size_t nb = num_batches();
        CollapsedStat stat(nrows_, num_groups_, nb);
        collect_basic_stat(stat);
        if (nb > 1) {
            std::vector<uint32_t> ref;
            if (reference_batch_names) {
                for (const auto& name : *reference_batch_names) {
                    auto it = std::find(batch_names_.begin(), batch_names_.end(), name);
                    if (it != batch_names_.end()) ref.push_back((uint32_t)(it - batch_names_.begin()));
                }
                if (ref.empty()) throw Error(LG_ERR_INVALID, "no reference batch names matched!");
            }
            collect_batch_stat(stat);
            collect_matched_stat(knn_batches.value_or(2), knn_cells.value_or(DEFAULT_KNN), reference_batch_names ? &ref : nullptr, stat);
        }
        CollapsedOut out = optimize(ctx_, stat, {1.0f, 1.0f}, num_opt_iter.value_or(DEFAULT_OPT_ITER), CalibrateTarget::All);
        if (stat_out) *stat_out = std::move(stat);
        return out;
    }

    // ---- MultilevelCollapsingOps::collapse_columns_multilevel_vec, un-refined path (collapse_data/mod.rs:867-1050);
    //      levels finest-first; per-level statistics are returned through stats_out when given ----
    template <typename T>
    std::vector<CollapsedOut> collapse_columns_multilevel_vec(const DMatrix& proj_kn, const std::vector<T>& batch_membership,
                                                              const MultilevelParams& params,
                                                              std::vector<CollapsedStat>* stats_out = nullptr) {
        register_batch_membership(batch_membership);
        const uint32_t nb = (uint32_t)num_batches();
        if (nb >= 2) build_hnsw_per_batch(proj_kn, batch_membership);
        const std::vector<size_t> level_dims = compute_level_sort_dims(params.sort_dim, params.num_levels);
        partition_columns_to_groups(proj_kn, level_dims[0]);
        const uint32_t ng = (uint32_t)num_groups_;
        if (params.refine) {  // mod.rs:914-941
            MultilevelCollapseOut out = refine_and_collect(proj_kn, level_dims, params, nullptr);
            if (stats_out) *stats_out = std::move(out.stats);
            return std::move(out.levels);
        }
        CollapsedStat fine(nrows_, ng, nb);
        collect_basic_stat(fine);
        if (nb >= 2) {
            collect_batch_stat(fine);
            const size_t cap = (size_t)ng * nb, K = proj_kn.nrows;
            std::vector<uint32_t> c2p(ncols_), pg(cap), pb(cap);
            std::vector<float> cnt(cap), cen(cap * K);
            uint32_t npb = 0;
            ctx_.check(lg_pb_layout(ctx_.get(), proj_kn.data.data(), (int)K, ncols_, col_to_group_.data(), ng, col_to_batch_.data(), nb,
                                    mult(), c2p.data(), pg.data(), pb.data(), cnt.data(), cen.data(), &npb));
            std::vector<float> gene_sums((size_t)nrows_ * npb), gsize(npb);
            ctx_.check(lg_collapse_basic(ctx_.get(), csc_, c2p.data(), mult(), npb, gene_sums.data(), gsize.data()));
            const uint32_t nslot = nb * (uint32_t)params.knn_pb_samples;
            std::vector<uint32_t> mp((size_t)npb * nslot);
            std::vector<float> md((size_t)npb * nslot);
            ctx_.check(lg_pb_match(ctx_.get(), proj_kn.data.data(), (int)K, ncols_, col_to_batch_.data(), nb, c2p.data(), cen.data(),
                                   pb.data(), npb, (int)params.knn_pb_samples, mp.data(), md.data()));
            ctx_.check(lg_collect_matched_stat_coarse(ctx_.get(), gene_sums.data(), nrows_, npb, cnt.data(), pg.data(), ng, mp.data(),
                                                      md.data(), nslot, fine.imputed_sum_ds.data.data(), fine.residual_sum_ds.data.data()));
        }
        std::vector<CollapsedOut> results;
        results.push_back(optimize(ctx_, fine, {1.0f, 1.0f}, params.num_opt_iter, CalibrateTarget::All));
        std::vector<CollapsedStat> stats;
        stats.push_back(std::move(fine));
        std::vector<uint32_t> prev_group = col_to_group_;
        uint32_t prev_n = ng;
        for (size_t level = 1; level < level_dims.size(); ++level) {
            std::vector<uint32_t> f2c(prev_n);
            uint32_t nc = 0;
            ctx_.check(lg_fine_to_coarse(ctx_.get(), binary_codes_.data(), prev_group.data(), ncols_, prev_n, (int)level_dims[level],
                                         f2c.data(), &nc));
            const CollapsedStat& prev = stats.back();
            CollapsedStat coarse(nrows_, nc, nb);
            ctx_.check(lg_merge_stat(ctx_.get(), prev.observed_sum_ds.data.data(), nrows_, prev_n, f2c.data(), nc, coarse.observed_sum_ds.data.data()));
            ctx_.check(lg_merge_stat(ctx_.get(), prev.imputed_sum_ds.data.data(), nrows_, prev_n, f2c.data(), nc, coarse.imputed_sum_ds.data.data()));
            ctx_.check(lg_merge_stat(ctx_.get(), prev.residual_sum_ds.data.data(), nrows_, prev_n, f2c.data(), nc, coarse.residual_sum_ds.data.data()));
            for (uint32_t f = 0; f < prev_n; ++f) {  // stats.rs:813-816
                coarse.size_s[f2c[f]] += prev.size_s[f];
                for (uint32_t b = 0; b < nb; ++b) coarse.n_bs(b, f2c[f]) += prev.n_bs(b, f);
            }
            coarse.observed_sum_db = prev.observed_sum_db;
            results.push_back(optimize(ctx_, coarse, {1.0f, 1.0f}, std::max<size_t>(params.num_opt_iter / 2, 10), CalibrateTarget::All));
            for (auto& g : prev_group) g = f2c[g];
            prev_n = nc;
            stats.push_back(std::move(coarse));
        }
        if (stats_out) *stats_out = std::move(stats);
        return results;
    }

    // ---- collapse_columns_multilevel_with_hierarchy (collapse_data/mod.rs:534-607) ----
    template <typename T>
    MultilevelCollapseOut collapse_columns_multilevel_with_hierarchy(const DMatrix& proj_kn, const std::vector<T>& batch_membership,
                                                                     const MultilevelParams& params) {
        if (!params.refine)
            throw Error(LG_ERR_INVALID, "collapse_columns_multilevel_with_hierarchy requires MultilevelParams.refine = Some(..); the "
                                        "legacy non-refinement path doesn't surface per-level cell->pb mappings");
        register_batch_membership(batch_membership);
        if (num_batches() >= 2) build_hnsw_per_batch(proj_kn, batch_membership);
        const std::vector<size_t> level_dims = compute_level_sort_dims(params.sort_dim, params.num_levels);
        partition_columns_to_groups(proj_kn, level_dims[0]);
        return refine_and_collect(proj_kn, level_dims, params, nullptr);
    }
    // ---- collapse_columns_multilevel_with_partition (collapse_data/mod.rs:617-821): every level's pb-sample -> group by
    //      majority vote of an inherited cell -> pb map (ties: the smallest label; the reference leaves them to its hash map) ----
    template <typename T>
    MultilevelCollapseOut collapse_columns_multilevel_with_partition(const DMatrix& proj_kn, const std::vector<T>& batch_membership,
                                                                     const MultilevelParams& params,
                                                                     const std::vector<std::vector<uint32_t>>& cell_to_pb_per_level) {
        register_batch_membership(batch_membership);
        if (num_batches() >= 2) build_hnsw_per_batch(proj_kn, batch_membership);
        const std::vector<size_t> level_dims = compute_level_sort_dims(params.sort_dim, params.num_levels);
        if (cell_to_pb_per_level.size() != level_dims.size())
            throw Error(LG_ERR_INVALID, "inherited cell_to_pb has " + std::to_string(cell_to_pb_per_level.size()) +
                                            " levels but --num-levels is " + std::to_string(level_dims.size()));
        for (const auto& lvl : cell_to_pb_per_level)
            if (lvl.size() != ncols_) throw Error(LG_ERR_INVALID, "inherited cell_to_pb level has the wrong number of cells");
        partition_columns_to_groups(proj_kn, level_dims[0]);
        MultilevelParams p = params;
        p.output_calibration = CalibrateTarget::All;
        return refine_and_collect(proj_kn, level_dims, p, &cell_to_pb_per_level);
    }

   private:
    // refine_and_collect_single_layer (refine.rs:264-500) and the tail of ..._with_partition (mod.rs:715-815): pb-samples
    // from the finest hash partition, every level's pb-sample -> group (hash-initialised and compacted, or inherited by
    // majority vote), finest statistics from one data pass, merge_stat descent along fine_to_coarse_from_refined
    MultilevelCollapseOut refine_and_collect(const DMatrix& proj_kn, const std::vector<size_t>& level_dims, const MultilevelParams& params,
                                             const std::vector<std::vector<uint32_t>>* inherited) {
        const uint32_t nb = (uint32_t)num_batches(), nbl = std::max<uint32_t>(nb, 1), ng = (uint32_t)num_groups_;
        const size_t cap = (size_t)ng * nbl, K = proj_kn.nrows;
        std::vector<uint32_t> c2p(ncols_), pg(cap), pb(cap), zero_batch;
        std::vector<float> cnt(cap), cen(cap * K);
        const uint32_t* bat = col_to_batch_.data();
        if (col_to_batch_.empty()) {
            zero_batch.assign(ncols_, 0u);
            bat = zero_batch.data();
        }
        uint32_t npb = 0;
        ctx_.check(lg_pb_layout(ctx_.get(), proj_kn.data.data(), (int)K, ncols_, col_to_group_.data(), ng, bat, nbl, mult(), c2p.data(),
                                pg.data(), pb.data(), cnt.data(), cen.data(), &npb));
        std::vector<float> gene_sums((size_t)nrows_ * npb), gsize(npb);
        ctx_.check(lg_collapse_basic(ctx_.get(), csc_, c2p.data(), mult(), npb, gene_sums.data(), gsize.data()));
        // every level's pb-sample -> group
        std::vector<std::vector<uint32_t>> p2g;
        std::vector<uint32_t> k_level;
        if (inherited) {
            for (const auto& lvl : *inherited) {
                std::vector<std::map<uint32_t, uint32_t>> votes(npb);
                for (size_t c = 0; c < ncols_; ++c)
                    if (c2p[c] != 0xFFFFFFFFu) votes[c2p[c]][lvl[c]]++;
                std::vector<uint64_t> modal(npb, 0);
                for (uint32_t p = 0; p < npb; ++p) {
                    uint32_t best = 0;
                    for (const auto& kv : votes[p])  // ascending labels: the first maximum is the smallest label
                        if (kv.second > best) {
                            best = kv.second;
                            modal[p] = kv.first;
                        }
                }
                auto cl = compact_labels(modal);
                p2g.push_back(std::move(cl.first));
                k_level.push_back(cl.second);
            }
        } else {
            std::vector<size_t> first(npb, ncols_);  // a pb-sample's first cell carries its finest code (refine.rs:68-88)
            for (size_t c = ncols_; c-- > 0;)
                if (c2p[c] != 0xFFFFFFFFu) first[c2p[c]] = c;
            for (size_t d : level_dims) {
                const uint64_t mask = d >= 64 ? ~0ull : ((1ull << d) - 1);
                std::vector<uint64_t> codes(npb);
                for (uint32_t p = 0; p < npb; ++p) codes[p] = binary_codes_[first[p]] & mask;
                auto cl = compact_labels(codes);
                p2g.push_back(std::move(cl.first));
                k_level.push_back(cl.second);
            }
            if (nb >= 2) {  // refine_or_identity(num_batches >= 2, ..) (refine.rs:329-345)
                const uint32_t nslot = nb * (uint32_t)params.knn_pb_samples;
                std::vector<uint32_t> mp((size_t)npb * nslot);
                std::vector<float> md((size_t)npb * nslot);
                ctx_.check(lg_pb_match(ctx_.get(), proj_kn.data.data(), (int)K, ncols_, col_to_batch_.data(), nb, c2p.data(), cen.data(), pb.data(),
                                       npb, (int)params.knn_pb_samples, mp.data(), md.data()));
                std::vector<std::vector<uint32_t>> bbknn(npb), offsets(level_dims.size());  // build_bbknn_neighbors (refine_multilevel.rs:60-83)
                for (uint32_t p = 0; p < npb; ++p)
                    for (uint32_t i = 0; i < nslot; ++i)
                        if (mp[(size_t)p * nslot + i] != 0xFFFFFFFFu) bbknn[p].push_back(mp[(size_t)p * nslot + i]);
                for (size_t level = 0; level + 1 < level_dims.size(); ++level) {  // build_reproject_offsets (refine.rs:95-125)
                    const size_t parent_dim = level_dims[level + 1], nbits = level_dims[level] > parent_dim ? level_dims[level] - parent_dim : 0;
                    const uint64_t mask = nbits >= 64 ? ~0ull : ((1ull << nbits) - 1);
                    offsets[level].resize(npb);
                    for (uint32_t p = 0; p < npb; ++p) offsets[level][p] = (uint32_t)((binary_codes_[first[p]] >> parent_dim) & mask);
                }
                RefinedAssignment ra = refine_assignments(ctx_, gene_sums, nrows_, npb, bbknn, p2g, offsets, params.refine_params);
                p2g = std::move(ra.pbsamp_to_group);
                k_level = std::move(ra.num_groups_per_level);
                refine_moves_ = ra.moves;
            }
        }
        // finest groups: pad_numeric_labels + assign_groups = the numeric id itself (refine.rs:21-35, 393-399)
        const uint32_t k0 = k_level[0];
        for (size_t c = 0; c < ncols_; ++c) col_to_group_[c] = p2g[0][c2p[c]];
        num_groups_ = k0;
        MultilevelCollapseOut out;
        CollapsedStat fine(nrows_, k0, nb);
        collect_basic_stat(fine);
        if (nb >= 2) {
            collect_batch_stat(fine);
            const uint32_t nslot = nb * (uint32_t)params.knn_pb_samples;
            std::vector<uint32_t> mp((size_t)npb * nslot);
            std::vector<float> md((size_t)npb * nslot);
            ctx_.check(lg_pb_match(ctx_.get(), proj_kn.data.data(), (int)K, ncols_, col_to_batch_.data(), nb, c2p.data(), cen.data(), pb.data(),
                                   npb, (int)params.knn_pb_samples, mp.data(), md.data()));
            ctx_.check(lg_collect_matched_stat_coarse(ctx_.get(), gene_sums.data(), nrows_, npb, cnt.data(), p2g[0].data(), k0, mp.data(),
                                                      md.data(), nslot, fine.imputed_sum_ds.data.data(), fine.residual_sum_ds.data.data()));
        }
        out.levels.push_back(optimize(ctx_, fine, {1.0f, 1.0f}, params.num_opt_iter, params.output_calibration));
        out.stats.push_back(std::move(fine));
        for (size_t level = 1; level < k_level.size(); ++level) {
            const uint32_t prev_n = k_level[level - 1], nc = k_level[level];
            const std::vector<uint32_t> f2c = fine_to_coarse_from_refined(p2g[level - 1], p2g[level], prev_n);
            const CollapsedStat& prev = out.stats.back();
            CollapsedStat coarse(nrows_, nc, nb);
            ctx_.check(lg_merge_stat(ctx_.get(), prev.observed_sum_ds.data.data(), nrows_, prev_n, f2c.data(), nc, coarse.observed_sum_ds.data.data()));
            ctx_.check(lg_merge_stat(ctx_.get(), prev.imputed_sum_ds.data.data(), nrows_, prev_n, f2c.data(), nc, coarse.imputed_sum_ds.data.data()));
            ctx_.check(lg_merge_stat(ctx_.get(), prev.residual_sum_ds.data.data(), nrows_, prev_n, f2c.data(), nc, coarse.residual_sum_ds.data.data()));
            for (uint32_t f = 0; f < prev_n; ++f) {
                coarse.size_s[f2c[f]] += prev.size_s[f];
                for (uint32_t b = 0; b < nb; ++b) coarse.n_bs(b, f2c[f]) += prev.n_bs(b, f);
            }
            coarse.observed_sum_db = prev.observed_sum_db;
            if (prev.size_ds) {
                coarse.size_ds.emplace(nrows_, nc);
                ctx_.check(lg_merge_stat(ctx_.get(), prev.size_ds->data.data(), nrows_, prev_n, f2c.data(), nc, coarse.size_ds->data.data()));
            }
            coarse.obs_mask_db = prev.obs_mask_db;
            out.levels.push_back(optimize(ctx_, coarse, {1.0f, 1.0f}, std::max<size_t>(params.num_opt_iter / 2, 10), params.output_calibration));
            out.stats.push_back(std::move(coarse));
        }
        for (const auto& lvl : p2g) {
            std::vector<uint32_t> c2g(ncols_);
            for (size_t c = 0; c < ncols_; ++c) c2g[c] = lvl[c2p[c]];
            out.cell_to_pb_per_level.push_back(std::move(c2g));
        }
        return out;
    }
    const float* mult() const { return multiplicity_.empty() ? nullptr : multiplicity_.data(); }
    uint64_t refine_moves_ = 0;
    const Context& ctx_;
    lg_csc* csc_ = nullptr;
    size_t nrows_ = 0, ncols_ = 0, num_groups_ = 0, sort_dim_ = 0;
    std::vector<uint32_t> col_to_group_, col_to_batch_, proximity_;
    std::vector<std::string> batch_names_;
    std::vector<float> multiplicity_;
    std::vector<uint64_t> binary_codes_;
    DMatrix batch_proj_;
};

// SplitMix64 avalanche of (base, salt): matrix-util/src/rand_util.rs:30-35
inline uint64_t mix_seed(uint64_t base, uint64_t salt) {
    uint64_t z = base ^ (salt * 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

// Several modalities over the same cells (data-beans/src/sparse_io_stack.rs): RandProjOps runs every modality with its
// own basis (the reference derives it from mix_seed(seed, m); here the bases are inputs) and stacks the results
// vertically (data-beans-alg/src/random_projection.rs:200-260)
class SparseIoStack {
   public:
    explicit SparseIoStack(std::vector<const SparseIoVec*> stack) : stack_(std::move(stack)) {
        if (stack_.empty()) throw Error(LG_ERR_INVALID, "SparseIoStack: empty stack");
    }
    size_t num_columns() const {
        size_t n = 0;
        for (auto* x : stack_) n = std::max(n, x->num_columns());
        return n;
    }
    size_t num_rows() const {
        size_t d = 0;
        for (auto* x : stack_) d += x->num_rows();
        return d;
    }
    template <typename T>
    RandColProjOut project_columns_with_batch_correction(const std::vector<DMatrix>& bases_dk, std::optional<size_t> block_size,
                                                         const std::vector<T>* batch_membership) const {
        if (bases_dk.size() != stack_.size()) throw Error(LG_ERR_INVALID, "SparseIoStack: one basis per modality");
        const size_t n = num_columns();
        std::vector<T> cut;
        if (batch_membership) cut.assign(batch_membership->begin(), batch_membership->begin() + std::min(n, batch_membership->size()));
        std::vector<RandColProjOut> parts;
        for (size_t m = 0; m < stack_.size(); ++m)
            parts.push_back(stack_[m]->project_columns_with_batch_correction(bases_dk[m], block_size, batch_membership ? &cut : nullptr));
        size_t Ktot = 0, Dtot = 0;
        const size_t K = parts[0].proj.nrows;
        for (auto& p : parts) {
            if (p.proj.ncols != n || p.proj.nrows != K) throw Error(LG_ERR_INVALID, "SparseIoStack: modalities disagree on the shape");
            Ktot += p.proj.nrows;
            Dtot += p.basis.nrows;
        }
        RandColProjOut out{DMatrix(Dtot, K), DMatrix(Ktot, n)};
        size_t r0 = 0, d0 = 0;
        for (auto& p : parts) {  // concatenate_vertical of the bases (rows) and of the projections (dims)
            for (size_t k = 0; k < K; ++k)
                for (size_t g = 0; g < p.basis.nrows; ++g) out.basis(d0 + g, k) = p.basis(g, k);
            for (size_t j = 0; j < n; ++j)
                for (size_t k = 0; k < K; ++k) out.proj(r0 + k, j) = p.proj(k, j);
            r0 += K;
            d0 += p.basis.nrows;
        }
        return out;
    }

   private:
    std::vector<const SparseIoVec*> stack_;
};

// matrix-util/src/knn/mod.rs:62-299 with the exact backend; names are the column indices 0..n-1 unless given
class ColumnDict {
   public:
    ColumnDict(const Context& ctx, DMatrix data_dn, std::vector<size_t> names = {}) : ctx_(ctx), data_(std::move(data_dn)), names_(std::move(names)) {
        if (names_.empty())
            for (size_t i = 0; i < data_.ncols; ++i) names_.push_back(i);
        if (names_.size() != data_.ncols) throw Error(LG_ERR_INVALID, "Data and names must have the same length");
        for (size_t i = 0; i < names_.size(); ++i) name2index_[names_[i]] = i;
    }
    size_t num_points() const { return data_.ncols; }
    size_t dim() const { return data_.nrows; }
    // search_by_query_data: (names, Euclidean distances), nearest first
    std::pair<std::vector<size_t>, std::vector<float>> search_by_query_data(const std::vector<float>& query, size_t knn) const {
        if (query.size() != data_.nrows) throw Error(LG_ERR_INVALID, "query's dim does not match");
        return search(query.data(), knn, nullptr, *this);
    }
    std::pair<std::vector<size_t>, std::vector<float>> search_others(size_t query_name, size_t knn) const {
        const uint32_t q = (uint32_t)index_of(query_name);
        return search(data_.column(q), knn, &q, *this);
    }
    std::pair<std::vector<size_t>, std::vector<float>> match_by_query_name_against(size_t query_name, size_t knn,
                                                                                   const ColumnDict& against) const {
        return search(data_.column(index_of(query_name)), knn, nullptr, against);
    }

   private:
    size_t index_of(size_t name) const {
        auto it = name2index_.find(name);
        if (it == name2index_.end()) throw Error(LG_ERR_INVALID, "name not found");
        return it->second;
    }
    std::pair<std::vector<size_t>, std::vector<float>> search(const float* q, size_t knn, const uint32_t* exclude,
                                                              const ColumnDict& against) const {
        std::pair<std::vector<size_t>, std::vector<float>> out;
        if (knn == 0 || against.data_.ncols == 0) return out;
        std::vector<uint32_t> idx(knn);
        std::vector<float> dist(knn);
        ctx_.check(lg_knn_topk(ctx_.get(), against.data_.data.data(), against.data_.ncols, q, 1, (int)data_.nrows, (int)knn, exclude,
                               idx.data(), dist.data()));
        for (size_t i = 0; i < knn; ++i)
            if (idx[i] != 0xFFFFFFFFu) {
                out.first.push_back(against.names_[idx[i]]);
                out.second.push_back(dist[i]);
            }
        return out;
    }
    const Context& ctx_;
    DMatrix data_;
    std::vector<size_t> names_;
    std::map<size_t, size_t> name2index_;
};

}  // namespace legume
#endif  // LEGUME_B200_HPP
