/*
 * oracle.cpp — CPU restatement of the legume-rs hot path.  TEST INFRASTRUCTURE ONLY
 * (see oracle.h).  Build: make -C oracle   (g++ -O2 -ffp-contract=off -fopenmp)
 *
 * -ffp-contract=off matters: Rust never contracts a*b+c into an FMA, so every
 * "mul then add" below must round twice, exactly as the reference does.
 *
 * All citations are relative to /root/reference (causalpathlab/legume-rs v0.3.2).
 */
#include "oracle.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>
#include <utility>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ============================================================================
 * Stage 1 — projection
 *   project_columns_visitor          data-beans-alg/src/random_projection.rs:169-199
 *   CscMatrix::normalize_columns     matrix-util/src/dmatrix_util.rs:766-784
 *   nalgebra axpy(a, x, 1.0)         y[i] = (a*x[i]) + y[i]   (two roundings)
 * ==========================================================================*/
extern "C" void orc_project_raw(const uint64_t* indptr, const uint64_t* indices, const float* data,
                                uint64_t ncols, const float* basis_kd, int K, float* proj_kn, int nthreads) {
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
    (void)nthreads;
#endif
#pragma omp parallel num_threads(nthreads > 0 ? nthreads : 1)
    {
        std::vector<float> x;
#pragma omp for schedule(dynamic, 64)
        for (int64_t j = 0; j < (int64_t)ncols; ++j) {
            const uint64_t lo = indptr[j], hi = indptr[j + 1];
            const uint64_t n = hi - lo;
            x.resize(n);
            // :181-183  x = ln_1p(y)
            for (uint64_t t = 0; t < n; ++t) x[t] = log1pf(data[lo + t]);
            // dmatrix_util.rs:770-778  denom = sqrt(sum x^2) (sequential), floor 1e-8, divide
            float denom = 0.0f;
            for (uint64_t t = 0; t < n; ++t) denom += x[t] * x[t];
            denom = std::max(std::sqrt(denom), (float)1e-8);
            for (uint64_t t = 0; t < n; ++t) x[t] /= denom;
            // :188-194  chunk[:, j] = x * basis_kd[:, i] + chunk[:, j], nnz in ascending row order
            float* out = proj_kn + (size_t)j * K;
            for (int k = 0; k < K; ++k) out[k] = 0.0f;
            for (uint64_t t = 0; t < n; ++t) {
                const float* b = basis_kd + (size_t)indices[lo + t] * K;
                const float v = x[t];
                for (int k = 0; k < K; ++k) {
                    float prod = v * b[k];
                    out[k] = prod + out[k];
                }
            }
        }
    }
}

/* nalgebra scale_columns_inplace on one contiguous column of length n
 * (matrix-util/src/dmatrix_util.rs:986-995; nalgebra mean = fold-sum / n,
 * variance = fold (x-mean)^2 / n) */
static void scale_column_f32(float* x, uint64_t n, uint64_t stride) {
    if (n == 0) return;
    const float nf = (float)(double)n;
    float s = 0.0f;
    for (uint64_t i = 0; i < n; ++i) s = s + x[i * stride];
    const float mu = s / nf;
    float v = 0.0f;
    for (uint64_t i = 0; i < n; ++i) {
        float d = x[i * stride] - mu;
        v = v + d * d;
    }
    v = v / nf;
    const float sig = std::sqrt(v);
    const float neg_mu = -mu;
    for (uint64_t i = 0; i < n; ++i) x[i * stride] += neg_mu;
    if (sig > 0.0f)
        for (uint64_t i = 0; i < n; ++i) x[i * stride] /= sig;
}

/* random_projection.rs:378-407 */
extern "C" void orc_project_finish(float* proj, int K, uint64_t ncols, const uint32_t* batch, uint32_t nbatch) {
    if (batch && nbatch > 0) {
        // :380-387 per batch, per dim: mean = left-fold over the batch's cells (ascending) / n
        std::vector<float> sum((size_t)nbatch * K, 0.0f);
        std::vector<uint64_t> cnt(nbatch, 0);
        for (uint64_t j = 0; j < ncols; ++j) {
            const uint32_t b = batch[j];
            cnt[b]++;
            float* sb = sum.data() + (size_t)b * K;
            const float* p = proj + (size_t)j * K;
            for (int k = 0; k < K; ++k) sb[k] = sb[k] + p[k];
        }
        for (uint32_t b = 0; b < nbatch; ++b) {
            if (!cnt[b]) continue;
            const float nf = (float)(double)cnt[b];
            for (int k = 0; k < K; ++k) sum[(size_t)b * K + k] = -(sum[(size_t)b * K + k] / nf);
        }
        for (uint64_t j = 0; j < ncols; ++j) {
            const float* nm = sum.data() + (size_t)batch[j] * K;
            float* p = proj + (size_t)j * K;
            for (int k = 0; k < K; ++k) p[k] += nm[k];
        }
    }
    // :399 per-cell standardise
    for (uint64_t j = 0; j < ncols; ++j) scale_column_f32(proj + (size_t)j * K, K, 1);
    // :401-407 global clamp decision
    float mx = -std::numeric_limits<float>::infinity(), mn = std::numeric_limits<float>::infinity();
    for (size_t e = 0; e < (size_t)ncols * K; ++e) {
        mx = std::max(mx, proj[e]);
        mn = std::min(mn, proj[e]);
    }
    if (mx > 4.0f || mn < -4.0f) {
        for (size_t e = 0; e < (size_t)ncols * K; ++e) proj[e] = std::min(std::max(proj[e], -4.0f), 4.0f);
        for (uint64_t j = 0; j < ncols; ++j) scale_column_f32(proj + (size_t)j * K, K, 1);
    }
}

/* ============================================================================
 * Stage 2 — binary codes
 *   binary_sort_columns   data-beans-alg/src/random_projection.rs:535-564
 *   _randomized_svd       matrix-util/src/dmatrix_rsvd.rs:134-180
 *   _subspace_iteration   matrix-util/src/dmatrix_rsvd.rs:85-132
 *
 * AS-WRITTEN SEMANTICS (interpretation A):
 *   `view_mut(..).lower_triangle().copy_from(..)` at dmatrix_rsvd.rs:110-112 and
 *   :121-125 writes into the owned temporary that nalgebra's
 *   `lower_triangle(&self) -> OMatrix` returns, so after every iteration
 *   ll = [I_r; 0] and qq = [I_r; 0].  Hence X*qq = X[:, 0..r] exactly,
 *   Qf = QR(X[:, 0..r]).q(), Q = Qf[:, 0..kk], B = Q^T X, then SVD(B).
 *
 * SVD(B) (nalgebra 0.34.2, source not on disk — PARITY UNPINNED): restated in
 * the algebraically equivalent Gram form so that the O(N) reductions have a
 * fixed, parallel-friendly order that the CUDA path mirrors bit for bit:
 *   G = B B^T (f64, per-1024-cell-block butterfly sums, blocks summed in order)
 *   G = U diag(s^2) U^T  (cyclic Jacobi, f64), singular values descending
 *   V[j,k] = (sum_i U[i,k] * B[i,j]) / s_k    (f32, sequential fma over i)
 * Sign convention (nalgebra's is unpinned by any reference test): the
 * largest-magnitude component of the left singular vector Q*u_k in R^K is made
 * positive (first index wins ties).
 * Standardisation of V's columns (random_projection.rs:549) only matters through
 * the sign of (V[j,k] - mean_k): sd > 0 never changes a sign, and with sd == 0
 * the reference leaves x - mean in place.  So bit k of code_j = [V[j,k] > mean_k],
 * with mean_k = f32(blocked-f64-sum / N).
 * ==========================================================================*/

/* Householder QR thin-Q, following nalgebra's householder::clear_column_unchecked /
 * reflection_axis_mut / QR::q() control flow (sign handling included); dot products
 * are plain sequential f32 folds (nalgebra's unrolled dot order is unpinned). */
extern "C" void orc_householder_q(const float* a_in, int K, int r, float* q) {
    const int dim = std::min(K, r);
    std::vector<float> a(a_in, a_in + (size_t)K * r);
    std::vector<float> diag(dim, 0.0f);
    auto A = [&](int i, int j) -> float& { return a[(size_t)j * K + i]; };
    for (int c = 0; c < dim; ++c) {
        // reflection_axis_mut on a[c.., c]
        float sq = 0.0f;
        for (int i = c; i < K; ++i) sq = sq + A(i, c) * A(i, c);
        const float norm = std::sqrt(sq);
        const float head = A(c, c);
        const float modulus = std::fabs(head);
        const float sgn = (head < 0.0f) ? -1.0f : 1.0f;  // to_exp(): sign of a real (0 -> +1)
        const float signed_norm = sgn * norm;
        const float factor = (sq + modulus * norm) * 2.0f;
        A(c, c) = head + signed_norm;
        bool not_zero = factor != 0.0f;
        if (not_zero) {
            const float fs = std::sqrt(factor);
            for (int i = c; i < K; ++i) A(i, c) /= fs;
            // normalize_mut again
            float n2 = 0.0f;
            for (int i = c; i < K; ++i) n2 = n2 + A(i, c) * A(i, c);
            const float nn = std::sqrt(n2);
            for (int i = c; i < K; ++i) A(i, c) /= nn;
            diag[c] = -signed_norm;
            // refl.reflect_with_sign(right, sign = signum(diag))
            const float sign = (diag[c] > 0.0f) ? 1.0f : ((diag[c] < 0.0f) ? -1.0f : 0.0f);
            const float m_two = sign * -2.0f;
            for (int j = c + 1; j < r; ++j) {
                float dot = 0.0f;
                for (int i = c; i < K; ++i) dot = dot + A(i, c) * A(i, j);
                const float f = dot * m_two;
                for (int i = c; i < K; ++i) A(i, j) = f * A(i, c) + sign * A(i, j);
            }
        } else {
            diag[c] = signed_norm;
        }
    }
    // QR::q(): res = I (K × dim); for i in (0..dim).rev(): reflect res[i.., i..] with sign signum(diag[i])
    for (int j = 0; j < dim; ++j)
        for (int i = 0; i < K; ++i) q[(size_t)j * K + i] = (i == j) ? 1.0f : 0.0f;
    for (int c = dim - 1; c >= 0; --c) {
        const float sign = (diag[c] > 0.0f) ? 1.0f : ((diag[c] < 0.0f) ? -1.0f : 0.0f);
        const float m_two = sign * -2.0f;
        for (int j = c; j < dim; ++j) {
            float dot = 0.0f;
            for (int i = c; i < K; ++i) dot = dot + A(i, c) * q[(size_t)j * K + i];
            const float f = dot * m_two;
            for (int i = c; i < K; ++i) q[(size_t)j * K + i] = f * A(i, c) + sign * q[(size_t)j * K + i];
        }
    }
    // columns beyond dim (r > K) do not exist in nalgebra's thin Q; zero them for safety
    for (int j = dim; j < r; ++j)
        for (int i = 0; i < K; ++i) q[(size_t)j * K + i] = 0.0f;
}

/* cyclic Jacobi eigen-decomposition, f64, fixed sweep order; eigenvalues sorted
 * descending (stable on ties), evecs column-major n×n (column k = k-th vector) */
extern "C" void orc_jacobi_eig(const double* g, int n, double* evals, double* evecs) {
    std::vector<double> a(g, g + (size_t)n * n), v((size_t)n * n, 0.0);
    for (int i = 0; i < n; ++i) v[(size_t)i * n + i] = 1.0;
    auto A = [&](int i, int j) -> double& { return a[(size_t)j * n + i]; };
    auto V = [&](int i, int j) -> double& { return v[(size_t)j * n + i]; };
    for (int sweep = 0; sweep < 64; ++sweep) {
        double off = 0.0, dg = 0.0;
        for (int p = 0; p < n; ++p) {
            dg += A(p, p) * A(p, p);
            for (int q = p + 1; q < n; ++q) off += A(p, q) * A(p, q);
        }
        if (off <= 1e-60 || off <= 1e-32 * dg) break;
        for (int p = 0; p < n - 1; ++p)
            for (int q = p + 1; q < n; ++q) {
                const double apq = A(p, q);
                if (apq == 0.0) continue;
                const double theta = (A(q, q) - A(p, p)) / (2.0 * apq);
                const double t = (theta >= 0.0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
                const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < n; ++k) {
                    const double akp = A(k, p), akq = A(k, q);
                    A(k, p) = c * akp - s * akq;
                    A(k, q) = s * akp + c * akq;
                }
                for (int k = 0; k < n; ++k) {
                    const double apk = A(p, k), aqk = A(q, k);
                    A(p, k) = c * apk - s * aqk;
                    A(q, k) = s * apk + c * aqk;
                }
                for (int k = 0; k < n; ++k) {
                    const double vkp = V(k, p), vkq = V(k, q);
                    V(k, p) = c * vkp - s * vkq;
                    V(k, q) = s * vkp + c * vkq;
                }
            }
    }
    std::vector<int> order(n);
    for (int i = 0; i < n; ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return A(x, x) > A(y, y); });
    for (int k = 0; k < n; ++k) {
        evals[k] = A(order[k], order[k]);
        for (int i = 0; i < n; ++i) evecs[(size_t)k * n + i] = V(i, order[k]);
    }
}

/* The fixed reduction tree the CUDA path uses: 32 values combined by an xor
 * butterfly with offsets 16,8,4,2,1 (IEEE add is commutative, so every lane
 * holds the same value after each level). */
static double butterfly32(double* v) {
    double t[32];
    for (int off = 16; off >= 1; off >>= 1) {
        for (int l = 0; l < 32; ++l) t[l] = v[l] + v[l ^ off];
        std::memcpy(v, t, sizeof(t));
    }
    return v[0];
}
/* sum of vals[0..1024) for one 1024-cell block: 32 warps of 32 lanes, warp
 * butterflies, then a butterfly over the 32 warp sums */
static double block_sum_1024(const double* vals) {
    double ws[32], w[32];
    for (int warp = 0; warp < 32; ++warp) {
        std::memcpy(w, vals + warp * 32, sizeof(w));
        ws[warp] = butterfly32(w);
    }
    return butterfly32(ws);
}

static const int ORC_BLOCK = 1024;

extern "C" int orc_binary_codes(const float* proj, int K, uint64_t N, int kk, uint64_t* codes, float* out_q,
                                float* out_u, float* out_sigma, float* out_mean) {
    if (kk <= 0 || kk > 31 || (uint64_t)kk > N || kk > K) return 1;
    // dmatrix_rsvd.rs:145-153: rank = min(K, N); if rank > kk { rank = kk; oversample = 5 }
    int rank = (int)std::min<uint64_t>((uint64_t)K, N);
    int oversample = 0;
    if (kk > 0 && rank > kk) {
        rank = kk;
        oversample = 5;
    }
    int r = rank + oversample;
    if ((uint64_t)r > N) r = (int)N;  // identity pad has only N rows: X*qq keeps min(r, N) columns
    // Qf = qr(X[:, 0..r]).q()  (K × min(K, r)); keep the first `rank` columns (:129-131, :157-159)
    std::vector<float> qf((size_t)K * r);
    orc_householder_q(proj, K, r, qf.data());
    rank = std::min(rank, std::min(K, r));
    const int m = rank;
    const float* Q = qf.data();  // first m columns
    if (out_q) std::memcpy(out_q, Q, sizeof(float) * (size_t)K * m);

    // B = Q^T X  (m × N), f32 sequential fma over the K axis (matrixmultiply's
    // fma micro-kernel accumulates along k in order for k <= kc)
    std::vector<float> B((size_t)m * N);
    for (uint64_t j = 0; j < N; ++j) {
        const float* x = proj + (size_t)j * K;
        for (int i = 0; i < m; ++i) {
            float acc = 0.0f;
            const float* qc = Q + (size_t)i * K;
            for (int k = 0; k < K; ++k) acc = fmaf(qc[k], x[k], acc);
            B[(size_t)j * m + i] = acc;
        }
    }
    // G = B B^T, blocked f64
    const uint64_t nblk = (N + ORC_BLOCK - 1) / ORC_BLOCK;
    std::vector<double> G((size_t)m * m, 0.0), vals(ORC_BLOCK);
    for (int a = 0; a < m; ++a)
        for (int b = a; b < m; ++b) {
            double tot = 0.0;
            for (uint64_t blk = 0; blk < nblk; ++blk) {
                for (int t = 0; t < ORC_BLOCK; ++t) {
                    uint64_t j = blk * ORC_BLOCK + t;
                    vals[t] = (j < N) ? (double)B[(size_t)j * m + a] * (double)B[(size_t)j * m + b] : 0.0;
                }
                tot = tot + block_sum_1024(vals.data());
            }
            G[(size_t)b * m + a] = tot;
            G[(size_t)a * m + b] = tot;
        }
    std::vector<double> ev(m), U((size_t)m * m);
    orc_jacobi_eig(G.data(), m, ev.data(), U.data());
    // sign convention on Q*u_k
    std::vector<float> Uf((size_t)m * m), sig(m);
    for (int k = 0; k < m; ++k) {
        double best = 0.0, bestv = 0.0;
        for (int d = 0; d < K; ++d) {
            double s = 0.0;
            for (int i = 0; i < m; ++i) s += (double)Q[(size_t)i * K + d] * U[(size_t)k * m + i];
            if (std::fabs(s) > best) {
                best = std::fabs(s);
                bestv = s;
            }
        }
        const double flip = (bestv < 0.0) ? -1.0 : 1.0;
        for (int i = 0; i < m; ++i) Uf[(size_t)k * m + i] = (float)(flip * U[(size_t)k * m + i]);
        sig[k] = (float)std::sqrt(std::max(ev[k], 0.0));
    }
    if (out_u) std::memcpy(out_u, Uf.data(), sizeof(float) * (size_t)m * m);
    if (out_sigma) std::memcpy(out_sigma, sig.data(), sizeof(float) * m);

    // V[j,k] and its blocked column means
    std::vector<float> V((size_t)m * N);
    for (uint64_t j = 0; j < N; ++j)
        for (int k = 0; k < m; ++k) {
            float acc = 0.0f;
            for (int i = 0; i < m; ++i) acc = fmaf(Uf[(size_t)k * m + i], B[(size_t)j * m + i], acc);
            V[(size_t)j * m + k] = (sig[k] > 0.0f) ? acc / sig[k] : 0.0f;
        }
    std::vector<float> mean(m);
    for (int k = 0; k < m; ++k) {
        double tot = 0.0;
        for (uint64_t blk = 0; blk < nblk; ++blk) {
            for (int t = 0; t < ORC_BLOCK; ++t) {
                uint64_t j = blk * ORC_BLOCK + t;
                vals[t] = (j < N) ? (double)V[(size_t)j * m + k] : 0.0;
            }
            tot = tot + block_sum_1024(vals.data());
        }
        mean[k] = (float)(tot / (double)N);
    }
    if (out_mean) std::memcpy(out_mean, mean.data(), sizeof(float) * m);
    // random_projection.rs:551-561
    for (uint64_t j = 0; j < N; ++j) {
        uint64_t c = 0;
        for (int k = 0; k < m; ++k)
            if (V[(size_t)j * m + k] > mean[k]) c |= (1ull << k);
        codes[j] = c;
    }
    return 0;
}

/* ============================================================================
 * Stage 3 — group ids   data-beans/src/sparse_io_vector/groups.rs:13-37
 *   groups sorted by key.to_string() (byte-wise), id = rank
 * ==========================================================================*/
static uint32_t assign_by_strings(const std::vector<std::string>& label, uint64_t n, uint32_t* out) {
    std::vector<std::string> keys(label);
    std::sort(keys.begin(), keys.end());
    keys.erase(std::unique(keys.begin(), keys.end()), keys.end());
    for (uint64_t j = 0; j < n; ++j)
        out[j] = (uint32_t)(std::lower_bound(keys.begin(), keys.end(), label[j]) - keys.begin());
    return (uint32_t)keys.size();
}
extern "C" uint32_t orc_assign_groups(const uint64_t* codes, uint64_t n, uint32_t* out) {
    std::vector<std::string> label(n);
    for (uint64_t j = 0; j < n; ++j) label[j] = std::to_string(codes[j]);
    return assign_by_strings(label, n, out);
}
/* refine.rs:21-35 pad_numeric_labels(k) then assign_groups */
extern "C" uint32_t orc_assign_groups_padded(const uint64_t* labels, uint64_t n, uint64_t k, uint32_t* out) {
    size_t width = 1;
    uint64_t m = std::max<uint64_t>(k, 1) - 1;
    while (m >= 10) {
        width++;
        m /= 10;
    }
    std::vector<std::string> label(n);
    for (uint64_t j = 0; j < n; ++j) {
        std::string s = std::to_string(labels[j]);
        if (s.size() < width) s = std::string(width - s.size(), '0') + s;
        label[j] = s;
    }
    return assign_by_strings(label, n, out);
}
/* refine.rs:718-734 */
extern "C" int orc_level_sort_dims(int finest, int num_levels, int* out) {
    if (num_levels <= 1) {
        out[0] = finest;
        return 1;
    }
    const int coarsest = std::min(7, finest);
    int n = 0;
    for (int level = 0; level < num_levels; ++level) {
        float t = (float)level / (float)(num_levels - 1);
        float dim = (float)finest - t * (float)(finest - coarsest);
        int d = (int)std::round(dim);  // f32::round: half away from zero
        if (n == 0 || out[n - 1] != d) out[n++] = d;
    }
    return n;
}

/* ============================================================================
 * Stage 4 — collapse   data-beans-alg/src/collapse_data/stats.rs:110-164
 *   groups visited in group order, cells ascending within a group, rows ascending
 * ==========================================================================*/
extern "C" void orc_collapse_basic(const uint64_t* indptr, const uint64_t* indices, const float* data,
                                   uint64_t D, uint64_t N, const uint32_t* grp, const float* mult, uint32_t S,
                                   float* sum_ds, float* size_s) {
    std::memset(sum_ds, 0, sizeof(float) * (size_t)D * S);
    std::memset(size_s, 0, sizeof(float) * S);
    // cells ascending within each group == one pass over cells in ascending order,
    // since every (gene, group) accumulator only ever sees its own group's cells.
    for (uint64_t j = 0; j < N; ++j) {
        const uint32_t s = grp[j];
        if (s >= S) continue;
        const float w = mult ? mult[j] : 1.0f;
        float* col = sum_ds + (size_t)s * D;
        for (uint64_t t = indptr[j]; t < indptr[j + 1]; ++t) col[indices[t]] += data[t] * w;
        size_s[s] += w;
    }
}
extern "C" void orc_collapse_batch(const uint64_t* indptr, const uint64_t* indices, const float* data,
                                   uint64_t D, uint64_t N, const uint32_t* grp, const uint32_t* bat,
                                   const float* mult, uint32_t S, uint32_t B, float* sum_db, float* n_bs) {
    std::memset(sum_db, 0, sizeof(float) * (size_t)D * B);
    std::memset(n_bs, 0, sizeof(float) * (size_t)B * S);
    // stats.rs:136-164 visits group by group; sum_db[g,b] therefore accumulates in
    // (group, cell) order, not plain cell order — mirror that.
    std::vector<std::vector<uint64_t>> cells(S);
    for (uint64_t j = 0; j < N; ++j)
        if (grp[j] < S) cells[grp[j]].push_back(j);
    for (uint32_t s = 0; s < S; ++s)
        for (uint64_t j : cells[s]) {
            const uint32_t b = bat[j];
            const float w = mult ? mult[j] : 1.0f;
            float* col = sum_db + (size_t)b * D;
            for (uint64_t t = indptr[j]; t < indptr[j + 1]; ++t) col[indices[t]] += data[t] * w;
            n_bs[(size_t)s * B + b] += w;
        }
}
extern "C" void orc_merge_stat(const float* fine, uint64_t D, uint32_t nfine, const uint32_t* f2c, uint32_t ncoarse,
                               float* coarse) {
    std::memset(coarse, 0, sizeof(float) * (size_t)D * ncoarse);
    for (uint32_t f = 0; f < nfine; ++f) {
        float* dst = coarse + (size_t)f2c[f] * D;
        const float* src = fine + (size_t)f * D;
        for (uint64_t g = 0; g < D; ++g) dst[g] += src[g];
    }
}

/* ============================================================================
 * Stage 5 — Poisson-Gamma posterior
 *   GammaMatrix           matrix-param/src/dmatrix_gamma.rs:41-123
 *   calibrate_with        matrix-param/src/traits.rs:61-77
 *   digamma / trigamma    crate `special` 0.13.1 (not on disk): Bernardo's AS 103
 *                         and Schneider's AS 121, evaluated in the argument's own
 *                         type (f32).  trigamma pinned by dmatrix_gamma_tests.rs:9-32;
 *                         digamma PARITY UNPINNED.
 * ==========================================================================*/
extern "C" float orc_digamma(float p) {
    const float C = 8.5f, S = 1e-5f, S3 = 8.333333333e-2f, S4 = 8.333333333e-3f, S5 = 3.968253968e-3f;
    const float EULER = 0.57721566490153286f;
    if (!(p > 0.0f)) return std::numeric_limits<float>::quiet_NaN();
    if (p <= S) return -EULER - 1.0f / p;
    float value = 0.0f, z = p;
    while (z < C) {
        value -= 1.0f / z;
        z += 1.0f;
    }
    float r = 1.0f / z;
    value += std::log(z) - 0.5f * r;
    r *= r;
    value -= r * (S3 - r * (S4 - r * S5));
    return value;
}
extern "C" float orc_trigamma(float x) {
    const float A = 1e-4f, Bc = 5.0f, B2 = 0.1666666667f, B4 = -0.03333333333f, B6 = 0.02380952381f,
                B8 = -0.03333333333f;
    if (!(x > 0.0f)) return std::numeric_limits<float>::quiet_NaN();
    if (x <= A) return 1.0f / (x * x);
    float value = 0.0f, z = x;
    while (z < Bc) {
        value += 1.0f / (z * z);
        z += 1.0f;
    }
    const float y = 1.0f / (z * z);
    value += 0.5f * y + (1.0f + y * (B2 + y * (B4 + y * (B6 + y * B8)))) / z;
    return value;
}

extern "C" void orc_gamma_calibrate(const float* num, const float* den, uint64_t n, float a0, float b0, int target,
                                    float* mean, float* sd, float* log_mean, float* log_sd) {
    for (uint64_t e = 0; e < n; ++e) {
        // update_stat: fill(a0) then += (dmatrix_gamma.rs:64-75)
        const float a = a0 + num[e], b = b0 + den[e];
        if (mean) mean[e] = a / b;
        if (target == 0 || target == 2)
            if (log_mean) log_mean[e] = orc_digamma(a) - std::log(b);
        if (target == 0) {
            if (sd) sd[e] = std::sqrt(a) / b;
            if (log_sd) log_sd[e] = std::sqrt(orc_trigamma(a));
        }
    }
}

/* stats.rs:351-368 (B <= 1 arm); size_ds (optional, D x S): add_effective_size with observability attached (:176-186) */
extern "C" void orc_optimize_single_obs(const float* sum_ds, const float* size_s, const float* size_ds, uint64_t D, uint32_t S,
                                        float a0, float b0, int target, float* mean, float* sd, float* log_mean, float* log_sd);
extern "C" void orc_optimize_single(const float* sum_ds, const float* size_s, uint64_t D, uint32_t S, float a0,
                                    float b0, int target, float* mean, float* sd, float* log_mean, float* log_sd) {
    orc_optimize_single_obs(sum_ds, size_s, nullptr, D, S, a0, b0, target, mean, sd, log_mean, log_sd);
}
extern "C" void orc_optimize_single_obs(const float* sum_ds, const float* size_s, const float* size_ds, uint64_t D, uint32_t S,
                                        float a0, float b0, int target, float* mean, float* sd, float* log_mean, float* log_sd) {
    std::vector<float> den(D);
    for (uint32_t s = 0; s < S; ++s) {
        // add_effective_size: denom (zeros) + size_s[s], or + size_ds[:, s]
        const size_t off = (size_t)s * D;
        for (uint64_t g = 0; g < D; ++g) den[g] = 0.0f + (size_ds ? size_ds[off + g] : size_s[s]);
        orc_gamma_calibrate(sum_ds + off, den.data(), D, a0, b0, target, mean ? mean + off : nullptr,
                            sd ? sd + off : nullptr, log_mean ? log_mean + off : nullptr,
                            log_sd ? log_sd + off : nullptr);
        if (target == 1 && mean)  // sparsify_mean_to_support (dmatrix_gamma.rs:247-257)
            for (uint64_t g = 0; g < D; ++g)
                if (sum_ds[off + g] == 0.0f) mean[off + g] = 0.0f;
    }
}

/* stats.rs:219-350 (B > 1 arm); size_ds (optional): per-(gene, sample) effective sizes (:176-204); mask_db (optional,
 * D x B): both sides of the delta ratio are multiplied by it (:299-322) */
extern "C" void orc_optimize_batched_obs(const float* obs, const float* imp, const float* res, const float* size_s,
                                         const float* size_ds, const float* obs_db, const float* n_bs, const float* mask_db,
                                         uint64_t D, uint32_t S, uint32_t B, float a0, float b0, int num_iter, int target,
                                         float* mu_obs, float* mu_adj, float* mu_res, float* gamma, float* delta,
                                         float* mu_adj_log_mean);
extern "C" void orc_optimize_batched(const float* obs, const float* imp, const float* res, const float* size_s,
                                     const float* obs_db, const float* n_bs, uint64_t D, uint32_t S, uint32_t B,
                                     float a0, float b0, int num_iter, int target, float* mu_obs, float* mu_adj,
                                     float* mu_res, float* gamma, float* delta, float* mu_adj_log_mean) {
    orc_optimize_batched_obs(obs, imp, res, size_s, nullptr, obs_db, n_bs, nullptr, D, S, B, a0, b0, num_iter, target, mu_obs,
                             mu_adj, mu_res, gamma, delta, mu_adj_log_mean);
}
extern "C" void orc_optimize_batched_obs(const float* obs, const float* imp, const float* res, const float* size_s,
                                         const float* size_ds, const float* obs_db, const float* n_bs, const float* mask_db,
                                         uint64_t D, uint32_t S, uint32_t B, float a0, float b0, int num_iter, int target,
                                         float* mu_obs, float* mu_adj, float* mu_res, float* gamma, float* delta,
                                         float* mu_adj_log_mean) {
    const size_t n = (size_t)D * S;
    // GammaMatrix::new starts estimated_mean at zero (dmatrix_gamma.rs:49-52) and a_stat/b_stat at (a0, b0)
    std::vector<float> m_res(n), m_gam(n, 0.0f), m_adj(n, 0.0f), a_adj(n, a0), b_adj(n, b0);
    auto sz = [&](size_t e) { return size_ds ? size_ds[e] : size_s[e / D]; };
    // :240-247 mu_resid = Gamma(a0 + residual, b0 + (0 + size))
    for (size_t e = 0; e < n; ++e) m_res[e] = (a0 + res[e]) / (b0 + (0.0f + sz(e)));
    for (int it = 0; it < num_iter; ++it) {
        for (size_t e = 0; e < n; ++e) {
            // :262-268 denom = (resid + gamma) * size ; mu_adj = (a0 + (obs + imp)) / (b0 + denom)
            float denom = (m_res[e] + m_gam[e]) * sz(e);
            a_adj[e] = a0 + (obs[e] + imp[e]);
            b_adj[e] = b0 + denom;
            m_adj[e] = a_adj[e] / b_adj[e];
            // :276-279 denom = mu * size ; gamma = (a0 + imp) / (b0 + denom)
            denom = m_adj[e] * sz(e);
            m_gam[e] = (a0 + imp[e]) / (b0 + denom);
        }
    }
    // :296-324 delta: denom_db = mu_adj (D×S) * n_bs^T (S×B), accumulated along s in order
    if (delta) {
        for (uint32_t b = 0; b < B; ++b)
            for (uint64_t g = 0; g < D; ++g) {
                float acc = 0.0f;
                for (uint32_t s = 0; s < S; ++s) acc = fmaf(m_adj[(size_t)s * D + g], n_bs[(size_t)s * B + b], acc);
                float num = obs_db[(size_t)b * D + g];
                if (mask_db) {  // denom_db.component_mul_assign(mask); num_db = observed_sum_db .* mask
                    acc = acc * mask_db[(size_t)b * D + g];
                    num = num * mask_db[(size_t)b * D + g];
                }
                delta[(size_t)b * D + g] = (a0 + num) / (b0 + acc);
            }
    }
    for (size_t e = 0; e < n; ++e) {
        if (mu_obs) mu_obs[e] = (a0 + obs[e]) / (b0 + (0.0f + sz(e)));  // :327-332
        if (mu_adj) mu_adj[e] = m_adj[e];
        if (mu_res) mu_res[e] = m_res[e];
        if (gamma) gamma[e] = m_gam[e];
        // :289 mu_adj.calibrate_with(out_target) on the a_stat/b_stat left by the last sweep
        if (mu_adj_log_mean && (target == 0 || target == 2)) mu_adj_log_mean[e] = orc_digamma(a_adj[e]) - std::log(b_adj[e]);
    }
    if (target == 1) {  // :339-344 sparsify_mean_to_support
        for (size_t e = 0; e < n; ++e) {
            if (mu_obs && obs[e] == 0.0f) mu_obs[e] = 0.0f;
            if (mu_adj && (obs[e] + imp[e]) == 0.0f) mu_adj[e] = 0.0f;
            if (gamma && imp[e] == 0.0f) gamma[e] = 0.0f;
            if (mu_res && res[e] == 0.0f) mu_res[e] = 0.0f;
        }
    }
}

/* ============================================================================
 * Stage 6 — exact kNN
 *   l2_sq_kernel   matrix-util/src/knn/metric.rs:19-45  (16 lanes, left-fold, tail)
 *   topk           matrix-util/src/knn/exact.rs:36-55   (rank by squared distance,
 *                  total_cmp; ties: the reference's unstable select leaves the order
 *                  of equal distances unspecified — the oracle fixes lower index first)
 *   search_indices matrix-util/src/knn/mod.rs:249-299   (exclude: fetch k+1, drop, truncate)
 * ==========================================================================*/
extern "C" float orc_l2_sq(const float* a, const float* b, int d) {
    float acc[16];
    for (int l = 0; l < 16; ++l) acc[l] = 0.0f;
    int c = 0;
    for (; c + 16 <= d; c += 16)
        for (int l = 0; l < 16; ++l) {
            float df = a[c + l] - b[c + l];
            acc[l] += df * df;
        }
    float sum = 0.0f;
    for (int l = 0; l < 16; ++l) sum += acc[l];
    for (; c < d; ++c) {
        float df = a[c] - b[c];
        sum += df * df;
    }
    return sum;
}

extern "C" void orc_knn_topk(const float* ref, uint64_t nr, const float* qry, uint64_t nq, int d, int k,
                             const uint32_t* exclude, uint32_t* out_idx, float* out_dist, int nthreads) {
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
    (void)nthreads;
#endif
#pragma omp parallel num_threads(nthreads > 0 ? nthreads : 1)
    {
        std::vector<std::pair<float, uint32_t>> scored(nr);
#pragma omp for schedule(dynamic, 16)
        for (int64_t q = 0; q < (int64_t)nq; ++q) {
            const float* qv = qry + (size_t)q * d;
            for (uint64_t i = 0; i < nr; ++i) scored[i] = {orc_l2_sq(ref + (size_t)i * d, qv, d), (uint32_t)i};
            const bool ex = exclude && exclude[q] != UINT32_MAX;
            const uint64_t fetch = std::min<uint64_t>(nr, (uint64_t)k + (ex ? 1 : 0));
            std::partial_sort(scored.begin(), scored.begin() + fetch, scored.end());
            uint32_t* oi = out_idx + (size_t)q * k;
            float* od = out_dist + (size_t)q * k;
            int w = 0;
            for (uint64_t t = 0; t < fetch && w < k; ++t) {
                if (ex && scored[t].second == exclude[q]) continue;
                oi[w] = scored[t].second;
                od[w] = std::sqrt(scored[t].first);
                ++w;
            }
            for (; w < k; ++w) {
                oi[w] = UINT32_MAX;
                od[w] = std::numeric_limits<float>::infinity();
            }
        }
    }
}

/* ============================================================================
 * Synthetic counts — data-beans-sim/src/core.rs:155-203 (`sample_poisson_triplets`)
 * restated: y_gj ~ Poisson(lambda_scale * delta[g,b(j)] * beta[g,k(j)]), keep y > 0.5.
 * The reference seeds StdRng per cell (core.rs:174); that generator is not on
 * disk, so a counter-based splitmix hash of (seed, cell, gene, piece) replaces it
 * — this makes the matrix reproducible shard by shard on CPU and GPU alike.
 * Poisson by CDF inversion in f32 with correctly-rounded mul/div/add only.
 * ==========================================================================*/
static inline uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static inline float u01(uint64_t cell_key, uint64_t g, uint32_t piece) {
    uint64_t h = mix64(cell_key ^ (g * 0xD1B54A32D192ED03ull + (uint64_t)piece * 0x8CB92BA72F3D8DD7ull));
    uint32_t m = (uint32_t)(h >> 40);  // 24 bits
    return ((float)m + 0.5f) * 5.9604644775390625e-08f;  // (m + 0.5) * 2^-24, exact in f32
}
static inline float poisson_inv(float lam, float p0, float u) {
    float y = 0.0f, p = p0, c = p0;
    while (u > c && y < 1024.0f) {
        y += 1.0f;
        p = (p * lam) / y;
        c = c + p;
    }
    return y;
}
extern "C" uint64_t orc_sim_poisson_csc(uint64_t seed, uint64_t D, uint64_t col_lo, uint64_t col_hi,
                                        const uint8_t* topic, const uint8_t* batch, uint32_t ntopic, uint32_t nbatch,
                                        const float* lam, const float* p0, const uint8_t* npiece, uint64_t* indptr,
                                        uint64_t* indices, float* data) {
    (void)ntopic;
    uint64_t nnz = 0;
    indptr[0] = 0;
    for (uint64_t j = col_lo; j < col_hi; ++j) {
        const uint64_t key = mix64(seed + j * 0x9E3779B97F4A7C15ull);
        const size_t base = ((size_t)topic[j - col_lo] * nbatch + batch[j - col_lo]) * D;
        for (uint64_t g = 0; g < D; ++g) {
            const size_t e = base + g;
            float y = 0.0f;
            for (uint32_t pc = 0; pc < npiece[e]; ++pc) y += poisson_inv(lam[e], p0[e], u01(key, g, pc));
            if (y > 0.5f) {
                if (indices) {
                    indices[nnz] = g;
                    data[nnz] = y;
                }
                ++nnz;
            }
        }
        indptr[j - col_lo + 1] = nnz;
    }
    return nnz;
}
