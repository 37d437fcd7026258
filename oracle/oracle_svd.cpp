/*
 * oracle_svd.cpp — the INDEPENDENT restatement of binary_sort_columns' factorisation.  TEST INFRASTRUCTURE ONLY.
 *
 * oracle.cpp::orc_binary_codes is the "mirror": it replaces nalgebra's SVD of the kk x N matrix B by the eigen-
 * decomposition of B B^T with the blocked f64 reductions the CUDA path uses, so that the GPU can be held to it bit for
 * bit.  That proves the kernels compute what the mirror says, not that the mirror says what the reference does.  This
 * file follows the reference's own route instead (matrix-util/src/dmatrix_rsvd.rs:129-171, as written — SURVEY
 * Appendix B, interpretation A):
 *     Q  = qr(X[:, 0..r]).q()[:, 0..kk]          (Householder QR, f32; orc_householder_q)
 *     B  = (X^T Q)^T                              (kk x N, f32 — :162)
 *     SVD(B) in f32 the way nalgebra 0.34's `svd(true, true)` goes about it: Householder bidiagonalisation followed
 *     by implicit-shift (Wilkinson) QR sweeps of Givens rotations on the bidiagonal (Golub-Kahan), singular values
 *     made non-negative and sorted descending (:164-171); V = v_t^T[:, 0..kk] (N x kk)
 *     scale_columns_inplace on V (f32 left folds per column, dmatrix_util.rs:986-995), bit k = [V[j,k] > 0]
 *
 * nalgebra's source is not on disk (PARITY UNPINNED): rounding order inside its reflectors and rotations cannot be
 * reproduced, and the sign of a singular-vector pair is a free choice.  So this oracle does not define bits; it
 * defines the PARTITION of the cells each bit makes, up to complement, and the tests report how many cells the mirror
 * (and the GPU) put on the other side: cells whose standardised V is within rounding noise of zero.
 */
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "oracle.h"

namespace {

// LAPACK-style elementary reflector for x = (alpha, x[1..m)): H = I - tau v v^T, v[0] = 1, H x = (beta, 0, ...)
float make_reflector(float* alpha, float* x, int m, size_t stride) {
    float xnorm2 = 0.0f;
    for (int i = 0; i < m; ++i) xnorm2 += x[i * stride] * x[i * stride];
    if (xnorm2 == 0.0f) return 0.0f;
    const float a = *alpha;
    float beta = std::sqrt(a * a + xnorm2);
    if (a > 0.0f) beta = -beta;
    const float tau = (beta - a) / beta;
    const float inv = 1.0f / (a - beta);
    for (int i = 0; i < m; ++i) x[i * stride] *= inv;
    *alpha = beta;
    return tau;
}

void givens(float f, float g, float* c, float* s) {
    if (g == 0.0f) {
        *c = 1.0f;
        *s = 0.0f;
    } else if (f == 0.0f) {
        *c = 0.0f;
        *s = 1.0f;
    } else {
        const float r = std::hypot(f, g);
        *c = f / r;
        *s = g / r;
    }
}

struct Dense {  // n x n, row-major, tiny
    int n;
    std::vector<float> a;
    explicit Dense(int n_) : n(n_), a((size_t)n_ * n_, 0.0f) {}
    float& operator()(int i, int j) { return a[(size_t)i * n + j]; }
    void identity() {
        std::fill(a.begin(), a.end(), 0.0f);
        for (int i = 0; i < n; ++i) (*this)(i, i) = 1.0f;
    }
    // columns (p, q) <- (c p + s q, -s p + c q)
    void rot_cols(int p, int q, float c, float s) {
        for (int i = 0; i < n; ++i) {
            const float x = (*this)(i, p), y = (*this)(i, q);
            (*this)(i, p) = c * x + s * y;
            (*this)(i, q) = -s * x + c * y;
        }
    }
    void rot_rows(int p, int q, float c, float s) {
        for (int j = 0; j < n; ++j) {
            const float x = (*this)(p, j), y = (*this)(q, j);
            (*this)(p, j) = c * x + s * y;
            (*this)(q, j) = -s * x + c * y;
        }
    }
};

// SVD of an upper-bidiagonal n x n matrix held densely: Bd = Ub diag(sig) Vb^T (Golub-Kahan implicit-shift QR)
bool bidiagonal_svd(Dense& Bd, Dense& Ub, Dense& Vb) {
    const int n = Bd.n;
    Ub.identity();
    Vb.identity();
    const float eps = 5.9604645e-8f;  // 2^-24
    for (int iter = 0; iter < 75 * n; ++iter) {
        for (int i = 0; i + 1 < n; ++i)
            if (std::fabs(Bd(i, i + 1)) <= eps * (std::fabs(Bd(i, i)) + std::fabs(Bd(i + 1, i + 1)))) Bd(i, i + 1) = 0.0f;
        int hi = n - 1;
        while (hi > 0 && Bd(hi - 1, hi) == 0.0f) --hi;
        if (hi == 0) return true;
        int lo = hi - 1;
        while (lo > 0 && Bd(lo - 1, lo) != 0.0f) --lo;
        // a zero on the diagonal of the unreduced block: rotate its row away (the block then splits)
        bool split = false;
        float scale = 0.0f;
        for (int i = lo; i <= hi; ++i) scale = std::max(scale, std::fabs(Bd(i, i)));
        for (int i = lo; i < hi && !split; ++i) {
            if (std::fabs(Bd(i, i)) <= eps * scale) {
                Bd(i, i) = 0.0f;
                for (int j = i + 1; j <= hi; ++j) {
                    float c, s;
                    givens(Bd(j, j), Bd(i, j), &c, &s);
                    // rows (j, i): zero Bd(i, j)
                    Bd.rot_rows(j, i, c, s);
                    Ub.rot_cols(j, i, c, s);
                    Bd(i, j) = 0.0f;
                }
                split = true;
            }
        }
        if (split) continue;
        // Wilkinson shift from the trailing 2 x 2 of B^T B
        const float dm = Bd(hi - 1, hi - 1), dn = Bd(hi, hi), fm = Bd(hi - 1, hi);
        const float fm1 = (hi - 2 >= lo) ? Bd(hi - 2, hi - 1) : 0.0f;
        const float t11 = dm * dm + fm1 * fm1, t12 = dm * fm, t22 = dn * dn + fm * fm;
        const float delta = 0.5f * (t11 - t22);
        float mu = t22;
        if (t12 != 0.0f) {
            const float den = delta + (delta >= 0.0f ? 1.0f : -1.0f) * std::sqrt(delta * delta + t12 * t12);
            if (den != 0.0f) mu = t22 - t12 * t12 / den;
        }
        float y = Bd(lo, lo) * Bd(lo, lo) - mu, z = Bd(lo, lo) * Bd(lo, lo + 1);
        for (int k = lo; k < hi; ++k) {
            float c, s;
            givens(y, z, &c, &s);
            Bd.rot_cols(k, k + 1, c, s);
            Vb.rot_cols(k, k + 1, c, s);
            if (k > lo) Bd(k - 1, k + 1) = 0.0f;
            y = Bd(k, k);
            z = Bd(k + 1, k);
            givens(y, z, &c, &s);
            Bd.rot_rows(k, k + 1, c, s);
            Ub.rot_cols(k, k + 1, c, s);
            Bd(k + 1, k) = 0.0f;
            if (k + 1 < hi) {
                y = Bd(k, k + 1);
                z = Bd(k, k + 2);
            }
        }
    }
    return false;
}

}  // namespace

/* returns 0 on success; out_v (may be NULL) receives the standardised N x kk factor, column-major (column k contiguous) */
extern "C" int orc_binary_codes_svd(const float* proj, int K, uint64_t N, int kk, uint64_t* codes, float* out_v, float* out_sigma) {
    if (kk <= 0 || kk > 31 || (uint64_t)kk > N || kk > K) return 1;
    int rank = (int)std::min<uint64_t>((uint64_t)K, N), oversample = 0;
    if (rank > kk) {
        rank = kk;
        oversample = 5;
    }
    int r = rank + oversample;
    if ((uint64_t)r > N) r = (int)N;
    std::vector<float> qf((size_t)K * r);
    orc_householder_q(proj, K, r, qf.data());
    const int n = std::min(rank, std::min(K, r));
    // A = B^T = X^T Q  (N x n, column-major): dmatrix_rsvd.rs:162 computes exactly this product before transposing
    std::vector<float> A((size_t)N * n);
    for (uint64_t j = 0; j < N; ++j) {
        const float* x = proj + (size_t)j * K;
        for (int i = 0; i < n; ++i) {
            const float* qc = qf.data() + (size_t)i * K;
            float acc = 0.0f;
            for (int k = 0; k < K; ++k) acc += x[k] * qc[k];
            A[(size_t)i * N + j] = acc;
        }
    }
    // Householder bidiagonalisation A = U1 Bd V1^T (upper bidiagonal: N >= n)
    std::vector<float> tauq(n, 0.0f), taup(n, 0.0f);
    auto a = [&](uint64_t i, int j) -> float& { return A[(size_t)j * N + i]; };
    for (int k = 0; k < n; ++k) {
        const int m = (int)(N - k - 1);
        tauq[k] = make_reflector(&a(k, k), m > 0 ? &a(k + 1, k) : nullptr, m, 1);
        for (int j = k + 1; j < n; ++j) {  // apply H_k to column j
            float w = a(k, j);
            for (uint64_t i = k + 1; i < N; ++i) w += a(i, k) * a(i, j);
            w *= tauq[k];
            a(k, j) -= w;
            for (uint64_t i = k + 1; i < N; ++i) a(i, j) -= w * a(i, k);
        }
        if (k + 2 < n) {  // zero row k beyond the superdiagonal
            const int mr = n - k - 2;
            taup[k] = make_reflector(&a(k, k + 1), &a(k, k + 2), mr, (size_t)N);
            for (uint64_t i = k + 1; i < N; ++i) {
                float w = a(i, k + 1);
                for (int j = k + 2; j < n; ++j) w += a(k, j) * a(i, j);
                w *= taup[k];
                a(i, k + 1) -= w;
                for (int j = k + 2; j < n; ++j) a(i, j) -= w * a(k, j);
            }
        }
    }
    Dense Bd(n), Ub(n), Vb(n);
    for (int k = 0; k < n; ++k) {
        Bd(k, k) = a(k, k);
        if (k + 1 < n) Bd(k, k + 1) = a(k, k + 1);
    }
    // U1 = H_0 H_1 ... H_{n-1} [I_n; 0]  (N x n), by backward accumulation
    std::vector<float> U1((size_t)N * n, 0.0f);
    auto u1 = [&](uint64_t i, int j) -> float& { return U1[(size_t)j * N + i]; };
    for (int j = 0; j < n; ++j) u1(j, j) = 1.0f;
    for (int k = n - 1; k >= 0; --k) {
        if (tauq[k] == 0.0f) continue;
        for (int j = k; j < n; ++j) {
            float w = u1(k, j);
            for (uint64_t i = k + 1; i < N; ++i) w += a(i, k) * u1(i, j);
            w *= tauq[k];
            u1(k, j) -= w;
            for (uint64_t i = k + 1; i < N; ++i) u1(i, j) -= w * a(i, k);
        }
    }
    if (!bidiagonal_svd(Bd, Ub, Vb)) return 2;
    // non-negative singular values, descending (nalgebra sorts: SVD::new)
    std::vector<float> sig(n);
    std::vector<int> order(n);
    for (int k = 0; k < n; ++k) {
        sig[k] = Bd(k, k);
        order[k] = k;
        if (sig[k] < 0.0f) sig[k] = -sig[k];  // the flip goes into the right factor, which is not needed here
    }
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return sig[x] > sig[y]; });
    // V_ref = left singular vectors of A = U1 Ub (N x n); column k of the result = order[k]
    std::vector<float> V((size_t)N * n);
    for (int k = 0; k < n; ++k) {
        const int src = order[k];
        float* dst = V.data() + (size_t)k * N;
        for (uint64_t i = 0; i < N; ++i) {
            float acc = 0.0f;
            for (int t = 0; t < n; ++t) acc += u1(i, t) * Ub(t, src);
            dst[i] = acc;
        }
        if (out_sigma) out_sigma[k] = sig[src];
    }
    // scale_columns_inplace (dmatrix_util.rs:986-995): per column mean / population sd with f32 left folds
    for (int k = 0; k < n; ++k) {
        float* x = V.data() + (size_t)k * N;
        const float nf = (float)(double)N;
        float s = 0.0f;
        for (uint64_t i = 0; i < N; ++i) s = s + x[i];
        const float mu = s / nf;
        float v = 0.0f;
        for (uint64_t i = 0; i < N; ++i) {
            const float d = x[i] - mu;
            v = v + d * d;
        }
        const float sd = std::sqrt(v / nf);
        for (uint64_t i = 0; i < N; ++i) x[i] += -mu;
        if (sd > 0.0f)
            for (uint64_t i = 0; i < N; ++i) x[i] /= sd;
    }
    for (uint64_t j = 0; j < N; ++j) {
        uint64_t c = 0;
        for (int k = 0; k < n; ++k)
            if (V[(size_t)k * N + j] > 0.0f) c |= (1ull << k);  // random_projection.rs:553-560
        codes[j] = c;
    }
    if (out_v) std::memcpy(out_v, V.data(), sizeof(float) * V.size());
    return 0;
}
