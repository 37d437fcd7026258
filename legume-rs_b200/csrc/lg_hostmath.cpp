// lg_hostmath.cpp — the two small dense factorizations that stay on the host:
// a K x r Householder QR (r <= 21) and a kk x kk symmetric eigen-decomposition (kk <= 16).
// They follow nalgebra's QR control flow (matrix-util/src/dmatrix_rsvd.rs:129-131 calls
// `.qr().q()`) and replace nalgebra's SVD of the kk x N matrix B by an eigen-decomposition of
// its Gram matrix.  Compile with -ffp-contract=off: the arithmetic order
// here is part of the parity surface for group codes.
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <numeric>
#include <vector>

void lgh_householder_q(const float* a_kr, int K, int r, float* q_kr);
void lgh_jacobi_eig(const double* g, int n, double* evals, double* evecs);

namespace {
inline float signum(float x) { return x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : 0.0f); }

struct ColMajor {
    float* p;
    int ld;
    float& operator()(int i, int j) const { return p[(size_t)j * ld + i]; }
};
}  // namespace

// Thin Q (K x min(K, r)) of the Householder QR of a K x r matrix; extra columns zeroed.
void lgh_householder_q(const float* a_kr, int K, int r, float* q_kr) {
    const int steps = std::min(K, r);
    std::vector<float> work(a_kr, a_kr + (size_t)K * r);
    std::vector<float> diag(steps, 0.0f);
    ColMajor W{work.data(), K};
    ColMajor Q{q_kr, K};

    // apply H = I - 2 v v^T (scaled by `sign`) to column `col` of M, rows [from, K)
    auto reflect = [&](const ColMajor& M, int col, int axis_col, int from, float sign) {
        float dot = 0.0f;
        for (int i = from; i < K; ++i) dot = dot + W(i, axis_col) * M(i, col);
        const float f = dot * (sign * -2.0f);
        for (int i = from; i < K; ++i) M(i, col) = f * W(i, axis_col) + sign * M(i, col);
    };

    for (int c = 0; c < steps; ++c) {
        float sq = 0.0f;
        for (int i = c; i < K; ++i) sq = sq + W(i, c) * W(i, c);
        const float norm = std::sqrt(sq);
        const float head = W(c, c);
        const float signed_norm = (head < 0.0f ? -1.0f : 1.0f) * norm;
        const float factor = (sq + std::fabs(head) * norm) * 2.0f;
        W(c, c) = head + signed_norm;
        if (factor != 0.0f) {
            const float root = std::sqrt(factor);
            for (int i = c; i < K; ++i) W(i, c) /= root;
            float n2 = 0.0f;
            for (int i = c; i < K; ++i) n2 = n2 + W(i, c) * W(i, c);
            const float nn = std::sqrt(n2);
            for (int i = c; i < K; ++i) W(i, c) /= nn;
            diag[c] = -signed_norm;
            const float s = signum(diag[c]);
            for (int j = c + 1; j < r; ++j) reflect(W, j, c, c, s);
        } else {
            diag[c] = signed_norm;
        }
    }
    for (int j = 0; j < r; ++j)
        for (int i = 0; i < K; ++i) Q(i, j) = (i == j && j < steps) ? 1.0f : 0.0f;
    for (int c = steps - 1; c >= 0; --c) {
        const float s = signum(diag[c]);
        for (int j = c; j < steps; ++j) reflect(Q, j, c, c, s);
    }
}

// Cyclic Jacobi for a symmetric n x n matrix (column-major f64).  Row-major sweep over (p, q),
// p < q; eigenvalues returned in descending order (stable), evecs[k*n + i] = i-th entry of the
// k-th eigenvector.
void lgh_jacobi_eig(const double* g, int n, double* evals, double* evecs) {
    std::vector<double> A(g, g + (size_t)n * n), V((size_t)n * n, 0.0);
    auto a = [&](int i, int j) -> double& { return A[(size_t)j * n + i]; };
    auto v = [&](int i, int j) -> double& { return V[(size_t)j * n + i]; };
    for (int i = 0; i < n; ++i) v(i, i) = 1.0;
    for (int sweep = 0; sweep < 64; ++sweep) {
        double off = 0.0, dg = 0.0;
        for (int p = 0; p < n; ++p) {
            dg += a(p, p) * a(p, p);
            for (int q = p + 1; q < n; ++q) off += a(p, q) * a(p, q);
        }
        if (off <= 1e-60 || off <= 1e-32 * dg) break;
        for (int p = 0; p < n - 1; ++p) {
            for (int q = p + 1; q < n; ++q) {
                const double apq = a(p, q);
                if (apq == 0.0) continue;
                const double theta = (a(q, q) - a(p, p)) / (2.0 * apq);
                const double t = (theta >= 0.0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
                const double cs = 1.0 / std::sqrt(t * t + 1.0), sn = t * cs;
                for (int k = 0; k < n; ++k) {  // columns p, q
                    const double x = a(k, p), y = a(k, q);
                    a(k, p) = cs * x - sn * y;
                    a(k, q) = sn * x + cs * y;
                }
                for (int k = 0; k < n; ++k) {  // rows p, q
                    const double x = a(p, k), y = a(q, k);
                    a(p, k) = cs * x - sn * y;
                    a(q, k) = sn * x + cs * y;
                }
                for (int k = 0; k < n; ++k) {
                    const double x = v(k, p), y = v(k, q);
                    v(k, p) = cs * x - sn * y;
                    v(k, q) = sn * x + cs * y;
                }
            }
        }
    }
    std::vector<int> order(n);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return a(x, x) > a(y, y); });
    for (int k = 0; k < n; ++k) {
        evals[k] = a(order[k], order[k]);
        for (int i = 0; i < n; ++i) evecs[(size_t)k * n + i] = v(i, order[k]);
    }
}

// knn/metric.rs:19-45 on the host (used for the B x B batch-centroid proximity, batch.rs:182-234):
// 16 lane accumulators, left fold, sequential tail.  Compiled with -ffp-contract=off.
float lgh_l2_sq(const float* a, const float* b, int d) {
    float acc[16];
    for (int l = 0; l < 16; ++l) acc[l] = 0.0f;
    int c = 0;
    for (; c + 16 <= d; c += 16)
        for (int l = 0; l < 16; ++l) {
            const float df = a[c + l] - b[c + l];
            acc[l] += df * df;
        }
    float sum = 0.0f;
    for (int l = 0; l < 16; ++l) sum += acc[l];
    for (; c < d; ++c) {
        const float df = a[c] - b[c];
        sum += df * df;
    }
    return sum;
}
