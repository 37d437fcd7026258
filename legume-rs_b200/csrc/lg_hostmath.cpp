// lg_hostmath.cpp — the one piece of host arithmetic on the path: the reference's 16-lane squared distance
// (matrix-util/src/knn/metric.rs:19-45), used to order the B batch centroids (sort_batch_proximity, B <= a few dozen).
// The K x r Householder QR and the kk x kk eigen-decomposition of K3 run on the device (lg_codes.cu).
// Compile with -ffp-contract=off: the arithmetic order is part of the parity surface.
#include <cstddef>

float lgh_l2_sq(const float* a, const float* b, int d);


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
