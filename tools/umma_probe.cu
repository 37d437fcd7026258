// umma_probe.cu — development probe (not part of the library): validates the tcgen05 operand
// layouts this repo relies on, against a CPU product, before the real kernels use them.
//   test 1: kind::i8,  A (u8) in TMEM, B (s8) in smem, K-major no-swizzle, s32 accumulate
//   test 2: kind::tf32, A and B in smem, K-major no-swizzle, f32 accumulate
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/umma_probe tools/umma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../legume-rs_b200/csrc/lg_umma.cuh"

using namespace umma;

#define CK(x)                                                                      \
    do {                                                                           \
        cudaError_t e = (x);                                                       \
        if (e != cudaSuccess) {                                                    \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
            exit(2);                                                               \
        }                                                                          \
    } while (0)

constexpr int M = 128;

// ---- test 1 -------------------------------------------------------------------------------
// a_rows: [128][KT] u8 row-major; b_tiled: KT/32 chunks of N*32 bytes in canonical layout
template <int N, int KT>
__global__ void k_probe_i8(const uint8_t* __restrict__ a_rows, const int8_t* __restrict__ b_tiled, int32_t* __restrict__ d_out,
                           uint32_t lbo, uint32_t sbo) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int BBYTES = N * KT;
    for (int i = threadIdx.x; i < BBYTES / 16; i += blockDim.x)
        reinterpret_cast<uint4*>(smem)[i] = reinterpret_cast<const uint4*>(b_tiled)[i];
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_barrier_init();
    }
    if (warp == 0) {
        tmem_alloc(&tmem_base_s, 512);
        tmem_relinquish();
    }
    fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = tmem_base_s;
    const uint32_t acc_col = 0, a_col = 256;
    // A: thread (row) writes KT bytes = KT/4 columns
    {
        const int row = threadIdx.x;  // 128 threads
        const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
        for (int c0 = 0; c0 < KT / 4; c0 += 16) {
            uint32_t r[16];
            for (int c = 0; c < 16; ++c) {
                const uint8_t* p = a_rows + (size_t)row * KT + (c0 + c) * 4;
                r[c] = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
            }
            tmem_st_x16(tbase + lane_base + a_col + c0, r);
        }
        tmem_wait_st();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 0 && elect_one()) {
        constexpr uint32_t idesc = make_idesc(CFMT_S32, FMT_U8, FMT_S8, M, N);
        for (int ch = 0; ch < KT / 32; ++ch) {
            const uint64_t db = make_smem_desc(smem_u32(smem) + ch * N * 32, lbo, sbo);
            mma_i8_ts(tbase + acc_col, tbase + a_col + ch * 8, db, idesc, ch > 0);
        }
        tc_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    {
        const int row = threadIdx.x;
        const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
        for (int c0 = 0; c0 < N; c0 += 16) {
            uint32_t r[16];
            tmem_ld_x16(tbase + lane_base + acc_col + c0, r);
            tmem_wait_ld();
            for (int c = 0; c < 16; ++c) d_out[(size_t)row * N + c0 + c] = (int32_t)r[c];
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, 512);
}

// ---- test 2 -------------------------------------------------------------------------------
// a_tiled / b_tiled: KT/8 chunks; each chunk = ROWS*32 bytes canonical (8x16B core matrices)
template <int N, int KT>
__global__ void k_probe_tf32(const float* __restrict__ a_tiled, const float* __restrict__ b_tiled, float* __restrict__ d_out,
                             uint32_t lbo, uint32_t sbo) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int warp = threadIdx.x >> 5;
    constexpr int ABYTES = M * KT * 4, BBYTES = N * KT * 4;
    uint8_t* sa = smem;
    uint8_t* sb = smem + ABYTES;
    for (int i = threadIdx.x; i < ABYTES / 16; i += blockDim.x) reinterpret_cast<uint4*>(sa)[i] = reinterpret_cast<const uint4*>(a_tiled)[i];
    for (int i = threadIdx.x; i < BBYTES / 16; i += blockDim.x) reinterpret_cast<uint4*>(sb)[i] = reinterpret_cast<const uint4*>(b_tiled)[i];
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_barrier_init();
    }
    if (warp == 0) {
        tmem_alloc(&tmem_base_s, 256);
        tmem_relinquish();
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = tmem_base_s;
    if (warp == 0 && elect_one()) {
        constexpr uint32_t idesc = make_idesc(CFMT_F32, FMT_TF32, FMT_TF32, M, N);
        for (int ch = 0; ch < KT / 8; ++ch) {
            const uint64_t da = make_smem_desc(smem_u32(sa) + ch * M * 32, lbo, sbo);
            const uint64_t db = make_smem_desc(smem_u32(sb) + ch * N * 32, lbo, sbo);
            mma_tf32_ss(tbase, da, db, idesc, ch > 0);
        }
        tc_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    {
        const int row = threadIdx.x;
        const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
        for (int c0 = 0; c0 < N; c0 += 16) {
            uint32_t r[16];
            tmem_ld_x16(tbase + lane_base + c0, r);
            tmem_wait_ld();
            for (int c = 0; c < 16; ++c) d_out[(size_t)row * N + c0 + c] = __uint_as_float(r[c]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, 256);
}

// canonical K-major no-swizzle placement of element (row, kbyte) inside one K-chunk of 32 bytes
static size_t canon(int row, int kbyte, int lbo, int sbo) { return (size_t)(row / 8) * sbo + (size_t)(kbyte / 16) * lbo + (row % 8) * 16 + (kbyte % 16); }

int main() {
    CK(cudaSetDevice(0));
    srand(1);
    {  // ---- test 1
        constexpr int N = 160, KT = 64;
        std::vector<uint8_t> a((size_t)M * KT);
        std::vector<int8_t> b((size_t)N * KT), bt((size_t)N * KT);
        for (auto& x : a) x = (rand() % 4 == 0) ? 128 : (rand() % 16 == 0 ? (uint8_t)(rand() % 256) : 0);
        for (auto& x : b) x = (int8_t)(rand() % 256 - 128);
        const int lbo = 128, sbo = 256;
        for (int ch = 0; ch < KT / 32; ++ch)
            for (int n = 0; n < N; ++n)
                for (int k = 0; k < 32; ++k) bt[(size_t)ch * N * 32 + canon(n, k, lbo, sbo)] = b[(size_t)n * KT + ch * 32 + k];
        uint8_t* da;
        int8_t* db;
        int32_t* dd;
        CK(cudaMalloc(&da, a.size()));
        CK(cudaMalloc(&db, bt.size()));
        CK(cudaMalloc(&dd, sizeof(int32_t) * M * N));
        CK(cudaMemcpy(da, a.data(), a.size(), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(db, bt.data(), bt.size(), cudaMemcpyHostToDevice));
        std::vector<int32_t> want((size_t)M * N), got((size_t)M * N);
        for (int m = 0; m < M; ++m)
            for (int n = 0; n < N; ++n) {
                int32_t s = 0;
                for (int k = 0; k < KT; ++k) s += (int32_t)a[(size_t)m * KT + k] * (int32_t)b[(size_t)n * KT + k];
                want[(size_t)m * N + n] = s;
            }
        for (int variant = 0; variant < 2; ++variant) {
            CK(cudaMemset(dd, 0xff, sizeof(int32_t) * M * N));
            k_probe_i8<N, KT><<<1, 128, N * KT>>>(da, db, dd, variant ? sbo : lbo, variant ? lbo : sbo);
            CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(got.data(), dd, sizeof(int32_t) * M * N, cudaMemcpyDeviceToHost));
            size_t bad = 0;
            for (size_t i = 0; i < got.size(); ++i) bad += got[i] != want[i];
            printf("i8 TS  N=%d KT=%d variant=%d (lbo=%d sbo=%d): mismatches %zu / %zu   got[0..3]=%d %d %d %d want=%d %d %d %d\n", N, KT,
                   variant, variant ? sbo : lbo, variant ? lbo : sbo, bad, got.size(), got[0], got[1], got[2], got[3], want[0],
                   want[1], want[2], want[3]);
        }
    }
    {  // ---- test 2
        constexpr int N = 128, KT = 16;
        std::vector<float> a((size_t)M * KT), b((size_t)N * KT), at((size_t)M * KT), bt((size_t)N * KT);
        auto rnd = [] { return (float)((rand() % 2001) - 1000) / 256.0f; };  // exactly representable in tf32
        for (auto& x : a) x = rnd();
        for (auto& x : b) x = rnd();
        const int lbo = 128, sbo = 256;
        for (int ch = 0; ch < KT / 8; ++ch) {
            for (int m = 0; m < M; ++m)
                for (int k = 0; k < 8; ++k) memcpy(reinterpret_cast<uint8_t*>(at.data()) + (size_t)ch * M * 32 + canon(m, k * 4, lbo, sbo), &a[(size_t)m * KT + ch * 8 + k], 4);
            for (int n = 0; n < N; ++n)
                for (int k = 0; k < 8; ++k) memcpy(reinterpret_cast<uint8_t*>(bt.data()) + (size_t)ch * N * 32 + canon(n, k * 4, lbo, sbo), &b[(size_t)n * KT + ch * 8 + k], 4);
        }
        float *da, *db, *dd;
        CK(cudaMalloc(&da, at.size() * 4));
        CK(cudaMalloc(&db, bt.size() * 4));
        CK(cudaMalloc(&dd, sizeof(float) * M * N));
        CK(cudaMemcpy(da, at.data(), at.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(db, bt.data(), bt.size() * 4, cudaMemcpyHostToDevice));
        std::vector<float> got((size_t)M * N);
        for (int variant = 0; variant < 2; ++variant) {
            k_probe_tf32<N, KT><<<1, 128, (M + N) * KT * 4>>>(da, db, dd, variant ? sbo : lbo, variant ? lbo : sbo);
            CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(got.data(), dd, sizeof(float) * M * N, cudaMemcpyDeviceToHost));
            size_t bad = 0;
            double maxerr = 0;
            for (int m = 0; m < M; ++m)
                for (int n = 0; n < N; ++n) {
                    double s = 0;
                    for (int k = 0; k < KT; ++k) s += (double)a[(size_t)m * KT + k] * b[(size_t)n * KT + k];
                    const double e = fabs(s - got[(size_t)m * N + n]);
                    maxerr = e > maxerr ? e : maxerr;
                    bad += e > 1e-3;
                }
            printf("tf32 SS N=%d KT=%d variant=%d: mismatches %zu / %d  max err %.3g\n", N, KT, variant, bad, M * N, maxerr);
        }
    }
    return 0;
}
