// lg_project_umma.cu — K1 on the Blackwell tensor path (project_columns_visitor,
// data-beans-alg/src/random_projection.rs:169-199).
//
// Why tensor cores here: per non-zero the projection needs one 200-byte basis row.  A CUDA-core
// SpMM has to move that row from shared memory (or L2) to registers for every non-zero, and the
// 128 B/clk/SM shared-memory pipe caps it near 2-2.5 clk per nnz (~15% of the HBM roofline; the
// warp-per-cell kernel in lg_project.cu measures 5%).  The tensor core reads a basis tile once per
// 128 cells instead, so the kernel becomes a dense int8 GEMM against a 0/1 indicator of the CSC
// pattern, which sm_100a runs at twice the f16 rate with EXACT s32 accumulation:
//
//   proj_raw[:, j] = ( ln2 * sum_{i in nnz(j)} B[i, :]  +  sum_{i: y_ij != 1} (ln(1+y_ij) - ln2) B[i, :] ) / ||x_j||
//                       `---- K1a: tcgen05 kind::i8 ----'   `---- K1b: CUDA cores, the few counts > 1 ----'
//
//   * the basis is quantised to 2^-20 fixed point (after a per-column power-of-two scale that brings the column's
//     largest magnitude to [4, 7.96)) and split into three signed 8-bit digits (N = 3K columns of the B operand).
//   * K1b (k_project_prep, one warp per cell, one coalesced pass over indices AND values) writes the
//     sparsity pattern as a bitmap (1 bit per gene: smaller than the u32 index stream above 3% density),
//     tiled per (256 cells x 2048 genes) so the tensor kernel fetches it with one bulk copy per chunk,
//     and leaves corr/norm and ln2/norm per cell.
//   * K1a never materialises the dense A operand in memory: expander warps turn 64-bit bitmap words
//     into 16 TMEM columns (u8 in {0, 128}) with two ALU ops per register and tcgen05.st them.
#include <cstdio>
#include <cstdlib>

#include "lg_common.cuh"
#include "lg_umma.cuh"

using namespace umma;

namespace {

constexpr int TILE_M = 128;          // cells per accumulator (UMMA M)
constexpr int NT = 2;                // accumulators (cell tiles) per CTA
constexpr int CELLS = TILE_M * NT;   // cells per CTA pass
constexpr int GS = 128;              // genes per pipeline stage (4 MMAs of K = 32)
constexpr int GC = 2048;             // genes per bitmap chunk: K1b stores one 256-byte row piece per chunk (1024 costs it 1 ms)
constexpr int BM_STRIDE = GC / 32 + 2;  // 66 words per cell row: 8-byte aligned, conflict-free LDS.64
static_assert(CELLS == LG_PAT_CELLS && GC == LG_PAT_GC && BM_STRIDE == LG_PAT_STRIDE, "lg_pattern describes this layout to the collapse");
constexpr int NBST = 4;              // B-operand ring depth (stages)
constexpr int NBM = 2;               // bitmap chunks in flight
constexpr int NAST = 3;              // A-operand ring depth in TMEM (stages per tile)
constexpr int A_COLS = GS / 4;       // 32 TMEM columns per A stage
constexpr int N_EXP_WARPS = 4 * NT;  // 8 expander warps (also the epilogue)
constexpr int WARP_MMA = N_EXP_WARPS;           // NT issuing warps, one per cell tile: a single thread cannot issue
                                                // 8 MMAs + their commits in the 640 clk the tensor pipe needs for them
constexpr int WARP_LOAD_B = N_EXP_WARPS + NT;       // basis stages
constexpr int WARP_LOAD_BM = N_EXP_WARPS + NT + 1;  // bitmap chunks
constexpr int THREADS = (N_EXP_WARPS + NT + 2) * 32;  // 384
constexpr uint32_t BM_CHUNK_BYTES = CELLS * BM_STRIDE * 4;  // 67584 bytes per (supertile, chunk)
constexpr int PREP_WARPS = 8;
constexpr float QSCALE = 1048576.0f;                       // 2^20
constexpr int QMAX = 127 * 65536 + 127 * 256 + 127;        // largest 3-digit signed base-256 value

struct Barriers {
    uint64_t b_full[NBST], b_empty[NBST];
    uint64_t a_full[NT][NAST], a_empty[NT][NAST];
    uint64_t bm_full[NBM], bm_empty[NBM];
    uint64_t acc_full[NT], acc_empty[NT];
};

// ---- basis -> three signed base-256 digits, laid out as the UMMA B operand ---------------------
// Every basis column (dim) gets its own power-of-two scale 2^e so that its largest magnitude lands in [4, 7.96) of the
// 2^-20 fixed-point grid: a relative resolution of ~1e-7 of the column's range whatever its size (row-weighted bases,
// U / sigma of a Nystrom basis), exact to undo in the epilogue.  A standard-normal basis keeps e = 0.
__global__ void __launch_bounds__(256) k_basis_colmax(const float* __restrict__ basis_kd, uint64_t D, int K,
                                                      unsigned int* __restrict__ colmax_bits) {
    // per-block maxima in shared memory first: one global atomic per (block, column) instead of one per element
    __shared__ unsigned int smax[128];
    if (threadIdx.x < 128) smax[threadIdx.x] = 0u;
    __syncthreads();
    const uint64_t total = D * (uint64_t)K;
    const uint64_t per_block = 256ull * 16ull;
    const uint64_t e0 = (uint64_t)blockIdx.x * per_block;
    for (uint64_t e = e0 + threadIdx.x; e < e0 + per_block && e < total; e += 256) {
        // non-negative floats order as uints; NaN (0x7fc00000) and +inf (0x7f800000) come out on top and are caught below
        atomicMax(&smax[e % K], __float_as_uint(fabsf(basis_kd[e])));
    }
    __syncthreads();
    if ((int)threadIdx.x < K && smax[threadIdx.x]) atomicMax(&colmax_bits[threadIdx.x], smax[threadIdx.x]);
}
// exponent e of the column's scale; sets *bad for a non-finite column
__device__ __forceinline__ int basis_col_exp(unsigned int max_bits, int* bad) {
    const float m = __uint_as_float(max_bits);
    if (!(m <= 3.0e38f)) {
        if (bad) *bad = 1;
        return 0;
    }
    if (m == 0.0f) return 0;
    int e = 2 - ilogbf(m);                       // m * 2^e in [4, 8)
    if (ldexpf(m, e) >= 7.96f) --e;              // keep clear of the largest three-digit value
    return e < -100 ? -100 : (e > 100 ? 100 : e);
}
// chunk = 32 genes x NB columns in the K-major no-swizzle canonical form:
//   byte(n, k) = (n/8)*256 + (k/16)*128 + (n%8)*16 + (k%16),  n = digit*K + dim
__global__ void k_quantize_basis(const float* __restrict__ basis_kd, uint64_t D, int K, int NB, uint64_t Dpad,
                                 const unsigned int* __restrict__ colmax_bits, int8_t* __restrict__ bq,
                                 int* __restrict__ too_large) {
    const uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= Dpad * (uint64_t)NB) return;
    const uint64_t gene = e / NB;
    const int n = (int)(e % NB);
    const int digit = n / K, dim = n % K;
    int8_t out = 0;
    if (gene < D && digit < 3) {
        const float b = basis_kd[gene * K + dim];
        const float qf = rintf(ldexpf(b, 20 + basis_col_exp(colmax_bits[dim], too_large)));
        if (!(fabsf(qf) <= (float)QMAX)) atomicOr(too_large, 1);
        int q = (int)fminf(fmaxf(qf, -(float)QMAX), (float)QMAX);
        int d0 = ((q + 128) & 255) - 128;
        q = (q - d0) >> 8;
        int d1 = ((q + 128) & 255) - 128;
        q = (q - d1) >> 8;
        const int d[3] = {d0, d1, q};
        out = (int8_t)d[digit];
    }
    const uint64_t chunk = gene >> 5;
    // K position of a gene inside its group of 32: the expander's (w << (7 - b)) & 0x80808080 puts bit b + 8j of a bitmap
    // word at K position 4b + j, and the scan keeps the NATURAL bit order (gene offset o at bit o: one shift per entry),
    // so the operand rows are permuted here instead, once per call
    const int o = (int)(gene & 31);
    const int k = 4 * (o & 7) + (o >> 3);
    bq[chunk * (uint64_t)(NB * 32) + (uint64_t)(n >> 3) * 256 + (k >> 4) * 128 + (n & 7) * 16 + (k & 15)] = out;
}

// ---- K1b: one pass over the CSC block (CUDA cores, HBM-bound) -----------------------------------
// per cell: (1) pattern bitmap row -> tiled global scratch, (2) norm, (3) correction for counts != 1.
// writes out[j*K + k] = corr_k / norm_j and scale[j] = ln2 / norm_j.
// bitmap word w of a cell covers genes 32w..32w+31, gene offset o at bit o; the expander's (w << (7-b)) & 0x80808080
// yields the four K-positions 4b..4b+3 of an MMA operand column, which k_quantize_basis fills with genes b, b+8, b+16, b+24.
constexpr int PREP_Q = 256;   // exception queue entries per warp (>= 31 + 128)
constexpr int PREP_LUT = 64;  // log1p look-up for integer counts below this

// HALF2 (K even, basis 8-byte aligned): the basis rows of TWO exceptions are gathered at once, one per half-warp,
// as float2 (lane l of a half holds dims 2l, 2l+1, 2l+32, 2l+33) — a third of the instructions per exception of the
// lanes-as-dims form, which stays as the fallback for odd K.
// VEC (index / value arrays 16-byte aligned): the aligned interior of a cell's nnz range is read with 128-bit loads (a lane
// holds 4 consecutive nnz, twice the bytes in flight per warp and a quarter of the load instructions); the <= 3 entries
// before and after it go through one masked round.
// MODE 0: projection (x = ln(1 + y), L2-normalised).  MODE 1: Nystrom re-projection without a batch divisor
// (senna/src/svd/fit.rs:433-466): x = ln(1 + y * csn / ||y||), standardised over the cell's stored entries; the common
// value is then z1 = ln(1 + csn / ||y||) and the pattern sum gets the weight (z1 - mean) / sd, so the same two-part
// split applies with exceptions weighted (x - z1) / sd.  ||y|| needs its own sweep over the cell's values first.
template <int NACC, bool HALF2, bool VEC, int MODE>
__global__ void __launch_bounds__(PREP_WARPS * 32, 4) k_project_prep(const uint64_t* __restrict__ indptr,
                                                                     const uint32_t* __restrict__ indices,
                                                                     const float* __restrict__ values, uint64_t ncols,
                                                                     const float* __restrict__ basis_kd, int K, uint32_t nchunks,
                                                                     uint32_t* __restrict__ bm_global, float* __restrict__ out,
                                                                     float* __restrict__ scale, float csn,
                                                                     uint32_t* __restrict__ exc, uint32_t* __restrict__ exc_cnt,
                                                                     int* __restrict__ exc_ovf) {
    extern __shared__ __align__(16) uint32_t rows[];  // PREP_WARPS bitmap rows, then the exception queues
    __shared__ float lut_x[PREP_LUT];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t row_words = nchunks * (GC / 32);
    uint32_t* row = rows + (size_t)warp * row_words;
    uint32_t* qg = rows + (size_t)PREP_WARPS * row_words + warp * (2 * PREP_Q);  // queued gene
    float* qv = reinterpret_cast<float*>(qg + PREP_Q);                            // queued raw value
    if (threadIdx.x < PREP_LUT) lut_x[threadIdx.x] = log1pf((float)threadIdx.x);
    __syncthreads();
    const uint64_t warp0 = (uint64_t)blockIdx.x * PREP_WARPS + warp;
    const uint64_t nwarps = (uint64_t)gridDim.x * PREP_WARPS;
    const float ln2 = lut_x[1];
    const unsigned lt_mask = (1u << lane) - 1u;
    for (uint64_t j = warp0; j < ncols; j += nwarps) {
        for (uint32_t i = lane; i < row_words / 4; i += 32) reinterpret_cast<uint4*>(row)[i] = make_uint4(0, 0, 0, 0);
        __syncwarp();
        const uint64_t lo = indptr[j];
        const uint32_t n = (uint32_t)(indptr[j + 1] - lo);
        const uint32_t* ip = indices + lo + lane;
        const float* vp = values + lo + lane;
        // lg_pattern: the cell's counts != 1 also leave as packed words for the collapse (exc != NULL only in projection mode)
        // in the slots [(lo >> 1) + j, ((lo + n) >> 1) + j] (recomputed in the drain: nothing extra stays live across the scan)
        float acc[HALF2 ? 4 : NACC];
#pragma unroll
        for (int a = 0; a < (HALF2 ? 4 : NACC); ++a) acc[a] = 0.0f;
        float nsq = 0.0f;            // MODE 0: sum x^2 over the exceptions
        double e1 = 0.0, e2 = 0.0;   // MODE 1: sum (x - z1), sum (x - z1)^2 over the exceptions (x - z1 is exact in f32)
        float a_j = 0.0f, base_x = ln2;  // MODE 1: csn / ||y|| and the common value z1 = ln(1 + a_j)
        if constexpr (MODE == 1) {
            float q = 0.0f;
            for (uint32_t t = lane; t < n; t += 32) {
                const float y = __ldg(vp + t - lane);
                q = fmaf(y, y, q);
            }
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) q += __shfl_xor_sync(0xffffffffu, q, off);
            a_j = __fdiv_rn(csn, fmaxf(sqrtf(q), 1e-8f));
            base_x = log1pf(a_j);
        }
        uint32_t qhead = 0, qtail = 0;  // warp-uniform ring cursors of the exception queue

        // entries != 1 are rare in count data: they are queued raw and handled 32 at a time, one queue entry per
        // lane for the log1p / norm part, then broadcast so that all lanes gather the basis row (lanes = dims)
        auto drain = [&](uint32_t cnt) {
            const bool on = (uint32_t)lane < cnt;
            const uint32_t g = on ? qg[(qhead + lane) & (PREP_Q - 1)] : 0u;
            float w = 0.0f;
            if (on) {
                const float val = qv[(qhead + lane) & (PREP_Q - 1)];
                float x;
                if constexpr (MODE == 1) {
                    x = log1pf(val * a_j);
                    w = x - base_x;
                    e1 += (double)w;
                    e2 = fma((double)w, (double)w, e2);
                } else {
                    const int vi = (int)val;
                    x = (val == (float)vi && vi >= 0 && vi < PREP_LUT) ? lut_x[vi] : log1pf(val);
                    nsq = fmaf(x, x, nsq);
                    w = x - base_x;
                    if (exc) {
                        const uint32_t idx = qhead + lane;  // qhead counts the cell's entries drained so far
                        const uint64_t e0 = (lo >> 1) + j;
                        if (val == (float)vi && vi >= 0 && vi <= 32767 && e0 + idx <= ((lo + n) >> 1) + j)
                            exc[e0 + idx] = g | ((vi == 0 ? LG_PAT_ZERO : (uint32_t)(vi - 1)) << 17);
                        else
                            *exc_ovf = 1;  // the collapse falls back to the CSC arrays
                    }
                }
            }
            if constexpr (HALF2) {
                const int half = lane >> 4, l = lane & 15;
                const bool hi_on = 2 * (l + 16) < K;
                for (uint32_t e0 = 0; e0 < cnt; e0 += 8) {  // 8 exceptions = 4 pairs per round
                    float2 b0[4], b1[4];
                    float ws[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int src = (int)e0 + 2 * e + half;  // lanes >= cnt carry w = 0 and gene 0
                        const uint32_t ge = __shfl_sync(0xffffffffu, g, src);
                        ws[e] = __shfl_sync(0xffffffffu, w, src);
                        const float2* brow = reinterpret_cast<const float2*>(basis_kd + (size_t)ge * K);
                        b0[e] = (2 * l < K) ? __ldg(brow + l) : make_float2(0.f, 0.f);
                        b1[e] = hi_on ? __ldg(brow + l + 16) : make_float2(0.f, 0.f);
                    }
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        acc[0] = fmaf(ws[e], b0[e].x, acc[0]);
                        acc[1] = fmaf(ws[e], b0[e].y, acc[1]);
                        acc[2] = fmaf(ws[e], b1[e].x, acc[2]);
                        acc[3] = fmaf(ws[e], b1[e].y, acc[3]);
                    }
                }
            } else {
            for (uint32_t e0 = 0; e0 < cnt; e0 += 8) {
                float bv[8][NACC], ws[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const uint32_t ge = __shfl_sync(0xffffffffu, g, e0 + e);
                    ws[e] = __shfl_sync(0xffffffffu, w, e0 + e);
                    const float* brow = basis_kd + (size_t)ge * K + lane;
#pragma unroll
                    for (int a = 0; a < NACC; ++a) bv[e][a] = (lane + 32 * a < K) ? __ldg(brow + 32 * a) : 0.0f;
                }
#pragma unroll
                for (int e = 0; e < 8; ++e)
#pragma unroll
                    for (int a = 0; a < NACC; ++a) acc[a] = fmaf(ws[e], bv[e][a], acc[a]);
            }
            }
            qhead += cnt;
        };
        auto consume = [&](uint32_t ixu, float vu, bool live) {
            if (live) atomicOr(row + (ixu >> 5), 1u << (ixu & 31));
            const bool exc = vu != 1.0f;  // dead lanes carry 1
            const unsigned m = __ballot_sync(0xffffffffu, exc);
            if (m) {
                if (exc) {
                    const uint32_t slot = (qtail + __popc(m & lt_mask)) & (PREP_Q - 1);
                    qg[slot] = ixu;
                    qv[slot] = vu;
                }
                qtail += __popc(m);
            }
        };

        // The queue must fill in ASCENDING POSITION inside the cell whatever the alignment of the cell's first entry:
        // the drain adds the exceptions in queue order, so any other order would make a cell's projection depend (in
        // the last bits) on where its column starts in the arrays, i.e. on how the cells were sharded.
        auto consume4 = [&](const uint4 g, const float4 x, bool live) {
            if (live) {
                atomicOr(row + (g.x >> 5), 1u << (g.x & 31));
                atomicOr(row + (g.y >> 5), 1u << (g.y & 31));
                atomicOr(row + (g.z >> 5), 1u << (g.z & 31));
                atomicOr(row + (g.w >> 5), 1u << (g.w & 31));
            }
            const bool e0 = x.x != 1.0f, e1 = x.y != 1.0f, e2 = x.z != 1.0f, e3 = x.w != 1.0f;  // dead lanes carry 1
            const uint32_t cnt = (uint32_t)e0 + (uint32_t)e1 + (uint32_t)e2 + (uint32_t)e3;
            const unsigned b0 = __ballot_sync(0xffffffffu, cnt & 1u), b1 = __ballot_sync(0xffffffffu, cnt & 2u),
                           b2 = __ballot_sync(0xffffffffu, cnt & 4u);
            if (b0 | b1 | b2) {
                // a lane's four entries are consecutive positions: lanes in order, entries in order inside a lane
                uint32_t slot = qtail + __popc(b0 & lt_mask) + 2 * __popc(b1 & lt_mask) + 4 * __popc(b2 & lt_mask);
                if (e0) {
                    qg[slot & (PREP_Q - 1)] = g.x;
                    qv[slot & (PREP_Q - 1)] = x.x;
                    ++slot;
                }
                if (e1) {
                    qg[slot & (PREP_Q - 1)] = g.y;
                    qv[slot & (PREP_Q - 1)] = x.y;
                    ++slot;
                }
                if (e2) {
                    qg[slot & (PREP_Q - 1)] = g.z;
                    qv[slot & (PREP_Q - 1)] = x.z;
                    ++slot;
                }
                if (e3) {
                    qg[slot & (PREP_Q - 1)] = g.w;
                    qv[slot & (PREP_Q - 1)] = x.w;
                }
                qtail += __popc(b0) + 2 * __popc(b1) + 4 * __popc(b2);
            }
        };
        if constexpr (VEC) {
            const uint64_t hi = lo + n;
            const uint64_t a_up = (lo + 3) & ~3ull, z_dn = hi & ~3ull;
            const uint64_t a = a_up < hi ? a_up : hi, z = z_dn > a ? z_dn : a;  // aligned interior [a, z)
            if (a > lo) {  // head [lo, a): at most 3 entries, one per lane
                const uint64_t e = lo + lane;
                const bool live = e < a;
                consume(live ? __ldg(indices + e) : 0u, live ? __ldg(values + e) : 1.0f, live);
            }
            const uint32_t nbat = (uint32_t)((z - a + 127) >> 7);
            const uint4* ip4 = reinterpret_cast<const uint4*>(indices + a) + lane;
            const float4* vp4 = reinterpret_cast<const float4*>(values + a) + lane;
            const uint32_t ngrp = (uint32_t)((z - a) >> 2);  // 128-bit groups in the interior
            uint4 g4 = make_uint4(0, 0, 0, 0);
            float4 x4 = make_float4(1.f, 1.f, 1.f, 1.f);
            if ((uint32_t)lane < ngrp) {
                g4 = __ldg(ip4);
                x4 = __ldg(vp4);
            }
            for (uint32_t b = 0; b < nbat; ++b) {
                // software pipeline: the next 128 nnz are in flight while this batch is consumed
                uint4 ng = make_uint4(0, 0, 0, 0);
                float4 nx = make_float4(1.f, 1.f, 1.f, 1.f);
                const bool nlive = 32u * (b + 1) + lane < ngrp;
                if (nlive) {
                    ng = __ldg(ip4 + 32 * (b + 1));
                    nx = __ldg(vp4 + 32 * (b + 1));
                }
                consume4(g4, x4, 32u * b + lane < ngrp);
                __syncwarp();
                while (qtail - qhead >= 32) drain(32);
                g4 = ng;
                x4 = nx;
            }
            if (hi > z) {  // tail [z, hi): at most 3 entries, after the interior
                const uint64_t e = z + lane;
                const bool live = e < hi;
                consume(live ? __ldg(indices + e) : 0u, live ? __ldg(values + e) : 1.0f, live);
                __syncwarp();
            }
        } else {
        const uint32_t nfull = n >> 7;
        uint32_t ix[4];
        float v[4];
        if (nfull) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                ix[u] = __ldg(ip + 32 * u);
                v[u] = __ldg(vp + 32 * u);
            }
        }
        for (uint32_t b = 0; b < nfull; ++b) {
            // software pipeline: the next 128 nnz are in flight while this batch is consumed
            uint32_t nix[4];
            float nv[4];
            const uint32_t nb = (b + 1 < nfull) ? (b + 1) : b;  // last iteration re-reads its own (cached) batch
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                nix[u] = __ldg(ip + 128 * nb + 32 * u);
                nv[u] = __ldg(vp + 128 * nb + 32 * u);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) consume(ix[u], v[u], true);
            __syncwarp();
            while (qtail - qhead >= 32) drain(32);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                ix[u] = nix[u];
                v[u] = nv[u];
            }
        }
        if (n & 127) {  // ragged tail
            const uint32_t base = nfull << 7;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const bool live = base + 32 * u + lane < n;
                const uint32_t ixu = live ? __ldg(ip + base + 32 * u) : 0u;
                const float vu = live ? __ldg(vp + base + 32 * u) : 1.0f;
                consume(ixu, vu, live);
            }
            __syncwarp();
        }
        }
        while (qtail != qhead) drain(min(32u, qtail - qhead));
        const uint32_t n_one = n - qtail;  // qtail counts every queued exception of this cell
        __syncwarp();
        // tiled store: (supertile, chunk) blocks of CELLS rows x BM_STRIDE words; one 256-byte chunk row per iteration
        {
            const uint64_t sup = j / CELLS;
            const uint32_t r = (uint32_t)(j % CELLS);
            uint32_t* dst = bm_global + ((sup * nchunks) * CELLS + r) * (uint64_t)BM_STRIDE;
            static_assert(GC / 32 == 32 || GC / 32 == 64, "one or two bitmap words per lane and chunk");
            for (uint32_t c = 0; c < nchunks; ++c) {
                if (GC / 32 == 64) {
                    const uint2 w2 = reinterpret_cast<const uint2*>(row + (size_t)c * (GC / 32))[lane];
                    reinterpret_cast<uint2*>(dst + (size_t)c * (CELLS * BM_STRIDE))[lane] = w2;
                } else {
                    dst[(size_t)c * (CELLS * BM_STRIDE) + lane] = row[(size_t)c * (GC / 32) + lane];
                }
            }
        }
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) nsq += __shfl_xor_sync(0xffffffffu, nsq, off);
        float denom, pat_scale;  // out = corr / denom + pattern sum * pat_scale
        if constexpr (MODE == 1) {
            e1 = lg_butterfly32(e1);
            e2 = lg_butterfly32(e2);
            // moments of d = x - z1 over the stored entries (d = 0 for every count of one): mean = z1 + sum d / n and
            // sd^2 = sum d^2 / n - (sum d / n)^2 has no cancellation, unlike the reference's s2/n - mean^2
            const double nn = n > 0 ? (double)n : 1.0;
            const double md = e1 / nn, var = e2 / nn - md * md;
            const double inv_sig = var > 1e-12 * (e2 / nn) ? 1.0 / sqrt(var) : 1.0;  // a constant column is only centred
            denom = (float)(1.0 / inv_sig);
            pat_scale = (float)(-md * inv_sig);  // (z1 - mean) / sd
        } else {
            nsq = fmaf((float)n_one, ln2 * ln2, nsq);
            denom = fmaxf(sqrtf(nsq), 1e-8f);
            pat_scale = ln2 / denom;
        }
        if constexpr (HALF2) {
#pragma unroll
            for (int a = 0; a < 4; ++a) acc[a] += __shfl_xor_sync(0xffffffffu, acc[a], 16);  // the two half-warps' shares
            const int l = lane & 15;
            if (lane < 16) {
                float* o = out + (size_t)j * K;
                if (2 * l < K) {
                    o[2 * l] = acc[0] / denom;
                    o[2 * l + 1] = acc[1] / denom;
                }
                if (2 * (l + 16) < K) {
                    o[2 * l + 32] = acc[2] / denom;
                    o[2 * l + 33] = acc[3] / denom;
                }
            }
        } else {
#pragma unroll
            for (int a = 0; a < NACC; ++a) {
                const int k = lane + 32 * a;
                if (k < K) out[(size_t)j * K + k] = acc[a] / denom;
            }
        }
        if (lane == 0) {
            scale[j] = pat_scale;
            if (exc) exc_cnt[j] = qtail;
        }
        __syncwarp();
    }
}

// ---- K1a: the tcgen05 kernel -----------------------------------------------------------------
__global__ void __launch_bounds__(THREADS, 1) k_project_umma(const uint32_t* __restrict__ bm_global, uint64_t ncols, uint64_t D,
                                                             const int8_t* __restrict__ bq, int K, int NB, uint32_t nstages,
                                                             const float* __restrict__ scale,
                                                             const unsigned int* __restrict__ colmax_bits, float* __restrict__ out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    // carve: [B ring][bitmap x NBM][barriers][tmem base]
    const uint32_t stage_bytes = (uint32_t)NB * 32u * (GS / 32);
    uint8_t* smem_b = smem;
    uint32_t* bitmap = reinterpret_cast<uint32_t*>(smem + (size_t)NBST * stage_bytes);
    Barriers* bars = reinterpret_cast<Barriers*>(bitmap + NBM * CELLS * BM_STRIDE);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 1);
    __shared__ double col_inv[64];  // 1 / (A scale 128 * 2^20 * 2^e) per basis column: an exact power of two
    if (threadIdx.x < 64) col_inv[threadIdx.x] = threadIdx.x < K ? ldexp(1.0, -(27 + basis_col_exp(colmax_bits[threadIdx.x], nullptr))) : 0.0;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t nchunks = (uint32_t)((D + GC - 1) / GC);
    const uint64_t nsuper = (ncols + CELLS - 1) / CELLS;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NBST; ++s) {
            mbar_init(&bars->b_full[s], 1);
            mbar_init(&bars->b_empty[s], NT);
        }
        for (int t = 0; t < NT; ++t) {
            for (int s = 0; s < NAST; ++s) {
                mbar_init(&bars->a_full[t][s], 4);
                mbar_init(&bars->a_empty[t][s], 1);
            }
            mbar_init(&bars->acc_full[t], 1);
            mbar_init(&bars->acc_empty[t], 4);
        }
        for (int b = 0; b < NBM; ++b) {
            mbar_init(&bars->bm_full[b], 1);
            mbar_init(&bars->bm_empty[b], N_EXP_WARPS);
        }
        fence_barrier_init();
    }
    if (warp == WARP_MMA) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = *tmem_slot;
    const uint32_t a_col0 = (uint32_t)NT * (uint32_t)NB;  // A ring starts after the accumulators

    if (warp < N_EXP_WARPS) {
        // ===== expanders / epilogue: one thread per cell row =====
        const int t = warp >> 2;                        // tile of this warp
        const int row = (warp & 3) * 32 + lane;         // TMEM lane == row inside the tile
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        const uint32_t acc_addr = tbase + lane_base + (uint32_t)t * NB;
        uint32_t a_it = 0, chunk_it = 0, super_it = 0;
        for (uint64_t sup = blockIdx.x; sup < nsuper; sup += gridDim.x, ++super_it) {
            uint32_t stage = 0;
            for (uint32_t c = 0; c < nchunks; ++c, ++chunk_it) {
                const uint32_t buf = chunk_it % NBM;
                mbar_wait(&bars->bm_full[buf], (chunk_it / NBM) & 1);
                const uint32_t* my = bitmap + (size_t)buf * CELLS * BM_STRIDE + (size_t)(t * TILE_M + row) * BM_STRIDE;
                const uint32_t st_end = min(nstages, (c + 1) * (GC / GS));
                for (uint32_t ls = 0; stage < st_end; ++stage, ++ls, ++a_it) {
                    const uint32_t slot = a_it % NAST;
                    mbar_wait(&bars->a_empty[t][slot], ((a_it / NAST) & 1) ^ 1);
                    tc_fence_after();
                    const uint32_t a_addr = tbase + lane_base + a_col0 + (uint32_t)(t * NAST + slot) * A_COLS;
#pragma unroll
                    for (int hh = 0; hh < GS / 64; ++hh) {
                        const uint2 w = *reinterpret_cast<const uint2*>(my + (GS / 32) * ls + 2 * hh);
                        uint32_t r[16];
#pragma unroll
                        for (int b = 0; b < 8; ++b) {
                            r[b] = (w.x << (7 - b)) & 0x80808080u;
                            r[8 + b] = (w.y << (7 - b)) & 0x80808080u;
                        }
                        tmem_st_x16(a_addr + 16u * hh, r);
                    }
                    tmem_wait_st();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bars->a_full[t][slot]);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars->bm_empty[buf]);
            }
            // ---- epilogue for this tile ----
            // the correction row and the scale are fetched while the last MMAs still run
            const uint64_t cell = sup * CELLS + (uint64_t)t * TILE_M + row;
            const bool live = cell < ncols;
            float* orow = out + (size_t)cell * K;
            // rows are 8-byte aligned when K is even (and `out` is): 64-bit loads / stores halve the scattered transactions
            const bool pair_io = (K % 2 == 0) && ((reinterpret_cast<uintptr_t>(out) & 7) == 0);
            float corr[64];
            float sc = 0.0f;
            if (live) {
                sc = __ldg(scale + cell);
                if (pair_io) {
#pragma unroll
                    for (int k = 0; k < 64; k += 2) {
                        float2 c2 = make_float2(0.f, 0.f);
                        if (k < K) c2 = __ldcs(reinterpret_cast<const float2*>(orow + k));
                        corr[k] = c2.x;
                        corr[k + 1] = c2.y;
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < 64; ++k) corr[k] = (k < K) ? __ldcs(orow + k) : 0.0f;
                }
            }
            mbar_wait(&bars->acc_full[t], super_it & 1);
            tc_fence_after();
#pragma unroll
            for (int kb = 0; kb < 64; kb += 16) {
                if (kb < K) {
                    uint32_t d0[16], d1[16], d2[16];
                    tmem_ld_x16(acc_addr + kb, d0);
                    tmem_ld_x16(acc_addr + K + kb, d1);
                    tmem_ld_x16(acc_addr + 2 * K + kb, d2);
                    tmem_wait_ld();
                    if (live) {
                        float res[16];
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            // digits carry the A scale of 128: total = 128 * sum(q_i), q in 2^-(20 + e) units of column kb + i
                            const long long tot = ((long long)(int)d2[i] << 16) + ((long long)(int)d1[i] << 8) + (long long)(int)d0[i];
                            const float sv = (float)((double)tot * col_inv[kb + i]);
                            res[i] = fmaf(sv, sc, corr[kb + i]);
                        }
                        if (pair_io) {
#pragma unroll
                            for (int i = 0; i < 16; i += 2)
                                if (kb + i < K) __stcs(reinterpret_cast<float2*>(orow + kb + i), make_float2(res[i], res[i + 1]));
                        } else {
#pragma unroll
                            for (int i = 0; i < 16; ++i)
                                if (kb + i < K) __stcs(orow + kb + i, res[i]);
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars->acc_empty[t]);
        }
    } else if (warp >= WARP_MMA && warp < WARP_MMA + NT) {
        // ===== MMA issuers: one elected thread per cell tile =====
        const int t = warp - WARP_MMA;
        if (elect_one()) {
            const uint32_t idesc = make_idesc(CFMT_S32, FMT_U8, FMT_S8, TILE_M, (uint32_t)NB);
            uint32_t b_it = 0, a_it = 0, super_it = 0;
            for (uint64_t sup = blockIdx.x; sup < nsuper; sup += gridDim.x, ++super_it) {
                mbar_wait(&bars->acc_empty[t], (super_it & 1) ^ 1);
                tc_fence_after();
                for (uint32_t stage = 0; stage < nstages; ++stage, ++b_it, ++a_it) {
                    const uint32_t bs = b_it % NBST, as = a_it % NAST;
                    mbar_wait(&bars->b_full[bs], (b_it / NBST) & 1);
                    const uint32_t b_addr = smem_u32(smem_b + (size_t)bs * stage_bytes);
                    mbar_wait(&bars->a_full[t][as], (a_it / NAST) & 1);
                    tc_fence_after();
#pragma unroll
                    for (int j = 0; j < GS / 32; ++j) {
                        const uint64_t db = make_smem_desc(b_addr + (uint32_t)j * NB * 32u, 128, 256);
                        mma_i8_ts(tbase + (uint32_t)t * NB, tbase + a_col0 + (uint32_t)(t * NAST + as) * A_COLS + 8u * j, db, idesc,
                                  stage > 0 || j > 0);
                    }
                    tc_commit(&bars->a_empty[t][as]);
                    tc_commit(&bars->b_empty[bs]);  // NT arrivals free the B stage
                }
                tc_commit(&bars->acc_full[t]);
            }
        }
    } else if (warp == WARP_LOAD_B) {
        // ===== B-operand loader: one bulk copy per stage =====
        if (elect_one()) {
            uint32_t b_it = 0;
            for (uint64_t sup = blockIdx.x; sup < nsuper; sup += gridDim.x) {
                for (uint32_t stage = 0; stage < nstages; ++stage, ++b_it) {
                    const uint32_t bs = b_it % NBST;
                    mbar_wait(&bars->b_empty[bs], ((b_it / NBST) & 1) ^ 1);
                    mbar_arrive_expect_tx(&bars->b_full[bs], stage_bytes);
                    bulk_g2s(smem_b + (size_t)bs * stage_bytes, bq + (size_t)stage * stage_bytes, stage_bytes, &bars->b_full[bs]);
                }
            }
        }
    } else if (warp == WARP_LOAD_BM) {
        // ===== bitmap loader: one bulk copy per (supertile, chunk) =====
        if (elect_one()) {
            uint32_t chunk_it = 0;
            for (uint64_t sup = blockIdx.x; sup < nsuper; sup += gridDim.x) {
                for (uint32_t c = 0; c < nchunks; ++c, ++chunk_it) {
                    const uint32_t buf = chunk_it % NBM;
                    mbar_wait(&bars->bm_empty[buf], ((chunk_it / NBM) & 1) ^ 1);
                    mbar_arrive_expect_tx(&bars->bm_full[buf], BM_CHUNK_BYTES);
                    bulk_g2s(bitmap + (size_t)buf * CELLS * BM_STRIDE,
                             bm_global + (sup * nchunks + c) * (uint64_t)(CELLS * BM_STRIDE), BM_CHUNK_BYTES, &bars->bm_full[buf]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == WARP_MMA) tmem_dealloc(tbase, 512);
}


// ---- K1 fused (LG_K1_FUSED=1, NOT the default): CSC stream -> bitmap chunks in shared memory -> tcgen05 -----------------
// The two-kernel form above writes the pattern bitmap to a 4.4 GB scratch and reads it back (1.83x the algorithmic
// traffic).  Here sixteen PRODUCER warps take the place of the bitmap loader: rows are ascending inside a column, so a
// cell's share of the 2048-gene chunk the tensor pipe needs next is ONE contiguous piece of its stream, found by a per-cell
// cursor that simply carries over from the previous chunk (no split table).  A producer warp owns 16 cells of the 256-cell
// supertile; per chunk and cell it fetches one 128-entry window at the cursor before the piece's length is known (four
// windows in flight per warp; what lies beyond the chunk is read again, from L2, one chunk later), sets the bits with
// shared-memory reductions and appends the counts != 1 to the cell's exception list.  k_project_finalize then turns
// the lists into the correction and the norm (the arithmetic of k_project_prep's drain, batch for batch) and combines them
// with the tensor sum the epilogue left in `out`: the result is BIT-IDENTICAL to the two-kernel form, which is what
// tests/test_gpu_parity.py::test_project_fused_form_is_bit_identical holds it to — an independent second scan of the
// same stream.
// Measured (262 144 cells x 30 000 genes, 369 M nnz): 2.65 + 0.47 ms against 1.09 + 0.69 ms.  The
// fused kernel issues 7 700 warp instructions per cell at IPC 2.9 — it is ISSUE-bound, the tensor pipe 12 % active, the
// expanders 75 % of their time waiting for bitmap chunks: the scan is instruction work (3 900 per cell here, 3 300 in
// k_project_prep) that the same four schedulers must issue either way, and inside this kernel only 16 of the SM's warps
// (the others feed the tensor pipe) share it.  Kept as the cross-check and as the record of the experiment.
constexpr int F_PW = 16;                               // producer warps
constexpr int F_CPW = CELLS / F_PW;                    // cells of a supertile per producer warp
constexpr int F_WARP_PROD = 12;                        // warps 0-7 expanders / epilogue, 8-9 MMA, 10 basis loader, 11 idle
constexpr int F_THREADS = (F_WARP_PROD + F_PW) * 32;   // 896
constexpr int F_CAPX = 128;                            // exception list entries per cell; a longer list: the cell is re-scanned

struct FBarriers {
    uint64_t b_full[NBST], b_empty[NBST];
    uint64_t a_full[NT][NAST], a_empty[NT][NAST];
    uint64_t bm_full[NBM], bm_empty[NBM];
    uint64_t acc_full[NT], acc_empty[NT];
};

// one batch of <= 32 exceptions (lane e holds entry e; lanes >= cnt carry gene 0, weight 0) into the HALF2 accumulators:
// the arithmetic of k_project_prep's drain, operation for operation
__device__ __forceinline__ void f_drain(uint32_t g, float val, bool on, uint32_t cnt, const float* __restrict__ basis_kd, int K,
                                        const float* lut_x, float ln2, int lane, float (&acc)[4], float& nsq) {
    float w = 0.0f;
    if (on) {
        const int vi = (int)val;
        const float x = (val == (float)vi && vi >= 0 && vi < PREP_LUT) ? lut_x[vi] : log1pf(val);
        nsq = fmaf(x, x, nsq);
        w = x - ln2;
    }
    const int half = lane >> 4, l = lane & 15;
    const bool hi_on = 2 * (l + 16) < K;
    for (uint32_t e0 = 0; e0 < cnt; e0 += 8) {
        float2 b0[4], b1[4];
        float ws[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int src = (int)e0 + 2 * e + half;
            const uint32_t ge = __shfl_sync(0xffffffffu, g, src);
            ws[e] = __shfl_sync(0xffffffffu, w, src);
            const float2* brow = reinterpret_cast<const float2*>(basis_kd + (size_t)ge * K);
            b0[e] = (2 * l < K) ? __ldg(brow + l) : make_float2(0.f, 0.f);
            b1[e] = hi_on ? __ldg(brow + l + 16) : make_float2(0.f, 0.f);
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            acc[0] = fmaf(ws[e], b0[e].x, acc[0]);
            acc[1] = fmaf(ws[e], b0[e].y, acc[1]);
            acc[2] = fmaf(ws[e], b1[e].x, acc[2]);
            acc[3] = fmaf(ws[e], b1[e].y, acc[3]);
        }
    }
}

// second half of the fused form: per cell (one warp), the exception list -> correction and norm in full batches of 32,
// exactly as k_project_prep drains its queue, then the combine with the tensor sum the fused kernel's epilogue left in
// `out`.  A kernel of its own because the gathers need an occupancy that the 16 producer warps of the fused kernel cannot
// give (~110 basis rows in flight per SM at L2 latency).
__global__ void __launch_bounds__(256) k_project_finalize(const uint64_t* __restrict__ indptr, const uint32_t* __restrict__ indices,
                                                          const float* __restrict__ values, uint64_t ncols,
                                                          const float* __restrict__ basis_kd, int K, const uint2* __restrict__ xl,
                                                          const uint32_t* __restrict__ xn, const int* __restrict__ bad_flag,
                                                          float* out) {
    if (*bad_flag) return;
    __shared__ float lut_x[PREP_LUT];
    if (threadIdx.x < PREP_LUT) lut_x[threadIdx.x] = log1pf((float)threadIdx.x);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const float ln2 = lut_x[1];
    const uint64_t nwarps = (uint64_t)gridDim.x * 8;
    for (uint64_t cell = (uint64_t)blockIdx.x * 8 + (threadIdx.x >> 5); cell < ncols; cell += nwarps) {
        const uint32_t cnt = xn[cell];
        const uint64_t lo = indptr[cell];
        const uint32_t n = (uint32_t)(indptr[cell + 1] - lo);
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        float nsq = 0.0f;
        if (cnt <= (uint32_t)F_CAPX) {
            const uint2* xlist = xl + cell * F_CAPX;
            uint2 e = make_uint2(0u, 0u);
            if ((uint32_t)lane < cnt) e = __ldcs(xlist + lane);
            for (uint32_t q0 = 0; q0 < cnt; q0 += 32) {
                const uint32_t c = min(32u, cnt - q0);
                uint2 ne = make_uint2(0u, 0u);
                if (q0 + 32 + lane < cnt) ne = __ldcs(xlist + q0 + 32 + lane);  // the next batch travels while this one is folded
                f_drain(e.x, __uint_as_float(e.y), (uint32_t)lane < c, c, basis_kd, K, lut_x, ln2, lane, acc, nsq);
                e = ne;
            }
        } else {
            // too many exceptions for the list (non-count data): stream the cell again and batch them 32 at a time
            uint32_t pg = 0, pc = 0;
            float pv = 1.0f;
            for (uint32_t t0 = 0; t0 < n; t0 += 32) {
                const bool live = t0 + lane < n;
                const uint32_t gi = live ? __ldg(indices + lo + t0 + lane) : 0u;
                const float vv = live ? __ldg(values + lo + t0 + lane) : 1.0f;
                unsigned m = __ballot_sync(0xffffffffu, vv != 1.0f);
                while (m) {
                    const int want = lane - (int)pc;  // this lane takes the want-th pending set bit of m
                    const unsigned src = want >= 0 ? __fns(m, 0, want + 1) : 0xffffffffu;
                    const uint32_t g_in = __shfl_sync(0xffffffffu, gi, src & 31u);
                    const float v_in = __shfl_sync(0xffffffffu, vv, src & 31u);
                    if (src != 0xffffffffu) {
                        pg = g_in;
                        pv = v_in;
                    }
                    const uint32_t ntake = min((uint32_t)__popc(m), 32u - pc);
                    const unsigned last = __fns(m, 0, (int)ntake);  // position of the last bit taken
                    m &= ~((2u << last) - 1u);
                    pc += ntake;
                    if (pc == 32) {
                        f_drain(pg, pv, true, 32, basis_kd, K, lut_x, ln2, lane, acc, nsq);
                        pc = 0;
                    }
                }
            }
            if (pc) {
                const bool on = (uint32_t)lane < pc;
                f_drain(on ? pg : 0u, pv, on, pc, basis_kd, K, lut_x, ln2, lane, acc, nsq);
            }
        }
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) nsq += __shfl_xor_sync(0xffffffffu, nsq, off);
        nsq = fmaf((float)(n - cnt), ln2 * ln2, nsq);
        const float denom = fmaxf(sqrtf(nsq), 1e-8f);
        const float pat_scale = ln2 / denom;
#pragma unroll
        for (int a = 0; a < 4; ++a) acc[a] += __shfl_xor_sync(0xffffffffu, acc[a], 16);
        if (lane < 16) {
            float* o = out + (size_t)cell * K;
            if (2 * lane < K) {
                const float2 sv = __ldcs(reinterpret_cast<const float2*>(o + 2 * lane));
                __stcs(reinterpret_cast<float2*>(o + 2 * lane), make_float2(fmaf(sv.x, pat_scale, acc[0] / denom), fmaf(sv.y, pat_scale, acc[1] / denom)));
            }
            if (2 * (lane + 16) < K) {
                const float2 sv = __ldcs(reinterpret_cast<const float2*>(o + 2 * lane + 32));
                __stcs(reinterpret_cast<float2*>(o + 2 * lane + 32), make_float2(fmaf(sv.x, pat_scale, acc[2] / denom), fmaf(sv.y, pat_scale, acc[3] / denom)));
            }
        }
    }
}

__global__ void __launch_bounds__(F_THREADS, 1)
    k_project_fused(const uint64_t* __restrict__ indptr, const uint32_t* __restrict__ indices, const float* __restrict__ values,
                    uint64_t ncols, uint64_t nnz, uint64_t D, const float* __restrict__ basis_kd, const int8_t* __restrict__ bq, int K,
                    int NB, uint32_t nstages, const unsigned int* __restrict__ colmax_bits, const int* __restrict__ bad_flag,
                    uint2* __restrict__ xl, uint32_t* __restrict__ xn, float* out) {
    if (*bad_flag) return;  // a non-finite basis column: the host falls back to the CUDA-core kernel
    extern __shared__ __align__(1024) uint8_t smem[];
    // carve: [B ring][bitmap x NBM][barriers][tmem base]
    const uint32_t stage_bytes = (uint32_t)NB * 32u * (GS / 32);
    uint8_t* smem_b = smem;
    uint32_t* bitmap = reinterpret_cast<uint32_t*>(smem + (size_t)NBST * stage_bytes);
    FBarriers* bars = reinterpret_cast<FBarriers*>(bitmap + NBM * CELLS * BM_STRIDE);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 1);
    __shared__ double col_inv[64];
    __shared__ float lut_x[PREP_LUT];
    if (threadIdx.x < 64) col_inv[threadIdx.x] = (int)threadIdx.x < K ? ldexp(1.0, -(27 + basis_col_exp(colmax_bits[threadIdx.x], nullptr))) : 0.0;
    if (threadIdx.x >= 64 && threadIdx.x < 64 + PREP_LUT) lut_x[threadIdx.x - 64] = log1pf((float)(threadIdx.x - 64));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t nchunks = (uint32_t)((D + GC - 1) / GC);
    const uint64_t nsuper = (ncols + CELLS - 1) / CELLS;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NBST; ++s) {
            mbar_init(&bars->b_full[s], 1);
            mbar_init(&bars->b_empty[s], NT);
        }
        for (int t = 0; t < NT; ++t) {
            for (int s = 0; s < NAST; ++s) {
                mbar_init(&bars->a_full[t][s], 4);
                mbar_init(&bars->a_empty[t][s], 1);
            }
            mbar_init(&bars->acc_full[t], 1);
            mbar_init(&bars->acc_empty[t], 4);
        }
        for (int b = 0; b < NBM; ++b) {
            mbar_init(&bars->bm_full[b], F_PW);
            mbar_init(&bars->bm_empty[b], N_EXP_WARPS);
        }
        fence_barrier_init();
    }
    if (warp == WARP_MMA) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = *tmem_slot;
    const uint32_t a_col0 = (uint32_t)NT * (uint32_t)NB;

    // register file: the kernel starts at 72 per thread (896 threads); the tensor-side warpgroups hand registers to the
    // producers, whose windows in flight are what hides the HBM latency of the stream.  Each setmaxnreg sits at the top
    // of its role's branch (ptxas takes the limit of a join to be the smaller one).
    if (warp < N_EXP_WARPS) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
        // ===== expanders / epilogue: one thread per cell row =====
        const int t = warp >> 2;
        const int row = (warp & 3) * 32 + lane;
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        const uint32_t acc_addr = tbase + lane_base + (uint32_t)t * NB;
        uint32_t a_it = 0, chunk_it = 0, super_it = 0;
        for (uint64_t sup = blockIdx.x; sup < nsuper; sup += gridDim.x, ++super_it) {
            uint32_t stage = 0;
            for (uint32_t c = 0; c < nchunks; ++c, ++chunk_it) {
                const uint32_t buf = chunk_it % NBM;
                mbar_wait_sleep(&bars->bm_full[buf], (chunk_it / NBM) & 1);
                const uint32_t* my = bitmap + (size_t)buf * CELLS * BM_STRIDE + (size_t)(t * TILE_M + row) * BM_STRIDE;
                const uint32_t st_end = min(nstages, (c + 1) * (GC / GS));
                for (uint32_t ls = 0; stage < st_end; ++stage, ++ls, ++a_it) {
                    const uint32_t slot = a_it % NAST;
                    mbar_wait(&bars->a_empty[t][slot], ((a_it / NAST) & 1) ^ 1);
                    tc_fence_after();
                    const uint32_t a_addr = tbase + lane_base + a_col0 + (uint32_t)(t * NAST + slot) * A_COLS;
#pragma unroll
                    for (int hh = 0; hh < GS / 64; ++hh) {
                        const uint2 w = *reinterpret_cast<const uint2*>(my + (GS / 32) * ls + 2 * hh);
                        uint32_t r[16];
#pragma unroll
                        for (int b = 0; b < 8; ++b) {
                            r[b] = (w.x << (7 - b)) & 0x80808080u;
                            r[8 + b] = (w.y << (7 - b)) & 0x80808080u;
                        }
                        tmem_st_x16(a_addr + 16u * hh, r);
                    }
                    tmem_wait_st();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bars->a_full[t][slot]);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars->bm_empty[buf]);
            }
            // ---- epilogue: the tensor sum of the pattern, unscaled; the cell's producer warp combines it later ----
            const uint64_t cell = sup * CELLS + (uint64_t)t * TILE_M + row;
            const bool live = cell < ncols;
            float* orow = out + (size_t)cell * K;
            mbar_wait(&bars->acc_full[t], super_it & 1);
            tc_fence_after();
#pragma unroll
            for (int kb = 0; kb < 64; kb += 8) {
                if (kb < K) {
                    uint32_t d0[8], d1[8], d2[8];
                    tmem_ld_x8(acc_addr + kb, d0);
                    tmem_ld_x8(acc_addr + K + kb, d1);
                    tmem_ld_x8(acc_addr + 2 * K + kb, d2);
                    tmem_wait_ld();
                    if (live) {
                        float res[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const long long tot = ((long long)(int)d2[i] << 16) + ((long long)(int)d1[i] << 8) + (long long)(int)d0[i];
                            res[i] = (float)((double)tot * col_inv[kb + i]);
                        }
#pragma unroll
                        for (int i = 0; i < 8; i += 2)
                            if (kb + i < K) __stcg(reinterpret_cast<float2*>(orow + kb + i), make_float2(res[i], res[i + 1]));
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars->acc_empty[t]);
        }
    } else if (warp < F_WARP_PROD) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");
      if (warp < WARP_MMA + NT) {
        // ===== MMA issuers: one elected thread per cell tile =====
        const int t = warp - WARP_MMA;
        if (elect_one()) {
            const uint32_t idesc = make_idesc(CFMT_S32, FMT_U8, FMT_S8, TILE_M, (uint32_t)NB);
            uint32_t b_it = 0, a_it = 0, super_it = 0;
            for (uint64_t sup = blockIdx.x; sup < nsuper; sup += gridDim.x, ++super_it) {
                mbar_wait(&bars->acc_empty[t], (super_it & 1) ^ 1);
                tc_fence_after();
                for (uint32_t stage = 0; stage < nstages; ++stage, ++b_it, ++a_it) {
                    const uint32_t bs = b_it % NBST, as = a_it % NAST;
                    mbar_wait(&bars->b_full[bs], (b_it / NBST) & 1);
                    const uint32_t b_addr = smem_u32(smem_b + (size_t)bs * stage_bytes);
                    mbar_wait_sleep(&bars->a_full[t][as], (a_it / NAST) & 1);
                    tc_fence_after();
#pragma unroll
                    for (int j = 0; j < GS / 32; ++j) {
                        const uint64_t db = make_smem_desc(b_addr + (uint32_t)j * NB * 32u, 128, 256);
                        mma_i8_ts(tbase + (uint32_t)t * NB, tbase + a_col0 + (uint32_t)(t * NAST + as) * A_COLS + 8u * j, db, idesc,
                                  stage > 0 || j > 0);
                    }
                    tc_commit(&bars->a_empty[t][as]);
                    tc_commit(&bars->b_empty[bs]);
                }
                tc_commit(&bars->acc_full[t]);
            }
        }
      } else if (warp == WARP_LOAD_B) {
        // ===== B-operand loader: one bulk copy per stage =====
        if (elect_one()) {
            uint32_t b_it = 0;
            for (uint64_t sup = blockIdx.x; sup < nsuper; sup += gridDim.x) {
                for (uint32_t stage = 0; stage < nstages; ++stage, ++b_it) {
                    const uint32_t bs = b_it % NBST;
                    mbar_wait_sleep(&bars->b_empty[bs], ((b_it / NBST) & 1) ^ 1);
                    mbar_arrive_expect_tx(&bars->b_full[bs], stage_bytes);
                    bulk_g2s(smem_b + (size_t)bs * stage_bytes, bq + (size_t)stage * stage_bytes, stage_bytes, &bars->b_full[bs]);
                }
            }
        }
      }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 88;");
        // ===== producers: CSC stream -> bitmap chunk + per-cell exception lists =====
        const int pw = warp - F_WARP_PROD;
        const int row0 = pw * F_CPW;
        const unsigned lt_mask = (1u << lane) - 1u;
        // lanes 0..15 hold the state of the warp's 16 cells (lanes 16..31 mirror them)
        uint32_t cur = 0, end = 0, nexc = 0;  // cursor / end of the stream (relative to the supertile's first entry), exceptions so far
        uint32_t chunk_it = 0;
        const uint32_t bm_s = smem_u32(bitmap);
        for (uint64_t sup = blockIdx.x; sup < nsuper; sup += gridDim.x) {
            const uint64_t first = sup * CELLS;
            const uint64_t base = indptr[first];  // first entry of the supertile (warp-uniform)
            const uint64_t mycell = first + (uint64_t)(row0 + (lane & 15));
            {
                uint64_t lo = base, hi = base;
                if (mycell < ncols) {
                    lo = indptr[mycell];
                    hi = indptr[mycell + 1];
                }
                cur = (uint32_t)(lo - base);
                end = (uint32_t)(hi - base);
                nexc = 0;
            }
            const uint32_t* ibase = indices + base;
            const float* vbase = values + base;
            const int lim = (int)(nnz - base < 0x7fffffffull ? nnz - base : 0x7fffffffull);  // entries left in the arrays
            const int mis0 = (int)(base & 3ull);
            uint2* xl_sup = xl + (first + (uint64_t)row0) * F_CAPX;
            for (uint32_t c = 0; c < nchunks; ++c, ++chunk_it) {
                const uint32_t buf = chunk_it % NBM;
                mbar_wait_sleep(&bars->bm_empty[buf], ((chunk_it / NBM) & 1) ^ 1);
                {
                    uint32_t* rows = bitmap + (size_t)buf * CELLS * BM_STRIDE + (size_t)row0 * BM_STRIDE;
                    for (int q = lane; q < F_CPW * BM_STRIDE / 4; q += 32) reinterpret_cast<uint4*>(rows)[q] = make_uint4(0, 0, 0, 0);
                }
                __syncwarp();
                // shared-window byte address of the warp's first row, moved back by the chunk's first word
                const uint32_t rows_s = bm_s + ((uint32_t)buf * CELLS * BM_STRIDE + (uint32_t)row0 * BM_STRIDE - c * (GC / 32)) * 4u;
                const uint32_t g1 = (c + 1 == nchunks) ? 0xffffffffu : (c + 1) * (uint32_t)GC;

                // A piece is fetched as ONE window of 32 x 128-bit groups (128 entries) from the 16-byte boundary at or below
                // the cursor, before its length is known; entries outside [cursor, end of the cell) are blanked, genes beyond
                // the chunk stay for the next chunk (that part of the window is read again, from L2).
                struct Win {
                    uint4 ix;
                    float4 v;
                    int a0;  // window start relative to the supertile's first entry (-3 .. -1 possible for its first cell)
                };
                auto fetch = [&](int a0, uint32_t en, Win& w) {
                    w.a0 = a0;
                    w.ix = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
                    w.v = make_float4(1.f, 1.f, 1.f, 1.f);
                    const int p0 = a0 + 4 * lane;
                    if (p0 < (int)en) {
                        if (p0 + 4 <= lim) {
                            w.ix = __ldg(reinterpret_cast<const uint4*>(ibase + p0));
                            w.v = __ldg(reinterpret_cast<const float4*>(vbase + p0));
                        } else {  // the last, partial group of the arrays
                            uint32_t gi[4] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};
                            float gv[4] = {1.f, 1.f, 1.f, 1.f};
                            for (int j = 0; j < 4; ++j)
                                if (p0 + j < lim) {
                                    gi[j] = __ldg(ibase + p0 + j);
                                    gv[j] = __ldg(vbase + p0 + j);
                                }
                            w.ix = make_uint4(gi[0], gi[1], gi[2], gi[3]);
                            w.v = make_float4(gv[0], gv[1], gv[2], gv[3]);
                        }
                    }
                };
                auto issue = [&](int i, Win& w) {
                    const uint32_t cu = __shfl_sync(0xffffffffu, cur, i), en = __shfl_sync(0xffffffffu, end, i);
                    fetch((int)cu - (int)((cu + mis0) & 3u), en, w);
                };
                auto consume = [&](int i, Win w) {
                    const uint32_t rowa = rows_s + (uint32_t)i * (BM_STRIDE * 4);
                    uint2* xlist = xl_sup + (size_t)i * F_CAPX;
                    const int cu = (int)__shfl_sync(0xffffffffu, cur, i), en = (int)__shfl_sync(0xffffffffu, end, i);
                    uint32_t xc = __shfl_sync(0xffffffffu, nexc, i);
                    uint32_t taken = 0;
                    for (;;) {
                        const int p0 = w.a0 + 4 * lane;
                        uint32_t gx[4] = {w.ix.x, w.ix.y, w.ix.z, w.ix.w};
                        const float vx[4] = {w.v.x, w.v.y, w.v.z, w.v.w};
                        if (p0 < cu) {  // the group the cursor sits in: its leading entries belong to the previous piece
#pragma unroll
                            for (int j = 0; j < 3; ++j)
                                if (p0 + j < cu) gx[j] = 0xffffffffu;
                        }
                        if (p0 + 4 > en) {  // the group the cell ends in
#pragma unroll
                            for (int j = 1; j < 4; ++j)
                                if (p0 + j >= en) gx[j] = 0xffffffffu;
                        }
                        bool in[4];
                        uint32_t cnt = 0, nx = 0, xg = 0;
                        float xv = 1.0f;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            in[j] = gx[j] < g1;
                            if (in[j]) {
                                asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(rowa + ((gx[j] >> 5) << 2)), "r"(1u << (gx[j] & 31)) : "memory");
                                ++cnt;
                                if (vx[j] != 1.0f) {
                                    ++nx;
                                    xg = gx[j];
                                    xv = vx[j];
                                }
                            }
                        }
                        // exceptions join the cell's list in ascending position: lanes in order, entries in order inside a lane
                        const unsigned b_any = __ballot_sync(0xffffffffu, nx != 0);
                        if (b_any) {
                            const unsigned b_multi = __ballot_sync(0xffffffffu, nx > 1);
                            if (!b_multi) {  // at most one per lane (the usual case): xg / xv hold it
                                const uint32_t slot = xc + __popc(b_any & lt_mask);
                                if (nx && slot < (uint32_t)F_CAPX) __stcg(xlist + slot, make_uint2(xg, __float_as_uint(xv)));
                                xc += __popc(b_any);
                            } else {
                                const unsigned b0 = __ballot_sync(0xffffffffu, nx & 1u), b1 = __ballot_sync(0xffffffffu, nx & 2u),
                                               b2 = __ballot_sync(0xffffffffu, nx & 4u);
                                uint32_t slot = xc + __popc(b0 & lt_mask) + 2 * __popc(b1 & lt_mask) + 4 * __popc(b2 & lt_mask);
#pragma unroll
                                for (int j = 0; j < 4; ++j)
                                    if (in[j] && vx[j] != 1.0f) {
                                        if (slot < (uint32_t)F_CAPX) __stcg(xlist + slot, make_uint2(gx[j], __float_as_uint(vx[j])));
                                        ++slot;
                                    }
                                xc += __popc(b0) + 2 * __popc(b1) + 4 * __popc(b2);
                            }
                        }
                        taken += __reduce_add_sync(0xffffffffu, cnt);
                        // the piece goes on beyond this window iff its last entry is still inside the chunk
                        if (!(__ballot_sync(0xffffffffu, in[3]) >> 31)) break;
                        fetch(w.a0 + 128, (uint32_t)en, w);
                    }
                    if ((lane & 15) == i) {
                        cur += taken;
                        nexc = xc;
                    }
                };

                Win wA, wB, wC, wD;
                issue(0, wA);
                issue(1, wB);
                issue(2, wC);
#pragma unroll 1
                for (int i = 0; i < F_CPW; i += 4) {
                    issue(i + 3, wD);
                    consume(i, wA);
                    if (i + 4 < F_CPW) issue(i + 4, wA);
                    consume(i + 1, wB);
                    if (i + 5 < F_CPW) issue(i + 5, wB);
                    consume(i + 2, wC);
                    if (i + 6 < F_CPW) issue(i + 6, wC);
                    consume(i + 3, wD);
                }
                // pull the lines that the chunk after the next one will start in into L2 (what the next chunk's windows
                // fetch lies behind this chunk's windows already): lanes 0-15 the index lines, 16-31 the value lines
                if (c + 2 < nchunks && cur + 128 < end) {
                    const char* pf = (lane < 16 ? reinterpret_cast<const char*>(ibase) : reinterpret_cast<const char*>(vbase)) + ((size_t)cur + 128) * 4;
#pragma unroll
                    for (int k = 0; k < 4; ++k) asm volatile("prefetch.global.L2 [%0];" ::"l"(pf + 128 * k));
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars->bm_full[buf]);
            }
            if (lane < F_CPW && mycell < ncols) xn[mycell] = nexc;  // the list's length (may exceed F_CAPX: that cell is re-scanned)
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == WARP_MMA) tmem_dealloc(tbase, 512);
}

}  // namespace

// returns LG_OK and sets *used = 1 when the tensor path ran, *used = 0 when the caller must fall back
int lg_project_raw_umma(lg_ctx* ctx, const lg_csc* m, const float* d_basis, int K, float* d_out, int* used, int mode, float csn) {
    *used = 0;
    const int NB = ((3 * K + 15) / 16) * 16;
    if (K < 1 || m->ncols == 0 || m->nrows == 0) return LG_OK;
    if (NT * NB + NT * NAST * A_COLS > 512 || m->nrows > 131072) {
        char b[200];
        snprintf(b, sizeof(b), "projection: K = %d, %llu genes is outside the tensor-core path (K <= 53, genes <= 131072): CUDA-core kernel", K,
                 (unsigned long long)m->nrows);
        lg_note_fallback(ctx, b);
        return LG_OK;
    }
    const uint64_t D = m->nrows;
    const uint64_t Dpad = ((D + GS - 1) / GS) * GS;
    const uint32_t nstages = (uint32_t)(Dpad / GS);
    LgStage st(ctx);
    int8_t* d_bq;
    int* d_flag;
    float* d_scale;
    LG_TRY(st.scratch((size_t)Dpad * NB, &d_bq));
    LG_TRY(st.scratch(1, &d_flag));
    LG_TRY(st.scratch((size_t)m->ncols, &d_scale));
    unsigned int* d_colmax;
    LG_TRY(st.scratch(64, &d_colmax));
    LG_CUDA(ctx, cudaMemsetAsync(d_flag, 0, sizeof(int), ctx->stream));
    LG_CUDA(ctx, cudaMemsetAsync(d_colmax, 0, 64 * sizeof(unsigned int), ctx->stream));
    {
        const uint64_t nb = D * (uint64_t)K;
        LG_LAUNCH(ctx, k_basis_colmax, (unsigned)((nb + 4095) / 4096), 256, 0, d_basis, D, K, d_colmax);
        const uint64_t tot = Dpad * (uint64_t)NB;
        LG_LAUNCH(ctx, k_quantize_basis, (unsigned)((tot + 255) / 256), 256, 0, d_basis, D, K, NB, Dpad, d_colmax, d_bq, d_flag);
    }
    // the flag travels home while K1b runs: the scan does not read the quantised basis, so the host round trip is hidden
    // behind it and only the tensor kernel waits for the verdict (the scan's work is wasted in the rare fall-back case)
    int* h_flag = static_cast<int*>(ctx->pinned);
    cudaEvent_t flag_ev;
    LG_CUDA(ctx, cudaEventCreateWithFlags(&flag_ev, cudaEventDisableTiming));
    LG_CUDA(ctx, cudaMemcpyAsync(h_flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    LG_CUDA(ctx, cudaEventRecord(flag_ev, ctx->stream));

    const char* tr = getenv("LG_K1_TRACE");
    const bool trace = tr && tr[0] == '1';
    cudaEvent_t ev[3];
    if (trace)
        for (int i = 0; i < 3; ++i) cudaEventCreate(&ev[i]);
    // the fused kernel (projection mode, even K, 8-byte aligned basis and output); LG_K1_FUSED=0 keeps the two-kernel form
    {
        const char* fz = getenv("LG_K1_FUSED");
        const bool fused = mode == 0 && fz && fz[0] == '1' && K % 2 == 0 && (((uintptr_t)d_basis | (uintptr_t)d_out) & 7) == 0 &&
                           (((uintptr_t)m->indices | (uintptr_t)m->values) & 15) == 0;
        if (fused) {
            const size_t stage_bytes = (size_t)NB * 32 * (GS / 32);
            const size_t smem = (size_t)NBST * stage_bytes + (size_t)NBM * BM_CHUNK_BYTES + sizeof(FBarriers) + 16;
            if (smem + 1024 > ctx->smem_optin) return lg_fail(ctx, LG_ERR_INTERNAL, "k_project_fused: shared memory budget exceeded");
            const uint64_t nsuper = (m->ncols + CELLS - 1) / CELLS;
            const unsigned grid = (unsigned)(nsuper < (uint64_t)ctx->num_sms ? nsuper : (uint64_t)ctx->num_sms);
            uint2* d_xl;
            uint32_t* d_xn;
            LG_TRY(st.scratch((size_t)m->ncols * F_CAPX, &d_xl));
            LG_TRY(st.scratch((size_t)m->ncols, &d_xn));
            LG_CUDA(ctx, cudaFuncSetAttribute(k_project_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            if (trace) cudaEventRecord(ev[0], ctx->stream);
            LG_LAUNCH(ctx, k_project_fused, grid, F_THREADS, smem, m->indptr, m->indices, m->values, m->ncols, m->nnz, D, d_basis, d_bq, K,
                      NB, nstages, d_colmax, d_flag, d_xl, d_xn, d_out);
            if (trace) cudaEventRecord(ev[1], ctx->stream);
            {
                uint64_t fb = (m->ncols + 7) / 8;
                const uint64_t cap = (uint64_t)ctx->num_sms * 8;
                if (fb > cap) fb = cap;
                LG_LAUNCH(ctx, k_project_finalize, (unsigned)fb, 256, 0, m->indptr, m->indices, m->values, m->ncols, d_basis, K, d_xl, d_xn,
                          d_flag, d_out);
            }
            if (trace) cudaEventRecord(ev[2], ctx->stream);
            const cudaError_t fe = cudaEventSynchronize(flag_ev);  // the quantiser's verdict: long home by now
            cudaEventDestroy(flag_ev);
            if (trace) {
                cudaEventSynchronize(ev[2]);
                float a = 0.f, b = 0.f;
                cudaEventElapsedTime(&a, ev[0], ev[1]);
                cudaEventElapsedTime(&b, ev[1], ev[2]);
                fprintf(stderr, "[lg_project] fused %.3f ms, finalize %.3f ms\n", a, b);
                for (int i = 0; i < 3; ++i) cudaEventDestroy(ev[i]);
            }
            LG_CUDA(ctx, fe);
            if (*h_flag) {  // the kernel returned at once; the caller falls back (and propagates the non-finite column)
                lg_note_fallback(ctx, "projection: a basis column is not finite or too large for the fixed-point grid: CUDA-core kernel");
                return LG_OK;
            }
            *used = 1;
            return LG_OK;
        }
    }
    // K1b first: bitmap + corr/norm in `out` + ln2/norm in `scale`; K1a then adds the tensor part
    const uint32_t nchunks = (uint32_t)((D + GC - 1) / GC);
    const uint64_t nsuper = (m->ncols + CELLS - 1) / CELLS;
    uint32_t* d_bm;
    // the hot path asks for the pattern to outlive this call (lg_pattern): it owns the buffers then
    lg_pattern* pat = (mode == 0) ? (ctx->pat ? ctx->pat : m->twin) : nullptr;
    if (pat && pat == m->twin) m->twin_ovf = -1;
    uint32_t *d_exc = nullptr, *d_exc_cnt = nullptr;
    int* d_exc_ovf = nullptr;
    if (pat) {
        pat->filled = false;
        pat->nchunks = nchunks;
        d_bm = pat->bm;
        d_exc = pat->exc;
        d_exc_cnt = pat->exc_cnt;
        d_exc_ovf = pat->ovf;
        LG_CUDA(ctx, cudaMemsetAsync(d_exc_ovf, 0, sizeof(int), ctx->stream));
    } else {
        LG_TRY(st.scratch((size_t)nsuper * nchunks * CELLS * BM_STRIDE, &d_bm));
    }
    if (m->ncols % CELLS)  // rows of the ragged last supertile that no cell writes must read as empty
        LG_CUDA(ctx, cudaMemsetAsync(d_bm + (nsuper - 1) * nchunks * (size_t)(CELLS * BM_STRIDE), 0,
                                     (size_t)nchunks * BM_CHUNK_BYTES, ctx->stream));
    {
        const size_t psmem = (size_t)PREP_WARPS * nchunks * (GC / 32) * 4 + (size_t)PREP_WARPS * 2 * PREP_Q * 4;
        int per_sm = (int)(ctx->smem_optin / (psmem + 1024));
        if (per_sm > 4) per_sm = 4;
        if (per_sm < 1) return lg_fail(ctx, LG_ERR_INTERNAL, "k_project_prep: shared memory budget exceeded");
        uint64_t blocks = (m->ncols + PREP_WARPS - 1) / PREP_WARPS;
        const uint64_t cap = (uint64_t)ctx->num_sms * per_sm;
        if (blocks > cap) blocks = cap;
        const int nacc = (K + 31) / 32;
        const bool half2 = (K % 2 == 0) && (((uintptr_t)d_basis & 7) == 0);
#define LG_PREP_LAUNCH_M(NA, H2, V, M)                                                                                                \
    do {                                                                                                                              \
        LG_CUDA(ctx, cudaFuncSetAttribute(k_project_prep<NA, H2, V, M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psmem));    \
        LG_LAUNCH(ctx, (k_project_prep<NA, H2, V, M>), (unsigned)blocks, PREP_WARPS * 32, psmem, m->indptr, m->indices, m->values,  \
                  m->ncols, d_basis, K, nchunks, d_bm, d_out, d_scale, csn, d_exc, d_exc_cnt, d_exc_ovf);                             \
    } while (0)
#define LG_PREP_LAUNCH(NA, H2, V)                 \
    do {                                          \
        if (mode == 1) LG_PREP_LAUNCH_M(NA, H2, V, 1); \
        else LG_PREP_LAUNCH_M(NA, H2, V, 0);      \
    } while (0)
        const bool vec = (((uintptr_t)m->indices | (uintptr_t)m->values) & 15) == 0;
        if (trace) cudaEventRecord(ev[0], ctx->stream);
        if (half2 && vec) LG_PREP_LAUNCH(2, true, true);
        else if (half2) LG_PREP_LAUNCH(2, true, false);
        else if (nacc == 1) LG_PREP_LAUNCH(1, false, false);
        else LG_PREP_LAUNCH(2, false, false);
#undef LG_PREP_LAUNCH
#undef LG_PREP_LAUNCH_M
    }
    const size_t stage_bytes = (size_t)NB * 32 * (GS / 32);
    const size_t smem = (size_t)NBST * stage_bytes + (size_t)NBM * BM_CHUNK_BYTES + sizeof(Barriers) + 16;
    if (smem > ctx->smem_optin) return lg_fail(ctx, LG_ERR_INTERNAL, "k_project_umma: shared memory budget exceeded");
    LG_CUDA(ctx, cudaFuncSetAttribute(k_project_umma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const unsigned grid = (unsigned)(nsuper < (uint64_t)ctx->num_sms ? nsuper : (uint64_t)ctx->num_sms);
    {
        const cudaError_t fe = cudaEventSynchronize(flag_ev);
        cudaEventDestroy(flag_ev);
        LG_CUDA(ctx, fe);
        if (*h_flag) {  // a non-finite basis column: fall back to the CUDA-core kernel (which propagates it)
            lg_note_fallback(ctx, "projection: a basis column is not finite or too large for the fixed-point grid: CUDA-core kernel");
            if (trace)
                for (int i = 0; i < 3; ++i) cudaEventDestroy(ev[i]);
            return LG_OK;
        }
    }
    if (trace) cudaEventRecord(ev[1], ctx->stream);
    LG_LAUNCH(ctx, k_project_umma, grid, THREADS, smem, d_bm, m->ncols, D, d_bq, K, NB, nstages, d_scale, d_colmax, d_out);
    if (trace) {  // LG_K1_TRACE=1: per-kernel device times of this call (diagnostic; synchronises)
        cudaEventRecord(ev[2], ctx->stream);
        cudaEventSynchronize(ev[2]);
        float a = 0.f, b = 0.f;
        cudaEventElapsedTime(&a, ev[0], ev[1]);
        cudaEventElapsedTime(&b, ev[1], ev[2]);
        fprintf(stderr, "[lg_project] prep %.3f ms, umma %.3f ms\n", a, b);
        for (int i = 0; i < 3; ++i) cudaEventDestroy(ev[i]);
    }
    if (pat) pat->filled = true;
    *used = 1;
    return LG_OK;
}
