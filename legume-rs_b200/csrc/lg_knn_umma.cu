// lg_knn_umma.cu — K7 on the tensor cores: exact k-nearest neighbours as filter + refine.
//   reference: ColumnDict exact backend, matrix-util/src/knn/{metric.rs:19-45, exact.rs:36-55, mod.rs:249-299}
//
// Phase A (tcgen05): for a tile of 128 queries and 256 reference points the dot products come from
//   a split-precision f16 GEMM (x*s = hi + lo; hi*hi + hi*lo + lo*hi, f32 accumulate in TMEM) whose K axis
//   is augmented by two slots ([1, 1] on the query side, [|r|^2 hi, lo] on the reference side, references
//   pre-scaled by -2), so the accumulator IS the ranking key  |r|^2 - 2 q.r.  The epilogue only takes
//   the minimum of 8 columns and votes; rows that can improve insert into their CAND-entry sorted list in
//   shared memory (fused top-k; the accumulator is double-buffered in TMEM so the filter of tile i
//   overlaps the MMAs of tile i+1).
// Phase B (CUDA cores): the CAND survivors of every list are re-scored with the reference's exact f32
//   arithmetic (16 lanes, left fold, tail) and the k smallest (distance, index) pairs are emitted.
//   A query is accepted only if its exact k-th distance is provably below everything Phase A pruned
//   (|key error| <= eps bound); otherwise it is re-run by the brute-force kernel in lg_knn.cu.  The result
//   is therefore always the reference's exact answer, ties included.
#include <cuda_fp16.h>

#include <cstdio>
#include <cstdlib>

#include "lg_common.cuh"
#include "lg_umma.cuh"

using namespace umma;

namespace {

constexpr int QT = 128;        // queries per CTA tile (UMMA M)
constexpr int RT = 256;        // reference points per tile (UMMA N)
constexpr int CAND = 16;       // candidates kept per (query, list)
constexpr int NRST = 2;        // reference-tile ring depth in shared memory (64 KB per tile; a third stage measured no gain and the
                               // space now holds the filter's append buffers)
constexpr int NBUF = 16;       // append-buffer slots per filter thread
constexpr int NPART = 2;        // filter warps per TMEM lane quadrant, each scanning RT / NPART columns of every tile
constexpr int EPI_WARPS = 4 * NPART;
constexpr int W_MMA = EPI_WARPS, W_LOAD = EPI_WARPS + 1;
constexpr int KTHREADS = (EPI_WARPS + 2) * 32;

// packed tile: [split hi|lo][kstep][rows x 32 B canonical K-major core matrices]
__host__ __device__ inline size_t tile_bytes(int rows, int ksteps) { return (size_t)2 * ksteps * rows * 32; }

__global__ void k_absmax(const float* __restrict__ x, uint64_t n, unsigned int* __restrict__ out_bits) {
    float m = 0.0f;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) m = fmaxf(m, fabsf(x[i]));
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
    if ((threadIdx.x & 31) == 0) atomicMax(out_bits, __float_as_uint(m));  // non-negative floats order as uints
}

// power-of-two scale that maps the largest magnitude to [4, 8): exact to apply; |x*s|^2 summed over d <= 126 dims fits f16
__device__ __forceinline__ float knn_scale(unsigned int absmax_bits) {
    const float m = __uint_as_float(absmax_bits);
    if (!(m > 0.0f) || !isfinite(m)) return 1.0f;
    int e;
    frexpf(m, &e);  // m = f * 2^e, f in [0.5, 1)
    return ldexpf(1.0f, 3 - e);
}

// one thread per (point, kstep): 16 consecutive K slots -> hi/lo f16, written as two 16-byte core-matrix rows each.
// K slots 0..d-1 hold the coordinates (references multiplied by -2), slots d and d+1 hold [1, 1] for queries and
// the f16 hi/lo split of |x*s|^2 for references (hi block only); the rest is zero.
__global__ void k_knn_pack(const float* __restrict__ pts, uint64_t n, int d, int ksteps, int rows, int is_ref,
                           const unsigned int* __restrict__ absmax_bits, uint8_t* __restrict__ tiles) {
    const uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t ntile = (n + rows - 1) / rows;
    if (e >= ntile * rows * (uint64_t)ksteps) return;
    const int ks = (int)(e % ksteps);
    const uint64_t p = e / ksteps;  // padded point index
    const uint64_t tile = p / rows;
    const int r = (int)(p % rows);
    const float s = knn_scale(*absmax_bits);
    const float mul = is_ref ? -2.0f * s : s;
    __align__(16) __half hi[16];
    __align__(16) __half lo[16];
    const bool aug = (ks * 16 <= d + 1) && (ks * 16 + 15 >= d);  // this kstep holds slot d and/or d+1
    float nn = 0.0f;
    if (aug && is_ref) {
        nn = INFINITY;  // padding rows can never win
        if (p < n) {
            nn = 0.0f;
            for (int c = 0; c < d; ++c) {
                const float x = pts[p * d + c] * s;
                nn = fmaf(x, x, nn);
            }
        }
    }
    const __half nn_hi = __float2half_rn(nn);
    const __half nn_lo = isinf(nn) ? __float2half_rn(0.0f) : __float2half_rn(nn - __half2float(nn_hi));
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int c = ks * 16 + i;
        const float x = (p < n && c < d) ? pts[p * d + c] * mul : 0.0f;
        hi[i] = __float2half_rn(x);
        lo[i] = __float2half_rn(x - __half2float(hi[i]));
        if (c == d) hi[i] = is_ref ? nn_hi : __float2half_rn(1.0f);
        if (c == d + 1) hi[i] = is_ref ? nn_lo : __float2half_rn(1.0f);
    }
    uint8_t* base = tiles + tile * tile_bytes(rows, ksteps);
    const size_t off = (size_t)ks * rows * 32 + (size_t)(r >> 3) * 256 + (size_t)(r & 7) * 16;
    uint4* dhi = reinterpret_cast<uint4*>(base + off);
    uint4* dlo = reinterpret_cast<uint4*>(base + (size_t)ksteps * rows * 32 + off);
    dhi[0] = reinterpret_cast<uint4*>(hi)[0];   // k bytes 0..15
    dhi[8] = reinterpret_cast<uint4*>(hi)[1];   // k bytes 16..31 live 128 B further (LBO)
    dlo[0] = reinterpret_cast<uint4*>(lo)[0];
    dlo[8] = reinterpret_cast<uint4*>(lo)[1];
}

struct KBars {
    uint64_t r_full[NRST], r_empty[NRST];
    uint64_t q_full;
    uint64_t d_full[2], d_empty[2];
};

// grid = (query tiles, reference splits).  Candidate lists: [query][split*NPART + part][CAND]
__global__ void __launch_bounds__(KTHREADS, 1) k_knn_umma(const uint8_t* __restrict__ qtiles, const uint8_t* __restrict__ rtiles,
                                                          uint64_t nq, uint64_t nr, int ksteps, uint32_t nlists,
                                                          float* __restrict__ cand_key, uint32_t* __restrict__ cand_idx) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const size_t qbytes = tile_bytes(QT, ksteps), rbytes = tile_bytes(RT, ksteps);
    uint8_t* sq = smem;
    uint8_t* sr = smem + ((qbytes + 127) / 128) * 128;
    const size_t rstride = ((rbytes + 127) / 128) * 128;
    KBars* bars = reinterpret_cast<KBars*>(sr + NRST * rstride);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 1);
    // the filter threads' append buffers, slot-major (conflict-free): [NBUF][EPI_WARPS * 32] keys, then indices
    float* buf_k = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 512);
    uint32_t* buf_i = reinterpret_cast<uint32_t*>(buf_k + NBUF * EPI_WARPS * 32);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint64_t nrt = (nr + RT - 1) / RT;
    const uint32_t split = blockIdx.y, nsplit = gridDim.y;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NRST; ++s) {
            mbar_init(&bars->r_full[s], 1);
            mbar_init(&bars->r_empty[s], 1);
        }
        mbar_init(&bars->q_full, 1);
        for (int b = 0; b < 2; ++b) {
            mbar_init(&bars->d_full[b], 1);
            mbar_init(&bars->d_empty[b], EPI_WARPS);
        }
        fence_barrier_init();
    }
    if (warp == W_MMA) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = *tmem_slot;

    if (warp < EPI_WARPS) {
        // ===== fused filter: thread = (query row, column part) =====
        const int quad = warp & 3, half = warp >> 2;  // `half` = this warp's part of the columns, 0 .. NPART-1
        const int row = quad * 32 + lane;
        const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
        // this thread's CAND best (key, index) pairs, ascending, in registers (static indexing only)
        float ck[CAND];
        uint32_t ci[CAND];
#pragma unroll
        for (int c = 0; c < CAND; ++c) {
            ck[c] = INFINITY;
            ci[c] = 0xffffffffu;
        }
        // Deferred insertion.  A key below the list's worst entry used to be inserted on the spot by a 16-stage register
        // network; one improving row out of a warp's 32 made all 32 lanes walk it, and with ~130 improvements per list
        // almost every 8-column group had one: the network's ~70 instructions per element, not the MMAs or the TMEM
        // reads, set the pace (4450 clk per tile against 1536 clk of MMAs).  Now a key below `tau` (the worst entry as
        // of the last merge: stale, hence an upper bound, so nothing is lost) is only APPENDED to the thread's
        // shared-memory buffer; when some lane's buffer could overflow on the next group the whole warp replays its
        // buffers through the network.  Arrival order is kept, so the final lists are exactly those of immediate insertion.
        float tau = INFINITY;
        uint32_t cnt = 0;
        float* bk = buf_k + threadIdx.x;
        uint32_t* bi = buf_i + threadIdx.x;
        auto merge = [&]() {
            const uint32_t mx = __reduce_max_sync(0xffffffffu, cnt);
            for (uint32_t sl = 0; sl < mx; ++sl) {
                if (sl < cnt) {
                    const float key = bk[sl * (EPI_WARPS * 32)];
                    const uint32_t id = bi[sl * (EPI_WARPS * 32)];
                    if (key < ck[CAND - 1]) {
                        // branch-free sorted insertion (ascending); ties keep the earlier (lower) index first
#pragma unroll
                        for (int c = CAND - 1; c >= 1; --c) {
                            const bool up = key < ck[c - 1];
                            const bool in = key < ck[c];
                            ck[c] = in ? (up ? ck[c - 1] : key) : ck[c];
                            ci[c] = in ? (up ? ci[c - 1] : id) : ci[c];
                        }
                        const bool first = key < ck[0];
                        ck[0] = first ? key : ck[0];
                        ci[0] = first ? id : ci[0];
                    }
                }
            }
            cnt = 0;
            tau = ck[CAND - 1];
        };
        uint32_t it = 0;
        for (uint64_t rt = split; rt < nrt; rt += nsplit, ++it) {
            const uint32_t buf = it & 1;
            mbar_wait(&bars->d_full[buf], (it >> 1) & 1);
            tc_fence_after();
            const uint32_t col0 = buf * RT + half * (RT / NPART);
            const uint32_t idx0 = (uint32_t)(rt * RT) + half * (RT / NPART);
            auto scan16 = [&](const uint32_t* dv, int c0) {
#pragma unroll
                for (int h8 = 0; h8 < 2; ++h8) {
                    // tree minimum of 8 keys, then one warp-uniform branch
                    const float m01 = fminf(__uint_as_float(dv[8 * h8 + 0]), __uint_as_float(dv[8 * h8 + 1]));
                    const float m23 = fminf(__uint_as_float(dv[8 * h8 + 2]), __uint_as_float(dv[8 * h8 + 3]));
                    const float m45 = fminf(__uint_as_float(dv[8 * h8 + 4]), __uint_as_float(dv[8 * h8 + 5]));
                    const float m67 = fminf(__uint_as_float(dv[8 * h8 + 6]), __uint_as_float(dv[8 * h8 + 7]));
                    const float m = fminf(fminf(m01, m23), fminf(m45, m67));
                    if (__any_sync(0xffffffffu, m < tau)) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const float key = __uint_as_float(dv[8 * h8 + i]);
                            if (key < tau) {  // cnt <= 8 on entry to a group: at most 16 after it
                                bk[cnt * (EPI_WARPS * 32)] = key;
                                bi[cnt * (EPI_WARPS * 32)] = idx0 + c0 + 8 * h8 + i;
                                ++cnt;
                            }
                        }
                        if (__any_sync(0xffffffffu, cnt > NBUF - 8)) merge();
                    }
                }
            };
            // software-pipelined TMEM reads: the next 16 columns are in flight while these are scanned
            uint32_t da[16], db[16];
            const uint32_t t0 = tbase + lane_base + col0;
            tmem_ld_x16(t0, da);
#pragma unroll 1
            for (int c0 = 0; c0 < RT / NPART; c0 += 32) {
                tmem_wait_ld_x16(da);
                tmem_ld_x16(t0 + c0 + 16, db);
                scan16(da, c0);
                tmem_wait_ld_x16(db);
                if (c0 + 32 < RT / NPART) tmem_ld_x16(t0 + c0 + 32, da);
                scan16(db, c0 + 16);
            }
            // this warp is done with accumulator `buf`
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars->d_empty[buf]);
        }
        if (__any_sync(0xffffffffu, cnt > 0)) merge();
        const uint64_t q = (uint64_t)blockIdx.x * QT + row;
        if (q < nq) {
            const size_t o = (q * nlists + (size_t)split * NPART + half) * CAND;
#pragma unroll
            for (int c = 0; c < CAND; ++c) {
                cand_key[o + c] = ck[c];
                cand_idx[o + c] = ci[c];
            }
        }
    } else if (warp == W_MMA) {
        if (elect_one()) {
            const uint32_t idesc = make_idesc(CFMT_F32, FMT_F16, FMT_F16, QT, RT);
            mbar_wait(&bars->q_full, 0);
            const uint32_t qa = smem_u32(sq);
            uint32_t it = 0;
            for (uint64_t rt = split; rt < nrt; rt += nsplit, ++it) {
                const uint32_t buf = it & 1, rs = it % NRST;
                mbar_wait(&bars->d_empty[buf], ((it >> 1) & 1) ^ 1);
                mbar_wait(&bars->r_full[rs], (it / NRST) & 1);
                tc_fence_after();
                const uint32_t ra = smem_u32(sr + rs * rstride);
                const uint32_t q_lo = qa + (uint32_t)ksteps * QT * 32, r_lo = ra + (uint32_t)ksteps * RT * 32;
                bool acc = false;
                for (int ks = 0; ks < ksteps; ++ks) {
                    const uint64_t qh = make_smem_desc(qa + ks * QT * 32, 128, 256), ql = make_smem_desc(q_lo + ks * QT * 32, 128, 256);
                    const uint64_t rh = make_smem_desc(ra + ks * RT * 32, 128, 256), rl = make_smem_desc(r_lo + ks * RT * 32, 128, 256);
                    mma_f16_ss(tbase + buf * RT, qh, rh, idesc, acc);
                    mma_f16_ss(tbase + buf * RT, qh, rl, idesc, true);
                    mma_f16_ss(tbase + buf * RT, ql, rh, idesc, true);
                    acc = true;
                }
                tc_commit(&bars->d_full[buf]);
                tc_commit(&bars->r_empty[rs]);
            }
        }
    } else {
        if (elect_one()) {
            mbar_arrive_expect_tx(&bars->q_full, (uint32_t)qbytes);
            bulk_g2s(sq, qtiles + (size_t)blockIdx.x * qbytes, (uint32_t)qbytes, &bars->q_full);
            uint32_t it = 0;
            for (uint64_t rt = split; rt < nrt; rt += nsplit, ++it) {
                const uint32_t rs = it % NRST;
                mbar_wait(&bars->r_empty[rs], ((it / NRST) & 1) ^ 1);
                mbar_arrive_expect_tx(&bars->r_full[rs], (uint32_t)rbytes);
                bulk_g2s(sr + rs * rstride, rtiles + rt * rbytes, (uint32_t)rbytes, &bars->r_full[rs]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == W_MMA) tmem_dealloc(tbase, 512);
}

__device__ __forceinline__ float exact_l2_sq(const float* __restrict__ r, const float* __restrict__ q, int d) {
    float acc[16];
#pragma unroll
    for (int l = 0; l < 16; ++l) acc[l] = 0.0f;
    int c = 0;
    for (; c + 16 <= d; c += 16) {
#pragma unroll
        for (int l = 0; l < 16; ++l) {
            const float df = __fsub_rn(r[c + l], q[c + l]);
            acc[l] = __fadd_rn(acc[l], __fmul_rn(df, df));
        }
    }
    float sum = 0.0f;
#pragma unroll
    for (int l = 0; l < 16; ++l) sum = __fadd_rn(sum, acc[l]);
    for (; c < d; ++c) {
        const float df = __fsub_rn(r[c], q[c]);
        sum = __fadd_rn(sum, __fmul_rn(df, df));
    }
    return sum;
}

// Phase B: one warp per query.  Exact re-score of all candidates, k smallest (dist, idx), verification.
__global__ void __launch_bounds__(256) k_knn_refine(const float* __restrict__ ref, const float* __restrict__ qry, uint64_t nq, int d,
                                                    int k, const uint32_t* __restrict__ exclude, uint32_t nlists,
                                                    const float* __restrict__ cand_key, const uint32_t* __restrict__ cand_idx,
                                                    const unsigned int* __restrict__ absmax_bits, uint32_t* __restrict__ out_idx,
                                                    float* __restrict__ out_dist, uint32_t* __restrict__ redo_list,
                                                    unsigned int* __restrict__ redo_count, int squared) {
    extern __shared__ unsigned long long keys_s[];  // per warp: nlists*CAND keys
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t q = (uint64_t)blockIdx.x * (blockDim.x >> 5) + warp;
    if (q >= nq) return;
    const uint32_t ncand = nlists * CAND;
    unsigned long long* keys = keys_s + (size_t)warp * ncand;
    const float* qv = qry + q * d;
    const uint32_t ex = exclude ? exclude[q] : 0xffffffffu;
    const float s = knn_scale(*absmax_bits);
    // lower bound (in key units) of every reference point Phase A dropped: the smallest "list maximum" over full lists
    float prune = INFINITY;
    for (uint32_t l = lane; l < nlists; l += 32) {
        const size_t o = (q * nlists + l) * CAND;
        if (cand_idx[o + CAND - 1] != 0xffffffffu) prune = fminf(prune, cand_key[o + CAND - 1]);
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) prune = fminf(prune, __shfl_xor_sync(0xffffffffu, prune, off));
    float qn = 0.0f;
    for (int c = lane; c < d; c += 32) qn = fmaf(qv[c], qv[c], qn);
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) qn += __shfl_xor_sync(0xffffffffu, qn, off);
    for (uint32_t c = lane; c < ncand; c += 32) {
        const uint32_t id = cand_idx[q * (size_t)ncand + c];
        unsigned long long key = ~0ull;
        if (id != 0xffffffffu && id != ex) {
            const float d2 = exact_l2_sq(ref + (size_t)id * d, qv, d);
            key = ((unsigned long long)__float_as_uint(d2) << 32) | id;
        }
        keys[c] = key;
    }
    __syncwarp();
    // k rounds of warp arg-min over the candidate keys
    unsigned long long kth = 0;
    int found = 0;
    for (int t = 0; t < k; ++t) {
        unsigned long long best = ~0ull;
        uint32_t bpos = 0;
        for (uint32_t c = lane; c < ncand; c += 32) {
            const unsigned long long v = keys[c];
            if (v < best) {
                best = v;
                bpos = c;
            }
        }
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) {
            const unsigned long long o = __shfl_xor_sync(0xffffffffu, best, off);
            const uint32_t op = __shfl_xor_sync(0xffffffffu, bpos, off);
            if (o < best) {
                best = o;
                bpos = op;
            }
        }
        if (lane == 0) {
            const bool empty = best == ~0ull;
            out_idx[q * k + t] = empty ? 0xffffffffu : (uint32_t)(best & 0xffffffffu);
            const float d2 = __uint_as_float((uint32_t)(best >> 32));
            out_dist[q * k + t] = empty ? INFINITY : (squared ? d2 : __fsqrt_rn(d2));
            if (!empty) keys[bpos] = ~0ull;
        }
        if (best != ~0ull) {
            kth = best;
            ++found;
        }
        __syncwarp();
    }
    if (lane == 0 && prune != INFINITY) {
        // accepted only if the exact k-th distance is below every pruned point's smallest possible distance
        const float kth_d2 = found == k ? __uint_as_float((uint32_t)(kth >> 32)) : INFINITY;
        const float lower = qn + prune / (s * s);
        // slack of the filter's keys: a RELATIVE part (2^-13 of the magnitudes: f16 hi/lo products, f32 key arithmetic)
        // and an ABSOLUTE part — a scaled coordinate below 2^-14 makes the f16 "lo" half subnormal, i.e. up to 2^-25 of
        // absolute error per coordinate, hence up to 2^-24 sqrt(d) (|q| + |r|) / s per dot product in distance units
        // whatever the norms are (data with one large outlier and a tight cluster near the origin).  |r| <= |q| + dist.
        const float rel = 1.220703125e-4f * (fabsf(qn) + fabsf(lower) + 1e-30f);
        const float abs_part = 4.76837158e-7f * __fsqrt_rn((float)d) * (2.0f * __fsqrt_rn(qn) + __fsqrt_rn(fmaxf(lower, 0.0f))) / s;  // 2^-21
        const float eps = rel + abs_part;
        if (!(kth_d2 < lower - eps)) redo_list[atomicAdd(redo_count, 1u)] = (uint32_t)q;
    }
}

}  // namespace

// brute-force kernel of lg_knn.cu on a list of query ids (device-side count)
int lg_knn_exact_list(lg_ctx* ctx, const float* d_ref, uint64_t nr, const float* d_qry, uint64_t nq, int d, int k,
                      const uint32_t* d_exclude, const uint32_t* d_qlist, const unsigned int* d_qcount, int squared,
                      uint32_t* d_idx, float* d_dist);

// returns LG_OK; *used = 0 means "not applicable, run the brute-force kernel"
int lg_knn_topk_umma(lg_ctx* ctx, const float* d_ref, uint64_t nr, const float* d_qry, uint64_t nq, int d, int k,
                     const uint32_t* d_ex, uint32_t* d_idx, float* d_dist, int squared, int* used) {
    *used = 0;
    if (nr < 4096 || nq == 0 || (double)nq * (double)nr < 5e7) return LG_OK;  // small searches: the brute-force kernel is the faster one
    if (k + 4 > CAND || d > 126 || nr >= 0xFFFFFF00ull) {
        char b[200];
        snprintf(b, sizeof(b), "kNN: k = %d, d = %d is outside the tensor-core filter (k <= %d, d <= 126): CUDA-core kernel", k, d, CAND - 4);
        lg_note_fallback(ctx, b);
        return LG_OK;
    }
    const int ksteps = (d + 2 + 15) / 16;
    const size_t qbytes = tile_bytes(QT, ksteps), rbytes = tile_bytes(RT, ksteps);
    static_assert(sizeof(KBars) + 16 <= 512, "barriers and the TMEM slot fit the 512 bytes ahead of the append buffers");
    const size_t smem = ((qbytes + 127) / 128) * 128 + NRST * (((rbytes + 127) / 128) * 128) + 512 + (size_t)2 * NBUF * EPI_WARPS * 32 * 4;
    if (smem > ctx->smem_optin) return LG_OK;
    const uint64_t nqt = (nq + QT - 1) / QT, nrt = (nr + RT - 1) / RT;
    uint32_t nsplit = 1;
    while (nqt * nsplit < (uint64_t)2 * ctx->num_sms && nsplit * 2 <= nrt && nsplit < 16) nsplit *= 2;
    const uint32_t nlists = nsplit * NPART;
    LgStage st(ctx);
    unsigned int *d_absmax, *d_redo_count;
    uint8_t *d_qt, *d_rt;
    float* d_ckey;
    uint32_t *d_cidx, *d_redo;
    LG_TRY(st.scratch(1, &d_absmax));
    LG_TRY(st.scratch(1, &d_redo_count));
    LG_TRY(st.scratch(nqt * qbytes, &d_qt));
    LG_TRY(st.scratch(nrt * rbytes, &d_rt));
    LG_TRY(st.scratch((size_t)nq * nlists * CAND, &d_ckey));
    LG_TRY(st.scratch((size_t)nq * nlists * CAND, &d_cidx));
    LG_TRY(st.scratch((size_t)nq, &d_redo));
    LG_CUDA(ctx, cudaMemsetAsync(d_absmax, 0, sizeof(unsigned int), ctx->stream));
    LG_CUDA(ctx, cudaMemsetAsync(d_redo_count, 0, sizeof(unsigned int), ctx->stream));
    LG_LAUNCH(ctx, k_absmax, ctx->num_sms * 4, 256, 0, d_ref, nr * (uint64_t)d, d_absmax);
    LG_LAUNCH(ctx, k_absmax, ctx->num_sms * 4, 256, 0, d_qry, nq * (uint64_t)d, d_absmax);
    {
        const uint64_t tq = nqt * QT * (uint64_t)ksteps, tr = nrt * RT * (uint64_t)ksteps;
        LG_LAUNCH(ctx, k_knn_pack, (unsigned)((tq + 127) / 128), 128, 0, d_qry, nq, d, ksteps, QT, 0, d_absmax, d_qt);
        LG_LAUNCH(ctx, k_knn_pack, (unsigned)((tr + 127) / 128), 128, 0, d_ref, nr, d, ksteps, RT, 1, d_absmax, d_rt);
    }
    LG_CUDA(ctx, cudaFuncSetAttribute(k_knn_umma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    {
        dim3 grid((unsigned)nqt, nsplit);
        k_knn_umma<<<grid, KTHREADS, smem, ctx->stream>>>(d_qt, d_rt, nq, nr, ksteps, nlists, d_ckey, d_cidx);
        ctx->launches++;
        LG_CUDA(ctx, cudaGetLastError());
    }
    {
        const int wpb = 8;
        const size_t rsmem = (size_t)wpb * nlists * CAND * sizeof(unsigned long long);
        LG_CUDA(ctx, cudaFuncSetAttribute(k_knn_refine, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rsmem));
        LG_LAUNCH(ctx, k_knn_refine, (unsigned)((nq + wpb - 1) / wpb), wpb * 32, rsmem, d_ref, d_qry, nq, d, k, d_ex, nlists, d_ckey,
                  d_cidx, d_absmax, d_idx, d_dist, d_redo, d_redo_count, squared);
    }
    if (const char* dbg = getenv("LG_KNN_STATS")) {
        if (dbg[0] == '1') {
            unsigned int h = 0;
            LG_CUDA(ctx, cudaMemcpyAsync(&h, d_redo_count, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
            LG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            fprintf(stderr, "[lg_knn] tensor path: nq=%llu nr=%llu d=%d k=%d splits=%u, %u queries re-run exactly\n",
                    (unsigned long long)nq, (unsigned long long)nr, d, k, nsplit, h);
        }
    }
    LG_TRY(lg_knn_exact_list(ctx, d_ref, nr, d_qry, nq, d, k, d_ex, d_redo, d_redo_count, squared, d_idx, d_dist));
    *used = 1;
    return LG_OK;
}
