// lg_knn.cu — stage 6: exact k-nearest neighbours (ColumnDict with the exact backend).
//   l2_sq_kernel     matrix-util/src/knn/metric.rs:19-45   16 lane accumulators, left-fold, tail
//   exact::topk      matrix-util/src/knn/exact.rs:36-55    rank by squared distance (total_cmp)
//   search_indices   matrix-util/src/knn/mod.rs:249-299    optional self-exclusion
//
// Baseline CUDA-core form: one warp per query, reference points streamed through a shared-memory
// tile, every lane evaluates the reference's exact f32 distance arithmetic for its own points and
// the warp keeps a k-entry candidate list ordered by the 64-bit key (distance bits << 32 | index),
// which is the reference's (distance, then lower index) order for non-negative distances.
#include <cstdlib>

#include "lg_common.cuh"

int lg_knn_topk_umma(lg_ctx* ctx, const float* d_ref, uint64_t nr, const float* d_qry, uint64_t nq, int d, int k,
                     const uint32_t* d_ex, uint32_t* d_idx, float* d_dist, int squared, int* used);

constexpr int KNN_WARPS = 8;      // queries per CTA
constexpr int KNN_TILE = 128;     // reference points per shared-memory tile
constexpr int KNN_KMAX = 1024;    // candidate list capacity per query

__device__ __forceinline__ float knn_l2_sq(const float* __restrict__ r, const float* __restrict__ q, int d) {
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

__device__ __forceinline__ unsigned long long warp_max_u64(unsigned long long v, int& arg, int mine) {
    arg = mine;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        const unsigned long long o = __shfl_xor_sync(0xffffffffu, v, off);
        const int oa = __shfl_xor_sync(0xffffffffu, arg, off);
        if (o > v || (o == v && oa < arg)) {
            v = o;
            arg = oa;
        }
    }
    return v;
}

__global__ void __launch_bounds__(KNN_WARPS * 32) k_knn_exact(const float* __restrict__ ref, uint64_t nr,
                                                              const float* __restrict__ qry, uint64_t nq, int d, int k,
                                                              const uint32_t* __restrict__ exclude,
                                                              const uint32_t* __restrict__ qlist,
                                                              const unsigned int* __restrict__ qcount, int squared,
                                                              uint32_t* __restrict__ out_idx, float* __restrict__ out_dist) {
    extern __shared__ unsigned char smem_raw[];
    const int ds = d | 1;  // odd row stride: lanes walking different rows hit different banks
    float* tile = reinterpret_cast<float*>(smem_raw);                      // KNN_TILE * ds
    float* qs = tile + (size_t)KNN_TILE * ds;                               // KNN_WARPS * d
    unsigned long long* lists = reinterpret_cast<unsigned long long*>(qs + (size_t)KNN_WARPS * d + ((KNN_WARPS * d) & 1));
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // optional indirection: only the queries listed in qlist[0 .. *qcount) are processed
    const uint64_t slot = (uint64_t)blockIdx.x * KNN_WARPS + warp;
    const uint64_t nslots = qlist ? (uint64_t)*qcount : nq;
    if ((uint64_t)blockIdx.x * KNN_WARPS >= nslots) return;  // CTA-uniform
    const bool live = slot < nslots;
    const uint64_t q = live ? (qlist ? (uint64_t)qlist[slot] : slot) : 0;
    unsigned long long* mine = lists + (size_t)warp * k;
    float* myq = qs + (size_t)warp * d;
    if (live)
        for (int c = lane; c < d; c += 32) myq[c] = qry[q * d + c];
    for (int t = lane; t < k; t += 32) mine[t] = ~0ull;
    const uint32_t ex = (live && exclude) ? exclude[q] : 0xffffffffu;
    unsigned long long thr = ~0ull;  // current worst key in the list
    int thr_pos = 0;
    for (uint64_t base = 0; base < nr; base += KNN_TILE) {
        const int nt = (nr - base) < (uint64_t)KNN_TILE ? (int)(nr - base) : KNN_TILE;
        __syncthreads();
        for (size_t e = threadIdx.x; e < (size_t)nt * d; e += blockDim.x) tile[(e / d) * ds + (e % d)] = ref[base * d + e];
        __syncthreads();
        if (!live) continue;
        for (int t0 = 0; t0 < nt; t0 += 32) {
            const int t = t0 + lane;
            unsigned long long key = ~0ull;
            if (t < nt) {
                const uint32_t idx = (uint32_t)(base + t);
                if (idx != ex) {
                    const float d2 = knn_l2_sq(tile + (size_t)t * ds, myq, d);
                    key = ((unsigned long long)__float_as_uint(d2) << 32) | idx;
                }
            }
            unsigned pend = __ballot_sync(0xffffffffu, key < thr);
            while (pend) {
                const int src = __ffs(pend) - 1;
                pend &= pend - 1;
                const unsigned long long cand = __shfl_sync(0xffffffffu, key, src);
                if (cand < thr) {
                    __syncwarp();
                    if (lane == 0) mine[thr_pos] = cand;
                    __syncwarp();
                    // recompute the worst entry (lexicographic max, lowest slot on ties)
                    unsigned long long best = 0;
                    int bpos = 0x7fffffff;
                    for (int s = lane; s < k; s += 32) {
                        const unsigned long long v = mine[s];
                        if (bpos == 0x7fffffff || v > best) {
                            best = v;
                            bpos = s;
                        }
                    }
                    int arg;
                    thr = warp_max_u64(best, arg, bpos);
                    thr_pos = arg;
                }
            }
        }
    }
    if (!live) return;
    __syncwarp();
    // rank sort: entry s goes to the slot equal to the number of smaller keys (keys are distinct or ~0)
    for (int s = lane; s < k; s += 32) {
        const unsigned long long v = mine[s];
        int rank = 0;
        for (int o = 0; o < k; ++o) {
            const unsigned long long w = mine[o];
            rank += (w < v) || (w == v && o < s);
        }
        const bool empty = v == ~0ull;
        out_idx[q * k + rank] = empty ? 0xffffffffu : (uint32_t)(v & 0xffffffffu);
        const float d2 = __uint_as_float((uint32_t)(v >> 32));
        out_dist[q * k + rank] = empty ? INFINITY : (squared ? d2 : __fsqrt_rn(d2));
    }
}

// device-pointer core shared by lg_knn_topk and the cross-batch matching of lg_adjust.cu:
// tensor-core filter + exact refine when applicable, else the CUDA-core kernel.  Same result either way.
int lg_knn_topk_device(lg_ctx* ctx, const float* d_ref, uint64_t nr, const float* d_qry, uint64_t nq, int d, int k,
                       const uint32_t* d_ex, uint32_t* d_idx, float* d_dist, int squared) {
    LG_REQUIRE(ctx, d >= 1 && d <= 256, "lg_knn_topk: d must be in [1, 256]");
    LG_REQUIRE(ctx, k >= 1 && k <= KNN_KMAX, "lg_knn_topk: k must be in [1, 1024]");
    LG_REQUIRE(ctx, nr < 0xFFFFFFFFull, "lg_knn_topk: reference set must have < 2^32-1 points");
    if (nq == 0) return LG_OK;
    const size_t fl = (size_t)KNN_TILE * (d | 1) + (size_t)KNN_WARPS * d + ((KNN_WARPS * d) & 1);
    const size_t smem = fl * sizeof(float) + (size_t)KNN_WARPS * k * sizeof(unsigned long long);
    LG_REQUIRE(ctx, smem <= ctx->smem_optin, "lg_knn_topk: d and k too large for shared memory");
    LG_CUDA(ctx, cudaFuncSetAttribute(k_knn_exact, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int used = 0;
    const char* force = getenv("LG_KNN_CUDA_CORES");
    if (!(force && force[0] == '1')) LG_TRY(lg_knn_topk_umma(ctx, d_ref, nr, d_qry, nq, d, k, d_ex, d_idx, d_dist, squared, &used));
    if (!used)
        LG_LAUNCH(ctx, k_knn_exact, (unsigned)((nq + KNN_WARPS - 1) / KNN_WARPS), KNN_WARPS * 32, smem, d_ref, nr, d_qry, nq, d, k,
                  d_ex, (const uint32_t*)nullptr, (const unsigned int*)nullptr, squared, d_idx, d_dist);
    return LG_OK;
}

static int knn_topk_entry(lg_ctx* ctx, const float* ref, uint64_t nr, const float* qry, uint64_t nq, int d, int k,
                          const uint32_t* exclude, uint32_t* out_idx, float* out_dist, int squared);

extern "C" int lg_knn_topk(lg_ctx* ctx, const float* ref, uint64_t nr, const float* qry, uint64_t nq, int d, int k,
                           const uint32_t* exclude, uint32_t* out_idx, float* out_dist) {
    return knn_topk_entry(ctx, ref, nr, qry, nq, d, k, exclude, out_idx, out_dist, 0);
}
extern "C" int lg_knn_topk_sq(lg_ctx* ctx, const float* ref, uint64_t nr, const float* qry, uint64_t nq, int d, int k,
                              const uint32_t* exclude, uint32_t* out_idx, float* out_sq) {
    return knn_topk_entry(ctx, ref, nr, qry, nq, d, k, exclude, out_idx, out_sq, 1);
}

static int knn_topk_entry(lg_ctx* ctx, const float* ref, uint64_t nr, const float* qry, uint64_t nq, int d, int k,
                          const uint32_t* exclude, uint32_t* out_idx, float* out_dist, int squared) {
    if (!ctx) return LG_ERR_INVALID;
    LG_REQUIRE(ctx, qry && out_idx && out_dist && (ref || nr == 0), "lg_knn_topk: null argument");
    LG_REQUIRE(ctx, d >= 1 && d <= 256, "lg_knn_topk: d must be in [1, 256]");
    LG_REQUIRE(ctx, k >= 1 && k <= KNN_KMAX, "lg_knn_topk: k must be in [1, 1024]");
    cudaSetDevice(ctx->device);
    LgStage st(ctx);
    const float *d_ref, *d_qry;
    const uint32_t* d_ex;
    uint32_t* d_idx;
    float* d_dist;
    LG_TRY(st.in(ref, (size_t)nr * d, &d_ref));
    LG_TRY(st.in(qry, (size_t)nq * d, &d_qry));
    LG_TRY(st.in(exclude, (size_t)nq, &d_ex));
    LG_TRY(st.out(out_idx, (size_t)nq * k, &d_idx));
    LG_TRY(st.out(out_dist, (size_t)nq * k, &d_dist));
    LG_TRY(lg_knn_topk_device(ctx, d_ref, nr, d_qry, nq, d, k, d_ex, d_idx, d_dist, squared));
    return st.finish();
}

// brute-force pass over the queries listed in d_qlist[0 .. *d_qcount) (used by the tensor path's verified fallback)
int lg_knn_exact_list(lg_ctx* ctx, const float* d_ref, uint64_t nr, const float* d_qry, uint64_t nq, int d, int k,
                      const uint32_t* d_exclude, const uint32_t* d_qlist, const unsigned int* d_qcount, int squared,
                      uint32_t* d_idx, float* d_dist) {
    if (nq == 0) return LG_OK;
    const size_t fl = (size_t)KNN_TILE * (d | 1) + (size_t)KNN_WARPS * d + ((KNN_WARPS * d) & 1);
    const size_t smem = fl * sizeof(float) + (size_t)KNN_WARPS * k * sizeof(unsigned long long);
    LG_CUDA(ctx, cudaFuncSetAttribute(k_knn_exact, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    LG_LAUNCH(ctx, k_knn_exact, (unsigned)((nq + KNN_WARPS - 1) / KNN_WARPS), KNN_WARPS * 32, smem, d_ref, nr, d_qry, nq, d, k,
              d_exclude, d_qlist, d_qcount, squared, d_idx, d_dist);
    return LG_OK;
}

// ---- top-k merge across reference-cell shards ------------------------------------------
// Every shard answered the same nq queries against its own slice of the reference cells with
// lg_knn_topk_sq (squared distances, local indices).  One thread per query merges the nshard sorted
// k-lists by the reference's order (squared distance, then lower GLOBAL index) and takes the root once.
__global__ void k_knn_merge(const uint32_t* __restrict__ shard_idx, const float* __restrict__ shard_sq, uint32_t nshard,
                            uint64_t nq, int k, const uint64_t* __restrict__ shard_offset, const uint32_t* __restrict__ exclude,
                            uint32_t* __restrict__ out_idx, float* __restrict__ out_dist) {
    const uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    uint32_t head[32];
    for (uint32_t s = 0; s < nshard; ++s) head[s] = 0;
    const uint32_t ex = exclude ? exclude[q] : 0xffffffffu;
    for (int r = 0; r < k;) {
        unsigned long long best = ~0ull;
        uint32_t arg = 0xffffffffu;
        for (uint32_t s = 0; s < nshard; ++s) {
            if (head[s] >= (uint32_t)k) continue;
            const size_t o = ((size_t)s * nq + q) * k + head[s];
            const uint32_t li = shard_idx[o];
            if (li == 0xffffffffu) {
                head[s] = (uint32_t)k;  // this shard's list is exhausted
                continue;
            }
            const unsigned long long key = ((unsigned long long)__float_as_uint(shard_sq[o]) << 32) | (uint32_t)(shard_offset[s] + li);
            if (key < best) {
                best = key;
                arg = s;
            }
        }
        if (arg == 0xffffffffu) {
            for (; r < k; ++r) {
                out_idx[q * k + r] = 0xffffffffu;
                out_dist[q * k + r] = INFINITY;
            }
            break;
        }
        head[arg]++;
        const uint32_t gi = (uint32_t)(best & 0xffffffffu);
        if (gi == ex) continue;  // self-exclusion by global index (knn/mod.rs:255-296)
        out_idx[q * k + r] = gi;
        out_dist[q * k + r] = __fsqrt_rn(__uint_as_float((uint32_t)(best >> 32)));
        ++r;
    }
}

extern "C" int lg_knn_merge_topk(lg_ctx* ctx, const uint32_t* shard_idx, const float* shard_sq, uint32_t nshard, uint64_t nq, int k,
                                 const uint64_t* shard_offset, const uint32_t* exclude, uint32_t* out_idx, float* out_dist) {
    if (!ctx) return LG_ERR_INVALID;
    LG_REQUIRE(ctx, shard_idx && shard_sq && shard_offset && out_idx && out_dist, "lg_knn_merge_topk: null argument");
    LG_REQUIRE(ctx, nshard >= 1 && nshard <= 32 && k >= 1 && k <= KNN_KMAX, "lg_knn_merge_topk: nshard in [1, 32], k in [1, 1024]");
    cudaSetDevice(ctx->device);
    LgStage st(ctx);
    const uint32_t *d_si, *d_ex;
    const float* d_ss;
    const uint64_t* d_off;
    uint32_t* d_oi;
    float* d_od;
    LG_TRY(st.in(shard_idx, (size_t)nshard * nq * k, &d_si));
    LG_TRY(st.in(shard_sq, (size_t)nshard * nq * k, &d_ss));
    LG_TRY(st.in(shard_offset, (size_t)nshard, &d_off));
    LG_TRY(st.in(exclude, (size_t)nq, &d_ex));
    LG_TRY(st.out(out_idx, (size_t)nq * k, &d_oi));
    LG_TRY(st.out(out_dist, (size_t)nq * k, &d_od));
    if (nq) LG_LAUNCH(ctx, k_knn_merge, (unsigned)((nq + 127) / 128), 128, 0, d_si, d_ss, nshard, nq, k, d_off, d_ex, d_oi, d_od);
    return st.finish();
}
