// lg_collapse.cu — stage 4: gene x group sums (pseudobulk collapse).
//   collect_basic_stat_visitor / collect_batch_stat_visitor   collapse_data/stats.rs:110-164
//   merge_stat                                                collapse_data/stats.rs:790-833
//
// The reference visits group by group under one global lock.  Here cells are stably sorted by
// their label (group or batch), the sorted order is cut into fixed chunks, and each CTA streams
// its chunk's columns with coalesced loads, scatter-adding into a D-long shared-memory
// accumulator that is flushed to the D x S output with one global atomic per touched gene when the
// label changes.  Sums of integer-valued counts are exact, hence order-independent and
// bit-identical to the reference's sequential fold.
#include <cub/cub.cuh>

#include "lg_common.cuh"

constexpr int COLLAPSE_THREADS = 1024;  // one CTA per SM (the accumulator fills most of shared memory)
constexpr int COLLAPSE_CHUNK = 256;     // sorted cells per work item
constexpr int COLLAPSE_UNROLL = 8;      // independent 32-nnz loads in flight per warp (the kernel is bound by bytes in flight)

__global__ void k_iota_u32(uint32_t* p, uint64_t n) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = (uint32_t)i;
}

// 1 if every stored value is a non-negative integer below 2^20 (then sums can be kept as u32)
__global__ void k_all_integral(const float* __restrict__ v, uint64_t n, int* __restrict__ not_integral) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    bool bad = false;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float x = __ldg(v + i);
        bad |= !(x >= 0.0f && x < 1048576.0f && x == truncf(x));
    }
    if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(not_integral, 1);
}

// Accumulator element.  Integer counts with unit multiplicity: u32 with native shared-memory ATOMS.ADD — exact, so the result
// does not depend on the order in which warps and CTAs arrive.  Anything else (fractional multiplicities, non-integer values):
// 64-bit FIXED POINT, `round(v * w * 2^e)` with e chosen per call so that no sum can overflow — integer adds again, so these
// sums are also bit-identical run to run and for any sharding (section 8b "Determinism"); f32 atomics in arrival order were not.
// They leave through a u64 twin of the output that one pass converts to f32 at the end.
template <bool INT>
struct Acc;
template <>
struct Acc<true> {
    using T = unsigned int;
    static __device__ __forceinline__ void add(T* a, float v, float, float) { atomicAdd(a, (unsigned int)v); }
};
template <>
struct Acc<false> {
    using T = unsigned long long;
    static __device__ __forceinline__ void add(T* a, float v, float w, float scale) {
        atomicAdd(a, (unsigned long long)__float2ll_rn(__fmul_rn(__fmul_rn(v, w), scale)));  // two's complement: negatives wrap and add up
    }
};

// largest |x| of an array as float bits (non-negative floats order like their bit patterns)
__global__ void k_max_abs_bits(const float* __restrict__ v, uint64_t n, unsigned int* __restrict__ out) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    float m = 0.0f;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float x = fabsf(__ldg(v + i));
        if (x > m && x <= 3.0e38f) m = x;  // non-finite values do not size the scale
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
    if ((threadIdx.x & 31) == 0 && m > 0.0f) atomicMax(out, __float_as_uint(m));
}

// fixed-point twin -> f32 sums
__global__ void k_fixed_to_f32(const unsigned long long* __restrict__ src, uint64_t n, double inv_scale, float* __restrict__ dst) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = (float)((double)(long long)src[i] * inv_scale);
}

// size_s[s] += w over the cells of group s in ascending cell order: the reference's own fold (stats.rs:126-131), one thread per
// group over its segment of the sorted cells (used when multiplicities are registered; with unit weights the in-kernel adds are exact)
__global__ void k_size_by_group(const uint32_t* __restrict__ sorted_label, const uint32_t* __restrict__ sorted_cell, uint32_t n,
                                const float* __restrict__ mult, uint32_t S, float* __restrict__ size_s) {
    const uint32_t sgrp = blockIdx.x * blockDim.x + threadIdx.x;
    if (sgrp >= S) return;
    auto lower = [&](uint32_t key) {
        uint32_t lo = 0, hi = n;
        while (lo < hi) {
            const uint32_t mid = lo + (hi - lo) / 2;
            if (sorted_label[mid] < key) lo = mid + 1;
            else hi = mid;
        }
        return lo;
    };
    float acc = 0.0f;
    for (uint32_t i = lower(sgrp), e = lower(sgrp + 1); i < e; ++i) acc = __fadd_rn(acc, mult[sorted_cell[i]]);
    size_s[sgrp] = acc;
}

// window [g0, g0 + W) of the gene axis lives in shared memory
// VEC: index / value arrays are 16-byte aligned, so a lane reads 4 consecutive nnz per 128-bit load (entries of the
// neighbouring columns that share the first / last group are masked): four times the bytes in flight per load
template <bool INT, bool VEC>
__global__ void __launch_bounds__(COLLAPSE_THREADS, 1) k_collapse_sorted(
    const uint64_t* __restrict__ indptr, const uint32_t* __restrict__ indices, const float* __restrict__ values, uint64_t nnz_total,
    const uint32_t* __restrict__ sorted_label, const uint32_t* __restrict__ sorted_cell, uint64_t ncells,
    const float* __restrict__ mult, uint32_t S, uint64_t D, uint32_t g0, uint32_t W, float* __restrict__ sum_ds,
    float* __restrict__ size_s, unsigned long long* __restrict__ next_chunk, const uint32_t* __restrict__ cell_range,
    unsigned long long* __restrict__ sum64, float scale) {
    using A = Acc<INT>;
    // cell_range (or NULL): this launch takes the sorted positions [cell_range[0], cell_range[1]) only — the sharded path
    // collapses the groups in two halves so that the first half's sums can be all-reduced while the second is summed
    uint64_t pos0 = 0;
    if (cell_range) {
        pos0 = cell_range[0];
        ncells = cell_range[1];
    }
    extern __shared__ __align__(16) unsigned char smem_raw[];
    typename A::T* acc = reinterpret_cast<typename A::T*>(smem_raw);  // W accumulators
    __shared__ unsigned long long s_chunk;
    __shared__ uint32_t s_label[COLLAPSE_CHUNK];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = COLLAPSE_THREADS / 32;
    const uint64_t nchunks = (ncells - pos0 + COLLAPSE_CHUNK - 1) / COLLAPSE_CHUNK;
    for (uint32_t g = threadIdx.x; g < W; g += COLLAPSE_THREADS) acc[g] = 0;
    while (true) {
        __syncthreads();
        if (threadIdx.x == 0) s_chunk = atomicAdd(next_chunk, 1ull);
        __syncthreads();
        const uint64_t chunk = s_chunk;
        if (chunk >= nchunks) break;
        const uint64_t p0 = pos0 + chunk * COLLAPSE_CHUNK;
        const int np = (int)((p0 + COLLAPSE_CHUNK) < ncells ? COLLAPSE_CHUNK : (ncells - p0));
        if ((int)threadIdx.x < np) s_label[threadIdx.x] = sorted_label[p0 + threadIdx.x];
        __syncthreads();
        int seg0 = 0;
        while (seg0 < np) {
            // segment of equal labels inside the chunk (labels are sorted)
            const uint32_t lab = s_label[seg0];
            int seg1 = seg0 + 1;
            while (seg1 < np && s_label[seg1] == lab) ++seg1;
            if (lab < S) {
                float wsum = 0.0f;
                // software-pipelined walk over this warp's cells: fetch the next cell's extent early
                int p = seg0 + warp;
                uint32_t cell = 0;
                uint64_t lo = 0, hi = 0;
                if (p < seg1) {
                    cell = sorted_cell[p0 + p];
                    lo = indptr[cell];
                    hi = indptr[cell + 1];
                }
                while (p < seg1) {
                    const int pn = p + nwarp;
                    uint32_t cell_n = 0;
                    uint64_t lo_n = 0, hi_n = 0;
                    if (pn < seg1) {
                        cell_n = sorted_cell[p0 + pn];
                        lo_n = indptr[cell_n];
                        hi_n = indptr[cell_n + 1];
                    }
                    const float w = (!INT && mult) ? mult[cell] : 1.0f;
                    if constexpr (VEC) {
                        constexpr int VU = 4;  // 128-bit groups in flight per lane and array
                        const uint64_t hi4 = nnz_total & ~3ull;  // groups at or beyond this one would cross the array end
                        for (uint64_t c = (lo & ~3ull) + 4ull * lane; c < hi; c += 128ull * VU) {
                            uint4 gq[VU];
                            float4 vq[VU];
#pragma unroll
                            for (int u = 0; u < VU; ++u) {
                                const uint64_t cu = c + 128ull * u;
                                gq[u] = make_uint4(0, 0, 0, 0);
                                vq[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                                if (cu < hi) {
                                    if (cu < hi4) {
                                        gq[u] = __ldg(reinterpret_cast<const uint4*>(indices + cu));
                                        vq[u] = __ldg(reinterpret_cast<const float4*>(values + cu));
                                    } else {  // the last, partial group of the whole array
                                        uint32_t gg[4] = {0, 0, 0, 0};
                                        float vv[4] = {0.f, 0.f, 0.f, 0.f};
                                        for (int e = 0; e < 4; ++e)
                                            if (cu + e < nnz_total) {
                                                gg[e] = indices[cu + e];
                                                vv[e] = values[cu + e];
                                            }
                                        gq[u] = make_uint4(gg[0], gg[1], gg[2], gg[3]);
                                        vq[u] = make_float4(vv[0], vv[1], vv[2], vv[3]);
                                    }
                                }
                            }
#pragma unroll
                            for (int u = 0; u < VU; ++u) {
                                const uint64_t cu = c + 128ull * u;
                                const uint32_t ge[4] = {gq[u].x - g0, gq[u].y - g0, gq[u].z - g0, gq[u].w - g0};
                                const float ve[4] = {vq[u].x, vq[u].y, vq[u].z, vq[u].w};
#pragma unroll
                                for (int e = 0; e < 4; ++e)
                                    if (cu + e >= lo && cu + e < hi && ge[e] < W) A::add(&acc[ge[e]], ve[e], w, scale);
                            }
                        }
                    } else {
                    uint64_t t = lo + lane;
                    for (; t + 32 * (COLLAPSE_UNROLL - 1) < hi; t += 32 * COLLAPSE_UNROLL) {
                        uint32_t gi[COLLAPSE_UNROLL];
                        float vv[COLLAPSE_UNROLL];
#pragma unroll
                        for (int u = 0; u < COLLAPSE_UNROLL; ++u) {
                            gi[u] = __ldg(indices + t + 32 * u) - g0;
                            vv[u] = __ldg(values + t + 32 * u);
                        }
#pragma unroll
                        for (int u = 0; u < COLLAPSE_UNROLL; ++u)
                            if (gi[u] < W) A::add(&acc[gi[u]], vv[u], w, scale);
                    }
                    for (; t + 96 < hi; t += 128) {  // lower tiers: keep several loads in flight for the ragged rest
                        uint32_t gi[4];
                        float vv[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            gi[u] = __ldg(indices + t + 32 * u) - g0;
                            vv[u] = __ldg(values + t + 32 * u);
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u)
                            if (gi[u] < W) A::add(&acc[gi[u]], vv[u], w, scale);
                    }
                    for (; t + 32 < hi; t += 64) {
                        const uint32_t ga = __ldg(indices + t) - g0, gb = __ldg(indices + t + 32) - g0;
                        const float va = __ldg(values + t), vb = __ldg(values + t + 32);
                        if (ga < W) A::add(&acc[ga], va, w, scale);
                        if (gb < W) A::add(&acc[gb], vb, w, scale);
                    }
                    for (; t < hi; t += 32) {
                        const uint32_t gi = __ldg(indices + t) - g0;
                        if (gi < W) A::add(&acc[gi], __ldg(values + t), w, scale);
                    }
                    }
                    wsum += (mult ? mult[cell] : 1.0f);
                    p = pn;
                    cell = cell_n;
                    lo = lo_n;
                    hi = hi_n;
                }
                if (lane == 0 && g0 == 0 && size_s && wsum != 0.0f) atomicAdd(&size_s[lab], wsum);
                __syncthreads();
                for (uint32_t g = threadIdx.x; g < W; g += COLLAPSE_THREADS) {
                    const typename A::T v = acc[g];
                    if (v != 0) {
                        if constexpr (INT) atomicAdd(sum_ds + (size_t)lab * D + g0 + g, (float)v);  // whole numbers: exact in any order
                        else atomicAdd(sum64 + (size_t)lab * D + g0 + g, v);                          // fixed point: exact in any order
                        acc[g] = 0;
                    }
                }
                __syncthreads();
            }
            seg0 = seg1;
        }
    }
}

// ---- K5 from K1's pattern (lg_pattern, lg_common.cuh) ----------------------------------------------------------------
// Inside one pass of the hot path the projection has already turned the CSC stream into a 1-bit pattern (3.75 KB per cell
// against 11.3 KB of indices + values) and a short list of the counts != 1.  The sum of a group is then
//     sum[g, s] = #{cells of s with bit g set}  +  sum over the listed entries (count - 1)
// The first term is a VERTICAL count: thread t owns the 32-gene word t of every cell's row and adds the words of 16 cells
// with a Harley-Seal carry-save tree (30 LOP3 per 16 cells) into nine bit planes; no index is ever decoded and no atomic is
// issued for the 94 % of the entries that are ones.  The listed entries ride along in the same memory round trip (two warps
// per cell of the round) and go to the D-long shared accumulator with ATOMS.ADD.  When the label changes (or the 256-cell
// chunk ends) the planes are expanded into the accumulator — four genes per multiply through the 0x00204081 bit spread —
// and the accumulator leaves with one RED.ADD.F32 per touched gene, exactly as in k_collapse_sorted.  Whole numbers
// throughout: bit-identical to the CSC kernel and to the reference's serial fold.
constexpr int CP_THREADS = 1024;
constexpr int CP_CHUNK = 256;  // sorted cells per work item: the largest count nine planes have to hold
constexpr int CP_ROUND = 16;   // cells per carry-save round (a double-buffered 8-cell form measured slower: 1.40 against 1.23 ms)
constexpr int CP_PAD = CP_CHUNK + CP_ROUND;
constexpr size_t CP_META_BYTES = (size_t)CP_PAD * (8 + 8 + 4 + 4) + (size_t)CP_CHUNK * (4 + 4 + 8 + 4);

// gene g lives at position cp_swz(g) of the accumulator: thread t's 32 genes then fall into 32 different banks for the 32
// threads of a warp (they would all hit bank b otherwise), and a position run of 32 is still one 128-byte line of the output
__device__ __forceinline__ uint32_t cp_swz(uint32_t g) { return (g & ~31u) | ((g + (g >> 5)) & 31u); }
__device__ __forceinline__ uint32_t cp_unswz(uint32_t q) { return (q & ~31u) | ((q - (q >> 5)) & 31u); }
// bits 0..3 of x to bit 0 of bytes 0..3
__device__ __forceinline__ uint32_t cp_spread4(uint32_t x) { return ((x & 0xfu) * 0x00204081u) & 0x01010101u; }
__device__ __forceinline__ void cp_async4(void* dst_smem, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async8(void* dst_smem, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
#define LG_CSA(h, l, a, b, c)                  \
    do {                                       \
        const uint32_t u__ = (a) ^ (b);        \
        h = ((a) & (b)) | (u__ & (c));         \
        l = u__ ^ (c);                         \
    } while (0)

__global__ void __launch_bounds__(CP_THREADS, 1) k_collapse_pattern(
    const uint32_t* __restrict__ bm, uint32_t nchunks_g, const uint64_t* __restrict__ indptr, const uint32_t* __restrict__ exc,
    const uint32_t* __restrict__ exc_cnt, const uint32_t* __restrict__ sorted_label, const uint32_t* __restrict__ sorted_cell,
    uint64_t ncells, uint32_t S, uint64_t D, float* __restrict__ sum_ds, float* __restrict__ size_s,
    unsigned long long* __restrict__ next_chunk) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const uint32_t Dpad = nchunks_g * LG_PAT_GC;
    uint32_t* acc = reinterpret_cast<uint32_t*>(smem_raw);                      // Dpad swizzled accumulators
    const char** s_base = reinterpret_cast<const char**>(acc + Dpad);           // the cell's bitmap row in chunk 0
    uint64_t* s_e0 = reinterpret_cast<uint64_t*>(s_base + CP_PAD);              // first list slot of the cell
    uint32_t* s_cnt = reinterpret_cast<uint32_t*>(s_e0 + CP_PAD);               // listed entries of the cell
    uint32_t* s_label = s_cnt + CP_PAD;
    // the NEXT chunk's metadata, brought in by cp.async while this chunk is summed (element t is written and read by thread t only)
    uint64_t* r_lo = reinterpret_cast<uint64_t*>(s_label + CP_PAD);
    uint32_t* r_cell = reinterpret_cast<uint32_t*>(r_lo + CP_CHUNK);
    uint32_t* r_label = r_cell + CP_CHUNK;
    uint32_t* r_cnt = r_label + CP_CHUNK;
    __shared__ unsigned long long s_chunk[2];
    __shared__ uint32_t s_last[CP_CHUNK / 32 + 1];  // bit p: position p is the last of its label inside the chunk
    const uint32_t t = threadIdx.x;
    const int lane = t & 31, warp = t >> 5;
    const bool active = (t >> 6) < nchunks_g;  // thread t owns bitmap word t & 63 of chunk t >> 6 = genes 32 t .. 32 t + 31
    const uint64_t toff = ((uint64_t)(t >> 6) * (LG_PAT_CELLS * LG_PAT_STRIDE) + (t & 63u)) * 4;  // bytes from the row start to this thread's word
    const uint64_t nchunks = (ncells + CP_CHUNK - 1) / CP_CHUNK;
    for (uint32_t g = t; g < Dpad; g += CP_THREADS) acc[g] = 0;
    auto add_listed = [&](uint32_t x) {
        const uint32_t f = x >> 17;
        atomicAdd(&acc[cp_swz(x & 0x1ffffu)], f == LG_PAT_ZERO ? 0xffffffffu : f);  // a stored zero takes its pattern bit back
    };
    auto cells_of = [&](uint64_t c) { return (int)((c * CP_CHUNK + CP_CHUNK) < ncells ? CP_CHUNK : (ncells - c * CP_CHUNK)); };
    auto meta_level1 = [&](uint64_t c) {  // which cells, which labels
        if (c < nchunks && (int)t < cells_of(c)) {
            cp_async4(&r_cell[t], sorted_cell + c * CP_CHUNK + t);
            cp_async4(&r_label[t], sorted_label + c * CP_CHUNK + t);
        }
    };
    auto meta_level2 = [&](uint64_t c) {  // after this thread's level-1 copies have landed: the cell's extent and list length
        if (c < nchunks && (int)t < cells_of(c)) {
            const uint32_t cell = r_cell[t];
            cp_async8(&r_lo[t], indptr + cell);
            cp_async4(&r_cnt[t], exc_cnt + cell);
        }
    };
    if (t == 0) s_chunk[0] = atomicAdd(next_chunk, 1ull);
    __syncthreads();
    uint64_t chunk = s_chunk[0];
    meta_level1(chunk);
    cp_async_wait_all();
    meta_level2(chunk);
    cp_async_wait_all();
    for (int it = 0; chunk < nchunks; ++it) {
        const int np = cells_of(chunk);
        if ((int)t < np) {  // this chunk's metadata out of the landing zone (own elements: no barrier needed before)
            const uint32_t cell = r_cell[t];
            s_label[t] = r_label[t];
            s_base[t] = reinterpret_cast<const char*>(bm + ((uint64_t)(cell >> 8) * nchunks_g) * (uint64_t)(LG_PAT_CELLS * LG_PAT_STRIDE) +
                                                      (uint64_t)(cell & 255u) * LG_PAT_STRIDE);
            s_e0[t] = (r_lo[t] >> 1) + cell;
            s_cnt[t] = r_cnt[t];
        }
        if (t == 0) s_chunk[(it + 1) & 1] = atomicAdd(next_chunk, 1ull);
        __syncthreads();
        const uint64_t nxt = s_chunk[(it + 1) & 1];
        meta_level1(nxt);
        bool level2_done = false;
        if (t < CP_CHUNK) {
            const bool last = (int)t < np && ((int)t == np - 1 || s_label[t + 1] != s_label[t]);
            const unsigned mk = __ballot_sync(0xffffffffu, last);
            if (lane == 0) s_last[warp] = mk;
        }
        __syncthreads();
        int seg0 = 0;
        while (seg0 < np) {
            const uint32_t lab = s_label[seg0];
            int seg1;
            {
                int w = seg0 >> 5;
                uint32_t mk = s_last[w] & (0xffffffffu << (seg0 & 31));
                while (!mk) mk = s_last[++w];  // position np - 1 is always marked
                seg1 = (w << 5) + __ffs(mk);
            }
            if (lab < S) {
                uint32_t ones = 0, twos = 0, fours = 0, eights = 0, h0 = 0, h1 = 0, h2 = 0, h3 = 0, h4 = 0;
                // the listed entries of cell b + (warp & 15) of a round: this half of the warp pair takes entries 32 half + lane (+ 64 i)
                const uint32_t k0 = ((uint32_t)(warp >> 4) << 5) + lane;
                for (int b = seg0; b < seg1; b += CP_ROUND) {
                    uint32_t w[CP_ROUND];
                    if (active && b + CP_ROUND <= seg1) {  // a full round: sixteen plain loads in flight
#pragma unroll
                        for (int u = 0; u < CP_ROUND; ++u) w[u] = __ldg(reinterpret_cast<const uint32_t*>(s_base[b + u] + toff));
                    } else {
#pragma unroll
                        for (int u = 0; u < CP_ROUND; ++u) {
                            w[u] = 0u;
                            if (active && b + u < seg1) w[u] = __ldg(reinterpret_cast<const uint32_t*>(s_base[b + u] + toff));
                        }
                    }
                    const int pe = b + (warp & 15);
                    uint32_t ecnt = 0;
                    uint64_t ee0 = 0;
                    if (pe < seg1) {
                        ecnt = s_cnt[pe];
                        ee0 = s_e0[pe];
                    }
                    uint32_t x0 = 0u, x1 = 0u;  // no packed word is 0 (field 0 would be a count of one)
                    if (k0 < ecnt) x0 = __ldg(exc + ee0 + k0);
                    if (k0 + 64 < ecnt) x1 = __ldg(exc + ee0 + k0 + 64);
                    uint32_t ta, tb, fa, fb, ea, eb, sx;
                    LG_CSA(ta, ones, ones, w[0], w[1]);
                    LG_CSA(tb, ones, ones, w[2], w[3]);
                    LG_CSA(fa, twos, twos, ta, tb);
                    LG_CSA(ta, ones, ones, w[4], w[5]);
                    LG_CSA(tb, ones, ones, w[6], w[7]);
                    LG_CSA(fb, twos, twos, ta, tb);
                    LG_CSA(ea, fours, fours, fa, fb);
                    LG_CSA(ta, ones, ones, w[8], w[9]);
                    LG_CSA(tb, ones, ones, w[10], w[11]);
                    LG_CSA(fa, twos, twos, ta, tb);
                    LG_CSA(ta, ones, ones, w[12], w[13]);
                    LG_CSA(tb, ones, ones, w[14], w[15]);
                    LG_CSA(fb, twos, twos, ta, tb);
                    LG_CSA(eb, fours, fours, fa, fb);
                    LG_CSA(sx, eights, eights, ea, eb);
                    // the sixteens ripple into the upper planes
                    uint32_t cy = sx, tt;
                    tt = h0 & cy; h0 ^= cy; cy = tt;
                    tt = h1 & cy; h1 ^= cy; cy = tt;
                    tt = h2 & cy; h2 ^= cy; cy = tt;
                    tt = h3 & cy; h3 ^= cy; cy = tt;
                    h4 ^= cy;
                    if (x0) add_listed(x0);
                    if (x1) add_listed(x1);
                    for (uint32_t k = k0 + 128; k < ecnt; k += 64) add_listed(__ldg(exc + ee0 + k));
                }
                if (!level2_done) {  // the next chunk's cells are known by now: ask for their extents (lands during the flushes)
                    cp_async_wait_all();
                    meta_level2(nxt);
                    level2_done = true;
                }
                __syncthreads();  // every listed entry is in
                if (active) {
                    const uint32_t gbase = t << 5;
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const int sh = 4 * q;
                        uint32_t pk = cp_spread4(ones >> sh);
                        pk += cp_spread4(twos >> sh) << 1;
                        pk += cp_spread4(fours >> sh) << 2;
                        pk += cp_spread4(eights >> sh) << 3;
                        pk += cp_spread4(h0 >> sh) << 4;
                        pk += cp_spread4(h1 >> sh) << 5;
                        pk += cp_spread4(h2 >> sh) << 6;
                        pk += cp_spread4(h3 >> sh) << 7;
                        const uint32_t top = (h4 >> sh) & 0xfu;  // a gene present in all 256 cells
                        if (pk | top) {
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const uint32_t cnt = ((pk >> (8 * i)) & 255u) + (((top >> i) & 1u) << 8);
                                if (cnt) acc[gbase + (((uint32_t)(sh + i) + t) & 31u)] += cnt;
                            }
                        }
                    }
                }
                __syncthreads();
                for (uint32_t q = t; q < Dpad; q += CP_THREADS) {
                    const uint32_t v = acc[q];
                    if (v != 0) {
                        const uint32_t gene = cp_unswz(q);
                        if (gene < D) atomicAdd(sum_ds + (size_t)lab * D + gene, (float)v);  // whole numbers: exact in any order
                        acc[q] = 0;
                    }
                }
                if (t == 0 && size_s) atomicAdd(&size_s[lab], (float)(seg1 - seg0));
                __syncthreads();
            }
            seg0 = seg1;
        }
        cp_async_wait_all();
        if (!level2_done) {  // a chunk whose labels were all out of range
            meta_level2(nxt);
            cp_async_wait_all();
        }
        __syncthreads();  // everybody is done with this chunk's tables before the next chunk's are written
        chunk = nxt;
    }
}
#undef LG_CSA

__global__ void k_count_bs(const uint32_t* __restrict__ group, const uint32_t* __restrict__ batch,
                           const float* __restrict__ mult, uint64_t ncols, uint32_t S, uint32_t B,
                           float* __restrict__ n_bs) {
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= ncols) return;
    const uint32_t s = group[j], b = batch[j];
    if (s < S && b < B) atomicAdd(&n_bs[(size_t)s * B + b], mult ? mult[j] : 1.0f);
}
// the same with registered multiplicities: fixed point (integer adds: any arrival order gives the same bits), converted after
__global__ void k_count_bs_fixed(const uint32_t* __restrict__ group, const uint32_t* __restrict__ batch,
                                 const float* __restrict__ mult, uint64_t ncols, uint32_t S, uint32_t B, float scale,
                                 unsigned long long* __restrict__ n64) {
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= ncols) return;
    const uint32_t s = group[j], b = batch[j];
    if (s < S && b < B) atomicAdd(&n64[(size_t)s * B + b], (unsigned long long)__float2ll_rn(__fmul_rn(mult[j], scale)));
}

// first sorted position whose label is >= split: out = {0, c, c, n} (the two position ranges of the split collapse)
__global__ void k_split_point(const uint32_t* __restrict__ sorted_label, uint32_t n, uint32_t split, uint32_t* __restrict__ out) {
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        const uint32_t mid = lo + (hi - lo) / 2;
        if (sorted_label[mid] < split) lo = mid + 1;
        else hi = mid;
    }
    out[0] = 0;
    out[1] = lo;
    out[2] = lo;
    out[3] = n;
}

static int collapse_by_label_pattern(lg_ctx* ctx, LgStage& st, const lg_csc* m, const uint32_t* d_label, uint32_t S, float* d_sum,
                                     float* d_size, const lg_pattern* pat);
static int twin_usable(lg_ctx* ctx, const lg_csc* m, bool* yes);

// generic "sum columns by label" driver shared by the basic (label = group) and batch (label = batch) stats
static int collapse_by_label(lg_ctx* ctx, LgStage& st, const lg_csc* m, const uint32_t* d_label, const float* d_mult,
                             uint32_t S, float* d_sum, float* d_size, uint32_t S_half = 0,
                             const std::function<int()>* after_first_half = nullptr) {
    const uint64_t N = m->ncols, D = m->nrows;
    if (!d_mult && !after_first_half && N && D && S) {  // the block keeps a pattern and a projection has filled it
        bool yes = false;
        LG_TRY(twin_usable(ctx, m, &yes));
        if (yes) return collapse_by_label_pattern(ctx, st, m, d_label, S, d_sum, d_size, m->twin);
    }
    LG_CUDA(ctx, cudaMemsetAsync(d_sum, 0, sizeof(float) * (size_t)D * S, ctx->stream));
    if (d_size) LG_CUDA(ctx, cudaMemsetAsync(d_size, 0, sizeof(float) * S, ctx->stream));
    if (N == 0 || D == 0 || S == 0) {  // an empty shard still takes its part in the caller's exchange
        if (after_first_half) LG_TRY((*after_first_half)());
        return LG_OK;
    }
    LG_REQUIRE(ctx, N < 0xFFFFFFFFull, "collapse: more than 2^32-1 cells in one block; shard the cells");
    // stable sort of cells by label (cells stay ascending inside a label)
    uint32_t *d_cell_in, *d_cell_out, *d_lab_out;
    LG_TRY(st.scratch(N, &d_cell_in));
    LG_TRY(st.scratch(N, &d_cell_out));
    LG_TRY(st.scratch(N, &d_lab_out));
    LG_LAUNCH(ctx, k_iota_u32, (unsigned)((N + 255) / 256), 256, 0, d_cell_in, N);
    size_t tmp_bytes = 0;
    LG_CUDA(ctx, cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_label, d_lab_out, d_cell_in, d_cell_out, (int)N, 0,
                                                 32, ctx->stream));
    char* d_tmp;
    LG_TRY(st.scratch(tmp_bytes, &d_tmp));
    LG_CUDA(ctx, cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, d_label, d_lab_out, d_cell_in, d_cell_out, (int)N, 0,
                                                 32, ctx->stream));
    ctx->launches += 4;  // cub's radix passes (histogram + onesweep), counted conservatively
    // integer-valued counts with unit multiplicity take the exact u32 accumulator
    bool use_int = false;
    if (!d_mult && m->nnz && m->int_valued >= 0) use_int = m->int_valued == 1;
    else if (!d_mult && m->nnz) {
        int* d_flag;
        LG_TRY(st.scratch(1, &d_flag));
        LG_CUDA(ctx, cudaMemsetAsync(d_flag, 0, sizeof(int), ctx->stream));
        LG_LAUNCH(ctx, k_all_integral, ctx->num_sms * 8, 256, 0, m->values, m->nnz, d_flag);
        int* h_flag = static_cast<int*>(ctx->pinned);
        LG_CUDA(ctx, cudaMemcpyAsync(h_flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        LG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        use_int = (*h_flag == 0);
        m->int_valued = use_int ? 1 : 0;
    }
    unsigned long long* d_next;
    LG_TRY(st.scratch(1, &d_next));
    const size_t smem_cap = ctx->smem_optin - 4096;
    // the fixed-point path: scale 2^e with max|v| * max|w| * N * 2^e < 2^62 (a (gene, group) cell gets at most one entry per cell)
    unsigned long long* d_sum64 = nullptr;
    float scale = 1.0f;
    double inv_scale = 1.0;
    float* d_size_kernel = d_size;
    if (!use_int) {
        unsigned int* d_mx;
        LG_TRY(st.scratch(2, &d_mx));
        LG_CUDA(ctx, cudaMemsetAsync(d_mx, 0, 2 * sizeof(unsigned int), ctx->stream));
        if (m->nnz) LG_LAUNCH(ctx, k_max_abs_bits, ctx->num_sms * 8, 256, 0, m->values, m->nnz, d_mx);
        if (d_mult) LG_LAUNCH(ctx, k_max_abs_bits, ctx->num_sms * 2, 256, 0, d_mult, N, d_mx + 1);
        unsigned int* h_mx = static_cast<unsigned int*>(ctx->pinned);
        LG_CUDA(ctx, cudaMemcpyAsync(h_mx, d_mx, 2 * sizeof(unsigned int), cudaMemcpyDeviceToHost, ctx->stream));
        LG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        float mv, mw = 1.0f;
        memcpy(&mv, &h_mx[0], 4);
        if (d_mult) memcpy(&mw, &h_mx[1], 4);
        const double bound = (double)mv * (double)mw * (double)N;
        int e = 24;
        if (bound > 0.0) {
            int be = 0;
            frexp(bound, &be);  // bound < 2^be
            e = 62 - be;
        }
        e = e > 100 ? 100 : (e < -100 ? -100 : e);
        scale = ldexpf(1.0f, e);
        inv_scale = ldexp(1.0, -e);
        LG_TRY(st.scratch((size_t)D * S, &d_sum64));
        LG_CUDA(ctx, cudaMemsetAsync(d_sum64, 0, sizeof(unsigned long long) * (size_t)D * S, ctx->stream));
        if (d_mult && d_size) {  // the reference's serial fold over the group's cells instead of f32 atomics in arrival order
            LG_LAUNCH(ctx, k_size_by_group, (S + 127) / 128, 128, 0, d_lab_out, d_cell_out, (uint32_t)N, d_mult, S, d_size);
            d_size_kernel = nullptr;
        }
    }
    const uint32_t Wmax = (uint32_t)(smem_cap / (use_int ? sizeof(unsigned int) : sizeof(unsigned long long)));
    // two launches over the sorted positions (labels below S_half, then the rest) when a caller wants the first half early and
    // the genes fit one pass; the position ranges stay on the device
    const bool split = after_first_half && S_half > 0 && S_half < S && D <= Wmax && use_int;
    uint32_t* d_range = nullptr;
    if (split) {
        LG_TRY(st.scratch(4, &d_range));
        LG_LAUNCH(ctx, k_split_point, 1, 1, 0, d_lab_out, (uint32_t)N, S_half, d_range);
    }
    for (uint64_t g0 = 0; g0 < D; g0 += Wmax) {
        const uint32_t W = (uint32_t)((D - g0) < Wmax ? (D - g0) : Wmax);
        const size_t smem = (size_t)W * (use_int ? sizeof(unsigned int) : sizeof(unsigned long long));
        const bool vec = (((uintptr_t)m->indices | (uintptr_t)m->values) & 15) == 0;
        for (int half = 0; half < (split ? 2 : 1); ++half) {
            LG_CUDA(ctx, cudaMemsetAsync(d_next, 0, sizeof(unsigned long long), ctx->stream));
            const uint32_t* rng = split ? d_range + 2 * half : nullptr;
#define LG_COLLAPSE_LAUNCH(I, V)                                                                                                   \
    do {                                                                                                                           \
        LG_CUDA(ctx, cudaFuncSetAttribute(k_collapse_sorted<I, V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));       \
        LG_LAUNCH(ctx, (k_collapse_sorted<I, V>), ctx->num_sms, COLLAPSE_THREADS, smem, m->indptr, m->indices, m->values, m->nnz, \
                  d_lab_out, d_cell_out, N, d_mult, S, D, (uint32_t)g0, W, d_sum, d_size_kernel, d_next, rng, d_sum64, scale);      \
    } while (0)
            if (use_int && vec) LG_COLLAPSE_LAUNCH(true, true);
            else if (use_int) LG_COLLAPSE_LAUNCH(true, false);
            else if (vec) LG_COLLAPSE_LAUNCH(false, true);
            else LG_COLLAPSE_LAUNCH(false, false);
#undef LG_COLLAPSE_LAUNCH
            if (split && half == 0) LG_TRY((*after_first_half)());
        }
    }
    if (d_sum64) {
        const uint64_t total = D * (uint64_t)S;
        LG_LAUNCH(ctx, k_fixed_to_f32, (unsigned)((total + 255) / 256), 256, 0, d_sum64, total, inv_scale, d_sum);
    }
    if (after_first_half && !split) LG_TRY((*after_first_half)());  // nothing was split: the callback still runs once, before the caller's second step
    return LG_OK;
}

// can k_collapse_pattern take this block?  (thread t of 1024 owns bitmap word t: at most 16 chunks of 2048 genes)
bool lg_collapse_pattern_fits(const lg_ctx* ctx, uint64_t D, uint64_t N) {
    const uint64_t nch = (D + LG_PAT_GC - 1) / LG_PAT_GC;
    const size_t smem = (size_t)nch * LG_PAT_GC * 4 + CP_META_BYTES;
    return D > 0 && N > 0 && N < 0xFFFFFFFFull && nch * (LG_PAT_GC / 32) <= CP_THREADS && smem + 4096 <= ctx->smem_optin;
}

// sums by label (unit multiplicities) from a filled pattern whose flag the caller has read as 0; device pointers
static int collapse_by_label_pattern(lg_ctx* ctx, LgStage& st, const lg_csc* m, const uint32_t* d_label, uint32_t S, float* d_sum,
                                     float* d_size, const lg_pattern* pat) {
    const uint64_t N = m->ncols, D = m->nrows;
    LG_REQUIRE(ctx, lg_collapse_pattern_fits(ctx, D, N) && pat->nchunks == (uint32_t)((D + LG_PAT_GC - 1) / LG_PAT_GC),
               "collapse: block outside the pattern kernel's range");
    LG_CUDA(ctx, cudaMemsetAsync(d_sum, 0, sizeof(float) * (size_t)D * S, ctx->stream));
    if (d_size) LG_CUDA(ctx, cudaMemsetAsync(d_size, 0, sizeof(float) * S, ctx->stream));
    if (S == 0) return LG_OK;
    uint32_t *d_cell_in, *d_cell_out, *d_lab_out;
    LG_TRY(st.scratch(N, &d_cell_in));
    LG_TRY(st.scratch(N, &d_cell_out));
    LG_TRY(st.scratch(N, &d_lab_out));
    LG_LAUNCH(ctx, k_iota_u32, (unsigned)((N + 255) / 256), 256, 0, d_cell_in, N);
    size_t tmp_bytes = 0;
    LG_CUDA(ctx, cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_label, d_lab_out, d_cell_in, d_cell_out, (int)N, 0, 32, ctx->stream));
    char* d_tmp;
    LG_TRY(st.scratch(tmp_bytes, &d_tmp));
    LG_CUDA(ctx, cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, d_label, d_lab_out, d_cell_in, d_cell_out, (int)N, 0, 32, ctx->stream));
    ctx->launches += 4;
    unsigned long long* d_next;
    LG_TRY(st.scratch(1, &d_next));
    LG_CUDA(ctx, cudaMemsetAsync(d_next, 0, sizeof(unsigned long long), ctx->stream));
    const size_t smem = (size_t)pat->nchunks * LG_PAT_GC * 4 + CP_META_BYTES;
    LG_CUDA(ctx, cudaFuncSetAttribute(k_collapse_pattern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const uint64_t nwork = (N + CP_CHUNK - 1) / CP_CHUNK;
    const unsigned grid = (unsigned)(nwork < (uint64_t)ctx->num_sms ? nwork : (uint64_t)ctx->num_sms);
    LG_LAUNCH(ctx, k_collapse_pattern, grid, CP_THREADS, smem, pat->bm, pat->nchunks, m->indptr, pat->exc, pat->exc_cnt, d_lab_out,
              d_cell_out, N, S, D, d_sum, d_size, d_next);
    ctx->pattern_collapses++;
    return LG_OK;
}

// lg_collapse_basic (unit multiplicities) from the pattern K1 left behind in this pass; the caller has checked pat->filled and
// read pat->ovf as 0
int lg_collapse_basic_pattern(lg_ctx* ctx, const lg_csc* m, const uint32_t* group_of_cell, uint32_t S, float* out_sum_ds,
                              float* out_size_s, const lg_pattern* pat) {
    LG_REQUIRE(ctx, m && group_of_cell && out_sum_ds && out_size_s && pat && pat->filled, "lg_collapse_basic_pattern: null argument");
    LgStage st(ctx);
    const uint32_t* d_label;
    float *d_sum, *d_size;
    LG_TRY(st.in(group_of_cell, (size_t)m->ncols, &d_label));
    LG_TRY(st.out(out_sum_ds, (size_t)m->nrows * S, &d_sum));
    LG_TRY(st.out(out_size_s, (size_t)S, &d_size));
    LG_TRY(collapse_by_label_pattern(ctx, st, m, d_label, S, d_sum, d_size, pat));
    return st.finish();
}

// the block's own pattern (lg_csc_keep_pattern), if a projection has filled it and it can express the block: 1, else 0
static int twin_usable(lg_ctx* ctx, const lg_csc* m, bool* yes) {
    *yes = false;
    const char* pz = getenv("LG_COLLAPSE_PATTERN");
    if (!m->twin || !m->twin->filled || (pz && pz[0] == '0')) return LG_OK;
    if (m->twin_ovf < 0) {  // once per projection
        int* h = reinterpret_cast<int*>(static_cast<char*>(ctx->pinned) + 256);
        LG_CUDA(ctx, cudaMemcpyAsync(h, m->twin->ovf, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        LG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        m->twin_ovf = *h ? 1 : 0;
    }
    *yes = m->twin_ovf == 0;
    return LG_OK;
}

int lg_collapse_basic_split(lg_ctx* ctx, const lg_csc* m, const uint32_t* d_group, uint32_t S, uint32_t S_half, float* d_sum_ds,
                            float* d_size_s, const std::function<int()>& after_first_half) {
    LgStage st(ctx);
    LG_TRY(collapse_by_label(ctx, st, m, d_group, nullptr, S, d_sum_ds, d_size_s, S_half, &after_first_half));
    return st.finish();
}

extern "C" int lg_collapse_basic(lg_ctx* ctx, const lg_csc* m, const uint32_t* group_of_cell, const float* mult, uint32_t S,
                                 float* out_sum_ds, float* out_size_s) {
    if (!ctx) return LG_ERR_INVALID;
    LG_REQUIRE(ctx, m && group_of_cell && out_sum_ds && out_size_s, "lg_collapse_basic: null argument");
    cudaSetDevice(ctx->device);
    LgStage st(ctx);
    const uint32_t* d_group;
    const float* d_mult;
    float *d_sum, *d_size;
    LG_TRY(st.in(group_of_cell, (size_t)m->ncols, &d_group));
    LG_TRY(st.in(mult, (size_t)m->ncols, &d_mult));
    LG_TRY(st.out(out_sum_ds, (size_t)m->nrows * S, &d_sum));
    LG_TRY(st.out(out_size_s, (size_t)S, &d_size));
    LG_TRY(collapse_by_label(ctx, st, m, d_group, d_mult, S, d_sum, d_size));
    return st.finish();
}

extern "C" int lg_collapse_batch(lg_ctx* ctx, const lg_csc* m, const uint32_t* group_of_cell, const uint32_t* batch_of_cell,
                                 const float* mult, uint32_t S, uint32_t B, float* out_sum_db, float* out_n_bs) {
    if (!ctx) return LG_ERR_INVALID;
    LG_REQUIRE(ctx, m && group_of_cell && batch_of_cell && out_sum_db && out_n_bs, "lg_collapse_batch: null argument");
    cudaSetDevice(ctx->device);
    LgStage st(ctx);
    const uint32_t *d_group, *d_batch;
    const float* d_mult;
    float *d_sum, *d_nbs;
    LG_TRY(st.in(group_of_cell, (size_t)m->ncols, &d_group));
    LG_TRY(st.in(batch_of_cell, (size_t)m->ncols, &d_batch));
    LG_TRY(st.in(mult, (size_t)m->ncols, &d_mult));
    LG_TRY(st.out(out_sum_db, (size_t)m->nrows * B, &d_sum));
    LG_TRY(st.out(out_n_bs, (size_t)B * S, &d_nbs));
    LG_TRY(collapse_by_label(ctx, st, m, d_batch, d_mult, B, d_sum, nullptr));
    LG_CUDA(ctx, cudaMemsetAsync(d_nbs, 0, sizeof(float) * (size_t)B * S, ctx->stream));
    if (m->ncols && !d_mult) {
        LG_LAUNCH(ctx, k_count_bs, (unsigned)((m->ncols + 255) / 256), 256, 0, d_group, d_batch, d_mult, m->ncols, S, B, d_nbs);
    } else if (m->ncols) {
        unsigned int* d_mx;
        unsigned long long* d_n64;
        LG_TRY(st.scratch(1, &d_mx));
        LG_TRY(st.scratch((size_t)B * S, &d_n64));
        LG_CUDA(ctx, cudaMemsetAsync(d_mx, 0, sizeof(unsigned int), ctx->stream));
        LG_CUDA(ctx, cudaMemsetAsync(d_n64, 0, sizeof(unsigned long long) * (size_t)B * S, ctx->stream));
        LG_LAUNCH(ctx, k_max_abs_bits, ctx->num_sms * 2, 256, 0, d_mult, m->ncols, d_mx);
        unsigned int* h_mx = static_cast<unsigned int*>(ctx->pinned);
        LG_CUDA(ctx, cudaMemcpyAsync(h_mx, d_mx, sizeof(unsigned int), cudaMemcpyDeviceToHost, ctx->stream));
        LG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        float mw;
        memcpy(&mw, h_mx, 4);
        int e = 24, be = 0;
        if (mw > 0.0f) {
            frexp((double)mw * (double)m->ncols, &be);
            e = 62 - be;
        }
        e = e > 100 ? 100 : (e < -100 ? -100 : e);
        LG_LAUNCH(ctx, k_count_bs_fixed, (unsigned)((m->ncols + 255) / 256), 256, 0, d_group, d_batch, d_mult, m->ncols, S, B,
                  ldexpf(1.0f, e), d_n64);
        LG_LAUNCH(ctx, k_fixed_to_f32, (unsigned)(((size_t)B * S + 255) / 256), 256, 0, d_n64, (uint64_t)B * S, ldexp(1.0, -e), d_nbs);
    }
    return st.finish();
}

__global__ void k_merge_stat(const float* __restrict__ fine, uint64_t D, const uint32_t* __restrict__ coarse_off,
                             const uint32_t* __restrict__ coarse_fine, uint32_t ncoarse, float* __restrict__ coarse) {
    // one thread per (gene, coarse) output; fine columns are summed in ascending fine index, as stats.rs:798-812 does
    const uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= D * ncoarse) return;
    const uint64_t g = e % D;
    const uint32_t c = (uint32_t)(e / D);
    float s = 0.0f;
    for (uint32_t i = coarse_off[c]; i < coarse_off[c + 1]; ++i) s = __fadd_rn(s, fine[(size_t)coarse_fine[i] * D + g]);
    coarse[e] = s;
}

extern "C" int lg_merge_stat(lg_ctx* ctx, const float* fine_ds, uint64_t nrows, uint32_t nfine, const uint32_t* fine_to_coarse,
                             uint32_t ncoarse, float* out_coarse_ds) {
    if (!ctx) return LG_ERR_INVALID;
    LG_REQUIRE(ctx, fine_ds && fine_to_coarse && out_coarse_ds, "lg_merge_stat: null argument");
    cudaSetDevice(ctx->device);
    LgStage st(ctx);
    const float* d_fine;
    float* d_coarse;
    LG_TRY(st.in(fine_ds, (size_t)nrows * nfine, &d_fine));
    LG_TRY(st.out(out_coarse_ds, (size_t)nrows * ncoarse, &d_coarse));
    // the map is tiny (<= 2^kk entries): bucket the fine columns by coarse column on the host, ascending inside a bucket
    std::vector<uint32_t> f2c(nfine);
    if (nfine) {
        LG_CUDA(ctx, cudaMemcpyAsync(f2c.data(), fine_to_coarse, sizeof(uint32_t) * nfine, cudaMemcpyDefault, ctx->stream));
        LG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    std::vector<uint32_t> off(ncoarse + 1, 0), lst(nfine);
    for (uint32_t f = 0; f < nfine; ++f) {
        LG_REQUIRE(ctx, f2c[f] < ncoarse, "lg_merge_stat: fine_to_coarse out of range");
        off[f2c[f] + 1]++;
    }
    for (uint32_t c = 0; c < ncoarse; ++c) off[c + 1] += off[c];
    {
        std::vector<uint32_t> cur(off.begin(), off.end() - 1);
        for (uint32_t f = 0; f < nfine; ++f) lst[cur[f2c[f]]++] = f;
    }
    uint32_t *d_off, *d_lst;
    LG_TRY(st.scratch(off.size(), &d_off));
    LG_TRY(st.scratch(lst.size(), &d_lst));
    LG_CUDA(ctx, cudaMemcpyAsync(d_off, off.data(), sizeof(uint32_t) * off.size(), cudaMemcpyHostToDevice, ctx->stream));
    if (nfine) LG_CUDA(ctx, cudaMemcpyAsync(d_lst, lst.data(), sizeof(uint32_t) * nfine, cudaMemcpyHostToDevice, ctx->stream));
    LG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // the host vectors go out of scope below
    const uint64_t total = nrows * ncoarse;
    if (total) LG_LAUNCH(ctx, k_merge_stat, (unsigned)((total + 255) / 256), 256, 0, d_fine, nrows, d_off, d_lst, ncoarse, d_coarse);
    return st.finish();
}
