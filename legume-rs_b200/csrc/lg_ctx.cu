// lg_ctx.cu — context lifecycle and the device-resident CSC container (the data feed).
#include <immintrin.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdlib>
#include <thread>

#include "lg_common.cuh"

extern "C" const char* lg_version(void) { return "legume-b200 0.1.0 (sm_100a)"; }

extern "C" int lg_ctx_create(int device, lg_ctx** out) {
    if (!out) return LG_ERR_INVALID;
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0 || device < 0 || device >= n) {
        // no CPU fallback: the product path needs a CUDA device
        return LG_ERR_CUDA;
    }
    lg_ctx* c = new lg_ctx();
    c->device = device;
    if (cudaSetDevice(device) != cudaSuccess) {
        delete c;
        return LG_ERR_CUDA;
    }
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, device);
    c->num_sms = p.multiProcessorCount;
    c->smem_optin = p.sharedMemPerBlockOptin;
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete c;
        return LG_ERR_CUDA;
    }
    c->own_stream = true;
    // keep stream-ordered scratch cached across synchronisations (the default threshold of 0
    // hands every freed block back to the driver at each sync, which costs milliseconds per call)
    {
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            uint64_t keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
    }
    c->pinned_bytes = 1 << 16;
    if (cudaMallocHost(&c->pinned, c->pinned_bytes) != cudaSuccess) {
        cudaStreamDestroy(c->stream);
        delete c;
        return LG_ERR_NOMEM;
    }
    *out = c;
    return LG_OK;
}

extern "C" int lg_ctx_destroy(lg_ctx* c) {
    if (!c) return LG_OK;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    if (c->own_stream) cudaStreamDestroy(c->stream);
    if (c->pinned) cudaFreeHost(c->pinned);
    if (c->ring) cudaFreeHost(c->ring);
    if (c->log1p_tab) cudaFree(c->log1p_tab);
    for (auto& b : c->cache) cudaFree(b.p);
    for (auto e : c->ring_ev) cudaEventDestroy(e);
    for (auto e : c->side_ev)
        if (e) cudaEventDestroy(e);
    if (c->side_stream) cudaStreamDestroy(c->side_stream);
    delete c;
    return LG_OK;
}

extern "C" const char* lg_last_error(const lg_ctx* c) { return c ? c->err.c_str() : "null context"; }

extern "C" int lg_ctx_set_stream(lg_ctx* c, void* s) {
    if (!c) return LG_ERR_INVALID;
    cudaSetDevice(c->device);
    if (c->own_stream) {
        cudaStreamSynchronize(c->stream);
        cudaStreamDestroy(c->stream);
        c->own_stream = false;
    } else if (c->stream != static_cast<cudaStream_t>(s)) {
        cudaStreamSynchronize(c->stream);  // cached scratch is handed out again in the order of ONE stream
    }
    // NULL is the legacy default stream (what torch reports for its default stream)
    c->stream = static_cast<cudaStream_t>(s);
    return LG_OK;
}

extern "C" int lg_ctx_sync(lg_ctx* c) {
    if (!c) return LG_ERR_INVALID;
    LG_CUDA(c, cudaStreamSynchronize(c->stream));
    return LG_OK;
}

extern "C" uint64_t lg_ctx_launch_count(const lg_ctx* c) { return c ? c->launches : 0; }
extern "C" uint64_t lg_ctx_h2d_bytes(const lg_ctx* c) { return c ? c->h2d_bytes : 0; }
extern "C" uint64_t lg_ctx_fallback_count(const lg_ctx* c) { return c ? c->fallbacks : 0; }
extern "C" uint64_t lg_ctx_pattern_collapse_count(const lg_ctx* c) { return c ? c->pattern_collapses : 0; }
extern "C" void lg_ctx_time_stages(lg_ctx* c, int on) {
    if (c) {
        c->time_stages = on != 0;
        c->stage_ms_valid = false;
    }
}
extern "C" int lg_hotpath_last_stage_ms(const lg_ctx* c, float* out6) {
    if (!c || !out6 || !c->stage_ms_valid) return LG_ERR_INVALID;
    for (int i = 0; i < 6; ++i) out6[i] = c->stage_ms[i];
    return LG_OK;
}
extern "C" const char* lg_ctx_last_fallback(const lg_ctx* c) { return c ? c->last_fallback.c_str() : ""; }

// ---- CSC container ----------------------------------------------------------------------------
// narrow u64 row indices to u32 (optionally through a row remap), flagging out-of-range rows
__global__ void k_narrow_indices(const uint64_t* __restrict__ src, uint32_t* __restrict__ dst, uint64_t n,
                                 uint64_t nrows_in, uint64_t nrows_out, const uint32_t* __restrict__ remap,
                                 int* __restrict__ flag) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        uint64_t r = src[i];
        if (r >= nrows_in) {
            atomicOr(flag, 1);
            r = 0;
        }
        uint32_t o = remap ? remap[r] : (uint32_t)r;
        if ((uint64_t)o >= nrows_out) {
            atomicOr(flag, 2);
            o = 0;
        }
        dst[i] = o;
    }
}
struct LgUpPatch {
    uint32_t pos;
    float val;
};
__global__ void k_expand_u8(const uint8_t* __restrict__ src, float* __restrict__ dst, uint64_t n,
                            const LgUpPatch* __restrict__ patch, uint32_t npatch) {
    // n4 full groups of four bytes, then the ragged tail; byte 255 marks a patched position (not written here)
    const uint64_t n4 = n >> 2;
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t g = i; g < n4; g += stride) {
        const uint32_t w = reinterpret_cast<const uint32_t*>(src)[g];
        const uint32_t b0 = w & 255u, b1 = (w >> 8) & 255u, b2 = (w >> 16) & 255u, b3 = w >> 24;
        if (b0 != 255u && b1 != 255u && b2 != 255u && b3 != 255u) {
            reinterpret_cast<float4*>(dst)[g] = make_float4((float)b0, (float)b1, (float)b2, (float)b3);
        } else {
            if (b0 != 255u) dst[4 * g] = (float)b0;
            if (b1 != 255u) dst[4 * g + 1] = (float)b1;
            if (b2 != 255u) dst[4 * g + 2] = (float)b2;
            if (b3 != 255u) dst[4 * g + 3] = (float)b3;
        }
    }
    for (uint64_t t = (n4 << 2) + i; t < n; t += stride)
        if (src[t] != 255) dst[t] = (float)src[t];
    for (uint64_t t = i; t < npatch; t += stride) dst[patch[t].pos] = patch[t].val;
}
// Row indices that travelled as one-byte gaps (upload_narrow_on_host): entry i is the entry before it plus byte i; byte 0
// marks an entry whose gap does not fit (a column start, a gap above 255) and is found in the patch list as
// (position, gap modulo 2^32).  Every tile of LG_UP_TILE entries carries the absolute value of the entry before it
// (its anchor), so tiles decode independently: one inclusive sum per tile, 8 consecutive entries per thread.
constexpr uint32_t LG_UP_TILE = 2048;
struct LgUpIPatch {
    uint32_t pos;
    uint32_t gap;
};
__global__ void __launch_bounds__(256) k_expand_idx_gaps(const uint8_t* __restrict__ bytes, const uint32_t* __restrict__ anchors,
                                                         const LgUpIPatch* __restrict__ patch, uint32_t npatch,
                                                         uint32_t* __restrict__ dst, uint32_t n) {
    __shared__ uint32_t wsum[8];
    const uint32_t base = blockIdx.x * LG_UP_TILE + threadIdx.x * 8;
    uint32_t d[8];
    if (base + 8 <= n) {
        const uint2 w = *reinterpret_cast<const uint2*>(bytes + base);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            d[k] = (w.x >> (8 * k)) & 255u;
            d[4 + k] = (w.y >> (8 * k)) & 255u;
        }
    } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) d[k] = base + k < n ? bytes[base + k] : 1u;
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        if (d[k] == 0u && base + k < n) {  // rare: look the gap up (positions ascend in the list)
            uint32_t lo = 0, hi = npatch;
            while (lo < hi) {
                const uint32_t mid = (lo + hi) >> 1;
                if (patch[mid].pos < base + k) lo = mid + 1;
                else hi = mid;
            }
            d[k] = patch[lo].gap;
        }
    }
#pragma unroll
    for (int k = 1; k < 8; ++k) d[k] += d[k - 1];
    uint32_t run = d[7];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, run, off);
        if (lane >= off) run += t;
    }
    if (lane == 31) wsum[warp] = run;
    __syncthreads();
    uint32_t before = anchors[blockIdx.x] + run - d[7];
    for (int w = 0; w < warp; ++w) before += wsum[w];
    if (base + 8 <= n) {
        reinterpret_cast<uint4*>(dst + base)[0] = make_uint4(before + d[0], before + d[1], before + d[2], before + d[3]);
        reinterpret_cast<uint4*>(dst + base)[1] = make_uint4(before + d[4], before + d[5], before + d[6], before + d[7]);
    } else {
#pragma unroll
        for (int k = 0; k < 8; ++k)
            if (base + k < n) dst[base + k] = before + d[k];
    }
}
__global__ void k_rebase_indptr(const uint64_t* __restrict__ src, uint64_t* __restrict__ dst, uint64_t n,
                                uint64_t base) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[i] - base;
}
__global__ void k_widen_indices(const uint32_t* __restrict__ src, uint64_t* __restrict__ dst, uint64_t n) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) dst[i] = src[i];
}

// ---- host-side narrowing ------------------------------------------------------------------------
// The reference hands out u64 row indices (csc_column_arrays, data-beans/src/sparse_io/traits.rs:98-100); on the wire
// they are two thirds of the upload.  Worker threads narrow them to u32 into a pinned ring, chunk by chunk, and queue
// the copies themselves, so the PCIe link carries 8 instead of 12 bytes per non-zero while the value chunks (which
// need no CPU work) keep it busy.  LG_UPLOAD_THREADS=0 selects the device-side narrowing of wide copies instead.
namespace {
constexpr uint64_t UP_CHUNK = 2ull << 20;          // non-zeros per ring slot
constexpr uint64_t UP_PATCH_BYTES = 8ull << 16;     // UP_PATCH_CAP (position, f32) pairs
constexpr uint64_t UP_SLOT_BYTES = UP_CHUNK * 5 + UP_PATCH_BYTES;       // u32 indices, u8 values, patch list
constexpr uint64_t UP_ANCHOR_BYTES = (UP_CHUNK / 2048) * 4;             // one u32 per LG_UP_TILE entries
constexpr uint64_t UP_GAP_SLOT = UP_ANCHOR_BYTES + UP_CHUNK + 8 + UP_PATCH_BYTES;  // device twin of a slot's index part as gaps
static_assert(UP_ANCHOR_BYTES + UP_CHUNK + 2 * UP_PATCH_BYTES + 8 <= UP_CHUNK * 4, "the gap form fits the slot's u32 index region");
constexpr uint64_t UP_SLOT_BYTES_RAW = UP_SLOT_BYTES + UP_CHUNK * 4;    // + raw f32 staging (pageable sources only)

__attribute__((target("avx2"))) void narrow_chunk_avx2(const uint64_t* s, uint32_t* d, uint64_t n, uint64_t* or_all,
                                                       uint32_t* max_lo) {
    const __m256i pick = _mm256_setr_epi32(0, 2, 4, 6, 0, 2, 4, 6);
    __m256i acc = _mm256_setzero_si256(), mx = _mm256_setzero_si256();
    uint64_t i = 0;
    const bool aligned = (reinterpret_cast<uintptr_t>(d) & 31) == 0;
    for (; i + 8 <= n; i += 8) {
        const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s + i));
        const __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s + i + 4));
        acc = _mm256_or_si256(acc, _mm256_or_si256(a, b));
        const __m256i pa = _mm256_permutevar8x32_epi32(a, pick), pb = _mm256_permutevar8x32_epi32(b, pick);
        const __m256i v = _mm256_blend_epi32(pa, pb, 0xF0);
        mx = _mm256_max_epu32(mx, v);
        if (aligned) _mm256_stream_si256(reinterpret_cast<__m256i*>(d + i), v);
        else _mm256_storeu_si256(reinterpret_cast<__m256i*>(d + i), v);
    }
    alignas(32) uint64_t a4[4];
    alignas(32) uint32_t m8[8];
    _mm256_store_si256(reinterpret_cast<__m256i*>(a4), acc);
    _mm256_store_si256(reinterpret_cast<__m256i*>(m8), mx);
    uint64_t o = a4[0] | a4[1] | a4[2] | a4[3];
    uint32_t m = 0;
    for (int k = 0; k < 8; ++k) m = m8[k] > m ? m8[k] : m;
    for (; i < n; ++i) {
        o |= s[i];
        const uint32_t v = (uint32_t)s[i];
        m = v > m ? v : m;
        d[i] = v;
    }
    _mm_sfence();
    *or_all |= o;
    *max_lo = m > *max_lo ? m : *max_lo;
}
void narrow_chunk_plain(const uint64_t* s, uint32_t* d, uint64_t n, uint64_t* or_all, uint32_t* max_lo) {
    uint64_t o = 0;
    uint32_t m = 0;
    for (uint64_t i = 0; i < n; ++i) {
        o |= s[i];
        const uint32_t v = (uint32_t)s[i];
        m = v > m ? v : m;
        d[i] = v;
    }
    *or_all |= o;
    *max_lo = m > *max_lo ? m : *max_lo;
}

// count data: values that are whole numbers in 0..254 travel as one byte each (exactly); anything else is marked 255
// and goes into the chunk's patch list as (position, f32).  Returns the number of patches, or -1 when the list is full
// (the chunk then goes up as raw f32).
struct UpPatch {
    uint32_t pos;
    float val;
};
constexpr int64_t UP_PATCH_CAP = 1 << 16;

inline int64_t pack_scalar(const float* s, uint8_t* d, uint64_t lo, uint64_t hi, UpPatch* patch, int64_t np) {
    for (uint64_t i = lo; i < hi; ++i) {
        const float f = s[i];
        const int q = (f >= 0.0f && f <= 254.0f) ? (int)f : -1;
        if (q < 0 || (float)q != f || std::signbit(f)) {
            if (np >= UP_PATCH_CAP) return -1;
            patch[np++] = UpPatch{(uint32_t)i, f};
            d[i] = 255;
        } else {
            d[i] = (uint8_t)q;
        }
    }
    return np;
}
__attribute__((target("avx2"))) int64_t pack_values_avx2(const float* s, uint8_t* d, uint64_t n, UpPatch* patch) {
    const __m256i hi = _mm256_set1_epi32(~255), esc = _mm256_set1_epi32(255);
    const __m256i sign = _mm256_set1_epi32((int)0x80000000u);
    const __m256i fix = _mm256_setr_epi32(0, 4, 1, 5, 2, 6, 3, 7);
    int64_t np = 0;
    uint64_t i = 0;
    for (; i + 32 <= n; i += 32) {
        __m256i q[4];
        __m256i bad = _mm256_setzero_si256();
        for (int k = 0; k < 4; ++k) {
            const __m256 f = _mm256_loadu_ps(s + i + 8 * k);
            q[k] = _mm256_cvttps_epi32(f);
            const __m256 back = _mm256_cvtepi32_ps(q[k]);
            bad = _mm256_or_si256(bad, _mm256_castps_si256(_mm256_cmp_ps(f, back, _CMP_NEQ_UQ)));
            bad = _mm256_or_si256(bad, _mm256_and_si256(q[k], hi));
            bad = _mm256_or_si256(bad, _mm256_cmpeq_epi32(q[k], esc));
            bad = _mm256_or_si256(bad, _mm256_and_si256(_mm256_castps_si256(f), sign));  // -0.0 would lose its sign
        }
        if (!_mm256_testz_si256(bad, bad)) {  // rare: redo this group one value at a time
            np = pack_scalar(s, d, i, i + 32, patch, np);
            if (np < 0) return -1;
            continue;
        }
        const __m256i w0 = _mm256_packus_epi32(q[0], q[1]), w1 = _mm256_packus_epi32(q[2], q[3]);
        const __m256i b = _mm256_permutevar8x32_epi32(_mm256_packus_epi16(w0, w1), fix);
        _mm256_storeu_si256(reinterpret_cast<__m256i*>(d + i), b);
    }
    return pack_scalar(s, d, i, n, patch, np);
}
int64_t pack_values_plain(const float* s, uint8_t* d, uint64_t n, UpPatch* patch) { return pack_scalar(s, d, 0, n, patch, 0); }

// Row indices as one-byte gaps: inside a column the rows ascend, and at the densities of count matrices (a few per cent)
// the gap to the previous entry is below 256 for all but one entry in ~10^5, so an index travels as ONE byte instead of
// four; the others (and every column's first entry, whose gap is negative) go into the chunk's patch list as (position,
// gap modulo 2^32).  anchors[t] = the entry before tile t (LG_UP_TILE entries), `prev0` = the entry before s[0].
// Returns the number of patches, or -1 when the list is full (the chunk then goes up as u32).
struct UpIPatch {
    uint32_t pos;
    uint32_t gap;
};
inline int64_t gaps_scalar(const uint64_t* s, uint8_t* d, uint64_t lo, uint64_t hi, uint32_t prev, UpIPatch* patch, int64_t np,
                           uint64_t* or_all, uint32_t* max_lo) {
    uint64_t o = *or_all;
    uint32_t m = *max_lo;
    for (uint64_t i = lo; i < hi; ++i) {
        o |= s[i];
        const uint32_t v = (uint32_t)s[i];
        m = v > m ? v : m;
        const uint32_t g = v - prev;
        prev = v;
        if (g - 1u < 255u) {
            d[i] = (uint8_t)g;
        } else {
            if (np >= UP_PATCH_CAP) return -1;
            patch[np++] = UpIPatch{(uint32_t)i, g};
            d[i] = 0;
        }
    }
    *or_all = o;
    *max_lo = m;
    return np;
}
__attribute__((target("avx2"))) int64_t gaps_chunk_avx2(const uint64_t* s, uint64_t prev0, uint8_t* d, uint32_t* anchors, uint64_t n,
                                                        UpIPatch* patch, uint64_t* or_all, uint32_t* max_lo) {
    for (uint64_t t = 0; t * LG_UP_TILE < n; ++t) anchors[t] = (uint32_t)(t ? s[t * LG_UP_TILE - 1] : prev0);
    const __m256i pick = _mm256_setr_epi32(0, 2, 4, 6, 0, 2, 4, 6);
    const __m256i fix = _mm256_setr_epi32(0, 4, 1, 5, 2, 6, 3, 7);
    const __m256i one = _mm256_set1_epi32(1), hi = _mm256_set1_epi32(~255);
    __m256i acc = _mm256_setzero_si256(), mx = _mm256_setzero_si256();
    int64_t np = 0;
    // the first group reads s[-1]: done one entry at a time
    const uint64_t head = n < 32 ? n : 32;
    np = gaps_scalar(s, d, 0, head, (uint32_t)prev0, patch, np, or_all, max_lo);
    if (np < 0) return -1;
    uint64_t i = head;
    for (; i + 32 <= n; i += 32) {
        __m256i g[4];
        __m256i bad = _mm256_setzero_si256();
        for (int k = 0; k < 4; ++k) {
            const uint64_t* p = s + i + 8 * k;
            const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(p));
            const __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(p + 4));
            const __m256i a1 = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(p - 1));
            const __m256i b1 = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(p + 3));
            acc = _mm256_or_si256(acc, _mm256_or_si256(a, b));
            const __m256i v = _mm256_blend_epi32(_mm256_permutevar8x32_epi32(a, pick), _mm256_permutevar8x32_epi32(b, pick), 0xF0);
            const __m256i v1 = _mm256_blend_epi32(_mm256_permutevar8x32_epi32(a1, pick), _mm256_permutevar8x32_epi32(b1, pick), 0xF0);
            mx = _mm256_max_epu32(mx, v);
            g[k] = _mm256_sub_epi32(v, v1);
            bad = _mm256_or_si256(bad, _mm256_and_si256(_mm256_or_si256(_mm256_sub_epi32(g[k], one), g[k]), hi));  // gap outside 1..255
        }
        if (!_mm256_testz_si256(bad, bad)) {  // a column start or a wide gap in this group: one entry at a time
            uint64_t o2 = 0;
            uint32_t m2 = 0;
            np = gaps_scalar(s, d, i, i + 32, (uint32_t)s[i - 1], patch, np, &o2, &m2);  // (range checks already taken above)
            if (np < 0) return -1;
            continue;
        }
        const __m256i w0 = _mm256_packus_epi32(g[0], g[1]), w1 = _mm256_packus_epi32(g[2], g[3]);
        const __m256i b = _mm256_permutevar8x32_epi32(_mm256_packus_epi16(w0, w1), fix);
        _mm256_storeu_si256(reinterpret_cast<__m256i*>(d + i), b);
    }
    alignas(32) uint64_t a4[4];
    alignas(32) uint32_t m8[8];
    _mm256_store_si256(reinterpret_cast<__m256i*>(a4), acc);
    _mm256_store_si256(reinterpret_cast<__m256i*>(m8), mx);
    uint64_t o = a4[0] | a4[1] | a4[2] | a4[3];
    uint32_t m = 0;
    for (int k = 0; k < 8; ++k) m = m8[k] > m ? m8[k] : m;
    *or_all |= o;
    *max_lo = m > *max_lo ? m : *max_lo;
    if (i < n) np = gaps_scalar(s, d, i, n, (uint32_t)s[i - 1], patch, np, or_all, max_lo);
    return np;
}
int64_t gaps_chunk_plain(const uint64_t* s, uint64_t prev0, uint8_t* d, uint32_t* anchors, uint64_t n, UpIPatch* patch,
                         uint64_t* or_all, uint32_t* max_lo) {
    for (uint64_t t = 0; t * LG_UP_TILE < n; ++t) anchors[t] = (uint32_t)(t ? s[t * LG_UP_TILE - 1] : prev0);
    return gaps_scalar(s, d, 0, n, (uint32_t)prev0, patch, 0, or_all, max_lo);
}

int upload_threads() {
    if (const char* e = getenv("LG_UPLOAD_THREADS")) return atoi(e);
    unsigned hc = std::thread::hardware_concurrency();
    // one process per GPU shares the host cores with its siblings (torchrun exports LOCAL_WORLD_SIZE)
    if (const char* lw = getenv("LOCAL_WORLD_SIZE")) {
        const int n = atoi(lw);
        if (n > 1) hc /= (unsigned)n;
    }
    if (hc < 4) return 0;  // too few cores to outrun the link: narrow on the device
    return (int)(hc > 16 ? 16 : hc);
}

// returns a CUDA error (cudaSuccess when every chunk was queued); *bad = 1 when an index was out of range.
// Chunks are claimed from both ends: the workers take them from the front and narrow on the host; the calling thread
// takes them from the back and sends them wide (12 bytes per non-zero) through two device staging buffers, but only
// while at most two of its chunks are in flight — the stream is a FIFO, so when the workers keep the link busy the
// wide chunks complete slowly and few are sent, and when the host cores cannot keep up the wide share grows.
cudaError_t upload_narrow_on_host(lg_ctx* ctx, const uint64_t* h_idx, const float* h_val, uint64_t nnz, uint64_t nrows,
                                  uint32_t* d_idx, float* d_val, int nthreads, int* bad) {
    const uint64_t nchunks = (nnz + UP_CHUNK - 1) / UP_CHUNK;
    if ((uint64_t)nthreads > nchunks) nthreads = (int)nchunks;
    const size_t want_slots = (size_t)2 * nthreads;
    // Are the caller's arrays page-locked?  Copies from ordinary pageable memory (a Rust Vec) are staged by the driver
    // and stall the stream, so then every byte goes through the pinned ring (measured on configs[1] from pageable
    // arrays: 1510 ms with wide copies narrowed on the device, 500-790 ms with a wide share, 182 ms through the ring).
    auto is_pinned = [](const void* p) {
        cudaPointerAttributes a;
        if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
            cudaGetLastError();
            return false;
        }
        return a.type == cudaMemoryTypeHost;
    };
    const bool src_pinned = is_pinned(h_idx) && is_pinned(h_val);
    const size_t want_bytes = src_pinned ? UP_SLOT_BYTES : UP_SLOT_BYTES_RAW;
    if (ctx->ring_slots < want_slots || ctx->ring_slot_bytes < want_bytes) {
        if (ctx->ring) cudaFreeHost(ctx->ring);
        ctx->ring = nullptr;
        ctx->ring_slots = 0;
        const size_t sb = want_bytes > ctx->ring_slot_bytes ? want_bytes : ctx->ring_slot_bytes;
        cudaError_t e = cudaHostAlloc(&ctx->ring, want_slots * sb, cudaHostAllocDefault);
        if (e != cudaSuccess) return e;
        ctx->ring_slots = want_slots;
        ctx->ring_slot_bytes = sb;
        while (ctx->ring_ev.size() < want_slots + 8) {
            cudaEvent_t ev;
            e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
            if (e != cudaSuccess) return e;
            ctx->ring_ev.push_back(ev);
        }
    }
    const size_t slot_bytes = ctx->ring_slot_bytes;
    const size_t nslots = ctx->ring_slots;
    uint8_t* ring = static_cast<uint8_t*>(ctx->ring);
    // device twin of the byte-value part of the ring (stream order makes a slot safe to reuse)
    uint8_t* d_bytes = nullptr;
    const bool pack_ok = getenv("LG_UPLOAD_NO_PACK") == nullptr;
    if (pack_ok) {
        cudaError_t e = cudaMallocAsync(&d_bytes, nslots * (UP_CHUNK + UP_PATCH_BYTES), ctx->stream);
        if (e != cudaSuccess) return e;
    }
    // ... and of the one-byte index gaps: [anchors][gap bytes][patches] per slot (LG_UPLOAD_NO_GAPS=1: indices travel as u32)
    uint8_t* d_gaps = nullptr;
    const bool gaps_ok = getenv("LG_UPLOAD_NO_GAPS") == nullptr;
    if (gaps_ok) {
        cudaError_t e = cudaMallocAsync(&d_gaps, nslots * UP_GAP_SLOT, ctx->stream);
        if (e != cudaSuccess) return e;
    }
    std::atomic<uint64_t> gap_chunks{0};
    std::atomic<uint64_t> extra_launches{0}, packed_chunks{0}, wire{0};
    std::vector<std::atomic<int>> queued(nchunks);
    for (auto& q : queued) q.store(0, std::memory_order_relaxed);
    // claim word: low 32 bits = chunks taken from the front, high 32 bits = chunks taken from the back
    std::atomic<uint64_t> claim{0};
    auto take = [&](bool front, uint64_t* idx) {
        uint64_t c = claim.load();
        for (;;) {
            const uint64_t f = c & 0xFFFFFFFFull, b = c >> 32;
            if (f + b >= nchunks) return false;
            if (claim.compare_exchange_weak(c, front ? c + 1 : c + (1ull << 32))) {
                *idx = front ? f : nchunks - 1 - b;
                return true;
            }
        }
    };
    std::atomic<int> first_err{(int)cudaSuccess};
    auto note = [&](cudaError_t e) {
        if (e != cudaSuccess) {
            int exp = (int)cudaSuccess;
            first_err.compare_exchange_strong(exp, (int)e);
        }
    };
    std::atomic<uint64_t> or_all{0};
    std::atomic<uint32_t> max_lo{0};
    const bool avx2 = __builtin_cpu_supports("avx2");
    auto worker = [&]() {
        cudaSetDevice(ctx->device);
        uint64_t my_or = 0, my_launches = 0, my_packed = 0, my_wire = 0, my_gaps = 0;
        uint32_t my_max = 0;
        uint64_t i;
        while (take(true, &i)) {
            const uint64_t off = i * UP_CHUNK;
            const uint64_t len = (nnz - off) < UP_CHUNK ? (nnz - off) : UP_CHUNK;
            const size_t slot = (size_t)(i % nslots);
            cudaError_t e = cudaSuccess;
            if (i >= nslots) {  // the slot's previous copy (chunk i - nslots) must have left the host
                while (!queued[i - nslots].load(std::memory_order_acquire)) std::this_thread::yield();
                e = cudaEventSynchronize(ctx->ring_ev[slot]);
            }
            uint32_t* dst = reinterpret_cast<uint32_t*>(ring + slot * slot_bytes);
            uint8_t* vdst = ring + slot * slot_bytes + UP_CHUNK * sizeof(uint32_t);
            UpPatch* pdst = reinterpret_cast<UpPatch*>(vdst + UP_CHUNK);
            const int64_t npatch = !pack_ok ? -1 : avx2 ? pack_values_avx2(h_val + off, vdst, len, pdst)
                                                        : pack_values_plain(h_val + off, vdst, len, pdst);
            if (e == cudaSuccess) {
                if (npatch >= 0) {
                    uint8_t* dv = d_bytes + slot * (UP_CHUNK + UP_PATCH_BYTES);
                    e = cudaMemcpyAsync(dv, vdst, len, cudaMemcpyHostToDevice, ctx->stream);
                    my_wire += len + (uint64_t)npatch * sizeof(UpPatch);
                    if (e == cudaSuccess && npatch)
                        e = cudaMemcpyAsync(dv + UP_CHUNK, pdst, (size_t)npatch * sizeof(UpPatch), cudaMemcpyHostToDevice, ctx->stream);
                    if (e == cudaSuccess) {
                        k_expand_u8<<<(unsigned)((len / 4 + 1023) / 1024 + 1), 256, 0, ctx->stream>>>(
                            dv, d_val + off, len, reinterpret_cast<const LgUpPatch*>(dv + UP_CHUNK), (uint32_t)npatch);
                        e = cudaGetLastError();
                        ++my_launches;
                        ++my_packed;
                    }
                } else {
                    const float* vsrc = h_val + off;
                    if (!src_pinned) {  // stage the raw values in the slot so that the copy is asynchronous
                        float* rdst = reinterpret_cast<float*>(vdst + UP_CHUNK + UP_PATCH_BYTES);
                        memcpy(rdst, vsrc, len * sizeof(float));
                        vsrc = rdst;
                    }
                    e = cudaMemcpyAsync(d_val + off, vsrc, len * sizeof(float), cudaMemcpyHostToDevice, ctx->stream);
                    my_wire += len * sizeof(float);
                }
            }
            // indices: one-byte gaps when they fit (the u32 region of the slot holds [anchors][gap bytes][patches]: ONE copy)
            int64_t nip = -1;
            if (gaps_ok) {
                uint8_t* gbase = reinterpret_cast<uint8_t*>(dst);
                uint32_t* anchors = reinterpret_cast<uint32_t*>(gbase);
                uint8_t* gbytes = gbase + UP_ANCHOR_BYTES;
                UpIPatch* ipatch = reinterpret_cast<UpIPatch*>(gbase + UP_ANCHOR_BYTES + UP_CHUNK + UP_PATCH_BYTES);  // scratch
                const uint64_t prev0 = off ? h_idx[off - 1] : 0;
                nip = avx2 ? gaps_chunk_avx2(h_idx + off, prev0, gbytes, anchors, len, ipatch, &my_or, &my_max)
                           : gaps_chunk_plain(h_idx + off, prev0, gbytes, anchors, len, ipatch, &my_or, &my_max);
                if (nip >= 0) {
                    const uint64_t pofs = UP_ANCHOR_BYTES + ((len + 7) & ~7ull);  // patches travel right behind the bytes
                    memcpy(gbase + pofs, ipatch, (size_t)nip * sizeof(UpIPatch));
                    const uint64_t total = pofs + (uint64_t)nip * sizeof(UpIPatch);
                    uint8_t* dg = d_gaps + slot * UP_GAP_SLOT;
                    if (e == cudaSuccess) e = cudaMemcpyAsync(dg, gbase, total, cudaMemcpyHostToDevice, ctx->stream);
                    if (e == cudaSuccess) {
                        k_expand_idx_gaps<<<(unsigned)((len + LG_UP_TILE - 1) / LG_UP_TILE), 256, 0, ctx->stream>>>(
                            dg + UP_ANCHOR_BYTES, reinterpret_cast<const uint32_t*>(dg), reinterpret_cast<const LgUpIPatch*>(dg + pofs),
                            (uint32_t)nip, d_idx + off, (uint32_t)len);
                        e = cudaGetLastError();
                        ++my_launches;
                        ++my_gaps;
                    }
                    my_wire += total;
                }
            }
            if (nip < 0) {
                if (avx2) narrow_chunk_avx2(h_idx + off, dst, len, &my_or, &my_max);
                else narrow_chunk_plain(h_idx + off, dst, len, &my_or, &my_max);
                if (e == cudaSuccess)
                    e = cudaMemcpyAsync(d_idx + off, dst, len * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream);
                my_wire += len * sizeof(uint32_t);
            }
            if (e == cudaSuccess) e = cudaEventRecord(ctx->ring_ev[slot], ctx->stream);
            note(e);
            queued[i].store(1, std::memory_order_release);  // set even on error so nobody waits forever
        }
        or_all.fetch_or(my_or);
        extra_launches.fetch_add(my_launches);
        wire.fetch_add(my_wire);
        packed_chunks.fetch_add(my_packed);
        gap_chunks.fetch_add(my_gaps);
        uint32_t cur = max_lo.load();
        while (my_max > cur && !max_lo.compare_exchange_weak(cur, my_max)) {}
    };
    std::vector<std::thread> pool;
    for (int t = 0; t < nthreads; ++t) pool.emplace_back(worker);
    // the calling thread: wide chunks from the back
    int h_flag = 0;
    {
        constexpr int WD_MAX = 8;
        int wdepth = 2;  // wide chunks in flight: sets the wide share the FIFO settles on (measured at 16 threads: depth 2 -> 14 %
                         // of the chunks, 177 ms; 3 -> 178 ms; 4 -> 184 ms; 8 -> 27 %, 193 ms)
        if (const char* wd = getenv("LG_UPLOAD_WIDE_DEPTH")) wdepth = atoi(wd);
        wdepth = wdepth < 1 ? 1 : (wdepth > WD_MAX ? WD_MAX : wdepth);
        uint64_t* stage[WD_MAX] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
        int* d_flag = nullptr;
        cudaError_t e = cudaMallocAsync(&d_flag, sizeof(int), ctx->stream);
        if (e == cudaSuccess) e = cudaMemsetAsync(d_flag, 0, sizeof(int), ctx->stream);
        for (int k = 0; k < wdepth && e == cudaSuccess; ++k) e = cudaMallocAsync(&stage[k], UP_CHUNK * sizeof(uint64_t), ctx->stream);
        note(e);
        uint64_t i, nwide = 0;
        // wide chunks are DMA'd straight from the caller's arrays: only when those are page-locked (see above)
        const bool wide_ok = getenv("LG_UPLOAD_NO_WIDE") == nullptr && src_pinned;
        while (e == cudaSuccess && wide_ok) {
            if (nwide >= (uint64_t)wdepth) e = cudaEventSynchronize(ctx->ring_ev[nslots + (nwide % wdepth)]);
            if (e != cudaSuccess || !take(false, &i)) break;
            const uint64_t off = i * UP_CHUNK;
            const uint64_t len = (nnz - off) < UP_CHUNK ? (nnz - off) : UP_CHUNK;
            uint64_t* sbuf = stage[nwide % wdepth];
            e = cudaMemcpyAsync(d_val + off, h_val + off, len * sizeof(float), cudaMemcpyHostToDevice, ctx->stream);
            if (e == cudaSuccess)
                e = cudaMemcpyAsync(sbuf, h_idx + off, len * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream);
            if (e == cudaSuccess) {
                k_narrow_indices<<<(unsigned)((len + 1023) / 1024), 256, 0, ctx->stream>>>(sbuf, d_idx + off, len, nrows, nrows,
                                                                                            nullptr, d_flag);
                ctx->launches++;
                e = cudaGetLastError();
            }
            if (e == cudaSuccess) e = cudaEventRecord(ctx->ring_ev[nslots + (nwide % wdepth)], ctx->stream);
            ++nwide;
            wire.fetch_add(len * (sizeof(float) + sizeof(uint64_t)));
        }
        note(e);
        for (auto& t : pool) t.join();
        ctx->launches += extra_launches.load();
        ctx->h2d_bytes += wire.load();
        if (d_bytes) cudaFreeAsync(d_bytes, ctx->stream);
        if (d_gaps) cudaFreeAsync(d_gaps, ctx->stream);
        if (d_flag) {
            note(cudaMemcpyAsync(&h_flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
            note(cudaStreamSynchronize(ctx->stream));
            cudaFreeAsync(d_flag, ctx->stream);
        }
        for (int k = 0; k < WD_MAX; ++k)
            if (stage[k]) cudaFreeAsync(stage[k], ctx->stream);
        if (getenv("LG_UPLOAD_TRACE"))
            fprintf(stderr, "[lg_csc_upload] %llu chunks: %llu narrowed on %d host threads (%llu with byte values, %llu with one-byte index gaps), %llu sent wide\n",
                    (unsigned long long)nchunks, (unsigned long long)(nchunks - nwide), nthreads,
                    (unsigned long long)packed_chunks.load(), (unsigned long long)gap_chunks.load(), (unsigned long long)nwide);
    }
    *bad = (h_flag != 0 || (or_all.load() >> 32) != 0 || (uint64_t)max_lo.load() >= nrows) ? 1 : 0;
    return (cudaError_t)first_err.load();
}
}  // namespace


// ---- canonical CSC (rows strictly ascending inside every column) -------------------------------------------------
// The kernels rely on it: K1's pattern bitmap and K10's packed fields need unique rows, K8 binary-searches a column's
// rows.  The reference guarantees it at the same place (read_columns_csc, data-beans/src/sparse_io_vector/read.rs:246-281).
// Entry t+1 may be <= entry t only when t+1 starts a column (empty columns share a start).
__global__ void k_check_canonical(const uint64_t* __restrict__ indptr, const uint32_t* __restrict__ indices, uint64_t ncols,
                                  uint64_t nnz, uint64_t nrows, int* __restrict__ flag) {
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; t < nnz; t += stride) {
        const uint32_t a = indices[t];
        if ((uint64_t)a >= nrows) atomicOr(flag, 2);
        if (t + 1 < nnz && indices[t + 1] <= a) {
            // upper_bound(indptr, t + 1) - 1 = the column holding entry t + 1; canonical iff that column starts there
            uint64_t lo = 0, hi = ncols + 1;
            while (lo < hi) {
                const uint64_t mid = (lo + hi) >> 1;
                if (indptr[mid] <= t + 1) lo = mid + 1;
                else hi = mid;
            }
            if (lo == 0 || indptr[lo - 1] != t + 1) atomicOr(flag, 1);
        }
    }
}

// LG_OK when the block is canonical (cached on the block: blocks are immutable); LG_ERR_INVALID otherwise
int lg_csc_require_canonical(lg_ctx* ctx, const lg_csc* m, const char* who) {
    if (m->canonical < 0) {
        int h = 0;
        if (m->nnz) {
            int* d_flag = nullptr;
            LG_CUDA(ctx, cudaMallocAsync(&d_flag, sizeof(int), ctx->stream));
            LG_CUDA(ctx, cudaMemsetAsync(d_flag, 0, sizeof(int), ctx->stream));
            uint64_t blocks = (m->nnz + 2047) / 2048;
            if (blocks > (uint64_t)ctx->num_sms * 16) blocks = (uint64_t)ctx->num_sms * 16;
            LG_LAUNCH(ctx, k_check_canonical, (unsigned)blocks, 256, 0, m->indptr, m->indices, m->ncols, m->nnz, m->nrows, d_flag);
            int* h_flag = static_cast<int*>(ctx->pinned) + 8;
            LG_CUDA(ctx, cudaMemcpyAsync(h_flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
            LG_CUDA(ctx, cudaFreeAsync(d_flag, ctx->stream));
            LG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            h = *h_flag;
        }
        m->canonical = h ? 0 : 1;
    }
    if (!m->canonical)
        return lg_fail(ctx, LG_ERR_INVALID, std::string(who) + ": the CSC block is not canonical (rows must be strictly ascending and "
                                                               "in range inside every column; lg_csc_upload_remap sorts and merges)");
    return LG_OK;
}

namespace {
// the per-backend row remap of read_columns_csc (read.rs:202-281) on the host: map every row, drop the rows the shared
// axis does not have (UINT32_MAX), and where a column is no longer strictly ascending sort it (stable) and fold equal
// rows by summing in that order.  Column ranges are independent, so they are cut over threads; two passes (sizes, fill).
struct RemapOut {
    std::vector<uint64_t> indptr;
    std::vector<uint32_t> idx;
    std::vector<float> val;
    int bad = 0;  // 1: backend row outside the remap, 2: mapped row outside the shared axis
};
void remap_columns(const uint64_t* indptr, const uint64_t* indices, const float* data, uint64_t col_lo, uint64_t col_hi,
                   const uint32_t* remap, uint64_t nrows_backend, uint64_t nrows_out, RemapOut* out) {
    const uint64_t ncols = col_hi - col_lo;
    out->indptr.assign(ncols + 1, 0);
    unsigned nt = std::thread::hardware_concurrency();
    nt = nt < 1 ? 1 : (nt > 16 ? 16 : nt);
    if (ncols < 4096) nt = 1;
    std::vector<std::vector<std::pair<uint32_t, float>>> cols(ncols);
    std::atomic<int> bad{0};
    auto work = [&](unsigned w) {
        const uint64_t a = ncols * w / nt, b = ncols * (w + 1) / nt;
        for (uint64_t j = a; j < b; ++j) {
            auto& c = cols[j];
            const uint64_t s = indptr[col_lo + j], e = indptr[col_lo + j + 1];
            c.reserve(e - s);
            bool sorted = true;
            for (uint64_t t = s; t < e; ++t) {
                const uint64_t r = indices[t];
                if (r >= nrows_backend) {
                    bad.fetch_or(1);
                    continue;
                }
                const uint32_t g = remap[r];
                if (g == 0xFFFFFFFFu) continue;  // g2c == None: the row is not part of the shared axis (read.rs:217)
                if ((uint64_t)g >= nrows_out) {
                    bad.fetch_or(2);
                    continue;
                }
                if (!c.empty() && c.back().first >= g) sorted = false;
                c.emplace_back(g, data[t]);
            }
            if (!sorted) {  // read.rs:262-277
                std::stable_sort(c.begin(), c.end(), [](const std::pair<uint32_t, float>& x, const std::pair<uint32_t, float>& y) {
                    return x.first < y.first;
                });
                size_t wr = 0, rd = 0;
                while (rd < c.size()) {
                    const uint32_t r = c[rd].first;
                    float v = c[rd].second;
                    ++rd;
                    while (rd < c.size() && c[rd].first == r) {
                        v += c[rd].second;
                        ++rd;
                    }
                    c[wr++] = {r, v};
                }
                c.resize(wr);
            }
            out->indptr[j + 1] = c.size();
        }
    };
    {
        std::vector<std::thread> pool;
        for (unsigned w = 1; w < nt; ++w) pool.emplace_back(work, w);
        work(0);
        for (auto& t : pool) t.join();
    }
    for (uint64_t j = 0; j < ncols; ++j) out->indptr[j + 1] += out->indptr[j];
    out->idx.resize(out->indptr[ncols]);
    out->val.resize(out->indptr[ncols]);
    auto fill = [&](unsigned w) {
        const uint64_t a = ncols * w / nt, b = ncols * (w + 1) / nt;
        for (uint64_t j = a; j < b; ++j) {
            uint64_t o = out->indptr[j];
            for (const auto& rv : cols[j]) {
                out->idx[o] = rv.first;
                out->val[o++] = rv.second;
            }
        }
    };
    {
        std::vector<std::thread> pool;
        for (unsigned w = 1; w < nt; ++w) pool.emplace_back(fill, w);
        fill(0);
        for (auto& t : pool) t.join();
    }
    out->bad = bad.load();
}
}  // namespace

extern "C" int lg_csc_upload(lg_ctx* ctx, const uint64_t* indptr, const uint64_t* indices, const float* data,
                             uint64_t nrows, uint64_t col_lo, uint64_t col_hi, const uint32_t* row_remap,
                             lg_csc** out) {
    // the remap's length is not part of this signature: the caller guarantees it covers every backend row that occurs
    return lg_csc_upload_remap(ctx, indptr, indices, data, nrows, col_lo, col_hi, row_remap, ~0ull, out);
}

extern "C" int lg_csc_upload_remap(lg_ctx* ctx, const uint64_t* indptr, const uint64_t* indices, const float* data,
                                   uint64_t nrows, uint64_t col_lo, uint64_t col_hi, const uint32_t* row_remap,
                                   uint64_t nrows_backend, lg_csc** out) {
    if (!ctx || !out) return LG_ERR_INVALID;
    *out = nullptr;
    LG_REQUIRE(ctx, indptr && col_hi >= col_lo, "lg_csc_upload: null indptr or empty column range");
    LG_REQUIRE(ctx, nrows < 0xFFFFFFFFull, "lg_csc_upload: nrows must fit in u32");
    LG_REQUIRE(ctx, !lg_is_device_ptr(indptr) && !lg_is_device_ptr(indices) && !lg_is_device_ptr(data),
               "lg_csc_upload takes host arrays; use lg_csc_wrap_device for device arrays");
    cudaSetDevice(ctx->device);
    const uint64_t ncols = col_hi - col_lo;
    const uint64_t base = indptr[col_lo], end = indptr[col_hi];
    LG_REQUIRE(ctx, end >= base, "lg_csc_upload: indptr not monotone");
    for (uint64_t j = col_lo; j < col_hi; ++j)  // a column pointer running backwards would send the kernels out of bounds
        if (indptr[j + 1] < indptr[j]) return lg_fail(ctx, LG_ERR_INVALID, "lg_csc_upload: indptr not monotone");
    const uint64_t nnz = end - base;
    LG_REQUIRE(ctx, nnz == 0 || (indices && data), "lg_csc_upload: null indices/data");
    lg_csc* m = new lg_csc();
    m->nrows = nrows;
    m->ncols = ncols;
    m->nnz = nnz;
    m->owned = true;
    auto fail = [&](int rc) {
        lg_csc_free(ctx, m);
        return rc;
    };
    cudaStream_t st = ctx->stream;
#define UP_CUDA(call)                                                                  \
    do {                                                                               \
        cudaError_t e__ = (call);                                                      \
        if (e__ != cudaSuccess) {                                                      \
            ctx->err = std::string("lg_csc_upload: ") + cudaGetErrorString(e__);       \
            return fail(e__ == cudaErrorMemoryAllocation ? LG_ERR_NOMEM : LG_ERR_CUDA); \
        }                                                                              \
    } while (0)
    // stream-ordered pool: a block freed by lg_csc_free is handed to the next upload without a device-wide
    // synchronisation or page (un)mapping (cudaMalloc / cudaFree of ~11 GB cost up to 200 ms per call)
    m->pooled = true;
    UP_CUDA(cudaMallocAsync(&m->indptr, (ncols + 1) * sizeof(uint64_t), st));
    UP_CUDA(cudaMallocAsync(&m->indices, (nnz ? nnz : 1) * sizeof(uint32_t), st));
    UP_CUDA(cudaMallocAsync(&m->values, (nnz ? nnz : 1) * sizeof(float), st));
    // indptr: copy then rebase to 0
    {
        uint64_t* tmp = nullptr;
        UP_CUDA(cudaMallocAsync(&tmp, (ncols + 1) * sizeof(uint64_t), st));
        UP_CUDA(cudaMemcpyAsync(tmp, indptr + col_lo, (ncols + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
        ctx->h2d_bytes += (ncols + 1) * sizeof(uint64_t);
        k_rebase_indptr<<<(unsigned)((ncols + 1 + 255) / 256), 256, 0, st>>>(tmp, m->indptr, ncols + 1, base);
        ctx->launches++;
        UP_CUDA(cudaGetLastError());
        UP_CUDA(cudaFreeAsync(tmp, st));
    }
    if (row_remap && nnz) {
        // remapped rows: canonicalised on the host (sorted, duplicates summed, absent rows dropped), then sent narrow
        RemapOut ro;
        remap_columns(indptr, indices, data, col_lo, col_hi, row_remap, nrows_backend, nrows, &ro);
        if (ro.bad) {
            UP_CUDA(cudaStreamSynchronize(st));
            ctx->err = ro.bad & 1 ? "lg_csc_upload: row index outside the backend's remap" : "lg_csc_upload: remapped row out of range";
            return fail(LG_ERR_INVALID);
        }
        const uint64_t nnz2 = ro.indptr[ncols];
        UP_CUDA(cudaMemcpyAsync(m->indptr, ro.indptr.data(), (ncols + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
        if (nnz2) {
            UP_CUDA(cudaMemcpyAsync(m->indices, ro.idx.data(), nnz2 * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
            UP_CUDA(cudaMemcpyAsync(m->values, ro.val.data(), nnz2 * sizeof(float), cudaMemcpyHostToDevice, st));
        }
        ctx->h2d_bytes += nnz2 * 8 + (ncols + 1) * sizeof(uint64_t);
        UP_CUDA(cudaStreamSynchronize(st));  // the host vectors die with this scope
        m->nnz = nnz2;
        m->canonical = 1;
        *out = m;
        return LG_OK;
    }
    const int up_threads = upload_threads();
    if (nnz >= 4 * UP_CHUNK && up_threads > 0) {  // small blocks: not worth the threads
        int bad = 0;
        UP_CUDA(upload_narrow_on_host(ctx, indices + base, data + base, nnz, nrows, m->indices, m->values, up_threads, &bad));
        UP_CUDA(cudaStreamSynchronize(st));
        if (bad) {
            ctx->err = "lg_csc_upload: row index out of range";
            return fail(LG_ERR_INVALID);
        }
    } else if (nnz) {
        UP_CUDA(cudaMemcpyAsync(m->values, data + base, nnz * sizeof(float), cudaMemcpyHostToDevice, st));
        ctx->h2d_bytes += nnz * (sizeof(float) + sizeof(uint64_t));
        const uint32_t* d_remap = nullptr;
        uint32_t* remap_buf = nullptr;
        const uint64_t nrows_in = nrows;
        int* d_flag = nullptr;
        UP_CUDA(cudaMallocAsync(&d_flag, sizeof(int), st));
        UP_CUDA(cudaMemsetAsync(d_flag, 0, sizeof(int), st));
        // two staging buffers so the copy of chunk i+1 can queue behind the narrowing of chunk i
        const uint64_t chunk = 32ull << 20;  // 32 Mi indices = 256 MiB per staging buffer
        uint64_t* stage[2] = {nullptr, nullptr};
        const uint64_t cap = nnz < chunk ? nnz : chunk;
        UP_CUDA(cudaMallocAsync(&stage[0], cap * sizeof(uint64_t), st));
        if (nnz > chunk) UP_CUDA(cudaMallocAsync(&stage[1], cap * sizeof(uint64_t), st));
        int which = 0;
        for (uint64_t off = 0; off < nnz; off += chunk, which ^= 1) {
            const uint64_t len = (nnz - off) < chunk ? (nnz - off) : chunk;
            uint64_t* sbuf = stage[stage[1] ? which : 0];
            UP_CUDA(cudaMemcpyAsync(sbuf, indices + base + off, len * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
            unsigned grid = (unsigned)((len + 1023) / 1024);
            if (grid > (unsigned)ctx->num_sms * 16) grid = ctx->num_sms * 16;
            k_narrow_indices<<<grid, 256, 0, st>>>(sbuf, m->indices + off, len, nrows_in, nrows, d_remap, d_flag);
            ctx->launches++;
            UP_CUDA(cudaGetLastError());
        }
        int* h_flag = static_cast<int*>(ctx->pinned);
        UP_CUDA(cudaMemcpyAsync(h_flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, st));
        UP_CUDA(cudaFreeAsync(stage[0], st));
        if (stage[1]) UP_CUDA(cudaFreeAsync(stage[1], st));
        UP_CUDA(cudaFreeAsync(d_flag, st));
        if (remap_buf) UP_CUDA(cudaFreeAsync(remap_buf, st));
        UP_CUDA(cudaStreamSynchronize(st));
        if (*h_flag) {
            ctx->err = "lg_csc_upload: row index out of range";
            return fail(LG_ERR_INVALID);
        }
    } else {
        UP_CUDA(cudaStreamSynchronize(st));
    }
#undef UP_CUDA
    // the arrays came from a backend: they must be canonical CSC (the reference's on-disk invariant); checked once here
    {
        const int rc = lg_csc_require_canonical(ctx, m, "lg_csc_upload");
        if (rc != LG_OK) return fail(rc);
    }
    *out = m;
    return LG_OK;
}

// columns of several blocks side by side (SparseIoVec::push of several backends over one feature axis)
__global__ void k_offset_indptr(const uint64_t* __restrict__ src, uint64_t* __restrict__ dst, uint64_t n, uint64_t add) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[i] + add;
}
extern "C" int lg_csc_concat(lg_ctx* ctx, const lg_csc* const* parts, uint32_t nparts, lg_csc** out) {
    if (!ctx || !out) return LG_ERR_INVALID;
    *out = nullptr;
    LG_REQUIRE(ctx, parts && nparts >= 1, "lg_csc_concat: no blocks");
    cudaSetDevice(ctx->device);
    uint64_t ncols = 0, nnz = 0;
    int canonical = 1;
    for (uint32_t p = 0; p < nparts; ++p) {
        LG_REQUIRE(ctx, parts[p] && parts[p]->nrows == parts[0]->nrows, "lg_csc_concat: blocks disagree on the number of rows");
        ncols += parts[p]->ncols;
        nnz += parts[p]->nnz;
        if (parts[p]->canonical != 1) canonical = -1;
    }
    lg_csc* m = new lg_csc();
    m->nrows = parts[0]->nrows;
    m->ncols = ncols;
    m->nnz = nnz;
    m->owned = m->pooled = true;
    m->canonical = canonical;
    cudaStream_t st = ctx->stream;
    auto fail = [&](cudaError_t e) {
        ctx->err = std::string("lg_csc_concat: ") + cudaGetErrorString(e);
        lg_csc_free(ctx, m);
        return e == cudaErrorMemoryAllocation ? LG_ERR_NOMEM : LG_ERR_CUDA;
    };
    cudaError_t e;
    if ((e = cudaMallocAsync(&m->indptr, (ncols + 1) * sizeof(uint64_t), st)) != cudaSuccess) return fail(e);
    if ((e = cudaMallocAsync(&m->indices, (nnz ? nnz : 1) * sizeof(uint32_t), st)) != cudaSuccess) return fail(e);
    if ((e = cudaMallocAsync(&m->values, (nnz ? nnz : 1) * sizeof(float), st)) != cudaSuccess) return fail(e);
    uint64_t c0 = 0, z0 = 0;
    for (uint32_t p = 0; p < nparts; ++p) {
        const lg_csc* q = parts[p];
        // every block's indptr starts at 0; the last entry of the previous block is overwritten by the same value
        k_offset_indptr<<<(unsigned)((q->ncols + 1 + 255) / 256), 256, 0, st>>>(q->indptr, m->indptr + c0, q->ncols + 1, z0);
        ctx->launches++;
        if ((e = cudaGetLastError()) != cudaSuccess) return fail(e);
        if (q->nnz) {
            if ((e = cudaMemcpyAsync(m->indices + z0, q->indices, q->nnz * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st)) != cudaSuccess) return fail(e);
            if ((e = cudaMemcpyAsync(m->values + z0, q->values, q->nnz * sizeof(float), cudaMemcpyDeviceToDevice, st)) != cudaSuccess) return fail(e);
        }
        c0 += q->ncols;
        z0 += q->nnz;
    }
    *out = m;
    return LG_OK;
}

extern "C" int lg_csc_wrap_device(lg_ctx* ctx, const uint64_t* d_indptr, const uint32_t* d_indices,
                                  const float* d_values, uint64_t nrows, uint64_t ncols, uint64_t nnz, lg_csc** out) {
    if (!ctx || !out) return LG_ERR_INVALID;
    *out = nullptr;
    LG_REQUIRE(ctx, d_indptr && (nnz == 0 || (d_indices && d_values)), "lg_csc_wrap_device: null array");
    LG_REQUIRE(ctx, lg_is_device_ptr(d_indptr), "lg_csc_wrap_device: indptr is not a device pointer");
    LG_REQUIRE(ctx, nrows < 0xFFFFFFFFull, "lg_csc_wrap_device: nrows must fit in u32");
    lg_csc* m = new lg_csc();
    m->nrows = nrows;
    m->ncols = ncols;
    m->nnz = nnz;
    m->indptr = const_cast<uint64_t*>(d_indptr);
    m->indices = const_cast<uint32_t*>(d_indices);
    m->values = const_cast<float*>(d_values);
    m->owned = false;
    *out = m;
    return LG_OK;
}

static void free_twin(lg_ctx* ctx, lg_csc* m) {
    if (!m->twin) return;
    if (ctx) {
        cudaSetDevice(ctx->device);
        cudaStreamSynchronize(ctx->stream);
    }
    if (m->twin->bm) cudaFree(m->twin->bm);
    if (m->twin->exc) cudaFree(m->twin->exc);
    if (m->twin->exc_cnt) cudaFree(m->twin->exc_cnt);
    if (m->twin->ovf) cudaFree(m->twin->ovf);
    delete m->twin;
    m->twin = nullptr;
    m->twin_ovf = -1;
}

extern "C" int lg_csc_keep_pattern(lg_ctx* ctx, lg_csc* m, int on) {
    if (!ctx) return LG_ERR_INVALID;
    LG_REQUIRE(ctx, m, "lg_csc_keep_pattern: null block");
    cudaSetDevice(ctx->device);
    if (!on) {
        free_twin(ctx, m);
        return LG_OK;
    }
    if (m->twin || !lg_collapse_pattern_fits(ctx, m->nrows, m->ncols)) return LG_OK;  // already there / outside the kernel's range
    const uint64_t nch = (m->nrows + LG_PAT_GC - 1) / LG_PAT_GC, nsup = (m->ncols + LG_PAT_CELLS - 1) / LG_PAT_CELLS;
    if (on == 2) {  // "if memory is plentiful": the buffers must fit four times into what is free
        const char* kz = getenv("LG_KEEP_PATTERN");
        if (kz && kz[0] == '0') return LG_OK;
        size_t fr = 0, tot = 0;
        const size_t need = ((size_t)(nsup * nch) * (LG_PAT_CELLS * LG_PAT_STRIDE) + (size_t)(m->nnz >> 1) + 2 * (size_t)m->ncols + 2) * 4;
        if (cudaMemGetInfo(&fr, &tot) != cudaSuccess || need * 4 > fr) {
            cudaGetLastError();
            return LG_OK;
        }
    }
    lg_pattern* t = new lg_pattern();
    cudaError_t e = cudaMalloc(&t->bm, (size_t)(nsup * nch) * (LG_PAT_CELLS * LG_PAT_STRIDE) * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMalloc(&t->exc, (size_t)((m->nnz >> 1) + m->ncols + 1) * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMalloc(&t->exc_cnt, (size_t)m->ncols * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMalloc(&t->ovf, sizeof(int));
    m->twin = t;
    if (e != cudaSuccess) {
        cudaGetLastError();
        free_twin(ctx, m);
        return lg_fail(ctx, e == cudaErrorMemoryAllocation ? LG_ERR_NOMEM : LG_ERR_CUDA, std::string("lg_csc_keep_pattern: ") + cudaGetErrorString(e));
    }
    return LG_OK;
}

extern "C" int lg_csc_free(lg_ctx* ctx, lg_csc* m) {
    if (!m) return LG_OK;
    free_twin(ctx, m);
    if (m->owned) {
        if (ctx) {
            cudaSetDevice(ctx->device);
            cudaStreamSynchronize(ctx->stream);
        }
        if (m->pooled && ctx) {
            if (m->indptr) cudaFreeAsync(m->indptr, ctx->stream);
            if (m->indices) cudaFreeAsync(m->indices, ctx->stream);
            if (m->values) cudaFreeAsync(m->values, ctx->stream);
        } else {
            if (m->indptr) cudaFree(m->indptr);
            if (m->indices) cudaFree(m->indices);
            if (m->values) cudaFree(m->values);
        }
    }
    delete m;
    return LG_OK;
}

extern "C" int lg_csc_shape(const lg_csc* m, uint64_t* nrows, uint64_t* ncols, uint64_t* nnz) {
    if (!m) return LG_ERR_INVALID;
    if (nrows) *nrows = m->nrows;
    if (ncols) *ncols = m->ncols;
    if (nnz) *nnz = m->nnz;
    return LG_OK;
}

extern "C" int lg_csc_device_arrays(const lg_csc* m, const uint64_t** ip, const uint32_t** ix, const float** v) {
    if (!m) return LG_ERR_INVALID;
    if (ip) *ip = m->indptr;
    if (ix) *ix = m->indices;
    if (v) *v = m->values;
    return LG_OK;
}

extern "C" int lg_csc_download(lg_ctx* ctx, const lg_csc* m, uint64_t* indptr, uint64_t* indices, float* data) {
    if (!ctx || !m) return LG_ERR_INVALID;
    cudaSetDevice(ctx->device);
    cudaStream_t st = ctx->stream;
    if (indptr)
        LG_CUDA(ctx, cudaMemcpyAsync(indptr, m->indptr, (m->ncols + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    if (data && m->nnz)
        LG_CUDA(ctx, cudaMemcpyAsync(data, m->values, m->nnz * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (indices && m->nnz) {
        const uint64_t chunk = 32ull << 20;
        uint64_t* wide = nullptr;
        const uint64_t cap = m->nnz < chunk ? m->nnz : chunk;
        LG_CUDA(ctx, cudaMallocAsync(&wide, cap * sizeof(uint64_t), st));
        for (uint64_t off = 0; off < m->nnz; off += chunk) {
            const uint64_t len = (m->nnz - off) < chunk ? (m->nnz - off) : chunk;
            unsigned grid = (unsigned)((len + 1023) / 1024);
            if (grid > (unsigned)ctx->num_sms * 16) grid = ctx->num_sms * 16;
            LG_LAUNCH(ctx, k_widen_indices, grid, 256, 0, m->indices + off, wide, len);
            LG_CUDA(ctx, cudaMemcpyAsync(indices + off, wide, len * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
        }
        LG_CUDA(ctx, cudaFreeAsync(wide, st));
    }
    LG_CUDA(ctx, cudaStreamSynchronize(st));
    return LG_OK;
}
