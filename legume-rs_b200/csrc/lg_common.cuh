// lg_common.cuh — context, error handling and host/device buffer staging shared by all kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <string>
#include <vector>

#include "../../include/legume_b200.h"

// K1's by-product for K5 (lg_project_umma.cu writes it, lg_collapse.cu reads it back inside ONE pass of the hot path): the
// 1-bit sparsity pattern the scan builds for the tensor kernel, tiled [N/256][D/2048][256 cells][66 words] with gene offset o
// of a 32-gene word at bit o, plus the entries whose count is not 1 as packed words `gene | field << 17` (field = count - 1,
// or 0x7fff for a stored zero) in the slots [(lo >> 1) + j, ...) of cell j (lo = indptr[j]; the slot count of a cell is
// (hi >> 1) - (lo >> 1) + 1: half of its entries may differ from one).  `ovf` is raised when a list did not fit its slots or a value is not a whole number in
// [0, 32767]: the collapse then streams the CSC arrays as it always did.
constexpr int LG_PAT_CELLS = 256;                    // cells per supertile
constexpr int LG_PAT_GC = 2048;                      // genes per bitmap chunk
constexpr int LG_PAT_STRIDE = LG_PAT_GC / 32 + 2;    // words per (cell, chunk) row
constexpr uint32_t LG_PAT_ZERO = 0x7fffu;            // field of a stored zero
struct lg_pattern {
    uint32_t* bm = nullptr;       // device, nsuper * nchunks * 256 * 66 words
    uint32_t* exc = nullptr;      // device, (nnz >> 1) + ncols + 1 words
    uint32_t* exc_cnt = nullptr;  // device, ncols: entries with a count != 1 per cell
    int* ovf = nullptr;           // device flag
    uint32_t nchunks = 0;
    bool filled = false;          // set by K1 when the tensor path ran and wrote all of the above
};

struct lg_ctx {
    int device = 0;
    int num_sms = 148;
    size_t smem_optin = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    uint64_t launches = 0;
    uint64_t h2d_bytes = 0;  // bytes lg_csc_upload put on the link
    std::string err;
    // pinned staging for small host<->device scalars
    void* pinned = nullptr;
    size_t pinned_bytes = 0;
    // pinned ring for host-side index narrowing in lg_csc_upload (allocated on first use)
    void* ring = nullptr;
    size_t ring_slots = 0;
    size_t ring_slot_bytes = 0;
    std::vector<cudaEvent_t> ring_ev;
    float* log1p_tab = nullptr;  // libm log1pf of 0 .. 65535 for the exact-order projection (built on first use)
    // NCCL communicator of the cell-sharded path (lg_comm.cu); world 1 = no communicator, every exchange is a no-op
    void* comm = nullptr;
    int comm_rank = 0, comm_world = 1;
    std::vector<uint32_t> lut_keep;  // host copy of the last group table (outlives its asynchronous upload)
    // second stream + two events of the sharded path: the all-reduce of the first half of the group sums runs there while the
    // second half is still being collapsed (lg_comm.cu); created on first use
    cudaStream_t side_stream = nullptr;
    cudaEvent_t side_ev[2] = {nullptr, nullptr};
    // Large scratch (>= LG_CACHE_MIN bytes) is kept by the context and handed out again, smallest fitting block first:
    // every use is ordered on ctx->stream, so a block may be reused by the next call while the previous one's kernels
    // are still queued.  The driver's stream-ordered pool does this too, but with the path's mix of one 4 GB block and
    // many 40 MB ones it fragments (small blocks land inside the big freed one, the next big request maps new memory)
    // until calls stall for hundreds of milliseconds; small scratch still comes from that pool.
    struct CacheBlk {
        void* p;
        size_t bytes;
        bool used;
    };
    std::vector<CacheBlk> cache;
    // tensor-core paths that could not take a call and handed it to the CUDA-core kernel (lg_note_fallback)
    lg_pattern* pat = nullptr;  // set by the hot path around its K1 call: "keep the pattern for the collapse"
    uint64_t pattern_collapses = 0;  // collapses that summed K1's pattern instead of the CSC arrays
    // lg_ctx_time_stages: device time of the six stages of the last lg_hotpath_run_sharded call (events around each stage)
    bool time_stages = false;
    bool stage_ms_valid = false;
    float stage_ms[6] = {0, 0, 0, 0, 0, 0};
    uint64_t fallbacks = 0;
    std::string last_fallback;
    std::vector<std::string> fallback_seen;
};

struct lg_csc {
    uint64_t nrows = 0, ncols = 0, nnz = 0;
    uint64_t* indptr = nullptr;  // device, ncols + 1
    uint32_t* indices = nullptr; // device, nnz
    float* values = nullptr;     // device, nnz
    bool owned = false;
    bool pooled = false;  // owned arrays came from the stream-ordered pool (cudaMallocAsync): freed with cudaFreeAsync
    // cached "all values are small non-negative integers" (-1 unknown); blocks are immutable
    mutable int int_valued = -1;
    // cached "rows strictly ascending and in range inside every column" (-1 unknown): see lg_csc_require_canonical
    mutable int canonical = -1;
    // lg_csc_keep_pattern: buffers owned by the block into which every projection of it leaves its pattern (lg_pattern), so
    // that the collapses that follow (lg_collapse_basic / lg_collapse_batch with unit multiplicities) sum the pattern instead
    // of streaming the arrays again; twin_ovf caches the device flag on the host (-1 = not read since the last projection)
    mutable lg_pattern* twin = nullptr;
    mutable int twin_ovf = -1;
};

constexpr size_t LG_CACHE_MIN = 1u << 20;

inline int lg_fail(lg_ctx* ctx, int code, const std::string& msg) {
    if (ctx) ctx->err = msg;
    return code;
}

// A tensor-core path declined a call for a reason of capability (not of size): counted on the context, remembered, and said
// once per distinct reason on stderr (LG_QUIET=1 silences it) — a 5x slower kernel must not be taken silently.
inline void lg_note_fallback(lg_ctx* ctx, const std::string& what) {
    if (!ctx) return;
    ctx->fallbacks++;
    ctx->last_fallback = what;
    for (const auto& s : ctx->fallback_seen)
        if (s == what) return;
    ctx->fallback_seen.push_back(what);
    const char* q = getenv("LG_QUIET");
    if (!(q && q[0] == '1')) fprintf(stderr, "[legume_b200] %s\n", what.c_str());
}

#define LG_CUDA(ctx, call)                                                                         \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess) {                                                                  \
            char b__[512];                                                                         \
            snprintf(b__, sizeof(b__), "%s:%d: %s: %s", __FILE__, __LINE__, #call,                 \
                     cudaGetErrorString(e__));                                                     \
            return lg_fail(ctx, e__ == cudaErrorMemoryAllocation ? LG_ERR_NOMEM : LG_ERR_CUDA, b__); \
        }                                                                                          \
    } while (0)

#define LG_REQUIRE(ctx, cond, msg)                                             \
    do {                                                                       \
        if (!(cond)) return lg_fail(ctx, LG_ERR_INVALID, std::string(msg));    \
    } while (0)

#define LG_TRY(expr)                    \
    do {                                \
        int rc__ = (expr);              \
        if (rc__ != LG_OK) return rc__; \
    } while (0)

// every kernel launch goes through this so ctx->launches is the library's own count
#define LG_LAUNCH(ctx, kernel, grid, block, smem, ...)                          \
    do {                                                                        \
        kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);        \
        (ctx)->launches++;                                                      \
        LG_CUDA(ctx, cudaGetLastError());                                       \
    } while (0)

inline bool lg_is_device_ptr(const void* p) {
    if (!p) return false;
    cudaPointerAttributes a;
    cudaError_t e = cudaPointerGetAttributes(&a, p);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// Scope object that gives device views of caller buffers.  Host inputs are copied up on
// construction; host outputs are copied back by finish().  Device pointers pass through.
class LgStage {
   public:
    explicit LgStage(lg_ctx* c) : ctx_(c) {}
    ~LgStage() {
        for (auto& t : tmp_) cudaFreeAsync(t, ctx_->stream);
        for (auto i : cached_) ctx_->cache[i].used = false;
    }
    // read-only input; bytes may be 0 / p may be NULL -> nullptr
    template <typename T>
    int in(const T* p, size_t count, const T** dev) {
        *dev = nullptr;
        if (!p || count == 0) return LG_OK;
        if (lg_is_device_ptr(p)) {
            *dev = p;
            return LG_OK;
        }
        any_host_ = true;
        void* d = nullptr;
        LG_CUDA(ctx_, cudaMallocAsync(&d, count * sizeof(T), ctx_->stream));
        tmp_.push_back(d);
        LG_CUDA(ctx_, cudaMemcpyAsync(d, p, count * sizeof(T), cudaMemcpyHostToDevice, ctx_->stream));
        *dev = static_cast<const T*>(d);
        return LG_OK;
    }
    // output; if host, a device scratch is returned and copied back at finish()
    template <typename T>
    int out(T* p, size_t count, T** dev, bool copy_in = false) {
        *dev = nullptr;
        if (!p || count == 0) return LG_OK;
        if (lg_is_device_ptr(p)) {
            *dev = p;
            return LG_OK;
        }
        any_host_ = true;
        void* d = nullptr;
        LG_CUDA(ctx_, cudaMallocAsync(&d, count * sizeof(T), ctx_->stream));
        tmp_.push_back(d);
        if (copy_in) LG_CUDA(ctx_, cudaMemcpyAsync(d, p, count * sizeof(T), cudaMemcpyHostToDevice, ctx_->stream));
        outs_.push_back({p, d, count * sizeof(T)});
        *dev = static_cast<T*>(d);
        return LG_OK;
    }
    // device scratch owned by the stage
    template <typename T>
    int scratch(size_t count, T** dev) {
        const size_t bytes = (count ? count : 1) * sizeof(T);
        if (bytes >= LG_CACHE_MIN) {
            int best = -1;
            for (size_t i = 0; i < ctx_->cache.size(); ++i) {
                const auto& b = ctx_->cache[i];
                if (!b.used && b.bytes >= bytes && (best < 0 || b.bytes < ctx_->cache[best].bytes)) best = (int)i;
            }
            // a much larger block is not worth tying up for a small request
            if (best >= 0 && ctx_->cache[best].bytes > 4 * bytes + (64u << 20)) best = -1;
            if (best < 0) {
                void* d = nullptr;
                LG_CUDA(ctx_, cudaMalloc(&d, bytes));
                ctx_->cache.push_back({d, bytes, false});
                best = (int)ctx_->cache.size() - 1;
            }
            ctx_->cache[best].used = true;
            cached_.push_back((size_t)best);
            *dev = static_cast<T*>(ctx_->cache[best].p);
            return LG_OK;
        }
        void* d = nullptr;
        LG_CUDA(ctx_, cudaMallocAsync(&d, bytes, ctx_->stream));
        tmp_.push_back(d);
        *dev = static_cast<T*>(d);
        return LG_OK;
    }
    int finish() {
        for (auto& o : outs_)
            LG_CUDA(ctx_, cudaMemcpyAsync(o.host, o.dev, o.bytes, cudaMemcpyDeviceToHost, ctx_->stream));
        if (any_host_) LG_CUDA(ctx_, cudaStreamSynchronize(ctx_->stream));
        outs_.clear();
        return LG_OK;
    }
    bool any_host() const { return any_host_; }
    void mark_host() { any_host_ = true; }

   private:
    struct Out {
        void* host;
        void* dev;
        size_t bytes;
    };
    lg_ctx* ctx_;
    std::vector<void*> tmp_;
    std::vector<size_t> cached_;
    std::vector<Out> outs_;
    bool any_host_ = false;
};

// ---- device helpers -------------------------------------------------------------------------
__device__ __forceinline__ double lg_butterfly32(double v) {
    // fixed tree: xor offsets 16, 8, 4, 2, 1 (the oracle mirrors this order)
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) v = v + __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

// Sum one value per thread over a 1024-thread block in the order fixed by LG_BLOCK_CELLS:
// warp butterflies, then a butterfly over the 32 warp sums.  Result valid in warp 0.
__device__ __forceinline__ double lg_block_sum_1024(double v, double* smem32) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = lg_butterfly32(v);
    __syncthreads();
    if (lane == 0) smem32[warp] = v;
    __syncthreads();
    double w = 0.0;
    if (warp == 0) {
        w = smem32[lane];
        w = lg_butterfly32(w);
    }
    return w;
}

// Many sums at once with the SAME tree as lg_block_sum_1024 (so the results are bit-identical to it): stage 1 — every
// warp butterflies value m and lane 0 parks it in stage[m * 32 + warp]; after ONE barrier, stage 2 — warp (m % 32)
// butterflies the 32 warp sums of value m.  lg_block_sum_1024 pays two barriers and an idle second stage per value.
constexpr int LG_SUMS_BATCH = 64;  // values per staging round: 64 * 32 doubles = 16 KB of shared memory
__device__ __forceinline__ void lg_block_sums_stage1(double v, int m, double* stage) {
    v = lg_butterfly32(v);
    if ((threadIdx.x & 31) == 0) stage[m * 32 + (threadIdx.x >> 5)] = v;
}
// call after __syncthreads(); writes out[m] for m in [0, count); ends with a barrier so that `stage` can be reused
__device__ __forceinline__ void lg_block_sums_stage2(const double* stage, int count, double* out) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int m = warp; m < count; m += 32) {
        double w = stage[m * 32 + lane];
        w = lg_butterfly32(w);
        if (lane == 0) out[m] = w;
    }
    __syncthreads();
}

// lg_collapse_basic on device pointers in two launches, groups below `S_half` first; `after_first_half` runs between them (the
// sums of those groups are complete and queued on ctx->stream when it is called).  Internal: the sharded path overlaps the
// all-reduce of the first half with the collapse of the second (lg_comm.cu).
int lg_collapse_basic_split(lg_ctx* ctx, const lg_csc* m, const uint32_t* d_group, uint32_t S, uint32_t S_half, float* d_sum_ds,
                            float* d_size_s, const std::function<int()>& after_first_half);
// K5 from the pattern K1 left behind in the same pass (lg_pattern; lg_collapse.cu)
bool lg_collapse_pattern_fits(const lg_ctx* ctx, uint64_t D, uint64_t N);
int lg_collapse_basic_pattern(lg_ctx* ctx, const lg_csc* m, const uint32_t* group_of_cell, uint32_t S, float* out_sum_ds,
                              float* out_size_s, const lg_pattern* pat);
// LG_OK when rows are strictly ascending and in range inside every column (checked once per block, cached)
int lg_csc_require_canonical(lg_ctx* ctx, const lg_csc* m, const char* who);
// rejects labels outside [0, bound) with LG_ERR_INVALID (one small kernel + a flag read-back)
int lg_check_labels(lg_ctx* ctx, const uint32_t* d_label, uint64_t n, uint32_t bound, const char* what);

// host-side small dense math (lg_hostmath.cpp) — plain C++, no CUDA, no oracle
float lgh_l2_sq(const float* a, const float* b, int d);
