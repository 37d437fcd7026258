// lg_comm.cu — the cell-sharded hot path behind the C ABI: one process per GPU, every
// rank holds a contiguous shard of the cells (its first cell a multiple of LG_BLOCK_CELLS), NCCL over NVLink only where
// the reference reduces over cells.  NCCL is bound at run time (dlopen of libnccl.so.2: the copy a host process already
// carries — torch's, say — is the one that gets used), so the library has no link-time dependency on it and a single-GPU
// host never loads it.
//
//   after K1   all-gather of per-1024-cell-block (batch, dim) partial sums, summed in GLOBAL block order; (min, max)
//   in K3      broadcast of the first r cells' K-vectors; all-gather of Gram / column-sum block partials
//   in K4      all-reduce(max) of the 2^kk code-presence flags
//   after K5   all-reduce(sum) of the gene x group sums and the group sizes (exact for count data in any order)
//
// Order-sensitive reductions never go through an all-reduce, so every result is bit-identical for any GPU count.
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <mutex>

#include <functional>

#include "lg_common.cuh"

namespace {

struct NcclApi {
    void* handle = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclAllReduce) AllReduce = nullptr;
    decltype(&ncclAllGather) AllGather = nullptr;
    decltype(&ncclBroadcast) Broadcast = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    std::string err;
};

NcclApi* nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (api.handle) break;
        }
        if (!api.handle) {
            api.err = std::string("NCCL not found (dlopen libnccl.so.2): ") + (dlerror() ? dlerror() : "");
            return;
        }
#define LG_NCCL_SYM(field, name)                                                      \
    api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.handle, name));        \
    if (!api.field) api.err = std::string("NCCL symbol missing: ") + name;
        LG_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
        LG_NCCL_SYM(CommInitRank, "ncclCommInitRank")
        LG_NCCL_SYM(CommDestroy, "ncclCommDestroy")
        LG_NCCL_SYM(AllReduce, "ncclAllReduce")
        LG_NCCL_SYM(AllGather, "ncclAllGather")
        LG_NCCL_SYM(Broadcast, "ncclBroadcast")
        LG_NCCL_SYM(GroupStart, "ncclGroupStart")
        LG_NCCL_SYM(GroupEnd, "ncclGroupEnd")
        LG_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef LG_NCCL_SYM
    });
    return &api;
}

#define LG_NCCL(ctx, call)                                                                                   \
    do {                                                                                                     \
        ncclResult_t r__ = (call);                                                                           \
        if (r__ != ncclSuccess) {                                                                            \
            char b__[512];                                                                                   \
            snprintf(b__, sizeof(b__), "%s:%d: %s: %s", __FILE__, __LINE__, #call, nccl_api()->GetErrorString(r__)); \
            return lg_fail(ctx, LG_ERR_CUDA, b__);                                                           \
        }                                                                                                    \
    } while (0)

inline ncclComm_t comm_of(lg_ctx* ctx) { return static_cast<ncclComm_t>(ctx->comm); }

// pad this rank's block partials to `mx` rows with +0.0 (exact no-ops in the ordered sum) — the copy is the all-gather's send buffer
__global__ void k_pad_rows(const double* __restrict__ src, uint64_t nsrc, uint64_t ntot, double* __restrict__ dst) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < ntot) dst[i] = i < nsrc ? src[i] : 0.0;
}

// partial sums of this rank's blocks -> sums over ALL ranks' blocks in global block order
int sum_partials(lg_ctx* ctx, LgStage& st, const double* d_part, uint64_t nblk_local, uint64_t nblk_max, uint32_t M, double* d_out) {
    if (ctx->comm_world == 1) return lg_block_partials_finalize(ctx, d_part, nblk_local, M, d_out);
    NcclApi* a = nccl_api();
    double *d_send, *d_all;
    const uint64_t per = nblk_max * M;
    LG_TRY(st.scratch((size_t)per, &d_send));
    LG_TRY(st.scratch((size_t)per * ctx->comm_world, &d_all));
    if (per) LG_LAUNCH(ctx, k_pad_rows, (unsigned)((per + 255) / 256), 256, 0, d_part, nblk_local * M, per, d_send);
    LG_NCCL(ctx, a->AllGather(d_send, d_all, per, ncclFloat64, comm_of(ctx), ctx->stream));
    return lg_block_partials_finalize(ctx, d_all, nblk_max * ctx->comm_world, M, d_out);
}

}  // namespace

extern "C" int lg_comm_unique_id(lg_ctx* ctx, void* out_id) {
    if (!ctx) return LG_ERR_INVALID;
    LG_REQUIRE(ctx, out_id, "lg_comm_unique_id: null argument");
    NcclApi* a = nccl_api();
    if (!a->err.empty()) return lg_fail(ctx, LG_ERR_INTERNAL, a->err);
    static_assert(sizeof(ncclUniqueId) == LG_COMM_ID_BYTES, "the id is 128 bytes");
    ncclUniqueId id;
    LG_NCCL(ctx, a->GetUniqueId(&id));
    memcpy(out_id, &id, sizeof(id));
    return LG_OK;
}

extern "C" int lg_comm_init(lg_ctx* ctx, const void* id, int rank, int world) {
    if (!ctx) return LG_ERR_INVALID;
    LG_REQUIRE(ctx, world >= 1 && rank >= 0 && rank < world, "lg_comm_init: bad rank / world");
    LG_REQUIRE(ctx, !ctx->comm, "lg_comm_init: the context already has a communicator");
    ctx->comm_rank = rank;
    ctx->comm_world = world;
    if (world == 1) return LG_OK;
    LG_REQUIRE(ctx, id, "lg_comm_init: null id");
    NcclApi* a = nccl_api();
    if (!a->err.empty()) return lg_fail(ctx, LG_ERR_INTERNAL, a->err);
    cudaSetDevice(ctx->device);
    ncclUniqueId nid;
    memcpy(&nid, id, sizeof(nid));
    ncclComm_t c = nullptr;
    LG_NCCL(ctx, a->CommInitRank(&c, world, nid, rank));
    ctx->comm = c;
    return LG_OK;
}

extern "C" int lg_comm_destroy(lg_ctx* ctx) {
    if (!ctx) return LG_ERR_INVALID;
    if (ctx->comm) {
        cudaSetDevice(ctx->device);
        cudaStreamSynchronize(ctx->stream);
        nccl_api()->CommDestroy(comm_of(ctx));
        ctx->comm = nullptr;
    }
    ctx->comm_rank = 0;
    ctx->comm_world = 1;
    return LG_OK;
}

extern "C" int lg_comm_info(lg_ctx* ctx, int* rank, int* world) {
    if (!ctx) return LG_ERR_INVALID;
    if (rank) *rank = ctx->comm_rank;
    if (world) *world = ctx->comm_world;
    return LG_OK;
}

// in-place all-reduce(sum) of the collapse statistics (device pointers; any of them may be NULL)
extern "C" int lg_allreduce_stats(lg_ctx* ctx, float* d_sum_ds, float* d_size_s, float* d_sum_db, float* d_n_bs, uint64_t D,
                                  uint32_t S, uint32_t B) {
    if (!ctx) return LG_ERR_INVALID;
    if (ctx->comm_world == 1) return LG_OK;
    cudaSetDevice(ctx->device);
    NcclApi* a = nccl_api();
    LG_NCCL(ctx, a->GroupStart());
    if (d_sum_ds) LG_NCCL(ctx, a->AllReduce(d_sum_ds, d_sum_ds, D * (uint64_t)S, ncclFloat32, ncclSum, comm_of(ctx), ctx->stream));
    if (d_size_s) LG_NCCL(ctx, a->AllReduce(d_size_s, d_size_s, S, ncclFloat32, ncclSum, comm_of(ctx), ctx->stream));
    if (d_sum_db) LG_NCCL(ctx, a->AllReduce(d_sum_db, d_sum_db, D * (uint64_t)B, ncclFloat32, ncclSum, comm_of(ctx), ctx->stream));
    if (d_n_bs) LG_NCCL(ctx, a->AllReduce(d_n_bs, d_n_bs, (uint64_t)B * S, ncclFloat32, ncclSum, comm_of(ctx), ctx->stream));
    LG_NCCL(ctx, a->GroupEnd());
    return LG_OK;
}

// the whole single-batch arm of the path over cell shards: projection -> codes -> groups -> collapse -> posterior
extern "C" int lg_hotpath_run_sharded(lg_ctx* ctx, const lg_csc* m, const float* d_basis_kd, int K, const uint32_t* d_batch,
                                      uint32_t nbatch, int kk, int target, float* d_proj, uint64_t* d_codes, uint32_t* d_group,
                                      uint32_t* out_num_groups, float* d_sum_ds, float* d_size_s, float* d_mean, float* d_sd,
                                      float* d_log_mean, float* d_log_sd) {
    if (!ctx) return LG_ERR_INVALID;
    LG_REQUIRE(ctx, m && d_basis_kd && d_proj && d_codes && d_group && out_num_groups && d_sum_ds && d_size_s,
               "lg_hotpath_run_sharded: null argument");
    LG_REQUIRE(ctx, K >= 1 && K <= 128 && kk >= 1 && kk <= 20 && kk <= K, "lg_hotpath_run_sharded: bad K / kk");
    LG_REQUIRE(ctx, !d_batch || nbatch >= 1, "lg_hotpath_run_sharded: batch labels need nbatch >= 1");
    cudaSetDevice(ctx->device);
    NcclApi* a = ctx->comm_world > 1 ? nccl_api() : nullptr;
    const int W = ctx->comm_world;
    const uint64_t n = m->ncols, D = m->nrows;
    LgStage st(ctx);
    // LG_TRACE=1: device time of every stage of this call on stderr (diagnostic; synchronises at the end)
    const char* tr = getenv("LG_TRACE");
    const bool trace = tr && tr[0] == '1';
    const bool timed = trace || ctx->time_stages;  // lg_ctx_time_stages: the same events, read back through lg_hotpath_last_stage_ms
    std::vector<cudaEvent_t> evs;
    auto mark = [&]() {
        if (!timed) return;
        cudaEvent_t e;
        cudaEventCreate(&e);
        cudaEventRecord(e, ctx->stream);
        evs.push_back(e);
    };
    mark();
    // ---- every rank's cell count (one small exchange; the only host read before the group table) ----
    std::vector<unsigned long long> counts(W, n);
    if (W > 1) {
        unsigned long long *d_mine, *d_allc;
        LG_TRY(st.scratch(1, &d_mine));
        LG_TRY(st.scratch((size_t)W, &d_allc));
        const unsigned long long mine = n;
        LG_CUDA(ctx, cudaMemcpyAsync(d_mine, &mine, sizeof(mine), cudaMemcpyHostToDevice, ctx->stream));
        LG_NCCL(ctx, a->AllGather(d_mine, d_allc, 1, ncclUint64, comm_of(ctx), ctx->stream));
        LG_CUDA(ctx, cudaMemcpyAsync(counts.data(), d_allc, sizeof(unsigned long long) * W, cudaMemcpyDeviceToHost, ctx->stream));
        LG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    uint64_t ntot = 0, nblk_max = 0;
    for (int r = 0; r < W; ++r) {
        ntot += counts[r];
        nblk_max = std::max<uint64_t>(nblk_max, (counts[r] + LG_BLOCK_CELLS - 1) / LG_BLOCK_CELLS);
        LG_REQUIRE(ctx, r + 1 == W || counts[r] % LG_BLOCK_CELLS == 0,
                   "lg_hotpath_run_sharded: every shard but the last must hold a multiple of 1024 cells");
    }
    const uint64_t nblk = (n + LG_BLOCK_CELLS - 1) / LG_BLOCK_CELLS;
    // ---- K1 + K2 ----
    // lg_pattern: the scan of K1 leaves its 1-bit pattern and the counts != 1 in buffers this call owns, and K5 sums the groups
    // from those (3.9 KB + ~0.35 KB per cell at configs[1]) instead of streaming the 8 bytes per non-zero of the CSC arrays a second time
    // (LG_COLLAPSE_PATTERN=0 keeps the CSC kernel)
    lg_pattern pat;
    int* h_ovf = reinterpret_cast<int*>(static_cast<char*>(ctx->pinned) + 256);
    {
        const char* pz = getenv("LG_COLLAPSE_PATTERN");
        if (!(pz && pz[0] == '0') && lg_collapse_pattern_fits(ctx, D, n)) {
            const uint64_t nch = (D + LG_PAT_GC - 1) / LG_PAT_GC, nsup = (n + LG_PAT_CELLS - 1) / LG_PAT_CELLS;
            LG_TRY(st.scratch((size_t)(nsup * nch) * (LG_PAT_CELLS * LG_PAT_STRIDE), &pat.bm));
            LG_TRY(st.scratch((size_t)((m->nnz >> 1) + n + 1), &pat.exc));
            LG_TRY(st.scratch((size_t)n, &pat.exc_cnt));
            LG_TRY(st.scratch(1, &pat.ovf));
            ctx->pat = &pat;
        }
    }
    {
        const int rc = lg_project_raw(ctx, m, d_basis_kd, K, d_proj);
        ctx->pat = nullptr;
        LG_TRY(rc);
    }
    *h_ovf = 1;
    if (pat.filled)  // home long before the group table's read-back below waits on the stream
        LG_CUDA(ctx, cudaMemcpyAsync(h_ovf, pat.ovf, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    double* d_sums = nullptr;
    if (d_batch) {
        const uint32_t M = nbatch * (uint32_t)(K + 1);
        double* d_part;
        LG_TRY(st.scratch((size_t)std::max<uint64_t>(nblk, 1) * M, &d_part));
        LG_TRY(st.scratch((size_t)M, &d_sums));
        LG_TRY(lg_proj_batch_partials(ctx, d_proj, K, n, d_batch, nbatch, d_part));
        LG_TRY(sum_partials(ctx, st, d_part, nblk, nblk_max, M, d_sums));
    }
    float* d_mm;
    LG_TRY(st.scratch(2, &d_mm));
    LG_TRY(lg_proj_centre_scale(ctx, d_proj, K, n, d_batch, d_batch ? nbatch : 0, d_sums, d_mm));
    if (W > 1) {  // the clamp is a global decision (random_projection.rs:401)
        LG_NCCL(ctx, a->GroupStart());
        LG_NCCL(ctx, a->AllReduce(d_mm, d_mm, 1, ncclFloat32, ncclMin, comm_of(ctx), ctx->stream));
        LG_NCCL(ctx, a->AllReduce(d_mm + 1, d_mm + 1, 1, ncclFloat32, ncclMax, comm_of(ctx), ctx->stream));
        LG_NCCL(ctx, a->GroupEnd());
    }
    LG_TRY(lg_proj_clamp_rescale_if(ctx, d_proj, K, n, d_mm));  // decided on the device: no read-back
    mark();
    // ---- K3 ----
    const uint64_t rank_all = std::min<uint64_t>((uint64_t)K, ntot);
    uint64_t r = rank_all > (uint64_t)kk ? (uint64_t)kk + 5 : rank_all;
    r = std::min<uint64_t>(r, ntot);
    LG_REQUIRE(ctx, counts[0] >= r, "lg_hotpath_run_sharded: rank 0 must hold at least kk + 5 cells");
    LG_REQUIRE(ctx, (uint64_t)kk <= r, "lg_hotpath_run_sharded: fewer cells than code bits");
    float *d_first, *d_q, *d_b, *d_u, *d_sig, *d_v, *d_cmean;
    LG_TRY(st.scratch((size_t)r * K, &d_first));
    if (ctx->comm_rank == 0) LG_CUDA(ctx, cudaMemcpyAsync(d_first, d_proj, sizeof(float) * r * K, cudaMemcpyDeviceToDevice, ctx->stream));
    if (W > 1) LG_NCCL(ctx, a->Broadcast(d_first, d_first, r * K, ncclFloat32, 0, comm_of(ctx), ctx->stream));
    LG_TRY(st.scratch((size_t)kk * K, &d_q));
    LG_TRY(lg_codes_basis(ctx, d_first, K, (int)r, kk, d_q));
    const uint32_t MG = (uint32_t)(kk * (kk + 1) / 2);
    double *d_gpart, *d_gram, *d_vpart, *d_vsum;
    LG_TRY(st.scratch((size_t)std::max<uint64_t>(n, 1) * kk, &d_b));
    LG_TRY(st.scratch((size_t)std::max<uint64_t>(nblk, 1) * MG, &d_gpart));
    LG_TRY(st.scratch((size_t)MG, &d_gram));
    LG_TRY(lg_codes_gram(ctx, d_proj, K, n, d_q, kk, d_b, d_gpart));
    LG_TRY(sum_partials(ctx, st, d_gpart, nblk, nblk_max, MG, d_gram));
    LG_TRY(st.scratch((size_t)kk * kk, &d_u));
    LG_TRY(st.scratch((size_t)kk, &d_sig));
    LG_TRY(lg_codes_factor(ctx, d_gram, d_q, K, kk, d_u, d_sig));
    LG_TRY(st.scratch((size_t)std::max<uint64_t>(n, 1) * kk, &d_v));
    LG_TRY(st.scratch((size_t)std::max<uint64_t>(nblk, 1) * kk, &d_vpart));
    LG_TRY(st.scratch((size_t)kk, &d_vsum));
    LG_TRY(lg_codes_vproj(ctx, d_b, kk, n, d_u, d_sig, d_v, d_vpart));
    LG_TRY(sum_partials(ctx, st, d_vpart, nblk, nblk_max, (uint32_t)kk, d_vsum));
    LG_TRY(st.scratch((size_t)kk, &d_cmean));
    LG_TRY(lg_codes_means(ctx, d_vsum, kk, ntot, d_cmean));
    LG_TRY(lg_codes_pack(ctx, d_v, kk, n, d_cmean, d_codes));
    mark();
    // ---- K4: presence flags -> (host) lexicographic group table -> group of every cell ----
    const size_t ncode = (size_t)1 << kk;
    uint32_t *d_present, *d_lut;
    LG_TRY(st.scratch(ncode, &d_present));
    LG_TRY(st.scratch(ncode, &d_lut));
    LG_TRY(lg_code_presence(ctx, d_codes, n, kk, d_present));
    if (W > 1) LG_NCCL(ctx, a->AllReduce(d_present, d_present, ncode, ncclUint32, ncclMax, comm_of(ctx), ctx->stream));
    std::vector<uint32_t> h_present(ncode);
    std::vector<uint32_t>& h_lut = ctx->lut_keep;
    h_lut.assign(ncode, 0u);
    LG_CUDA(ctx, cudaMemcpyAsync(h_present.data(), d_present, sizeof(uint32_t) * ncode, cudaMemcpyDeviceToHost, ctx->stream));
    LG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    uint32_t S = 0;
    LG_TRY(lg_group_lut(ctx, h_present.data(), kk, 0, h_lut.data(), &S));
    LG_CUDA(ctx, cudaMemcpyAsync(d_lut, h_lut.data(), sizeof(uint32_t) * ncode, cudaMemcpyHostToDevice, ctx->stream));
    LG_TRY(lg_codes_to_groups(ctx, d_codes, n, kk, d_lut, d_group));
    *out_num_groups = S;
    mark();
    // ---- K5 + the all-reduce of the sums ----
    // Cells are sorted by group inside the collapse, so the sums of the groups below a split point are final once that share of
    // the sorted cells has been walked.  LG_ALLREDUCE_OVERLAP=1 runs the collapse as two launches and all-reduces the first share
    // (three quarters of the groups) on a second stream while the rest is still being summed.  Measured on 8 x B200 (1.25M cells
    // per GPU): the exposed all-reduce drops 0.39 -> 0.21 ms, but the two launches (two ramps and tails of a persistent grid, the
    // NCCL kernel beside the second one) cost the collapse 2.72 -> 2.95 ms — 14.19 against 14.13 ms per pass, so it is NOT the
    // default; results are bit-identical either way (tools/check_multi_gpu.py on 2 and 8 GPUs).
    const char* ov = getenv("LG_ALLREDUCE_OVERLAP");
    if (W > 1 && S >= 2 && ov && ov[0] == '1' && lg_is_device_ptr(d_sum_ds) && lg_is_device_ptr(d_size_s) && lg_is_device_ptr(d_group)) {
        if (!ctx->side_stream) {
            LG_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->side_stream, cudaStreamNonBlocking));
            for (auto& e : ctx->side_ev) LG_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        }
        const uint32_t S_half = S - S / 4;  // three quarters overlapped: the exposed all-reduce is the last quarter, the hidden one fits the last launch
        const std::function<int()> first_half_done = [&]() -> int {
            LG_CUDA(ctx, cudaEventRecord(ctx->side_ev[0], ctx->stream));
            LG_CUDA(ctx, cudaStreamWaitEvent(ctx->side_stream, ctx->side_ev[0], 0));
            LG_NCCL(ctx, a->AllReduce(d_sum_ds, d_sum_ds, D * (uint64_t)S_half, ncclFloat32, ncclSum, comm_of(ctx), ctx->side_stream));
            LG_CUDA(ctx, cudaEventRecord(ctx->side_ev[1], ctx->side_stream));
            return LG_OK;
        };
        LG_TRY(lg_collapse_basic_split(ctx, m, d_group, S, S_half, d_sum_ds, d_size_s, first_half_done));
        mark();
        LG_NCCL(ctx, a->GroupStart());
        LG_NCCL(ctx, a->AllReduce(d_sum_ds + D * (uint64_t)S_half, d_sum_ds + D * (uint64_t)S_half, D * (uint64_t)(S - S_half), ncclFloat32, ncclSum,
                                  comm_of(ctx), ctx->stream));
        LG_NCCL(ctx, a->AllReduce(d_size_s, d_size_s, S, ncclFloat32, ncclSum, comm_of(ctx), ctx->stream));
        LG_NCCL(ctx, a->GroupEnd());
        LG_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->side_ev[1], 0));
    } else {
        if (pat.filled && *h_ovf == 0) LG_TRY(lg_collapse_basic_pattern(ctx, m, d_group, S, d_sum_ds, d_size_s, &pat));
        else LG_TRY(lg_collapse_basic(ctx, m, d_group, nullptr, S, d_sum_ds, d_size_s));
        mark();
        LG_TRY(lg_allreduce_stats(ctx, d_sum_ds, d_size_s, nullptr, nullptr, D, S, 0));
    }
    mark();
    // ---- K6 (replicated) ----
    if (d_mean || d_sd || d_log_mean || d_log_sd)
        LG_TRY(lg_optimize_single(ctx, d_sum_ds, d_size_s, D, S, 1.0f, 1.0f, target, d_mean, d_sd, d_log_mean, d_log_sd));
    mark();
    if (timed && !trace && evs.size() == 7) {
        cudaEventSynchronize(evs.back());
        for (int i = 0; i < 6; ++i) cudaEventElapsedTime(&ctx->stage_ms[i], evs[i], evs[i + 1]);
        ctx->stage_ms_valid = true;
    }
    if (trace && evs.size() == 7) {
        static cudaEvent_t prev_end = nullptr;  // diagnostic only: the stream time between two consecutive calls
        cudaEventSynchronize(evs.back());
        if (prev_end) {
            float gap = 0.f;
            if (cudaEventElapsedTime(&gap, prev_end, evs[0]) == cudaSuccess) fprintf(stderr, "[lg_hotpath] since the previous call ended: %.3f ms; ", gap);
            cudaEventDestroy(prev_end);
        }
        prev_end = evs.back();
        evs.pop_back();
        evs.push_back(nullptr);
        const char* names[6] = {"project", "codes", "groups", "collapse", "allreduce", "posterior"};
        fprintf(stderr, "[lg_hotpath rank %d]", ctx->comm_rank);
        for (int i = 0; i < 6; ++i) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, evs[i], i + 1 == 6 ? prev_end : evs[i + 1]);
            fprintf(stderr, " %s %.3f ms", names[i], ms);
        }
        fprintf(stderr, "\n");
    }
    for (auto e : evs)
        if (e) cudaEventDestroy(e);
    return LG_OK;
}
