// lg_sim.cu — synthetic count generator for benchmarks and tests, restating data-beans-sim's
// `sample_poisson_triplets` (data-beans-sim/src/core.rs:155-203): y ~ Poisson(rate), keep y > 0.5.
// A counter-based hash of (seed, cell, gene, piece) replaces the per-cell StdRng so any column
// range can be produced independently on any GPU (and identically by the CPU twin in oracle/).
#include <cub/cub.cuh>

#include "lg_common.cuh"

__device__ __forceinline__ uint64_t sim_mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__device__ __forceinline__ float sim_u01(uint64_t key, uint64_t g, uint32_t piece) {
    const uint64_t h = sim_mix64(key ^ (g * 0xD1B54A32D192ED03ull + (uint64_t)piece * 0x8CB92BA72F3D8DD7ull));
    const uint32_t m = (uint32_t)(h >> 40);
    return __fmul_rn(__fadd_rn((float)m, 0.5f), 5.9604644775390625e-08f);
}
__device__ __forceinline__ float sim_poisson_inv(float lam, float p0, float u) {
    float y = 0.0f, p = p0, c = p0;
    while (u > c && y < 1024.0f) {
        y += 1.0f;
        p = __fdiv_rn(__fmul_rn(p, lam), y);
        c = __fadd_rn(c, p);
    }
    return y;
}
__device__ __forceinline__ float sim_draw(uint64_t key, uint64_t g, float lam, float p0, uint32_t np) {
    float y = 0.0f;
    for (uint32_t pc = 0; pc < np; ++pc) y = __fadd_rn(y, sim_poisson_inv(lam, p0, sim_u01(key, g, pc)));
    return y;
}

// one warp per cell; WRITE = false counts nnz, WRITE = true emits rows in ascending order
template <bool WRITE>
__global__ void __launch_bounds__(256) k_sim_cells(uint64_t seed, uint64_t D, uint64_t col_lo, uint64_t ncols,
                                                   const uint8_t* __restrict__ topic, const uint8_t* __restrict__ batch,
                                                   uint32_t nbatch, const float* __restrict__ lam,
                                                   const float* __restrict__ p0, const uint8_t* __restrict__ npiece,
                                                   uint64_t* __restrict__ counts, const uint64_t* __restrict__ indptr,
                                                   uint32_t* __restrict__ indices, float* __restrict__ values) {
    const int lane = threadIdx.x & 31;
    const uint64_t w0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t nw = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t c = w0; c < ncols; c += nw) {
        const uint64_t j = col_lo + c;
        const uint64_t key = sim_mix64(seed + j * 0x9E3779B97F4A7C15ull);
        const size_t base = ((size_t)topic[c] * nbatch + batch[c]) * D;
        uint64_t pos = WRITE ? indptr[c] : 0;
        for (uint64_t g0 = 0; g0 < D; g0 += 32) {
            const uint64_t g = g0 + lane;
            float y = 0.0f;
            if (g < D) y = sim_draw(key, g, lam[base + g], p0[base + g], npiece[base + g]);
            const bool keep = y > 0.5f;
            const unsigned mask = __ballot_sync(0xffffffffu, keep);
            if (WRITE && keep) {
                const uint64_t at = pos + __popc(mask & ((1u << lane) - 1u));
                indices[at] = (uint32_t)g;
                values[at] = y;
            }
            pos += __popc(mask);
        }
        if (!WRITE && lane == 0) counts[c] = pos;
    }
}

__global__ void k_set_last(uint64_t* indptr, const uint64_t* counts, uint64_t ncols) {
    // indptr[0..ncols) holds the exclusive scan; close it
    if (threadIdx.x == 0 && blockIdx.x == 0) indptr[ncols] = ncols ? indptr[ncols - 1] + counts[ncols - 1] : 0;
}

extern "C" int lg_sim_poisson_csc(lg_ctx* ctx, uint64_t seed, uint64_t D, uint64_t col_lo, uint64_t col_hi,
                                  const uint8_t* topic_of_cell, const uint8_t* batch_of_cell, uint32_t ntopic, uint32_t nbatch,
                                  const float* lam, const float* p0, const uint8_t* npiece, lg_csc** out) {
    if (!ctx || !out) return LG_ERR_INVALID;
    *out = nullptr;
    LG_REQUIRE(ctx, topic_of_cell && batch_of_cell && lam && p0 && npiece, "lg_sim_poisson_csc: null argument");
    LG_REQUIRE(ctx, col_hi >= col_lo && D < 0xFFFFFFFFull && ntopic >= 1 && nbatch >= 1, "lg_sim_poisson_csc: bad shape");
    cudaSetDevice(ctx->device);
    const uint64_t ncols = col_hi - col_lo;
    const size_t tab = (size_t)ntopic * nbatch * D;
    LgStage st(ctx);
    const uint8_t *d_topic, *d_batch, *d_np;
    const float *d_lam, *d_p0;
    LG_TRY(st.in(topic_of_cell, (size_t)ncols, &d_topic));
    LG_TRY(st.in(batch_of_cell, (size_t)ncols, &d_batch));
    LG_TRY(st.in(lam, tab, &d_lam));
    LG_TRY(st.in(p0, tab, &d_p0));
    LG_TRY(st.in(npiece, tab, &d_np));
    lg_csc* m = new lg_csc();
    m->nrows = D;
    m->ncols = ncols;
    m->owned = true;
    auto bail = [&](int rc) {
        lg_csc_free(ctx, m);
        return rc;
    };
    if (cudaMalloc(&m->indptr, (ncols + 1) * sizeof(uint64_t)) != cudaSuccess) return bail(lg_fail(ctx, LG_ERR_NOMEM, "lg_sim: indptr alloc"));
    uint64_t* d_counts;
    if (int rc = st.scratch((size_t)ncols + 1, &d_counts)) return bail(rc);
    const unsigned grid = (unsigned)std::min<uint64_t>((ncols + 7) / 8 ? (ncols + 7) / 8 : 1, (uint64_t)ctx->num_sms * 32);
    uint64_t nnz = 0;
    if (ncols) {
        k_sim_cells<false><<<grid, 256, 0, ctx->stream>>>(seed, D, col_lo, ncols, d_topic, d_batch, nbatch, d_lam, d_p0, d_np,
                                                          d_counts, nullptr, nullptr, nullptr);
        ctx->launches++;
        size_t tmp_bytes = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_counts, m->indptr, (int)ncols, ctx->stream);
        char* d_tmp;
        if (int rc = st.scratch(tmp_bytes, &d_tmp)) return bail(rc);
        cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, d_counts, m->indptr, (int)ncols, ctx->stream);
        ctx->launches += 2;
        k_set_last<<<1, 32, 0, ctx->stream>>>(m->indptr, d_counts, ncols);
        ctx->launches++;
        uint64_t* h = static_cast<uint64_t*>(ctx->pinned);
        if (cudaMemcpyAsync(h, m->indptr + ncols, sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
            cudaStreamSynchronize(ctx->stream) != cudaSuccess)
            return bail(lg_fail(ctx, LG_ERR_CUDA, std::string("lg_sim: count pass failed: ") + cudaGetErrorString(cudaGetLastError())));
        nnz = *h;
    } else {
        cudaMemsetAsync(m->indptr, 0, sizeof(uint64_t), ctx->stream);
    }
    m->nnz = nnz;
    if (cudaMalloc(&m->indices, (nnz ? nnz : 1) * sizeof(uint32_t)) != cudaSuccess ||
        cudaMalloc(&m->values, (nnz ? nnz : 1) * sizeof(float)) != cudaSuccess)
        return bail(lg_fail(ctx, LG_ERR_NOMEM, "lg_sim: nnz arrays alloc"));
    if (ncols) {
        k_sim_cells<true><<<grid, 256, 0, ctx->stream>>>(seed, D, col_lo, ncols, d_topic, d_batch, nbatch, d_lam, d_p0, d_np,
                                                         nullptr, m->indptr, m->indices, m->values);
        ctx->launches++;
    }
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) return bail(lg_fail(ctx, LG_ERR_CUDA, std::string("lg_sim: ") + cudaGetErrorString(e)));
    st.mark_host();
    int rc = st.finish();
    if (rc) return bail(rc);
    *out = m;
    return LG_OK;
}
