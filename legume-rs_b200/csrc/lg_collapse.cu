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

constexpr int COLLAPSE_THREADS = 512;
constexpr int COLLAPSE_CHUNK = 128;  // sorted cells per work item

__global__ void k_iota_u32(uint32_t* p, uint64_t n) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = (uint32_t)i;
}

// window [g0, g0 + W) of the gene axis lives in shared memory
__global__ void __launch_bounds__(COLLAPSE_THREADS) k_collapse_sorted(
    const uint64_t* __restrict__ indptr, const uint32_t* __restrict__ indices, const float* __restrict__ values,
    const uint32_t* __restrict__ sorted_label, const uint32_t* __restrict__ sorted_cell, uint64_t ncells,
    const float* __restrict__ mult, uint32_t S, uint64_t D, uint32_t g0, uint32_t W, float* __restrict__ sum_ds,
    float* __restrict__ size_s, unsigned long long* __restrict__ next_chunk) {
    extern __shared__ float acc[];  // W floats
    __shared__ unsigned long long s_chunk;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = COLLAPSE_THREADS / 32;
    const uint64_t nchunks = (ncells + COLLAPSE_CHUNK - 1) / COLLAPSE_CHUNK;
    for (uint32_t g = threadIdx.x; g < W; g += COLLAPSE_THREADS) acc[g] = 0.0f;
    while (true) {
        __syncthreads();
        if (threadIdx.x == 0) s_chunk = atomicAdd(next_chunk, 1ull);
        __syncthreads();
        const uint64_t chunk = s_chunk;
        if (chunk >= nchunks) break;
        const uint64_t p0 = chunk * COLLAPSE_CHUNK;
        const uint64_t p1 = (p0 + COLLAPSE_CHUNK) < ncells ? (p0 + COLLAPSE_CHUNK) : ncells;
        uint64_t seg0 = p0;
        while (seg0 < p1) {
            // segment of equal labels inside the chunk (labels are sorted)
            const uint32_t lab = sorted_label[seg0];
            uint64_t seg1 = seg0 + 1;
            while (seg1 < p1 && sorted_label[seg1] == lab) ++seg1;
            if (lab < S) {
                float wsum = 0.0f;
                for (uint64_t p = seg0 + warp; p < seg1; p += nwarp) {
                    const uint32_t cell = sorted_cell[p];
                    const float w = mult ? mult[cell] : 1.0f;
                    const uint64_t lo = indptr[cell], hi = indptr[cell + 1];
                    for (uint64_t t = lo + lane; t < hi; t += 32) {
                        const uint32_t gi = __ldg(indices + t) - g0;
                        if (gi < W) atomicAdd(&acc[gi], __ldg(values + t) * w);
                    }
                    wsum += w;
                }
                if (lane == 0 && g0 == 0 && size_s && wsum != 0.0f) atomicAdd(&size_s[lab], wsum);
                __syncthreads();
                float* col = sum_ds + (size_t)lab * D + g0;
                for (uint32_t g = threadIdx.x; g < W; g += COLLAPSE_THREADS) {
                    const float v = acc[g];
                    if (v != 0.0f) {
                        atomicAdd(col + g, v);
                        acc[g] = 0.0f;
                    }
                }
                __syncthreads();
            }
            seg0 = seg1;
        }
    }
}

__global__ void k_count_bs(const uint32_t* __restrict__ group, const uint32_t* __restrict__ batch,
                           const float* __restrict__ mult, uint64_t ncols, uint32_t S, uint32_t B,
                           float* __restrict__ n_bs) {
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= ncols) return;
    const uint32_t s = group[j], b = batch[j];
    if (s < S && b < B) atomicAdd(&n_bs[(size_t)s * B + b], mult ? mult[j] : 1.0f);
}

// generic "sum columns by label" driver shared by the basic (label = group) and batch (label = batch) stats
static int collapse_by_label(lg_ctx* ctx, LgStage& st, const lg_csc* m, const uint32_t* d_label, const float* d_mult,
                             uint32_t S, float* d_sum, float* d_size) {
    const uint64_t N = m->ncols, D = m->nrows;
    LG_CUDA(ctx, cudaMemsetAsync(d_sum, 0, sizeof(float) * (size_t)D * S, ctx->stream));
    if (d_size) LG_CUDA(ctx, cudaMemsetAsync(d_size, 0, sizeof(float) * S, ctx->stream));
    if (N == 0 || D == 0 || S == 0) return LG_OK;
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
    unsigned long long* d_next;
    LG_TRY(st.scratch(1, &d_next));
    const size_t smem_cap = ctx->smem_optin - 1024;
    const uint32_t Wmax = (uint32_t)(smem_cap / sizeof(float));
    for (uint64_t g0 = 0; g0 < D; g0 += Wmax) {
        const uint32_t W = (uint32_t)((D - g0) < Wmax ? (D - g0) : Wmax);
        const size_t smem = (size_t)W * sizeof(float);
        LG_CUDA(ctx, cudaMemsetAsync(d_next, 0, sizeof(unsigned long long), ctx->stream));
        LG_CUDA(ctx, cudaFuncSetAttribute(k_collapse_sorted, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int per_sm = (smem * 2 + 2048 <= ctx->smem_optin) ? 2 : 1;
        LG_LAUNCH(ctx, k_collapse_sorted, ctx->num_sms * per_sm, COLLAPSE_THREADS, smem, m->indptr, m->indices, m->values,
                  d_lab_out, d_cell_out, N, d_mult, S, D, (uint32_t)g0, W, d_sum, d_size, d_next);
    }
    return LG_OK;
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
    if (m->ncols)
        LG_LAUNCH(ctx, k_count_bs, (unsigned)((m->ncols + 255) / 256), 256, 0, d_group, d_batch, d_mult, m->ncols, S, B, d_nbs);
    return st.finish();
}

__global__ void k_merge_stat(const float* __restrict__ fine, uint64_t D, uint32_t nfine, const uint32_t* __restrict__ f2c,
                             uint32_t ncoarse, float* __restrict__ coarse) {
    // one thread per (gene, coarse) output; fine columns are summed in ascending fine index, as stats.rs:798-812 does
    const uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= D * ncoarse) return;
    const uint64_t g = e % D;
    const uint32_t c = (uint32_t)(e / D);
    float s = 0.0f;
    for (uint32_t f = 0; f < nfine; ++f)
        if (f2c[f] == c) s = __fadd_rn(s, fine[(size_t)f * D + g]);
    coarse[e] = s;
}

extern "C" int lg_merge_stat(lg_ctx* ctx, const float* fine_ds, uint64_t nrows, uint32_t nfine, const uint32_t* fine_to_coarse,
                             uint32_t ncoarse, float* out_coarse_ds) {
    if (!ctx) return LG_ERR_INVALID;
    LG_REQUIRE(ctx, fine_ds && fine_to_coarse && out_coarse_ds, "lg_merge_stat: null argument");
    cudaSetDevice(ctx->device);
    LgStage st(ctx);
    const float* d_fine;
    const uint32_t* d_f2c;
    float* d_coarse;
    LG_TRY(st.in(fine_ds, (size_t)nrows * nfine, &d_fine));
    LG_TRY(st.in(fine_to_coarse, (size_t)nfine, &d_f2c));
    LG_TRY(st.out(out_coarse_ds, (size_t)nrows * ncoarse, &d_coarse));
    const uint64_t total = nrows * ncoarse;
    if (total) LG_LAUNCH(ctx, k_merge_stat, (unsigned)((total + 255) / 256), 256, 0, d_fine, nrows, nfine, d_f2c, ncoarse, d_coarse);
    return st.finish();
}
