// lg_ctx.cu — context lifecycle and the device-resident CSC container (the data feed).
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

extern "C" int lg_csc_upload(lg_ctx* ctx, const uint64_t* indptr, const uint64_t* indices, const float* data,
                             uint64_t nrows, uint64_t col_lo, uint64_t col_hi, const uint32_t* row_remap,
                             lg_csc** out) {
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
        k_rebase_indptr<<<(unsigned)((ncols + 1 + 255) / 256), 256, 0, st>>>(tmp, m->indptr, ncols + 1, base);
        ctx->launches++;
        UP_CUDA(cudaGetLastError());
        UP_CUDA(cudaFreeAsync(tmp, st));
    }
    if (nnz) {
        UP_CUDA(cudaMemcpyAsync(m->values, data + base, nnz * sizeof(float), cudaMemcpyHostToDevice, st));
        const uint32_t* d_remap = nullptr;
        uint32_t* remap_buf = nullptr;
        uint64_t nrows_in = nrows;
        if (row_remap) {
            // the remap is indexed by backend row; its length is not known here, so the caller
            // guarantees it covers every index that occurs (read.rs:202-219 builds it that way)
            uint64_t mx = 0;
            for (uint64_t t = base; t < end; ++t) mx = indices[t] > mx ? indices[t] : mx;
            nrows_in = mx + 1;
            UP_CUDA(cudaMallocAsync(&remap_buf, nrows_in * sizeof(uint32_t), st));
            UP_CUDA(cudaMemcpyAsync(remap_buf, row_remap, nrows_in * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
            d_remap = remap_buf;
        }
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

extern "C" int lg_csc_free(lg_ctx* ctx, lg_csc* m) {
    if (!m) return LG_OK;
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
