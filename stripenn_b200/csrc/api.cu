// C ABI entry points (include/stripenn_b200.h): context life cycle, error reporting, stage
// timers, argument checking.  The numerical work lives in band.cu, quantile.cu,
// background.cu, image.cu, search.cu and score.cu; dedup.cpp is host-only.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <string.h>

#include "score_warp.cuh"

static thread_local std::string g_last_error;

int snn_fail(const snn_ctx* ctx, int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_last_error = buf;
    if (ctx) ctx->err = buf;
    return code;
}

int snn_cuda_fail(const snn_ctx* ctx, cudaError_t e, const char* what) {
    int code = (e == cudaErrorMemoryAllocation) ? SNN_ENOMEM : SNN_ECUDA;
    return snn_fail(ctx, code, "CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
}

SnnChrom* snn_find_chrom(snn_ctx* ctx, int chrom_id) {
    auto it = ctx->chroms.find(chrom_id);
    return it == ctx->chroms.end() ? nullptr : &it->second;
}

int snn_materialize(snn_ctx* ctx, SnnChrom* c) {
    if (!c->pending) return SNN_OK;
    SNN_CUDA(cudaStreamWaitEvent(ctx->stream, c->ready, 0));
    cudaEventDestroy(c->ready);
    c->ready = nullptr;
    c->pending = false;
    SNN_CUDA(cudaMallocAsync(&c->band, (size_t)c->nbins * c->bw * sizeof(double), ctx->stream));
    return snn_band_build(ctx, c);
}

cudaError_t snn_stream_sync(snn_ctx* ctx) {
    if (!ctx->blocking_sync) return cudaStreamSynchronize(ctx->stream);
    if (!ctx->sync_event) {
        cudaError_t e = cudaEventCreateWithFlags(&ctx->sync_event, cudaEventBlockingSync | cudaEventDisableTiming);
        if (e != cudaSuccess) return e;
    }
    cudaError_t e = cudaEventRecord(ctx->sync_event, ctx->stream);
    if (e != cudaSuccess) return e;
    return cudaEventSynchronize(ctx->sync_event);
}

int snn_scratch(snn_ctx* ctx, size_t bytes, void** out) {
    if (bytes > ctx->scratch_bytes) {
        if (ctx->scratch) cudaFree(ctx->scratch);
        ctx->scratch = nullptr;
        ctx->scratch_bytes = 0;
        SNN_CUDA(cudaMalloc(&ctx->scratch, bytes));
        // image buffers are read a little beyond what was written (discarded results): never NaN bits
        SNN_CUDA(cudaMemsetAsync(ctx->scratch, 0, bytes, ctx->stream));
        ctx->scratch_bytes = bytes;
    }
    *out = ctx->scratch;
    return SNN_OK;
}

// ---- timers ------------------------------------------------------------------------------
static cudaEvent_t take_event(snn_ctx* ctx) {
    if (!ctx->event_pool.empty()) {
        cudaEvent_t e = ctx->event_pool.back();
        ctx->event_pool.pop_back();
        return e;
    }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}

SnnScope::SnnScope(snn_ctx* c, int s, int launches) : ctx(c), stage(s), n(launches) {
    if (!ctx->timing) return;
    a = take_event(ctx);
    b = take_event(ctx);
    cudaEventRecord(a, ctx->stream);
}

SnnScope::~SnnScope() {
    if (!a) return;
    cudaEventRecord(b, ctx->stream);
    ctx->timers[stage].pending.push_back({a, b, n});
}

static void drain_timers(snn_ctx* ctx) {
    for (int s = 0; s < ST_COUNT; ++s) {
        SnnTimer& t = ctx->timers[s];
        for (auto& p : t.pending) {
            float ms = 0.f;
            if (cudaEventSynchronize(p.second) == cudaSuccess &&
                cudaEventElapsedTime(&ms, p.first, p.second) == cudaSuccess) {
                t.ms += ms;
                t.launches += p.n;
            }
            ctx->event_pool.push_back(p.first);
            ctx->event_pool.push_back(p.second);
        }
        t.pending.clear();
    }
}

static const char* kStageNames[ST_COUNT] = {"band_build", "quantile_select", "expected",   "row_flags",
                                            "background", "background_sort", "frame_prepare", "gray",
                                            "gaussian",   "sobel_nms",       "search",     "finish",
                                            "score",      "k_gray_gv",       "k_gh_nms"};

extern "C" {

int snn_version(void) { return SNN_VERSION; }

const char* snn_last_error(const snn_ctx* ctx) {
    if (ctx) return ctx->err.c_str();
    return g_last_error.c_str();
}

int snn_create(int device, snn_ctx** out) {
    if (!out) return snn_fail(nullptr, SNN_EINVAL, "snn_create: out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0)
        return snn_fail(nullptr, SNN_ENODEV,
                        "snn_create: no CUDA device (%s); stripenn_b200 has no CPU fallback",
                        e != cudaSuccess ? cudaGetErrorString(e) : "count = 0");
    if (device < 0 || device >= n)
        return snn_fail(nullptr, SNN_EINVAL, "snn_create: device %d out of range [0,%d)", device, n);
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess)
        return snn_cuda_fail(nullptr, e, "cudaGetDeviceProperties");
    if (prop.major != 10)
        return snn_fail(nullptr, SNN_ENODEV,
                        "snn_create: device %d is sm_%d%d; this library is built for sm_100a only",
                        device, prop.major, prop.minor);
    if ((e = cudaSetDevice(device)) != cudaSuccess) return snn_cuda_fail(nullptr, e, "cudaSetDevice");
    snn_ctx* ctx = new snn_ctx();
    ctx->device = device;
    if ((e = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking)) != cudaSuccess) {
        delete ctx;
        return snn_cuda_fail(nullptr, e, "cudaStreamCreate");
    }
    ctx->stream = ctx->own_stream;
    if ((e = cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking)) != cudaSuccess) {
        cudaStreamDestroy(ctx->own_stream);
        delete ctx;
        return snn_cuda_fail(nullptr, e, "cudaStreamCreate");
    }
    {
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
    }
    {
        const char* g = getenv("SNN_FORCE_GENERIC");
        ctx->force_generic = g && g[0] == '1';
        const char* b = getenv("SNN_BLOCKING_SYNC");
        ctx->blocking_sync = b && b[0] == '1';
    }
    *out = ctx;
    return SNN_OK;
}

// Chromosome buffers come from the device's stream-ordered pool (cudaMallocAsync) with the
// release threshold lifted, so re-loading chromosomes re-uses HBM instead of paying a
// synchronising cudaFree/cudaMalloc (~10 ms per 70 MB band) each time.
static void free_chrom(snn_ctx* ctx, SnnChrom& c) {
    if (c.ready) {   // copies of an asynchronous load may still be in flight
        cudaStreamWaitEvent(ctx->stream, c.ready, 0);
        cudaEventDestroy(c.ready);
        c.ready = nullptr;
        c.pending = false;
    }
    if (c.own_coo) {
        if (c.bin1) cudaFreeAsync(c.bin1, ctx->stream);
        if (c.bin2) cudaFreeAsync(c.bin2, ctx->stream);
        if (c.count) cudaFreeAsync(c.count, ctx->stream);
        if (c.weight) cudaFreeAsync(c.weight, ctx->stream);
    }
    if (c.band) cudaFreeAsync(c.band, ctx->stream);
    if (c.fband) cudaFreeAsync(c.fband, ctx->stream);
    if (c.nanbin) cudaFreeAsync(c.nanbin, ctx->stream);
    if (c.nanpre) cudaFreeAsync(c.nanpre, ctx->stream);
    if (c.rowptr) cudaFreeAsync(c.rowptr, ctx->stream);
}

void snn_destroy(snn_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    snn_stream_sync(ctx);
    if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
    drain_timers(ctx);
    for (auto& kv : ctx->chroms) free_chrom(ctx, kv.second);
    snn_stream_sync(ctx);
    for (cudaEvent_t e : ctx->event_pool) cudaEventDestroy(e);
    cudaFree(ctx->bg_sorted);
    cudaFree(ctx->bg_valid);
    cudaFree(ctx->bg_rcp);
    cudaFree(ctx->bg_tree);
    cudaFree(ctx->scratch);
    if (ctx->sync_event) cudaEventDestroy(ctx->sync_event);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

int snn_set_stream(snn_ctx* ctx, void* s) {
    if (!ctx) return snn_fail(nullptr, SNN_EINVAL, "ctx is NULL");
    snn_stream_sync(ctx);
    ctx->stream = s ? (cudaStream_t)s : ctx->own_stream;
    return SNN_OK;
}

int snn_synchronize(snn_ctx* ctx) {
    if (!ctx) return snn_fail(nullptr, SNN_EINVAL, "ctx is NULL");
    SNN_CUDA(cudaSetDevice(ctx->device));
    SNN_CUDA(snn_stream_sync(ctx));
    SNN_CUDA(cudaStreamSynchronize(ctx->copy_stream));
    return SNN_OK;
}

static int load_common(snn_ctx* ctx, int chrom_id, int64_t nbins, int64_t binsize, int64_t chromsize,
                       int64_t nnz, const int32_t* b1, const int32_t* b2, const int32_t* cnt,
                       const double* w, int bw, bool device_ptrs, bool async = false) {
    if (!ctx) return snn_fail(nullptr, SNN_EINVAL, "ctx is NULL");
    if (nbins <= 0 || binsize <= 0 || chromsize <= 0 || nnz < 0 || bw < 1)
        return snn_fail(ctx, SNN_EINVAL, "snn_load_chrom: bad sizes (nbins=%lld binsize=%lld nnz=%lld bw=%d)",
                        (long long)nbins, (long long)binsize, (long long)nnz, bw);
    if ((chromsize + binsize - 1) / binsize != nbins)
        return snn_fail(ctx, SNN_EINVAL, "snn_load_chrom: nbins %lld != ceil(%lld / %lld)", (long long)nbins,
                        (long long)chromsize, (long long)binsize);
    if (nnz > 0 && (!b1 || !b2 || !cnt)) return snn_fail(ctx, SNN_EINVAL, "snn_load_chrom: NULL pixel arrays");
    SNN_CUDA(cudaSetDevice(ctx->device));
    if (snn_find_chrom(ctx, chrom_id)) snn_unload_chrom(ctx, chrom_id);
    SnnChrom c;
    c.id = chrom_id;
    c.nbins = nbins;
    c.binsize = binsize;
    c.chromsize = chromsize;
    c.nnz = nnz;
    c.bw = (int)(bw < nbins ? bw : nbins);
    c.own_coo = !device_ptrs;
    if (device_ptrs) {
        c.bin1 = const_cast<int32_t*>(b1);
        c.bin2 = const_cast<int32_t*>(b2);
        c.count = const_cast<int32_t*>(cnt);
        c.weight = const_cast<double*>(w);
    } else {
        // every failure below releases what was allocated so far (free_chrom takes null members)
#define LOAD_CUDA(call)                                                  \
    do {                                                                 \
        cudaError_t e__ = (call);                                        \
        if (e__ != cudaSuccess) {                                        \
            if (alloc_done) cudaEventDestroy(alloc_done);                \
            free_chrom(ctx, c);                                          \
            return snn_cuda_fail(ctx, e__, #call);                       \
        }                                                                \
    } while (0)
        cudaEvent_t alloc_done = nullptr;
        size_t nb = (size_t)(nnz > 0 ? nnz : 1) * sizeof(int32_t);
        LOAD_CUDA(cudaMallocAsync(&c.bin1, nb, ctx->stream));
        LOAD_CUDA(cudaMallocAsync(&c.bin2, nb, ctx->stream));
        LOAD_CUDA(cudaMallocAsync(&c.count, nb, ctx->stream));
        if (w) LOAD_CUDA(cudaMallocAsync(&c.weight, (size_t)nbins * sizeof(double), ctx->stream));
        cudaStream_t cs = ctx->stream;
        if (async) {
            // the allocations are ordered on the compute stream: the copy stream starts after them
            LOAD_CUDA(cudaEventCreateWithFlags(&alloc_done, cudaEventDisableTiming));
            LOAD_CUDA(cudaEventRecord(alloc_done, ctx->stream));
            LOAD_CUDA(cudaStreamWaitEvent(ctx->copy_stream, alloc_done, 0));
            cudaEventDestroy(alloc_done);
            alloc_done = nullptr;
            cs = ctx->copy_stream;
        }
        if (nnz > 0) {
            LOAD_CUDA(cudaMemcpyAsync(c.bin1, b1, nb, cudaMemcpyHostToDevice, cs));
            LOAD_CUDA(cudaMemcpyAsync(c.bin2, b2, nb, cudaMemcpyHostToDevice, cs));
            LOAD_CUDA(cudaMemcpyAsync(c.count, cnt, nb, cudaMemcpyHostToDevice, cs));
        }
        if (w) LOAD_CUDA(cudaMemcpyAsync(c.weight, w, (size_t)nbins * sizeof(double), cudaMemcpyHostToDevice, cs));
        if (async) {
            LOAD_CUDA(cudaEventCreateWithFlags(&c.ready, cudaEventDisableTiming));
            LOAD_CUDA(cudaEventRecord(c.ready, ctx->copy_stream));
            c.pending = true;
            ctx->chroms[chrom_id] = c;      // the band is built on first use (snn_materialize)
            return SNN_OK;
        }
#undef LOAD_CUDA
    }
    {
        cudaError_t e = cudaMallocAsync(&c.band, (size_t)nbins * c.bw * sizeof(double), ctx->stream);
        if (e != cudaSuccess) {
            free_chrom(ctx, c);
            return snn_cuda_fail(ctx, e, "cudaMallocAsync(band)");
        }
    }
    int rc = snn_band_build(ctx, &c);
    if (rc != SNN_OK) {
        free_chrom(ctx, c);
        return rc;
    }
    if (!device_ptrs) {  // host buffers may go away
        cudaError_t e = snn_stream_sync(ctx);
        if (e != cudaSuccess) {
            free_chrom(ctx, c);
            return snn_cuda_fail(ctx, e, "cudaStreamSynchronize");
        }
    }
    ctx->chroms[chrom_id] = c;
    return SNN_OK;
}

int snn_load_chrom(snn_ctx* ctx, int chrom_id, int64_t nbins, int64_t binsize, int64_t chromsize,
                   int64_t nnz, const int32_t* b1, const int32_t* b2, const int32_t* cnt, const double* w,
                   int bw) {
    return load_common(ctx, chrom_id, nbins, binsize, chromsize, nnz, b1, b2, cnt, w, bw, false);
}

int snn_load_chrom_async(snn_ctx* ctx, int chrom_id, int64_t nbins, int64_t binsize, int64_t chromsize,
                         int64_t nnz, const int32_t* b1, const int32_t* b2, const int32_t* cnt, const double* w,
                         int bw) {
    return load_common(ctx, chrom_id, nbins, binsize, chromsize, nnz, b1, b2, cnt, w, bw, false, true);
}

int snn_load_chrom_device(snn_ctx* ctx, int chrom_id, int64_t nbins, int64_t binsize, int64_t chromsize,
                          int64_t nnz, const int32_t* b1, const int32_t* b2, const int32_t* cnt,
                          const double* w, int bw) {
    return load_common(ctx, chrom_id, nbins, binsize, chromsize, nnz, b1, b2, cnt, w, bw, true);
}

int snn_unload_chrom(snn_ctx* ctx, int chrom_id) {
    if (!ctx) return snn_fail(nullptr, SNN_EINVAL, "ctx is NULL");
    SnnChrom* c = snn_find_chrom(ctx, chrom_id);
    if (!c) return snn_fail(ctx, SNN_ESTATE, "chromosome %d is not loaded", chrom_id);
    cudaSetDevice(ctx->device);
    free_chrom(ctx, *c);   // stream-ordered: later kernels of this context cannot see the memory
    ctx->chroms.erase(chrom_id);
    return SNN_OK;
}

#define GET_CHROM(name)                                                                  \
    if (!ctx) return snn_fail(nullptr, SNN_EINVAL, "ctx is NULL");                       \
    SnnChrom* name = snn_find_chrom(ctx, chrom_id);                                      \
    if (!name) return snn_fail(ctx, SNN_ESTATE, "chromosome %d is not loaded", chrom_id); \
    SNN_CUDA(cudaSetDevice(ctx->device));                                                 \
    if (name->pending) {                                                                  \
        int rc__ = snn_materialize(ctx, name);                                            \
        if (rc__) return rc__;                                                            \
    }

int snn_quantiles(snn_ctx* ctx, int chrom_id, int nq, const double* q, double* out) {
    GET_CHROM(c);
    if (nq < 1 || nq > 64 || !q || !out) return snn_fail(ctx, SNN_EINVAL, "snn_quantiles: bad arguments");
    for (int i = 0; i < nq; ++i)
        if (!(q[i] >= 0.0 && q[i] <= 1.0))
            return snn_fail(ctx, SNN_EINVAL, "Quantiles must be in the range [0, 1]");
    return snn_quantiles_impl(ctx, c, nq, q, out);
}

int snn_quantiles_multi(snn_ctx* ctx, int nchrom, const int32_t* chrom_ids, int nq, const double* q, double* out) {
    if (!ctx) return snn_fail(nullptr, SNN_EINVAL, "ctx is NULL");
    if (nchrom < 0 || (nchrom > 0 && !chrom_ids) || nq < 1 || nq > 64 || !q || !out)
        return snn_fail(ctx, SNN_EINVAL, "snn_quantiles_multi: bad arguments");
    for (int i = 0; i < nq; ++i)
        if (!(q[i] >= 0.0 && q[i] <= 1.0)) return snn_fail(ctx, SNN_EINVAL, "Quantiles must be in the range [0, 1]");
    SNN_CUDA(cudaSetDevice(ctx->device));
    std::vector<SnnChrom*> cs;
    for (int k = 0; k < nchrom; ++k) {
        SnnChrom* c = snn_find_chrom(ctx, chrom_ids[k]);
        if (!c) return snn_fail(ctx, SNN_ESTATE, "chromosome %d is not loaded", chrom_ids[k]);
        if (int rc = snn_materialize(ctx, c)) return rc;
        cs.push_back(c);
    }
    return snn_quantiles_multi_impl(ctx, cs, nq, q, out);
}

int snn_quantiles_select(snn_ctx* ctx, int nsrc, const int32_t* chrom_ids, const int32_t* job, const int64_t* sel,
                         int njob, int nq, const double* q, double* out) {
    if (!ctx) return snn_fail(nullptr, SNN_EINVAL, "ctx is NULL");
    if (nsrc < 0 || njob < 0 || (nsrc > 0 && (!chrom_ids || !job || !sel)) || nq < 1 || nq > 64 || !q || !out)
        return snn_fail(ctx, SNN_EINVAL, "snn_quantiles_select: bad arguments");
    for (int i = 0; i < nq; ++i)
        if (!(q[i] >= 0.0 && q[i] <= 1.0)) return snn_fail(ctx, SNN_EINVAL, "Quantiles must be in the range [0, 1]");
    SNN_CUDA(cudaSetDevice(ctx->device));
    std::vector<QSource> src((size_t)nsrc);
    for (int k = 0; k < nsrc; ++k) {
        SnnChrom* c = snn_find_chrom(ctx, chrom_ids[k]);
        if (!c) return snn_fail(ctx, SNN_ESTATE, "chromosome %d is not loaded", chrom_ids[k]);
        if (int rc = snn_materialize(ctx, c)) return rc;
        if (job[k] < 0 || job[k] >= njob) return snn_fail(ctx, SNN_EINVAL, "snn_quantiles_select: job out of range");
        if (sel[k * 5 + 1] < 1) return snn_fail(ctx, SNN_EINVAL, "snn_quantiles_select: block height must be >= 1");
        src[k].c = c;
        src[k].job = job[k];
        src[k].sel = true;
        src[k].s = sel[k * 5 + 0];
        src[k].wrows = sel[k * 5 + 1];
        src[k].F = sel[k * 5 + 2];
        src[k].c0 = sel[k * 5 + 3];
        src[k].c1 = sel[k * 5 + 4];
    }
    if (nsrc == 0) {   // every job is empty: np.quantile of an empty array
        for (int64_t i = 0; i < (int64_t)njob * nq; ++i) out[i] = NAN;
        return SNN_OK;
    }
    return snn_quantiles_sources_impl(ctx, src, njob, nq, q, out);
}

int snn_expected(snn_ctx* ctx, int chrom_id, double* out) {
    GET_CHROM(c);
    if (!out) return snn_fail(ctx, SNN_EINVAL, "snn_expected: out is NULL");
    return snn_expected_impl(ctx, c, out);
}

int snn_row_flags(snn_ctx* ctx, int chrom_id, int nunits, const int64_t* geom, uint8_t* avail,
                  uint8_t* live) {
    GET_CHROM(c);
    if (nunits < 0 || (nunits > 0 && !geom) || !avail || !live)
        return snn_fail(ctx, SNN_EINVAL, "snn_row_flags: bad arguments");
    return snn_row_flags_impl(ctx, c, nunits, geom, avail, live);
}

int snn_background(snn_ctx* ctx, int chrom_id, int64_t nsamp, const int32_t* spec, double* out) {
    GET_CHROM(c);
    if (nsamp < 0 || (nsamp > 0 && (!spec || !out)))
        return snn_fail(ctx, SNN_EINVAL, "snn_background: bad arguments");
    if (nsamp == 0) return SNN_OK;
    return snn_background_impl(ctx, c, nsamp, spec, out);
}

int snn_set_background(snn_ctx* ctx, int64_t ncol, const double* tables) {
    if (!ctx) return snn_fail(nullptr, SNN_EINVAL, "ctx is NULL");
    if (ncol < 0 || (ncol > 0 && !tables)) return snn_fail(ctx, SNN_EINVAL, "snn_set_background: bad arguments");
    SNN_CUDA(cudaSetDevice(ctx->device));
    return snn_set_background_impl(ctx, ncol, tables);
}

int snn_background_dev(snn_ctx* ctx, int chrom_id, int64_t nsamp, const int32_t* spec, double* d_buf,
                       int64_t slot0) {
    GET_CHROM(c);
    if (nsamp < 0 || slot0 < 0 || (nsamp > 0 && (!spec || !d_buf)))
        return snn_fail(ctx, SNN_EINVAL, "snn_background_dev: bad arguments");
    if (nsamp == 0) return SNN_OK;
    return snn_background_dev_impl(ctx, c, nsamp, spec, d_buf, slot0);
}

int snn_set_background_dev(snn_ctx* ctx, int64_t ncol, const int32_t* slots, const double* d_buf) {
    if (!ctx) return snn_fail(nullptr, SNN_EINVAL, "ctx is NULL");
    if (ncol < 0 || (ncol > 0 && !d_buf))
        return snn_fail(ctx, SNN_EINVAL, "snn_set_background_dev: bad arguments");
    for (int64_t t = 0; slots && t < ncol; ++t)
        if (slots[t] < 0) return snn_fail(ctx, SNN_EINVAL, "snn_set_background_dev: negative slot %lld", (long long)t);
    SNN_CUDA(cudaSetDevice(ctx->device));
    return snn_set_background_dev_impl(ctx, ncol, slots, d_buf);
}

int snn_device_alloc(snn_ctx* ctx, int64_t bytes, void** d_out) {
    if (!ctx) return snn_fail(nullptr, SNN_EINVAL, "ctx is NULL");
    if (!d_out || bytes < 0) return snn_fail(ctx, SNN_EINVAL, "snn_device_alloc: bad arguments");
    *d_out = nullptr;
    SNN_CUDA(cudaSetDevice(ctx->device));
    SNN_CUDA(cudaMalloc(d_out, (size_t)(bytes > 0 ? bytes : 1)));
    return SNN_OK;
}

int snn_device_free(snn_ctx* ctx, void* d_ptr) {
    if (!ctx) return snn_fail(nullptr, SNN_EINVAL, "ctx is NULL");
    SNN_CUDA(cudaSetDevice(ctx->device));
    SNN_CUDA(snn_stream_sync(ctx));
    if (d_ptr) SNN_CUDA(cudaFree(d_ptr));
    return SNN_OK;
}

int snn_device_read(snn_ctx* ctx, void* dst_host, const void* d_src, int64_t bytes) {
    if (!ctx) return snn_fail(nullptr, SNN_EINVAL, "ctx is NULL");
    if (bytes < 0 || (bytes > 0 && (!dst_host || !d_src)))
        return snn_fail(ctx, SNN_EINVAL, "snn_device_read: bad arguments");
    SNN_CUDA(cudaSetDevice(ctx->device));
    SNN_CUDA(snn_stream_sync(ctx));
    if (bytes) SNN_CUDA(cudaMemcpy(dst_host, d_src, (size_t)bytes, cudaMemcpyDeviceToHost));
    return SNN_OK;
}

int snn_extract(snn_ctx* ctx, int chrom_id, int nq, const double* mp, const double* gw, int radius,
                int minH, int maxW, int bfilter, int gray_variant, int frame_lo, int frame_hi,
                snn_candidate* out, int64_t cap, int64_t* n_out) {
    GET_CHROM(c);
    if (nq < 1 || nq > 64 || !mp || !gw || radius < 0 || radius > 32 || bfilter < 1 || bfilter > 15 ||
        gray_variant < 0 || gray_variant > 2 || !n_out || cap < 0 || (cap > 0 && !out))
        return snn_fail(ctx, SNN_EINVAL, "snn_extract: bad arguments");
    int nframes = (int)((c->nbins + 199) / 200);
    if (frame_lo < 0 || frame_hi > nframes || frame_lo > frame_hi)
        return snn_fail(ctx, SNN_EINVAL, "snn_extract: frame range [%d,%d) outside [0,%d)", frame_lo,
                        frame_hi, nframes);
    if (c->bw < (c->nbins < SNN_SMAX ? c->nbins : SNN_SMAX))
        return snn_fail(ctx, SNN_ERANGE, "snn_extract: band width %d < %d", c->bw, SNN_SMAX);
    return snn_extract_impl(ctx, c, nq, mp, gw, radius, minH, maxW, bfilter, gray_variant, frame_lo,
                            frame_hi, out, cap, n_out);
}

int snn_debug_stages(snn_ctx* ctx, int chrom_id, double mp, const double* gw, int radius, int bfilter,
                     int gray_variant, int frame, int bright, int32_t* S, int32_t* nz, float* gray,
                     float* smoothed, float* mag, uint8_t* low, uint8_t* edges, uint8_t* vert) {
    GET_CHROM(c);
    int nframes = (int)((c->nbins + 199) / 200);
    if (!gw || radius < 0 || radius > 32 || frame < 0 || frame >= nframes || bright < 0 ||
        bright >= SNN_NBRIGHT || !S || bfilter < 1 || bfilter > 15)
        return snn_fail(ctx, SNN_EINVAL, "snn_debug_stages: bad arguments");
    return snn_debug_stages_impl(ctx, c, mp, gw, radius, bfilter, gray_variant, frame, bright, S, nz, gray,
                                 smoothed, mag, low, edges, vert);
}

int snn_score(snn_ctx* ctx, int chrom_id, int64_t n, const int64_t* pos, const int32_t* stale,
              const double* expected,
              double* pvalue, double* stripiness, double* oe_mean, double* oe_total, double* o_mean,
              double* o_sum) {
    GET_CHROM(c);
    if (n < 0 || (n > 0 && !pos)) return snn_fail(ctx, SNN_EINVAL, "snn_score: bad arguments");
    if (n == 0) return SNN_OK;
    if (pvalue && !ctx->bg_sorted)
        return snn_fail(ctx, SNN_ESTATE, "snn_score: p-values requested before snn_set_background");
    return snn_score_impl(ctx, c, n, pos, stale, expected, pvalue, stripiness, oe_mean, oe_total, o_mean, o_sum);
}

int snn_score_multi(snn_ctx* ctx, int nchrom, const int32_t* chrom_ids, const double* expected, int64_t n,
                    const int32_t* chrom_index, const int64_t* pos, const int32_t* stale, double* pvalue,
                    double* stripiness, double* oe_mean, double* oe_total, double* o_mean, double* o_sum) {
    if (!ctx) return snn_fail(nullptr, SNN_EINVAL, "ctx is NULL");
    if (nchrom < 1 || !chrom_ids || n < 0 || (n > 0 && !pos))
        return snn_fail(ctx, SNN_EINVAL, "snn_score_multi: bad arguments");
    SNN_CUDA(cudaSetDevice(ctx->device));
    std::vector<ScoreSlot> slots((size_t)nchrom);
    for (int k = 0; k < nchrom; ++k) {
        SnnChrom* c = snn_find_chrom(ctx, chrom_ids[k]);
        if (!c) return snn_fail(ctx, SNN_ESTATE, "chromosome %d is not loaded", chrom_ids[k]);
        if (int rc = snn_materialize(ctx, c)) return rc;
        slots[k].c = c;
        slots[k].expected = expected ? expected + (size_t)k * SNN_NDIST : nullptr;
    }
    if (n == 0) return SNN_OK;
    if (pvalue && !ctx->bg_sorted)
        return snn_fail(ctx, SNN_ESTATE, "snn_score_multi: p-values requested before snn_set_background");
    return snn_score_multi_impl(ctx, slots, n, chrom_index, pos, stale, pvalue, stripiness, oe_mean, oe_total,
                                o_mean, o_sum);
}

int snn_set_mask(snn_ctx* ctx, int chrom_id, int64_t start_bp, int64_t end_bp) {
    if (!ctx) return snn_fail(nullptr, SNN_EINVAL, "ctx is NULL");
    if (chrom_id >= 0 && (start_bp < 0 || end_bp < 0))
        return snn_fail(ctx, SNN_EINVAL, "snn_set_mask: negative mask coordinates");
    ctx->mask_chrom = chrom_id < 0 ? -1 : chrom_id;
    ctx->mask_start = start_bp;
    ctx->mask_end = end_bp;
    return SNN_OK;
}

int snn_stage_times(snn_ctx* ctx, int cap, const char** names, double* ms, int64_t* launches) {
    if (!ctx) return snn_fail(nullptr, SNN_EINVAL, "ctx is NULL");
    cudaSetDevice(ctx->device);
    drain_timers(ctx);
    int n = 0;
    for (int s = 0; s < ST_COUNT && n < cap; ++s, ++n) {
        if (names) names[n] = kStageNames[s];
        if (ms) ms[n] = ctx->timers[s].ms;
        if (launches) launches[n] = ctx->timers[s].launches;
    }
    return n;
}

int snn_reset_stage_times(snn_ctx* ctx) {
    if (!ctx) return snn_fail(nullptr, SNN_EINVAL, "ctx is NULL");
    cudaSetDevice(ctx->device);
    drain_timers(ctx);
    for (int s = 0; s < ST_COUNT; ++s) {
        ctx->timers[s].ms = 0;
        ctx->timers[s].launches = 0;
    }
    return SNN_OK;
}

int snn_enable_stage_timing(snn_ctx* ctx, int on) {
    if (!ctx) return snn_fail(nullptr, SNN_EINVAL, "ctx is NULL");
    ctx->timing = on != 0;
    return SNN_OK;
}

int snn_score_routing(snn_ctx* ctx, int64_t out[3], int reset) {
    if (!ctx || !out) return snn_fail(ctx, SNN_EINVAL, "snn_score_routing: NULL argument");
    for (int k = 0; k < 3; ++k) out[k] = ctx->score_routed[k];
    if (reset)
        for (int k = 0; k < 3; ++k) ctx->score_routed[k] = 0;
    return SNN_OK;
}

}  // extern "C"
