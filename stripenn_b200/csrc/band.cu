// K1 band builder (replaces cooler's sparse->dense + balancing), dense
// window fetch for tests, K-mp expected-by-distance (getStripe.py:178-235) and the row flags
// the background sampling plan needs (getStripe.py:249-273, :328-335).
//
// HBM layout: band[i * bw + d] = value of pixel (i, i + d), d in [0, bw); float64.
//   value = count * (w[i] * w[i+d])   (cooler: counts * outer(w, w), outer product first)
//   bins with NaN weight give NaN for every d, zeros included (0 * NaN = NaN).
// Roofline: HBM.  Algorithmic bytes: 12 B per stored pixel read + 8 B * nbins * bw written.
#include "snn_internal.cuh"

// band[i][d] = 0 * (w_i * w_{i+d}) : 0 for ordinary bins, NaN where a weight is NaN/inf.
__global__ void k_band_init(double* __restrict__ band, const double* __restrict__ w, int64_t nbins, int bw) {
    int64_t total = nbins * bw;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
        double v = 0.0;
        if (w) {
            int64_t i = t / bw;
            int64_t j = i + (t - i * bw);
            if (j < nbins) v = 0.0 * (w[i] * w[j]);
        }
        band[t] = v;
    }
}

__global__ void k_band_scatter(double* __restrict__ band, const int32_t* __restrict__ b1,
                               const int32_t* __restrict__ b2, const int32_t* __restrict__ cnt,
                               const double* __restrict__ w, int64_t nnz, int64_t nbins, int bw) {
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < nnz; t += stride) {
        int64_t i = b1[t], j = b2[t];
        if (i > j) { int64_t s = i; i = j; j = s; }
        int64_t d = j - i;
        if (i < 0 || j >= nbins || d >= bw) continue;
        double v = (double)cnt[t];
        if (w) v = v * (w[i] * w[j]);
        band[i * bw + d] = v;
    }
}

// CSR row pointers of the (bin1, bin2)-sorted pixel table + sortedness check.
__global__ void k_rowptr(const int32_t* __restrict__ b1, const int32_t* __restrict__ b2, int64_t nnz,
                         int64_t nbins, int64_t* __restrict__ rowptr, int* __restrict__ bad) {
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t <= nnz; t += stride) {
        int64_t prev = (t > 0) ? b1[t - 1] : -1;
        int64_t cur = (t < nnz) ? b1[t] : nbins;
        if (t < nnz) {
            int64_t c2 = b2[t];
            if (cur < 0 || cur >= nbins || c2 < cur || c2 >= nbins) *bad = 1;       // upper triangle, in range
            if (t > 0 && (prev > cur || (prev == cur && b2[t - 1] >= c2))) *bad = 2;  // sorted, unique
        }
        if (cur > nbins) cur = nbins;
        if (prev < -1) prev = -1;    // malformed input (negative bin, reported through `bad`): no write below row 0
        for (int64_t r = prev + 1; r <= cur && r <= nbins; ++r) rowptr[r] = t;
    }
}

int snn_band_build(snn_ctx* ctx, SnnChrom* c) {
    {
        void* scratch = nullptr;
        int rc0 = snn_scratch(ctx, 64, &scratch);
        if (rc0) return rc0;
        int* d_bad = (int*)scratch;
        SNN_CUDA(cudaMemsetAsync(d_bad, 0, sizeof(int), ctx->stream));
        SNN_CUDA(cudaMallocAsync(&c->rowptr, (size_t)(c->nbins + 1) * sizeof(int64_t), ctx->stream));
        int blocks0 = (int)((c->nnz + 256) / 256);
        if (blocks0 > 148 * 16) blocks0 = 148 * 16;
        {
            SnnScope scope(ctx, ST_BAND, 1);
            k_rowptr<<<blocks0, 256, 0, ctx->stream>>>(c->bin1, c->bin2, c->nnz, c->nbins, c->rowptr, d_bad);
            SNN_LAUNCH_CHECK("k_rowptr");
        }
        int bad = 0;
        SNN_CUDA(cudaMemcpyAsync(&bad, d_bad, sizeof bad, cudaMemcpyDeviceToHost, ctx->stream));
        SNN_CUDA(snn_stream_sync(ctx));
        if (bad == 1)
            return snn_fail(ctx, SNN_EINVAL, "snn_load_chrom: pixels must satisfy 0 <= bin1 <= bin2 < nbins");
        if (bad == 2)
            return snn_fail(ctx, SNN_EINVAL,
                            "snn_load_chrom: pixels must be sorted by (bin1, bin2) without duplicates, like a "
                            "cooler pixel table");
    }
    SnnScope scope(ctx, ST_BAND, 2);
    int threads = 256;
    int64_t total = c->nbins * c->bw;
    int blocks = (int)((total + threads - 1) / threads);
    if (blocks > 148 * 16) blocks = 148 * 16;
    k_band_init<<<blocks, threads, 0, ctx->stream>>>(c->band, c->weight, c->nbins, c->bw);
    SNN_LAUNCH_CHECK("k_band_init");
    if (c->nnz > 0) {
        blocks = (int)((c->nnz + threads - 1) / threads);
        if (blocks > 148 * 16) blocks = 148 * 16;
        k_band_scatter<<<blocks, threads, 0, ctx->stream>>>(c->band, c->bin1, c->bin2, c->count, c->weight,
                                                            c->nnz, c->nbins, c->bw);
        SNN_LAUNCH_CHECK("k_band_scatter");
    }
    return SNN_OK;
}

// ---- column-major full band for the fast scoring kernel (score_fast.cu) -------------------------
// fband[col * S + (row - col) + bw - 1] = nan0(pixel (row, col)), S = 2*bw - 1, |row - col| < bw: the
// rows of one matrix column are contiguous on BOTH sides of the diagonal, so a warp whose lanes are
// the rows of a stripe reads each window column as one coalesced run (the row-major band gives that
// only below the diagonal).  NaN pixels are stored as 0 — every consumer sums nan0(pixel) — and the
// NaN bookkeeping of scoringstripes (dead columns, finite counts) comes from the weight vector:
// the chromosome is eligible only if every weight is NaN or |w| < 1e140, so that a pixel is NaN
// exactly when one of its two bins has a NaN weight (products of moderate weights and int32 counts
// stay finite).  nanbin[i] = weight i is NaN, nanpre[i] = #(NaN bins < i).
// Algorithmic bytes: 12 B per stored pixel read + 8 B * nbins * (2*bw - 1) written.
__global__ void k_fband_scatter(double* __restrict__ fb, const int32_t* __restrict__ b1,
                                const int32_t* __restrict__ b2, const int32_t* __restrict__ cnt,
                                const double* __restrict__ w, int64_t nnz, int64_t nbins, int bw) {
    const int S = 2 * bw - 1;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < nnz; t += stride) {
        int64_t i = b1[t], j = b2[t];
        if (i > j) { int64_t s = i; i = j; j = s; }
        const int64_t d = j - i;
        if (i < 0 || j >= nbins || d >= bw) continue;
        double v = (double)cnt[t];
        if (w) v = v * (w[i] * w[j]);
        v = snn_nan0(v);
        fb[j * S + (bw - 1) - d] = v;      // (row i, col j): above the diagonal
        fb[i * S + (bw - 1) + d] = v;      // (row j, col i): below
    }
}

__global__ void k_nan_bins(const double* __restrict__ w, int64_t nbins, uint8_t* __restrict__ flag,
                           int* __restrict__ ineligible) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < nbins; t += stride) {
        const double x = w[t];
        flag[t] = (x != x) ? 1 : 0;
        if (x == x && !(fabs(x) < 1e140)) *ineligible = 1;
    }
}

// exclusive prefix counts of the NaN flags, one CTA: per-thread segment sums, a block scan, a second walk
__global__ void __launch_bounds__(1024) k_nan_prefix(const uint8_t* __restrict__ flag, int64_t nbins,
                                                     int32_t* __restrict__ pre) {
    __shared__ int part[1024];
    const int64_t seg = (nbins + 1023) / 1024;
    const int64_t a = (int64_t)threadIdx.x * seg;
    const int64_t b = a + seg < nbins ? a + seg : nbins;
    int s = 0;
    for (int64_t i = a; i < b; ++i) s += flag[i];
    part[threadIdx.x] = s;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        const int v = threadIdx.x >= o ? part[threadIdx.x - o] : 0;
        __syncthreads();
        part[threadIdx.x] += v;
        __syncthreads();
    }
    int run = part[threadIdx.x] - s;
    for (int64_t i = a; i < b; ++i) {
        pre[i] = run;
        run += flag[i];
    }
    if (threadIdx.x == 1023) pre[nbins] = part[1023];
}

// Builds the full band on first use.  c->fband_state: 0 not built, 1 ready, 2 not eligible (weights
// outside the moderate range: the fast kernel is not used for this chromosome).
int snn_fband_build(snn_ctx* ctx, SnnChrom* c) {
    if (c->fband_state) return SNN_OK;
    if (c->weight) {
        void* scratch = nullptr;
        if (int rc0 = snn_scratch(ctx, 64, &scratch)) return rc0;
        int* d_flag = (int*)scratch;
        SNN_CUDA(cudaMemsetAsync(d_flag, 0, sizeof(int), ctx->stream));
        SNN_CUDA(cudaMallocAsync(&c->nanbin, (size_t)c->nbins, ctx->stream));
        SNN_CUDA(cudaMallocAsync(&c->nanpre, (size_t)(c->nbins + 1) * sizeof(int32_t), ctx->stream));
        int blocks = (int)((c->nbins + 255) / 256);
        if (blocks > 148 * 8) blocks = 148 * 8;
        {
            SnnScope scope(ctx, ST_BAND, 2);
            k_nan_bins<<<blocks, 256, 0, ctx->stream>>>(c->weight, c->nbins, c->nanbin, d_flag);
            SNN_LAUNCH_CHECK("k_nan_bins");
            k_nan_prefix<<<1, 1024, 0, ctx->stream>>>(c->nanbin, c->nbins, c->nanpre);
            SNN_LAUNCH_CHECK("k_nan_prefix");
        }
        int inel = 0;
        SNN_CUDA(cudaMemcpyAsync(&inel, d_flag, sizeof inel, cudaMemcpyDeviceToHost, ctx->stream));
        SNN_CUDA(snn_stream_sync(ctx));
        if (inel) {
            c->fband_state = 2;
            return SNN_OK;
        }
    }
    const int64_t total = c->nbins * (2 * (int64_t)c->bw - 1);
    {
        // an optional layout: when HBM cannot hold it the chromosome simply stays with the warp kernel
        const cudaError_t e = cudaMallocAsync(&c->fband, (size_t)total * sizeof(double), ctx->stream);
        if (e == cudaErrorMemoryAllocation) {
            cudaGetLastError();
            c->fband = nullptr;
            c->fband_state = 2;
            return SNN_OK;
        }
        if (e != cudaSuccess) return snn_cuda_fail(ctx, e, "cudaMallocAsync(fband)");
    }
    SnnScope scope(ctx, ST_BAND, 1);
    const int threads = 256;
    SNN_CUDA(cudaMemsetAsync(c->fband, 0, (size_t)total * sizeof(double), ctx->stream));
    if (c->nnz > 0) {
        int blocks = (int)((c->nnz + threads - 1) / threads);
        if (blocks > 148 * 16) blocks = 148 * 16;
        k_fband_scatter<<<blocks, threads, 0, ctx->stream>>>(c->fband, c->bin1, c->bin2, c->count, c->weight, c->nnz,
                                                             c->nbins, c->bw);
        SNN_LAUNCH_CHECK("k_fband_scatter");
    }
    c->fband_state = 1;
    return SNN_OK;
}

// ---- dense window (tests / debugging) -------------------------------------------------------
__global__ void k_fetch(const double* __restrict__ band, int64_t nbins, int bw, int64_t i0, int64_t i1,
                        int64_t j0, int64_t j1, double* __restrict__ out) {
    int64_t nc = j1 - j0;
    int64_t total = (i1 - i0) * nc;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (int64_t)gridDim.x * blockDim.x) {
        int64_t i = i0 + t / nc, j = j0 + t % nc;
        out[t] = snn_band_at(band, nbins, bw, i, j);
    }
}

extern "C" int snn_fetch(snn_ctx* ctx, int chrom_id, int64_t i0, int64_t i1, int64_t j0, int64_t j1,
                         double* out) {
    if (!ctx) return snn_fail(nullptr, SNN_EINVAL, "ctx is NULL");
    SnnChrom* c = snn_find_chrom(ctx, chrom_id);
    if (!c) return snn_fail(ctx, SNN_ESTATE, "chromosome %d is not loaded", chrom_id);
    SNN_CUDA(cudaSetDevice(ctx->device));
    if (int rc = snn_materialize(ctx, c)) return rc;
    if (i0 < 0 || j0 < 0 || i1 > c->nbins || j1 > c->nbins || i1 < i0 || j1 < j0 || !out)
        return snn_fail(ctx, SNN_EINVAL, "snn_fetch: bad window");
    int64_t far = (j1 - 1 - i0) > (i1 - 1 - j0) ? (j1 - 1 - i0) : (i1 - 1 - j0);
    if (far >= c->bw) return snn_fail(ctx, SNN_ERANGE, "snn_fetch: window reaches offset %lld >= band width %d",
                                      (long long)far, c->bw);
    int64_t total = (i1 - i0) * (j1 - j0);
    if (total == 0) return SNN_OK;
    SNN_CUDA(cudaSetDevice(ctx->device));
    double* d = nullptr;
    SNN_CUDA(cudaMalloc(&d, total * sizeof(double)));
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    k_fetch<<<blocks, 256, 0, ctx->stream>>>(c->band, c->nbins, c->bw, i0, i1, j0, j1, d);
    cudaError_t e = cudaMemcpyAsync(out, d, total * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = snn_stream_sync(ctx);
    cudaFree(d);
    if (e != cudaSuccess) return snn_cuda_fail(ctx, e, "snn_fetch");
    return SNN_OK;
}

// ---- seeimage: saturated heat map of a square window (seeimage.py:77-90) ------------------------
// blue = 255*(M - A)/M, negatives -> 0; img = (1, blue/255, blue/255) clipped to [0, 1]
// (same saturation as StripeSearch, getStripe.py:885-895).  `chan` gets the green/blue channel
// as the reference's float64 (NaN pixels stay NaN), `rgb` the 8-bit image the way matplotlib's
// imshow quantises a float RGB array ((255*v) truncated; NaN -> 0).  Pixels beyond the band
// come from the CSR table (snn_mat_at), so any window of the chromosome can be drawn.
__global__ void k_heatmap(MatView m, int64_t i0, int64_t n, double M, double* __restrict__ chan,
                          unsigned char* __restrict__ rgb) {
    const int64_t total = n * n;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = i0 + t / n, j = i0 + t % n;
        const double a = snn_mat_at(m, i, j);
        double blue = (255.0 * (M - a)) / M;
        if (blue < 0.0) blue = 0.0;
        double v = blue / 255.0;
        if (v < 0.0) v = 0.0;           // np.clip keeps NaN (both comparisons are false)
        if (v > 1.0) v = 1.0;
        if (chan) chan[t] = v;
        if (rgb) {
            const unsigned char g = (v != v) ? (unsigned char)0 : (unsigned char)(int)(255.0 * v);
            rgb[3 * t] = 255;
            rgb[3 * t + 1] = g;
            rgb[3 * t + 2] = g;
        }
    }
}

extern "C" int snn_heatmap(snn_ctx* ctx, int chrom_id, int64_t i0, int64_t i1, double maxpixel_value,
                           double* chan, unsigned char* rgb) {
    if (!ctx) return snn_fail(nullptr, SNN_EINVAL, "ctx is NULL");
    SnnChrom* c = snn_find_chrom(ctx, chrom_id);
    if (!c) return snn_fail(ctx, SNN_ESTATE, "chromosome %d is not loaded", chrom_id);
    SNN_CUDA(cudaSetDevice(ctx->device));
    if (int rc = snn_materialize(ctx, c)) return rc;
    if (i0 < 0 || i1 > c->nbins || i1 < i0 || (!chan && !rgb))
        return snn_fail(ctx, SNN_EINVAL, "snn_heatmap: bad window");
    const int64_t n = i1 - i0, total = n * n;
    if (total == 0) return SNN_OK;
    double* dch = nullptr;
    unsigned char* drgb = nullptr;
    cudaError_t e = cudaSuccess;
    if (chan) e = cudaMalloc(&dch, total * sizeof(double));
    if (e == cudaSuccess && rgb) e = cudaMalloc(&drgb, total * 3);
    if (e == cudaSuccess) {
        int64_t blocks = (total + 255) / 256;
        if (blocks > 148 * 8) blocks = 148 * 8;
        k_heatmap<<<(int)blocks, 256, 0, ctx->stream>>>(snn_view(c), i0, n, maxpixel_value, dch, drgb);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess && chan)
        e = cudaMemcpyAsync(chan, dch, total * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess && rgb) e = cudaMemcpyAsync(rgb, drgb, total * 3, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = snn_stream_sync(ctx);
    cudaFree(dch);
    cudaFree(drgb);
    if (e != cudaSuccess) return snn_cuda_fail(ctx, e, "snn_heatmap");
    return SNN_OK;
}

// ---- K-mp: expected value per diagonal -----------------------------------------------------
// ex[d] = sum_i nan0(band[i][d]) / (N - d), d < 400.  Two deterministic stages: per row-chunk
// partial column sums (coalesced along d), then one fixed-order reduction over chunks.
// Algorithmic bytes: nbins * min(400, bw) * 8.
#define EXP_ROWS 64    // rows per partial sum: 4x the CTAs of the first version (latency-bound at 2 CTAs per SM)

__global__ void k_expected_partial(const double* __restrict__ band, int64_t nbins, int bw, int nd,
                                   double* __restrict__ partial) {
    int d = blockIdx.y * blockDim.x + threadIdx.x;
    if (d >= nd) return;
    int64_t r0 = (int64_t)blockIdx.x * EXP_ROWS;
    int64_t r1 = r0 + EXP_ROWS;
    if (r1 > nbins - d) r1 = nbins - d;  // pixel (i, i+d) must exist
    double s = 0.0;
    for (int64_t i = r0; i < r1; ++i) s += snn_nan0(band[i * bw + d]);
    partial[(int64_t)blockIdx.x * nd + d] = s;
}

__global__ void k_expected_final(const double* __restrict__ partial, int nchunk, int nd, int64_t nbins,
                                 double* __restrict__ out) {
    int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= SNN_NDIST) return;
    double s = 0.0;
    if (d < nd)
        for (int k = 0; k < nchunk; ++k) s += partial[(int64_t)k * nd + d];
    out[d] = s / (double)(nbins - d);  // d >= nbins: 0 / non-positive, like the reference
}

int snn_expected_impl(snn_ctx* ctx, SnnChrom* c, double* out) {
    int nd = c->bw < SNN_NDIST ? c->bw : SNN_NDIST;
    if (c->bw < SNN_NDIST && c->bw < c->nbins)
        return snn_fail(ctx, SNN_ERANGE, "snn_expected: band width %d < %d", c->bw, SNN_NDIST);
    int nchunk = (int)((c->nbins + EXP_ROWS - 1) / EXP_ROWS);
    void* scratch = nullptr;
    int rc = snn_scratch(ctx, ((size_t)nchunk * nd + SNN_NDIST) * sizeof(double), &scratch);
    if (rc) return rc;
    double* partial = (double*)scratch;
    double* dout = partial + (size_t)nchunk * nd;
    {
        SnnScope scope(ctx, ST_EXPECTED, 2);
        dim3 grid(nchunk, (nd + 127) / 128);
        k_expected_partial<<<grid, 128, 0, ctx->stream>>>(c->band, c->nbins, c->bw, nd, partial);
        SNN_LAUNCH_CHECK("k_expected_partial");
        k_expected_final<<<(SNN_NDIST + 127) / 128, 128, 0, ctx->stream>>>(partial, nchunk, nd, c->nbins, dout);
        SNN_LAUNCH_CHECK("k_expected_final");
    }
    SNN_CUDA(cudaMemcpyAsync(out, dout, SNN_NDIST * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    SNN_CUDA(snn_stream_sync(ctx));
    return SNN_OK;
}

// ---- row flags for the background sampling plan ------------------------------------------------
// One pass over the COO pixels (exact for any off-diagonal distance, not only the band):
// a pixel (a, b) with a finite non-zero balanced value marks row a if column b lies in a's
// block / window, and row b if column a lies in b's (the matrix is symmetric).
// "sum != 0" of the reference is evaluated as "has a non-zero, non-NaN entry": counts are
// non-negative, so the two agree.
__global__ void k_row_flags(const int32_t* __restrict__ b1, const int32_t* __restrict__ b2,
                            const int32_t* __restrict__ cnt, const double* __restrict__ w, int64_t nnz,
                            const int64_t* __restrict__ geom, int nunits, uint8_t* __restrict__ avail,
                            uint8_t* __restrict__ live) {
    extern __shared__ int64_t g[];
    for (int t = threadIdx.x; t < nunits * 6; t += blockDim.x) g[t] = geom[t];
    __syncthreads();
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < nnz; t += stride) {
        int64_t a = b1[t], b = b2[t];
        double v = (double)cnt[t];
        if (w) v = v * (w[a] * w[b]);
        if (v != v || v == 0.0) continue;
        for (int u = 0; u < nunits; ++u) {
            const int64_t* q = g + u * 6;
            bool a_in = a >= q[0] && a < q[1], b_in = b >= q[0] && b < q[1];
            if (a_in && b_in) { avail[a] = 1; avail[b] = 1; }
            if (a >= q[2] && a < q[3] && b >= q[4] && b < q[5]) live[a] = 1;
            if (b >= q[2] && b < q[3] && a >= q[4] && a < q[5]) live[b] = 1;
        }
    }
}

int snn_row_flags_impl(snn_ctx* ctx, SnnChrom* c, int nunits, const int64_t* geom, uint8_t* avail,
                       uint8_t* live) {
    size_t nb = (size_t)c->nbins;
    size_t gbytes = (size_t)(nunits > 0 ? nunits : 1) * 6 * sizeof(int64_t);
    void* scratch = nullptr;
    int rc = snn_scratch(ctx, gbytes + 2 * nb + 64, &scratch);
    if (rc) return rc;
    int64_t* dgeom = (int64_t*)scratch;
    uint8_t* davail = (uint8_t*)scratch + gbytes;
    uint8_t* dlive = davail + nb;
    SNN_CUDA(cudaMemsetAsync(davail, 0, 2 * nb, ctx->stream));
    if (nunits > 0 && c->nnz > 0) {
        SNN_CUDA(cudaMemcpyAsync(dgeom, geom, (size_t)nunits * 6 * sizeof(int64_t), cudaMemcpyHostToDevice,
                                 ctx->stream));
        SnnScope scope(ctx, ST_FLAGS, 1);
        int blocks = (int)((c->nnz + 255) / 256);
        if (blocks > 148 * 8) blocks = 148 * 8;
        k_row_flags<<<blocks, 256, gbytes, ctx->stream>>>(c->bin1, c->bin2, c->count, c->weight, c->nnz, dgeom,
                                                          nunits, davail, dlive);
        SNN_LAUNCH_CHECK("k_row_flags");
    }
    SNN_CUDA(cudaMemcpyAsync(avail, davail, nb, cudaMemcpyDeviceToHost, ctx->stream));
    SNN_CUDA(cudaMemcpyAsync(live, dlive, nb, cudaMemcpyDeviceToHost, ctx->stream));
    SNN_CUDA(snn_stream_sync(ctx));
    return SNN_OK;
}
