// K8: the sampled background distribution of getStripe.nulldist / main_null_calc
// (getStripe.py:285-479) and its preparation for the p-value kernel.
//
// The host draws the samples (Python random.Random, like the reference) and passes, per
// sample, the extent of the sub-matrix the reference fetched for the sample's unit
// (rows [i0,i1), cols [j0,j1) in bins), the local row x and the column offset actually
// applied (+400 or 0 — the reference's stale-`it` quirk lives in that number).  For each
// distance j < 400 and both directions the kernel evaluates
//     mean(centre block) - mean(block above)   /   mean(centre block) - mean(block below)
// with numpy slice semantics ON THE FETCHED SUB-MATRIX (stops clipped to the extent, negative
// starts wrapped, empty slice -> NaN), exactly like getStripe.py:347-368.
//
// Roofline: L1/L2-resident gathers (each sample touches a 3*bs x (800+bs) strip of the band,
// ≈194 KB at 5 kb, re-read bs times by neighbouring distances); HBM sees the strip once.
#include "snn_internal.cuh"

__device__ __forceinline__ double block_mean(const MatView& m, int64_t i0, int64_t j0, int64_t ra, int64_t rb,
                                             int64_t ca, int64_t cb) {
    const int64_t nr = rb - ra, nc = cb - ca;
    if (nr <= 0 || nc <= 0) return __longlong_as_double(0x7ff8000000000000LL);  // np.mean of empty
    double s = 0.0;
    for (int64_t r = ra; r < rb; ++r)
        for (int64_t c = ca; c < cb; ++c) s += snn_nan0(snn_mat_at(m, i0 + r, j0 + c));
    return s / (double)(nr * nc);
}

// sample_major = 0: out[table][distance][sample] (snn_background, what the host API returns);
// sample_major = 1: out[slot0 + sample][table][distance] (snn_background_dev: the columns of one
// chromosome are a contiguous slot range, the unit the GPUs exchange)
__global__ void __launch_bounds__(256) k_background(MatView m, const int32_t* __restrict__ spec, int64_t nsamp,
                                                    int bs, int bu, int bd, double* __restrict__ out,
                                                    int sample_major, int64_t slot0) {
    const int64_t k = blockIdx.x;
    const int32_t* sp = spec + k * 6;
    const int64_t i0 = sp[0], i1 = sp[1], j0 = sp[2], j1 = sp[3], x = sp[4], shift = sp[5];
    const int64_t nr = i1 - i0, nc = j1 - j0;
    int64_t ra[3], rb[3];
    snn_pyslice(x - bu - bs, x - bu, nr, &ra[0], &rb[0]);
    snn_pyslice(x - bu, x + bd, nr, &ra[1], &rb[1]);
    snn_pyslice(x + bd, x + bd + bs, nr, &ra[2], &rb[2]);
    for (int t = threadIdx.x; t < 2 * SNN_NDIST; t += blockDim.x) {
        const int j = t % SNN_NDIST;
        const int down = t / SNN_NDIST;  // 0: towards the diagonal (up), 1: away (down)
        const int64_t y = x + (down ? j : -j) + shift;
        int64_t ca, cb;
        snn_pyslice(y - bu, y + bd, nc, &ca, &cb);
        const double top = block_mean(m, i0, j0, ra[0], rb[0], ca, cb);
        const double mid = block_mean(m, i0, j0, ra[1], rb[1], ca, cb);
        const double bot = block_mean(m, i0, j0, ra[2], rb[2], ca, cb);
        // tables: 0 bgleft_up, 1 bgright_up, 2 bgleft_down, 3 bgright_down
        if (sample_major) {
            double* o = out + (slot0 + k) * (4 * SNN_NDIST);
            o[(down ? 2 : 0) * SNN_NDIST + j] = mid - top;
            o[(down ? 3 : 1) * SNN_NDIST + j] = mid - bot;
        } else {
            out[((int64_t)(down ? 2 : 0) * SNN_NDIST + j) * nsamp + k] = mid - top;
            out[((int64_t)(down ? 3 : 1) * SNN_NDIST + j) * nsamp + k] = mid - bot;
        }
    }
}

static int check_specs(snn_ctx* ctx, SnnChrom* c, int64_t nsamp, const int32_t* spec) {
    for (int64_t k = 0; k < nsamp; ++k) {
        const int32_t* sp = spec + k * 6;
        if (sp[0] < 0 || sp[1] > c->nbins || sp[2] < 0 || sp[3] > c->nbins || sp[1] < sp[0] || sp[3] < sp[2])
            return snn_fail(ctx, SNN_EINVAL, "snn_background: sample %lld has a window outside the chromosome",
                            (long long)k);
    }
    return SNN_OK;
}

// d_ext = null: table-major result copied to the host array `out`; else sample-major into d_ext
static int background_run(snn_ctx* ctx, SnnChrom* c, int64_t nsamp, const int32_t* spec, double* out,
                          double* d_ext, int64_t slot0) {
    int rc = check_specs(ctx, c, nsamp, spec);
    if (rc) return rc;
    const double bsf = 50000.0 / (double)c->binsize;  // getStripe.py:287-290
    const int bs = (int)bsf;
    const int bu = (int)floor(bsf / 2.0);
    const int bd = (int)(bsf - floor(bsf / 2.0));
    void* scratch = nullptr;
    size_t sbytes = ((size_t)nsamp * 6 * sizeof(int32_t) + 255) & ~(size_t)255;
    size_t obytes = d_ext ? 0 : (size_t)4 * SNN_NDIST * nsamp * sizeof(double);
    rc = snn_scratch(ctx, sbytes + obytes, &scratch);
    if (rc) return rc;
    int32_t* dspec = (int32_t*)scratch;
    double* dout = d_ext ? d_ext : (double*)((char*)scratch + sbytes);
    SNN_CUDA(cudaMemcpyAsync(dspec, spec, (size_t)nsamp * 6 * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    {
        SnnScope scope(ctx, ST_BACKGROUND, 1);
        k_background<<<(unsigned)nsamp, 256, 0, ctx->stream>>>(snn_view(c), dspec, nsamp, bs, bu, bd, dout,
                                                                d_ext ? 1 : 0, slot0);
        SNN_LAUNCH_CHECK("k_background");
    }
    if (!d_ext) SNN_CUDA(cudaMemcpyAsync(out, dout, obytes, cudaMemcpyDeviceToHost, ctx->stream));
    SNN_CUDA(snn_stream_sync(ctx));
    return SNN_OK;
}

int snn_background_impl(snn_ctx* ctx, SnnChrom* c, int64_t nsamp, const int32_t* spec, double* out) {
    return background_run(ctx, c, nsamp, spec, out, nullptr, 0);
}

int snn_background_dev_impl(snn_ctx* ctx, SnnChrom* c, int64_t nsamp, const int32_t* spec, double* d_buf,
                            int64_t slot0) {
    return background_run(ctx, c, nsamp, spec, nullptr, d_buf, slot0);
}

// ---- sorted copy of the genome-wide tables for the p-value kernel ---------------------------------
// One CTA per (table, distance) row: bitonic sort in shared memory with NaNs sent to the end;
// #(bg >= v) then becomes n_valid - lower_bound(v).
// slots = null: in[row][col] (the host tables); else sample-major in[slots[col]][row]
__global__ void __launch_bounds__(512) k_bg_sort(const double* __restrict__ in, int64_t ncol, int npow2,
                                                 double* __restrict__ out, int32_t* __restrict__ nvalid,
                                                 double* __restrict__ nrcp, float4* __restrict__ tree,
                                                 int levels,
                                                 const int32_t* __restrict__ slots, int use_slots) {
    extern __shared__ double s[];
    const int64_t row = blockIdx.x;
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    int local = 0;
    for (int t = threadIdx.x; t < npow2; t += blockDim.x) {
        double v = inf;
        bool pad = true;
        if (t < ncol) {
            v = use_slots ? in[(int64_t)(slots ? slots[t] : t) * (4 * SNN_NDIST) + row] : in[row * ncol + t];
            if (v == v) { ++local; pad = false; }
        }
        // NaNs and padding sort last: keep real +inf values ahead of them with a flag trick —
        // encode padding as NaN-free +inf and fix the count afterwards (valid = non-NaN count).
        s[t] = pad ? inf : v;
    }
    __shared__ int total;
    if (threadIdx.x == 0) total = 0;
    __syncthreads();
    if (local) atomicAdd(&total, local);
    for (int k = 2; k <= npow2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            __syncthreads();
            for (int t = threadIdx.x; t < npow2; t += blockDim.x) {
                int p = t ^ j;
                if (p > t) {
                    double a = s[t], b = s[p];
                    bool up = (t & k) == 0;
                    if ((a > b) == up) { s[t] = b; s[p] = a; }
                }
            }
        }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < ncol; t += blockDim.x) out[row * ncol + t] = s[t];
    if (threadIdx.x == 0) {
        nvalid[row] = total;
        nrcp[row] = 1.0 / (double)total;
    }
    // 5-ary search tree for the fast scoring kernel: node q of level l holds the keys of the sorted
    // positions (5q + r + 1) * 5^(levels-1-l) - 1, r = 0..3, as float32 rounded toward -inf (so that
    // key < RU(v) whenever entry < v), +inf past the row.  Descending with q' = 5q + #(key < RU(v)) ends
    // at q = #(key < RU(v)) over the whole row, an upper bound of #(entry < v) that is exact unless an
    // entry lies within one float32 ulp above v (score_fast.cu checks one float64 entry).
    if (tree) {
        int nodes = 1, total_nodes = 0;
        for (int l = 0; l < levels; ++l) { total_nodes += nodes; nodes *= 5; }
        float4* trow = tree + row * total_nodes;
        int base = 0, cnt = 1, stride = nodes / 5;          // stride of level 0 = 5^(levels-1)
        for (int l = 0; l < levels; ++l) {
            for (int q = threadIdx.x; q < cnt; q += blockDim.x) {
                float k[4];
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const int64_t posn = (int64_t)(5 * q + r + 1) * stride - 1;
                    k[r] = posn < ncol ? __double2float_rd(s[posn]) : __int_as_float(0x7f800000);
                }
                trow[base + q] = make_float4(k[0], k[1], k[2], k[3]);
            }
            base += cnt;
            cnt *= 5;
            stride /= 5;
        }
    }
}

// tables != null: host tables [4][400][ncol]; else sample-major device buffer d_buf with the host
// slot list `slots` (null = identity)
static int set_background_run(snn_ctx* ctx, int64_t ncol, const double* tables, const int32_t* slots,
                              const double* d_buf) {
    if (ncol != ctx->bg_ncol) {   // keep the buffers when the shape repeats (every compute step)
        cudaFree(ctx->bg_sorted);
        cudaFree(ctx->bg_valid);
        cudaFree(ctx->bg_rcp);
        cudaFree(ctx->bg_tree);
        ctx->bg_tree = nullptr;
        ctx->bg_tree_levels = 0;
        ctx->bg_sorted = nullptr;
        ctx->bg_valid = nullptr;
        ctx->bg_rcp = nullptr;
    }
    ctx->bg_ncol = 0;
    if (ncol == 0) return SNN_OK;
    int npow2 = 1;
    while (npow2 < ncol) npow2 <<= 1;
    if ((size_t)npow2 * sizeof(double) > 200 * 1024)
        return snn_fail(ctx, SNN_EINVAL, "snn_set_background: %lld columns exceed the supported 25600",
                        (long long)ncol);
    const int64_t rows = 4 * SNN_NDIST;
    size_t bytes = (size_t)rows * ncol * sizeof(double);
    void* scratch = nullptr;
    int rc = snn_scratch(ctx, tables ? bytes : (size_t)ncol * sizeof(int32_t) + 256, &scratch);
    if (rc) return rc;
    if (!ctx->bg_sorted) {
        SNN_CUDA(cudaMalloc(&ctx->bg_sorted, bytes));
        SNN_CUDA(cudaMalloc(&ctx->bg_valid, rows * sizeof(int32_t)));
        SNN_CUDA(cudaMalloc(&ctx->bg_rcp, rows * sizeof(double)));
        int levels = 1;
        int64_t reach = 5;                                   // 5^levels positions
        while (reach < ncol + 1) { reach *= 5; ++levels; }
        if (levels <= 5) {                                   // rows up to 3124 entries: 12.5 KB of keys per row
            SNN_CUDA(cudaMalloc(&ctx->bg_tree, (size_t)rows * (size_t)((reach - 1) / 4) * sizeof(float4)));
            ctx->bg_tree_levels = levels;
        }
    }
    const double* src = d_buf;
    const int32_t* dslots = nullptr;
    if (tables) {
        SNN_CUDA(cudaMemcpyAsync(scratch, tables, bytes, cudaMemcpyHostToDevice, ctx->stream));
        src = (const double*)scratch;
    } else if (slots) {
        SNN_CUDA(cudaMemcpyAsync(scratch, slots, (size_t)ncol * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
        dslots = (const int32_t*)scratch;
    }
    size_t smem = (size_t)npow2 * sizeof(double);
    // constant across calls (host threads may race on the attribute otherwise)
    SNN_CUDA(cudaFuncSetAttribute(k_bg_sort, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    {
        SnnScope scope(ctx, ST_BGSORT, 1);
        k_bg_sort<<<(unsigned)rows, 512, smem, ctx->stream>>>(src, ncol, npow2, ctx->bg_sorted, ctx->bg_valid, ctx->bg_rcp, ctx->bg_tree,
                                                              ctx->bg_tree_levels, dslots,
                                                              tables ? 0 : 1);
        SNN_LAUNCH_CHECK("k_bg_sort");
    }
    SNN_CUDA(snn_stream_sync(ctx));
    ctx->bg_ncol = ncol;
    return SNN_OK;
}

int snn_set_background_impl(snn_ctx* ctx, int64_t ncol, const double* tables) {
    return set_background_run(ctx, ncol, tables, nullptr, nullptr);
}

int snn_set_background_dev_impl(snn_ctx* ctx, int64_t ncol, const int32_t* slots, const double* d_buf) {
    return set_background_run(ctx, ncol, nullptr, slots, d_buf);
}
