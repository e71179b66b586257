// K2: exact order statistics of the positive pixels of the full symmetric matrix — replaces
// np.quantile(mat[mat > 0], q) of getStripe.getQuantile_original (getStripe.py:160-176).
//
// The population is never materialised: it is the multiset {v_ij x2 (i<j), v_ii x1} of the COO
// pixels with v = count * (w_i * w_j) > 0.  Positive doubles order like their bit patterns, so
// an 8-bit MSD radix select over the uint64 keys finds the ranks floor(vi) and floor(vi)+1 of
// every requested quantile in the same 8 passes (shared-memory histograms, one per rank).
// numpy's 'linear' method (numpy/lib/_function_base_impl.py: 'linear' -> (n-1)*q, _get_gamma,
// _lerp) is then applied on the host in the same floating-point order.
//
// Roofline: HBM.  Algorithmic bytes: 12 B per stored pixel (one pass); the 8 radix passes and
// the counting pass re-read the COO arrays (9 x 12 B per pixel of real traffic).
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "snn_internal.cuh"

#define QMAXK 32  // ranks handled per sweep (2 per quantile)

// One chromosome's pixel table and its slice of the per-chromosome work arrays.
struct QChrom {
    const int32_t* b1;
    const int32_t* b2;
    const int32_t* cnt;
    const double* w;
    int64_t nnz;
    int job;            // work area this source adds to (several sources may share one)
    int sel;            // 0: the whole symmetric matrix; 1: the row/column selection below
    // getQuantile_slow (getStripe.py:107-158): matrix entry (row i, column j) belongs to the
    // population iff c0 <= j < c1 and a = i - s satisfies 0 <= a < F and a % w != w - 1
    // (blocks of w rows, the last row of every block is dropped by the reference's slices)
    int64_t s, wrows, F, c0, c1;
};

__device__ __forceinline__ unsigned slow_in(const QChrom& c, int64_t row, int64_t col) {
    const int64_t a = row - c.s;
    return (col >= c.c0 && col < c.c1 && a >= 0 && a < c.F && (a % c.wrows) != c.wrows - 1) ? 1u : 0u;
}
// Per-chromosome work area (device): total, state[2*QMAXK], group[QMAXK+1] (+pad), hist[QMAXK*256]
#define QW_TOTAL 0
#define QW_STATE 1
#define QW_GROUP (1 + 2 * QMAXK)                       /* ints packed in 64-bit words */
#define QW_HIST (1 + 2 * QMAXK + (QMAXK + 2) / 2)
#define QW_WORDS (QW_HIST + QMAXK * 256)

__device__ __forceinline__ bool pixel_key(const QChrom& c, int64_t t, unsigned long long* key, unsigned* mult) {
    int32_t a = c.b1[t], b = c.b2[t];
    double v = (double)c.cnt[t];
    if (c.w) v = v * (c.w[a] * c.w[b]);
    if (!(v > 0.0)) return false;  // also drops NaN
    *key = (unsigned long long)__double_as_longlong(v);
    if (c.sel) {
        *mult = slow_in(c, a, b) + ((a != b) ? slow_in(c, b, a) : 0u);
        return *mult != 0u;
    }
    *mult = (a == b) ? 1u : 2u;
    return true;
}

// grid = (blocks per chromosome, chromosomes): every kernel below handles all chromosomes of a
// call in ONE launch (a chromosome's table is only ~35 MB: per-chromosome launches were
// latency-bound, 18 launches of ~15 us each).
__global__ void k_count_positive(const QChrom* __restrict__ chroms, unsigned long long* __restrict__ work) {
    const QChrom c = chroms[blockIdx.y];
    unsigned long long local = 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < c.nnz; t += stride) {
        unsigned long long key;
        unsigned mult;
        if (pixel_key(c, t, &key, &mult)) local += mult;
    }
    for (int o = 16; o > 0; o >>= 1) local += __shfl_down_sync(0xffffffffu, local, o);
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(work + (size_t)c.job * QW_WORDS + QW_TOTAL, local);
}

// state[k] = {prefix, remaining rank}; hist[g][256] 64-bit counters, one histogram per GROUP of
// ranks that still share their prefix (group[k], group[QMAXK] = number of groups: written by
// k_radix_pick — the ranks of neighbouring quantiles walk down the same digits for most of the
// passes, and in the first pass all of them share the empty prefix).  Lanes of a warp that hit
// the same (group, digit) counter are merged with __match_any_sync before the shared-memory
// atomic.
__global__ void k_radix_hist(const QChrom* __restrict__ chroms, int K, int shift,
                             unsigned long long* __restrict__ work) {
    __shared__ unsigned sh[QMAXK * 256];
    __shared__ unsigned long long pre[QMAXK];   // prefix (bits above the current digit) of each group
    const QChrom c = chroms[blockIdx.y];
    unsigned long long* wk = work + (size_t)c.job * QW_WORDS;
    const unsigned long long* state = wk + QW_STATE;
    const int* group = (const int*)(wk + QW_GROUP);
    unsigned long long* hist = wk + QW_HIST;
    const int ng = group[QMAXK];
    if (ng == 0) return;                         // chromosome without positive pixels
    for (int t = threadIdx.x; t < ng * 256; t += blockDim.x) sh[t] = 0;
    if (threadIdx.x < K) {
        const unsigned long long p = state[2 * threadIdx.x];
        pre[group[threadIdx.x]] = (shift == 56) ? 0ull : (p >> (shift + 8));   // same value from every member
    }
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t nround = (c.nnz + stride - 1) / stride;
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (int64_t it = 0; it < nround; ++it, t += stride) {     // whole warps stay in the loop (match_any)
        unsigned long long key = 0;
        unsigned mult = 0;
        const bool ok = t < c.nnz && pixel_key(c, t, &key, &mult);
        const unsigned digit = (unsigned)(key >> shift) & 255u;
        const unsigned long long hi = (shift == 56) ? 0ull : (key >> (shift + 8));
        for (int g = 0; g < ng; ++g) {
            const bool hit = ok && hi == pre[g];
            const unsigned act = __ballot_sync(0xffffffffu, hit);
            if (!act) continue;
            if (hit) {
                const unsigned peers = __match_any_sync(act, digit);
                const unsigned twos = __ballot_sync(act, mult == 2u) & peers;
                if ((int)(threadIdx.x & 31) == __ffs(peers) - 1)
                    atomicAdd(&sh[g * 256 + digit], (unsigned)__popc(peers) + (unsigned)__popc(twos));
            }
        }
    }
    __syncthreads();
    for (int q = threadIdx.x; q < ng * 256; q += blockDim.x)
        if (sh[q]) atomicAdd(&hist[q], (unsigned long long)sh[q]);
}

// ---- raw counts (--norm None): one-pass value histogram ----------------------------------------
// Without weights the population is made of small positive integers, so the order statistics
// can be read off an exact histogram of the VALUES: one pass over the pixel table instead of
// the counting pass + 8 radix passes (58 M pixels of the mouse genome: 0.5 instead of 4.3 ms per
// genome).  Bins 1..QI_NB-1 are exact; larger counts fall into one overflow bin, and a rank that
// lands there sends the whole call to the radix select.  Lanes of a warp holding the same value
// (1, 2, 3 dominate away from the diagonal) are merged with __match_any_sync before the
// shared-memory atomic, like in k_radix_hist.  hist[job][v] (64-bit) lives in the job's work area.
#define QI_NB 4096
__global__ void __launch_bounds__(256) k_int_hist(const QChrom* __restrict__ chroms,
                                                  unsigned long long* __restrict__ work) {
    __shared__ unsigned sh[QI_NB + 1];
    const QChrom c = chroms[blockIdx.y];
    unsigned long long* hist = work + (size_t)c.job * QW_WORDS;
    for (int t = threadIdx.x; t <= QI_NB; t += blockDim.x) sh[t] = 0;
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t nround = (c.nnz + stride - 1) / stride;
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (int64_t it = 0; it < nround; ++it, t += stride) {     // whole warps stay in the loop (match_any)
        unsigned mult = 0;
        int v = 0;
        if (t < c.nnz) {
            v = c.cnt[t];
            if (v > 0) {
                if (c.sel) {
                    const int32_t a = c.b1[t], b = c.b2[t];
                    mult = slow_in(c, a, b) + ((a != b) ? slow_in(c, b, a) : 0u);
                } else {
                    mult = (c.b1[t] == c.b2[t]) ? 1u : 2u;
                }
            }
        }
        const bool hit = mult != 0u;
        const unsigned act = __ballot_sync(0xffffffffu, hit);
        if (hit) {
            const unsigned bin = v < QI_NB ? (unsigned)v : (unsigned)QI_NB;
            const unsigned peers = __match_any_sync(act, bin);
            const unsigned twos = __ballot_sync(act, mult == 2u) & peers;
            if ((int)(threadIdx.x & 31) == __ffs(peers) - 1)
                atomicAdd(&sh[bin], (unsigned)__popc(peers) + (unsigned)__popc(twos));
        }
    }
    __syncthreads();
    for (int q = threadIdx.x; q <= QI_NB; q += blockDim.x)
        if (sh[q]) atomicAdd(&hist[q], (unsigned long long)sh[q]);
}

// One CTA per chromosome: one thread per rank picks its digit from its group's histogram;
// thread 0 then regroups the ranks by their (longer) prefixes for the next pass.
__global__ void k_radix_pick(int K, int shift, unsigned long long* __restrict__ work) {
    __shared__ unsigned long long npre[QMAXK];
    unsigned long long* wk = work + (size_t)blockIdx.x * QW_WORDS;
    unsigned long long* state = wk + QW_STATE;
    int* group = (int*)(wk + QW_GROUP);
    unsigned long long* hist = wk + QW_HIST;
    if (group[QMAXK] == 0) return;
    const int k = threadIdx.x;
    if (k < K) {
        const unsigned long long* h = hist + (size_t)group[k] * 256;
        unsigned long long rank = state[2 * k + 1], cum = 0;
        int digit = 255;
        for (int d = 0; d < 256; ++d) {
            const unsigned long long cc = h[d];
            if (cum + cc > rank) { digit = d; break; }
            cum += cc;
        }
        npre[k] = state[2 * k] | (((unsigned long long)digit) << shift);
        state[2 * k] = npre[k];
        state[2 * k + 1] = rank - cum;
    }
    __syncthreads();
    if (k == 0) {
        int ng = 0;
        for (int a = 0; a < K; ++a) {
            int g = -1;
            for (int b = 0; b < a; ++b)
                if (npre[b] == npre[a]) { g = group[b]; break; }
            if (g < 0) g = ng++;
            group[a] = g;
        }
        group[QMAXK] = ng;
    }
    __syncthreads();
    for (int t = threadIdx.x; t < QMAXK * 256; t += blockDim.x) hist[t] = 0;
}

// numpy 'linear': virtual index (n - 1) * q ; _get_indexes: floor, +1, clipped to the ends
static void np_ranks(unsigned long long n, double q, unsigned long long* prev, unsigned long long* next,
                     double* gamma, char* top) {
    const double vi = (double)(long long)(n - 1) * q;
    const double fl = floor(vi);
    *top = 0;
    if (vi >= (double)(long long)(n - 1)) {
        *prev = *next = n - 1;
        *top = 1;
    } else if (vi < 0) {
        *prev = *next = 0;
    } else {
        *prev = (unsigned long long)fl;
        *next = *prev + 1;
    }
    *gamma = vi - fl;
}

// numpy _lerp, in its floating-point order
static double np_lerp(double a, double b, double t, char top) {
    if (top) t = 0.0, b = a;
    volatile double diff = b - a;
    volatile double prod = diff * t;
    double r = a + prod;
    if (t >= 0.5) {
        volatile double omt = 1 - t;
        volatile double prod2 = diff * omt;
        r = b - prod2;
    }
    return r;
}

static_assert(QI_NB + 1 <= QW_WORDS, "the value histogram must fit the job's work area");

int snn_quantiles_sources_impl(snn_ctx* ctx, const std::vector<QSource>& src, int njob, int nq, const double* q,
                               double* out) {
    const int ns = (int)src.size();
    const int nc = njob;
    if (ns == 0 || nc == 0) return SNN_OK;
    void* scratch = nullptr;
    const size_t wbytes = (size_t)nc * QW_WORDS * sizeof(unsigned long long);
    const size_t cbytes = (((size_t)ns * sizeof(QChrom)) + 255) & ~(size_t)255;
    int rc = snn_scratch(ctx, cbytes + wbytes, &scratch);
    if (rc) return rc;
    QChrom* d_chroms = (QChrom*)scratch;
    unsigned long long* d_work = (unsigned long long*)((char*)scratch + cbytes);
    std::vector<QChrom> hc((size_t)ns);
    int64_t maxnnz = 1;
    for (int i = 0; i < ns; ++i) {
        const SnnChrom* c = src[i].c;
        hc[i].b1 = c->bin1;
        hc[i].b2 = c->bin2;
        hc[i].cnt = c->count;
        hc[i].w = c->weight;
        hc[i].nnz = c->nnz;
        hc[i].job = src[i].job;
        hc[i].sel = src[i].sel ? 1 : 0;
        hc[i].s = src[i].s;
        hc[i].wrows = src[i].wrows > 0 ? src[i].wrows : 1;
        hc[i].F = src[i].F;
        hc[i].c0 = src[i].c0;
        hc[i].c1 = src[i].c1;
        if (c->nnz > maxnnz) maxnnz = c->nnz;
    }
    const int threads = 256;
    int blocks = (int)((maxnnz + threads * 8 - 1) / (threads * 8));    // ~8 pixels per thread
    const int cap = (148 * 8 + ns - 1) / ns;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    const dim3 grid(blocks, ns);

    SNN_CUDA(cudaMemcpyAsync(d_chroms, hc.data(), (size_t)ns * sizeof(QChrom), cudaMemcpyHostToDevice, ctx->stream));
    SNN_CUDA(cudaMemsetAsync(d_work, 0, wbytes, ctx->stream));
    bool raw_counts = getenv("SNN_NO_INT_HIST") == nullptr;      // A/B switch and test hook
    for (int i = 0; i < ns; ++i) raw_counts = raw_counts && hc[i].w == nullptr;
    if (raw_counts) {
        {
            SnnScope scope(ctx, ST_QUANTILE, 1);
            k_int_hist<<<grid, threads, 0, ctx->stream>>>(d_chroms, d_work);
            SNN_LAUNCH_CHECK("k_int_hist");
        }
        std::vector<unsigned long long> hh((size_t)nc * (QI_NB + 1));
        SNN_CUDA(cudaMemcpy2DAsync(hh.data(), (QI_NB + 1) * 8, d_work, QW_WORDS * 8, (QI_NB + 1) * 8, nc,
                                   cudaMemcpyDeviceToHost, ctx->stream));
        SNN_CUDA(snn_stream_sync(ctx));
        bool overflow = false;
        for (int ci = 0; ci < nc && !overflow; ++ci) {
            const unsigned long long* h = hh.data() + (size_t)ci * (QI_NB + 1);
            unsigned long long n = 0;
            for (int v = 1; v <= QI_NB; ++v) n += h[v];
            for (int i = 0; i < nq; ++i) {
                double* dst = out + (size_t)ci * nq + i;
                if (!n) { *dst = NAN; continue; }              // np.quantile of an empty array
                unsigned long long prev, next;
                double gamma;
                char top;
                np_ranks(n, q[i], &prev, &next, &gamma, &top);
                unsigned long long cum = 0;
                int va = -1, vb = -1;
                for (int v = 1; v < QI_NB && vb < 0; ++v) {
                    cum += h[v];
                    if (va < 0 && cum > prev) va = v;
                    if (cum > next) vb = v;
                }
                if (va < 0 || vb < 0) { overflow = true; break; }   // a rank beyond the exact bins
                *dst = np_lerp((double)va, (double)vb, gamma, top);
            }
        }
        if (!overflow) return SNN_OK;
        SNN_CUDA(cudaMemsetAsync(d_work, 0, wbytes, ctx->stream));   // large counts: radix select below
    }
    {
        SnnScope scope(ctx, ST_QUANTILE, 1);
        k_count_positive<<<grid, threads, 0, ctx->stream>>>(d_chroms, d_work);
        SNN_LAUNCH_CHECK("k_count_positive");
    }
    std::vector<unsigned long long> hw((size_t)nc * QW_WORDS);
    // only the heads of the work areas travel: total, state, group
    SNN_CUDA(cudaMemcpy2DAsync(hw.data(), QW_WORDS * 8, d_work, QW_WORDS * 8, QW_HIST * 8, nc, cudaMemcpyDeviceToHost,
                               ctx->stream));
    SNN_CUDA(snn_stream_sync(ctx));
    for (int base = 0; base < nq; base += QMAXK / 2) {
        const int m = nq - base < QMAXK / 2 ? nq - base : QMAXK / 2;
        const int K = 2 * m;
        std::vector<double> gamma((size_t)nc * m);
        std::vector<char> top((size_t)nc * m, 0);
        for (int ci = 0; ci < nc; ++ci) {
            unsigned long long* wk = hw.data() + (size_t)ci * QW_WORDS;
            const unsigned long long n = wk[QW_TOTAL];
            unsigned long long* state = wk + QW_STATE;
            int* group = (int*)(wk + QW_GROUP);
            for (int i = 0; i < 2 * QMAXK; ++i) state[i] = 0ull;
            for (int i = 0; i < QMAXK; ++i) group[i] = 0;      // every rank starts with the empty prefix
            group[QMAXK] = n ? 1 : 0;
            if (!n) continue;
            for (int i = 0; i < m; ++i) {
                unsigned long long prev, next;
                np_ranks(n, q[base + i], &prev, &next, &gamma[(size_t)ci * m + i], &top[(size_t)ci * m + i]);
                state[2 * (2 * i) + 1] = prev;
                state[2 * (2 * i + 1) + 1] = next;
            }
        }
        SNN_CUDA(cudaMemcpy2DAsync(d_work, QW_WORDS * 8, hw.data(), QW_WORDS * 8, QW_HIST * 8, nc,
                                   cudaMemcpyHostToDevice, ctx->stream));
        {
            SnnScope scope(ctx, ST_QUANTILE, 16);
            for (int shift = 56; shift >= 0; shift -= 8) {
                k_radix_hist<<<grid, threads, 0, ctx->stream>>>(d_chroms, K, shift, d_work);
                SNN_LAUNCH_CHECK("k_radix_hist");
                k_radix_pick<<<nc, 256, 0, ctx->stream>>>(K, shift, d_work);
                SNN_LAUNCH_CHECK("k_radix_pick");
            }
        }
        SNN_CUDA(cudaMemcpy2DAsync(hw.data(), QW_WORDS * 8, d_work, QW_WORDS * 8, QW_HIST * 8, nc,
                                   cudaMemcpyDeviceToHost, ctx->stream));
        SNN_CUDA(snn_stream_sync(ctx));
        for (int ci = 0; ci < nc; ++ci) {
            const unsigned long long* wk = hw.data() + (size_t)ci * QW_WORDS;
            const unsigned long long* state = wk + QW_STATE;
            for (int i = 0; i < m; ++i) {
                double* dst = out + (size_t)ci * nq + base + i;
                if (!wk[QW_TOTAL]) { *dst = NAN; continue; }   // np.quantile of an empty array
                double a, b;
                memcpy(&a, &state[2 * (2 * i)], 8);
                memcpy(&b, &state[2 * (2 * i + 1)], 8);
                const double r = np_lerp(a, b, gamma[(size_t)ci * m + i], top[(size_t)ci * m + i]);
                *dst = r;
            }
        }
    }
    return SNN_OK;
}

int snn_quantiles_multi_impl(snn_ctx* ctx, const std::vector<SnnChrom*>& cs, int nq, const double* q,
                             double* out) {
    std::vector<QSource> src(cs.size());
    for (size_t i = 0; i < cs.size(); ++i) {
        src[i].c = cs[i];
        src[i].job = (int)i;
    }
    return snn_quantiles_sources_impl(ctx, src, (int)cs.size(), nq, q, out);
}

int snn_quantiles_impl(snn_ctx* ctx, SnnChrom* c, int nq, const double* q, double* out) {
    std::vector<SnnChrom*> one(1, c);
    return snn_quantiles_multi_impl(ctx, one, nq, q, out);
}
