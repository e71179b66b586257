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
#include <string.h>

#include "snn_internal.cuh"

#define QMAXK 32  // ranks handled per sweep (2 per quantile)

__device__ __forceinline__ bool pixel_key(const int32_t* b1, const int32_t* b2, const int32_t* cnt,
                                          const double* w, int64_t t, unsigned long long* key, unsigned* mult) {
    int32_t a = b1[t], b = b2[t];
    double v = (double)cnt[t];
    if (w) v = v * (w[a] * w[b]);
    if (!(v > 0.0)) return false;  // also drops NaN
    *key = (unsigned long long)__double_as_longlong(v);
    *mult = (a == b) ? 1u : 2u;
    return true;
}

__global__ void k_count_positive(const int32_t* __restrict__ b1, const int32_t* __restrict__ b2,
                                 const int32_t* __restrict__ cnt, const double* __restrict__ w, int64_t nnz,
                                 unsigned long long* __restrict__ total) {
    unsigned long long local = 0;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < nnz; t += stride) {
        unsigned long long key;
        unsigned mult;
        if (pixel_key(b1, b2, cnt, w, t, &key, &mult)) local += mult;
    }
    for (int o = 16; o > 0; o >>= 1) local += __shfl_down_sync(0xffffffffu, local, o);
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(total, local);
}

// state[k] = {prefix, remaining rank}; hist[k][256] global 64-bit counters.
__global__ void k_radix_hist(const int32_t* __restrict__ b1, const int32_t* __restrict__ b2,
                             const int32_t* __restrict__ cnt, const double* __restrict__ w, int64_t nnz, int K,
                             int shift, const unsigned long long* __restrict__ state,
                             unsigned long long* __restrict__ hist) {
    __shared__ unsigned sh[QMAXK * 256];
    __shared__ unsigned long long pre[QMAXK];
    for (int t = threadIdx.x; t < K * 256; t += blockDim.x) sh[t] = 0;
    if (threadIdx.x < K) pre[threadIdx.x] = state[2 * threadIdx.x];
    __syncthreads();
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < nnz; t += stride) {
        unsigned long long key;
        unsigned mult;
        if (!pixel_key(b1, b2, cnt, w, t, &key, &mult)) continue;
        unsigned digit = (unsigned)(key >> shift) & 255u;
        unsigned long long hi = (shift == 56) ? 0ull : (key >> (shift + 8));
        for (int k = 0; k < K; ++k) {
            unsigned long long ph = (shift == 56) ? 0ull : (pre[k] >> (shift + 8));
            if (hi == ph) atomicAdd(&sh[k * 256 + digit], mult);
        }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < K * 256; t += blockDim.x)
        if (sh[t]) atomicAdd(&hist[t], (unsigned long long)sh[t]);
}

__global__ void k_radix_pick(int K, int shift, unsigned long long* __restrict__ state,
                             unsigned long long* __restrict__ hist) {
    int k = threadIdx.x;
    if (k < K) {
        unsigned long long rank = state[2 * k + 1], cum = 0;
        int digit = 255;
        for (int d = 0; d < 256; ++d) {
            unsigned long long h = hist[k * 256 + d];
            if (cum + h > rank) { digit = d; break; }
            cum += h;
        }
        state[2 * k] |= ((unsigned long long)digit) << shift;
        state[2 * k + 1] = rank - cum;
    }
    __syncthreads();
    for (int t = threadIdx.x; t < K * 256; t += blockDim.x) hist[t] = 0;
}

int snn_quantiles_impl(snn_ctx* ctx, SnnChrom* c, int nq, const double* q, double* out) {
    void* scratch = nullptr;
    size_t bytes = (1 + 2 * QMAXK + (size_t)QMAXK * 256) * sizeof(unsigned long long);
    int rc = snn_scratch(ctx, bytes, &scratch);
    if (rc) return rc;
    unsigned long long* d_total = (unsigned long long*)scratch;
    unsigned long long* d_state = d_total + 1;
    unsigned long long* d_hist = d_state + 2 * QMAXK;
    int threads = 256;
    int blocks = (int)((c->nnz + threads - 1) / threads);
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks < 1) blocks = 1;

    SNN_CUDA(cudaMemsetAsync(scratch, 0, bytes, ctx->stream));
    unsigned long long n = 0;
    {
        SnnScope scope(ctx, ST_QUANTILE, 1);
        k_count_positive<<<blocks, threads, 0, ctx->stream>>>(c->bin1, c->bin2, c->count, c->weight, c->nnz,
                                                              d_total);
        SNN_LAUNCH_CHECK("k_count_positive");
    }
    SNN_CUDA(cudaMemcpyAsync(&n, d_total, sizeof n, cudaMemcpyDeviceToHost, ctx->stream));
    SNN_CUDA(cudaStreamSynchronize(ctx->stream));
    if (n == 0) {  // np.quantile of an empty array
        for (int i = 0; i < nq; ++i) out[i] = NAN;
        return SNN_OK;
    }
    for (int base = 0; base < nq; base += QMAXK / 2) {
        int m = nq - base < QMAXK / 2 ? nq - base : QMAXK / 2;
        int K = 2 * m;
        std::vector<unsigned long long> state(2 * K, 0ull);
        std::vector<double> gamma(m);
        std::vector<char> top(m, 0);
        for (int i = 0; i < m; ++i) {
            // 'linear': virtual index (n - 1) * q ; _get_indexes: floor, +1, clipped to the ends
            double vi = (double)(long long)(n - 1) * q[base + i];
            double fl = floor(vi);
            unsigned long long prev, next;
            if (vi >= (double)(long long)(n - 1)) {
                prev = next = n - 1;
                top[i] = 1;
            } else if (vi < 0) {
                prev = next = 0;
            } else {
                prev = (unsigned long long)fl;
                next = prev + 1;
            }
            gamma[i] = vi - fl;
            state[2 * (2 * i) + 1] = prev;
            state[2 * (2 * i + 1) + 1] = next;
        }
        SNN_CUDA(cudaMemcpyAsync(d_state, state.data(), state.size() * sizeof(unsigned long long),
                                 cudaMemcpyHostToDevice, ctx->stream));
        {
            SnnScope scope(ctx, ST_QUANTILE, 16);
            for (int shift = 56; shift >= 0; shift -= 8) {
                k_radix_hist<<<blocks, threads, 0, ctx->stream>>>(c->bin1, c->bin2, c->count, c->weight, c->nnz,
                                                                  K, shift, d_state, d_hist);
                SNN_LAUNCH_CHECK("k_radix_hist");
                k_radix_pick<<<1, 256, 0, ctx->stream>>>(K, shift, d_state, d_hist);
                SNN_LAUNCH_CHECK("k_radix_pick");
            }
        }
        SNN_CUDA(cudaMemcpyAsync(state.data(), d_state, state.size() * sizeof(unsigned long long),
                                 cudaMemcpyDeviceToHost, ctx->stream));
        SNN_CUDA(cudaStreamSynchronize(ctx->stream));
        for (int i = 0; i < m; ++i) {
            double a, b;
            memcpy(&a, &state[2 * (2 * i)], 8);
            memcpy(&b, &state[2 * (2 * i + 1)], 8);
            double t = gamma[i];
            if (top[i]) t = 0.0, b = a;
            // numpy _lerp
            volatile double diff = b - a;
            volatile double prod = diff * t;
            double r = a + prod;
            if (t >= 0.5) {
                volatile double omt = 1 - t;
                volatile double prod2 = diff * omt;
                r = b - prod2;
            }
            out[base + i] = r;
        }
    }
    return SNN_OK;
}
