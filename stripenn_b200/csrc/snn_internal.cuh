// Internal declarations shared by the translation units of libstripenn_b200.so.
// Everything here is device-side plumbing for the C ABI in include/stripenn_b200.h.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/stripenn_b200.h"

#define SNN_SMAX 400  // largest search window (getStripe.py:794-801)
#define SNN_WPR 13    // 32-bit words per bit-packed window row (13*32 = 416 >= 400)

// ---- stage ids for the CUDA-event timers (snn_stage_times) -------------------------------
enum SnnStage {
    ST_BAND = 0,
    ST_QUANTILE,
    ST_EXPECTED,
    ST_FLAGS,
    ST_BACKGROUND,
    ST_BGSORT,
    ST_FRAME,
    ST_GRAY,
    ST_GAUSS,
    ST_NMS,
    ST_SEARCH,
    ST_FINISH,
    ST_SCORE,
    ST_GRAY_GV,   // k_gray_gv (fused gray + axis-0 Gaussian)
    ST_GH_NMS,    // k_gh_nms (fused axis-1 Gaussian + Sobel + NMS)
    ST_COUNT
};

struct SnnChrom {
    int id = -1;
    int64_t nbins = 0, binsize = 0, chromsize = 0, nnz = 0;
    int bw = 0;
    bool own_coo = true;
    int32_t* bin1 = nullptr;   // device
    int32_t* bin2 = nullptr;   // device
    int32_t* count = nullptr;  // device
    double* weight = nullptr;  // device or null
    double* band = nullptr;    // device, nbins * bw
    // device, nbins * (2*bw - 1): the band stored by COLUMN with both sides of the diagonal,
    // fband[col * (2*bw-1) + (row - col) + bw - 1] = pixel (row, col); built on the first fast-scoring
    // call (snn_fband_build), so a stripe's column is one contiguous run of rows (score_fast.cu)
    double* fband = nullptr;
    uint8_t* nanbin = nullptr;  // device, nbins: weight is NaN (null without weights)
    int32_t* nanpre = nullptr;  // device, nbins + 1: exclusive prefix counts of nanbin
    int fband_state = 0;        // 0 not built, 1 ready, 2 not eligible (band.cu: snn_fband_build)
    int64_t* rowptr = nullptr; // device, nbins + 1: CSR row pointers into the (bin1, bin2)-sorted pixels
    // snn_load_chrom_async: the pixel arrays are still travelling on the copy stream; `ready` fires
    // when they have landed and the band is built on first use (snn_materialize)
    bool pending = false;
    cudaEvent_t ready = nullptr;
};

// Read-only view of one chromosome's matrix for kernels that may touch ANY pixel: the band
// answers |i - j| < bw, a binary search in the CSR row answers the (rare) far pixels.
struct MatView {
    const double* band;
    int64_t nbins;
    int bw;
    const int64_t* rowptr;
    const int32_t* bin2;
    const int32_t* count;
    const double* weight;
};

inline MatView snn_view(const SnnChrom* c) {
    MatView v;
    v.band = c->band;
    v.nbins = c->nbins;
    v.bw = c->bw;
    v.rowptr = c->rowptr;
    v.bin2 = c->bin2;
    v.count = c->count;
    v.weight = c->weight;
    return v;
}

struct SnnPending {
    cudaEvent_t first, second;
    int n;
};
struct SnnTimer {
    std::vector<SnnPending> pending;
    double ms = 0.0;
    int64_t launches = 0;
};

struct snn_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t own_stream = nullptr;
    cudaStream_t copy_stream = nullptr;   // H2D copies of snn_load_chrom_async
    std::map<int, SnnChrom> chroms;
    mutable std::string err;
    // background tables, sorted per (table, distance): 4 * 400 rows of `bg_ncol` doubles,
    // NaNs moved to the end; bg_valid[row] = number of non-NaN entries.
    double* bg_sorted = nullptr;
    int32_t* bg_valid = nullptr;
    double* bg_rcp = nullptr;      // 1 / bg_valid[row] (correctly rounded; inf for an empty row)
    // 5-ary search tree over each sorted row with float32 keys rounded DOWN (background.cu: k_bg_sort):
    // bg_tree_levels levels, (5^levels - 1) / 4 nodes of four keys per row; null when the rows are too
    // long for five levels (the fast scoring kernel then searches the float64 rows directly)
    float4* bg_tree = nullptr;
    int bg_tree_levels = 0;
    int64_t bg_ncol = 0;
    // --mask of scoringstripes (snn_set_mask): chromosome id (< 0: none), start/end in bp
    int mask_chrom = -1;
    int64_t mask_start = 0, mask_end = 0;
    // scratch of snn_extract, grown on demand
    void* scratch = nullptr;
    size_t scratch_bytes = 0;
    // SNN_FORCE_GENERIC=1: run the unspecialised kernels of image.cu (A/B parity tests)
    bool force_generic = false;
    // SNN_BLOCKING_SYNC=1: wait for the stream on a blocking event instead of cudaStreamSynchronize's
    // spin loop — for hosts where the waiting threads of all processes outnumber the cores (8 GPUs x 2
    // device threads on a 16-core box: the spinners starved each other, sporadic +40..70 ms steps)
    bool blocking_sync = false;
    cudaEvent_t sync_event = nullptr;
    // stripes routed to the fast / warp / CTA scoring kernel (snn_score_routing)
    int64_t score_routed[3] = {0, 0, 0};
    // timers
    bool timing = false;
    SnnTimer timers[ST_COUNT];
    std::vector<cudaEvent_t> event_pool;
};

int snn_fail(const snn_ctx* ctx, int code, const char* fmt, ...);
int snn_cuda_fail(const snn_ctx* ctx, cudaError_t e, const char* what);
SnnChrom* snn_find_chrom(snn_ctx* ctx, int chrom_id);
// Finish an asynchronous load: make the compute stream wait for the copy, build the band.
int snn_materialize(snn_ctx* ctx, SnnChrom* c);
int snn_scratch(snn_ctx* ctx, size_t bytes, void** out);
// Wait until everything queued on the context's stream has finished (see snn_ctx::blocking_sync).
cudaError_t snn_stream_sync(snn_ctx* ctx);

#define SNN_CUDA(call)                                                   \
    do {                                                                 \
        cudaError_t e__ = (call);                                        \
        if (e__ != cudaSuccess) return snn_cuda_fail(ctx, e__, #call);   \
    } while (0)

#define SNN_LAUNCH_CHECK(what)                                             \
    do {                                                                   \
        cudaError_t e__ = cudaGetLastError();                              \
        if (e__ != cudaSuccess) return snn_cuda_fail(ctx, e__, what);      \
    } while (0)

// RAII stage timer: records two events around the launches issued in its scope.
struct SnnScope {
    snn_ctx* ctx;
    int stage;
    int n;  // kernel launches issued inside the scope
    cudaEvent_t a = nullptr, b = nullptr;
    SnnScope(snn_ctx* c, int s, int launches = 1);
    ~SnnScope();
};

// ---- device helpers ---------------------------------------------------------------------
__device__ __forceinline__ double snn_nan0(double v) { return (v != v) ? 0.0 : v; }

// Symmetric matrix element from the band; (i, j) chromosome-local bins, must satisfy
// |i - j| < bw (callers guarantee it; out-of-chromosome reads return 0).
__device__ __forceinline__ double snn_band_at(const double* __restrict__ band, int64_t nbins, int bw,
                                               int64_t i, int64_t j) {
    if (i < 0 || j < 0 || i >= nbins || j >= nbins) return 0.0;
    int64_t lo = i < j ? i : j;
    int64_t d = i < j ? j - i : i - j;
    if (d >= bw) return 0.0;
    return band[lo * bw + d];
}

#ifdef __CUDACC__
// Pixel (i, j) of the symmetric balanced matrix, NaN kept; 0 <= i, j < nbins required.
__device__ __forceinline__ double snn_mat_at(const MatView& m, int64_t i, int64_t j) {
    int64_t lo = i < j ? i : j, hi = i < j ? j : i;
    int64_t d = hi - lo;
    if (d < m.bw) return m.band[lo * m.bw + d];
    int64_t a = m.rowptr[lo], b = m.rowptr[lo + 1];
    double cnt = 0.0;
    while (a < b) {
        int64_t mid = (a + b) >> 1;
        int32_t c = m.bin2[mid];
        if (c == hi) { cnt = (double)m.count[mid]; break; }
        if (c < hi) a = mid + 1; else b = mid;
    }
    if (m.weight) cnt = cnt * (m.weight[lo] * m.weight[hi]);
    return cnt;
}
#endif

// Python slice(a, b) on an axis of length n -> [lo, hi) (hi <= lo means empty).
__host__ __device__ __forceinline__ void snn_pyslice(int64_t a, int64_t b, int64_t n, int64_t* lo,
                                                      int64_t* hi) {
    if (a < 0) { a += n; if (a < 0) a = 0; } else if (a > n) a = n;
    if (b < 0) { b += n; if (b < 0) b = 0; } else if (b > n) b = n;
    *lo = a;
    *hi = b;
}

// ---- per-TU entry points used by api.cu ---------------------------------------------------
int snn_band_build(snn_ctx* ctx, SnnChrom* c);
int snn_fband_build(snn_ctx* ctx, SnnChrom* c);
int snn_quantiles_impl(snn_ctx* ctx, SnnChrom* c, int nq, const double* q, double* out);
// One source of a quantile job: the positive pixels of chromosome c, whole (sel = false) or
// restricted like getQuantile_slow's genome-wide slices (quantile.cu: QChrom).
struct QSource {
    SnnChrom* c = nullptr;
    int job = 0;
    bool sel = false;
    int64_t s = 0, wrows = 1, F = 0, c0 = 0, c1 = 0;
};
int snn_quantiles_sources_impl(snn_ctx* ctx, const std::vector<QSource>& src, int njob, int nq, const double* q,
                               double* out);
int snn_quantiles_multi_impl(snn_ctx* ctx, const std::vector<SnnChrom*>& cs, int nq, const double* q, double* out);
int snn_expected_impl(snn_ctx* ctx, SnnChrom* c, double* out);
int snn_row_flags_impl(snn_ctx* ctx, SnnChrom* c, int nunits, const int64_t* geom, uint8_t* avail,
                       uint8_t* live);
int snn_background_impl(snn_ctx* ctx, SnnChrom* c, int64_t nsamp, const int32_t* spec, double* out);
int snn_set_background_impl(snn_ctx* ctx, int64_t ncol, const double* tables);
int snn_background_dev_impl(snn_ctx* ctx, SnnChrom* c, int64_t nsamp, const int32_t* spec, double* d_buf,
                            int64_t slot0);
int snn_set_background_dev_impl(snn_ctx* ctx, int64_t ncol, const int32_t* slots, const double* d_buf);
int snn_extract_impl(snn_ctx* ctx, SnnChrom* c, int nq, const double* mp, const double* gw, int radius,
                     int minH, int maxW, int bfilter, int gray_variant, int frame_lo, int frame_hi,
                     snn_candidate* out, int64_t cap, int64_t* n_out);
int snn_debug_stages_impl(snn_ctx* ctx, SnnChrom* c, double mp, const double* gw, int radius,
                          int bfilter, int gray_variant, int frame, int bright, int32_t* S, int32_t* nz,
                          float* gray, float* smoothed, float* mag, uint8_t* low, uint8_t* edges,
                          uint8_t* vert);
int snn_score_impl(snn_ctx* ctx, SnnChrom* c, int64_t n, const int64_t* pos, const int32_t* stale,
                   const double* expected,
                   double* pvalue, double* stripiness, double* oe_mean, double* oe_total,
                   double* o_mean, double* o_sum);
