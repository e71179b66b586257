/* stripenn_b200 — C ABI of the B200-native `stripenn compute` / `score` hot path.
 *
 * The reference (ysora/stripenn 1.1.65.22) has no FFI: its hot path is the Python class
 * `getStripe` (src/stripenn/getStripe.py:17-33) on top of `cooler.matrix(balance).fetch`.
 * Every entry point below replaces one method of that class (cited per function); the
 * Python host in stripenn_b200/ binds them with ctypes (see INTEGRATION.md for the stub a
 * maintainer of the reference would add).
 *
 * Conventions
 *  - every function returns 0 on success, a negative SNN_E* code on failure; the message is
 *    available from snn_last_error(ctx) (ctx may be NULL for errors of snn_create);
 *  - the caller owns every host buffer; the library owns device memory behind `snn_ctx`;
 *  - one context per GPU, used from one host thread at a time; all work of a context is
 *    issued on one CUDA stream (its own, or the one given to snn_set_stream) and every
 *    function that returns data to host memory synchronises that stream before returning;
 *  - there is NO CPU fallback: without a CUDA device snn_create fails with SNN_ENODEV.
 */
#ifndef STRIPENN_B200_H
#define STRIPENN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SNN_VERSION 100

#define SNN_OK 0
#define SNN_EINVAL (-1)   /* bad argument                                  */
#define SNN_ENODEV (-2)   /* no usable CUDA device / wrong architecture     */
#define SNN_ECUDA (-3)    /* CUDA runtime error (message has the details)   */
#define SNN_ENOMEM (-4)   /* device or host allocation failed               */
#define SNN_ECAP (-5)     /* caller's output buffer too small (n_out = need) */
#define SNN_ESTATE (-6)   /* call order violated (e.g. chromosome not loaded) */
#define SNN_ERANGE (-7)   /* a window needs matrix entries outside the band  */

#define SNN_NDIST 400     /* distances 0..399 (framesize of getStripe.py:794) */
#define SNN_NBRIGHT 6     /* np.arange(0.5, 1.01, 0.1) (getStripe.py:898)      */

/* cv.cvtColor float32 RGB2GRAY rounding variants */
#define SNN_GRAY_IPP 0    /* fmaf(B,cb, fl(fl(R*cr) + fl(G*cg)))  — OpenCV with IPP (default wheels) */
#define SNN_GRAY_FMA2 1   /* fmaf(B,cb, fmaf(G,cg, fl(R*cr)))     — universal intrinsics, FMA3        */
#define SNN_GRAY_PLAIN 2  /* fl(fl(fl(R*cr)+fl(G*cg)) + fl(B*cb)) — cv.setUseOptimized(False)          */

typedef struct snn_ctx snn_ctx;

/* One candidate stripe as emitted by StripeSearch (getStripe.py:1081-1110), before any
 * RemoveRedundant.  Records come back ordered exactly like the reference emits them:
 * maxpixel index, window id, brightness pass, direction (up then down), pair index. */
typedef struct snn_candidate {
    int32_t q;        /* index into the maxpixel list                          */
    int32_t frame;    /* window id `num` (getStripe.py:842)                    */
    int32_t bright;   /* brightness pass 0..5                                  */
    int32_t ud;       /* 1 = up, 2 = down (getStripe.py:937-940)               */
    int32_t k;        /* emission index inside (q, frame, bright, ud)          */
    int32_t x, y, w, h;   /* window-local, after zero-column compaction        */
    int32_t start, end;   /* first/last bin of the window (getStripe.py:794-801) */
    int32_t S;        /* compacted window size                                 */
    int64_t pos1, pos2, pos3, pos4;   /* bp, 1-based inclusive                 */
    double total;     /* submat[y:y+h, x:x+w].sum() (getStripe.py:1095)         */
} snn_candidate;

/* ---- life cycle ------------------------------------------------------------------- */
int snn_version(void);
int snn_create(int device, snn_ctx** out);
void snn_destroy(snn_ctx* ctx);
const char* snn_last_error(const snn_ctx* ctx);
/* Issue all work on an existing CUDA stream (cudaStream_t passed as void*); NULL restores
 * the context's own stream.  Lets a host that times with its own events see the kernels. */
int snn_set_stream(snn_ctx* ctx, void* cuda_stream);
int snn_synchronize(snn_ctx* ctx);

/* ---- a1: the data boundary — replaces cooler.Cooler.matrix(balance).fetch
 * (stripenn.py:118; call sites getStripe.py:172,193,266,326,432,518,560,684,690,696,808).
 * Upper-triangle COO pixels (bin1 <= bin2, chromosome-local bin ids) and the per-bin
 * MULTIPLICATIVE weight vector (NULL = raw counts; NaN = masked bin; a divisive column such
 * as KR is inverted by the host) are uploaded once and turned into a diagonal band
 * band[i][d] = count(i,i+d) * (w[i]*w[i+d]), d in [0, band_width), resident in HBM.
 * Entries with d >= band_width stay in the COO arrays (used by the quantiles and the row
 * flags) but are not addressable as windows (SNN_ERANGE). */
int snn_load_chrom(snn_ctx* ctx, int chrom_id, int64_t nbins, int64_t binsize,
                   int64_t chromsize_bp, int64_t nnz, const int32_t* bin1, const int32_t* bin2,
                   const int32_t* count, const double* weight, int band_width);
/* Asynchronous snn_load_chrom: the pixel arrays are copied on the context's own copy stream and
 * the call returns at once, so the upload of the next chromosomes overlaps the kernels of the
 * first ones (the band is built on the compute stream the first time the chromosome is used).
 * The caller must keep the four host arrays alive and unchanged until a later call that uses
 * this chromosome has returned data, or until snn_synchronize / snn_unload_chrom; pinned host
 * memory is what makes the copy truly asynchronous. */
int snn_load_chrom_async(snn_ctx* ctx, int chrom_id, int64_t nbins, int64_t binsize,
                         int64_t chromsize_bp, int64_t nnz, const int32_t* bin1, const int32_t* bin2,
                         const int32_t* count, const double* weight, int band_width);
int snn_unload_chrom(snn_ctx* ctx, int chrom_id);
/* Same, but the four arrays are DEVICE pointers already resident in HBM (zero-copy: the
 * library keeps using them, the caller must keep them alive until snn_unload_chrom). */
int snn_load_chrom_device(snn_ctx* ctx, int chrom_id, int64_t nbins, int64_t binsize,
                          int64_t chromsize_bp, int64_t nnz, const int32_t* d_bin1,
                          const int32_t* d_bin2, const int32_t* d_count, const double* d_weight,
                          int band_width);
/* Dense window rows [i0,i1) x cols [j0,j1) in bins (NaN kept) — debugging / tests only. */
int snn_fetch(snn_ctx* ctx, int chrom_id, int64_t i0, int64_t i1, int64_t j0, int64_t j1,
              double* out);

/* seeimage (src/stripenn/seeimage.py:77-90): saturated heat map of the square
 * window of bins [i0,i1): blue = 255*(M - A)/M, negatives -> 0, img = (1, blue/255, blue/255)
 * clipped to [0,1].  chan (may be NULL) receives the green/blue channel as float64, row-major
 * n x n, NaN kept; rgb (may be NULL) the n x n x 3 uint8 image as matplotlib's imshow quantises
 * it ((255*v) truncated, NaN -> 0).  Any window of the chromosome (far pixels read from the
 * pixel table). */
int snn_heatmap(snn_ctx* ctx, int chrom_id, int64_t i0, int64_t i1, double maxpixel_value,
                double* chan, unsigned char* rgb);

/* ---- a2: getStripe.getQuantile_original (getStripe.py:160-176):
 * np.quantile(mat[mat > 0], q) over the full symmetric matrix (off-diagonal pixels count
 * twice), numpy 'linear' interpolation reproduced bit-exactly.  Raw counts (no weights) are
 * selected from a one-pass histogram of the integer values, balanced values by an 8-bit radix
 * select over the float64 bit patterns (also the fallback for counts >= 4096); same results. */
int snn_quantiles(snn_ctx* ctx, int chrom_id, int nq, const double* q, double* out);
/* The same for several resident chromosomes in one call (getStripe.py:166 loops over ChrList):
 * out = nchrom x nq, row k = chrom_ids[k]. */
int snn_quantiles_multi(snn_ctx* ctx, int nchrom, const int32_t* chrom_ids, int nq, const double* q,
                        double* out);

/* -s / getStripe.getQuantile_slow (getStripe.py:107-158): np.quantile over a ROW/COLUMN-SELECTED
 * population.  The reference slices the matrix selector in genome-wide bin coordinates, block by
 * block: rows r_start + [k*w, min((k+1)*w - 1, floor(chromsize/r))) (the last row of every block is
 * lost), columns [r_start, r_end), with r_start = ceil(cumsize_before/r) + 1 — which skips the
 * first bin and, after several chromosomes, starts inside the PREVIOUS chromosome.  The host
 * (stripenn_b200.getstripe.getQuantile_slow) turns that into sources: source k adds to job[k] the
 * positive entries (row i, column j) — both triangles of the symmetric matrix — of resident
 * chromosome chrom_ids[k] with c0 <= j < c1 and a = i - s satisfying 0 <= a < F and
 * a % w != w - 1; sel = nsrc x 5 int64 {s, w, F, c0, c1} in chromosome-local bins (s may be
 * negative).  out = njob x nq (NaN for an empty population). */
int snn_quantiles_select(snn_ctx* ctx, int nsrc, const int32_t* chrom_ids, const int32_t* job,
                         const int64_t* sel, int njob, int nq, const double* q, double* out);

/* ---- a3: getStripe.mpmean (getStripe.py:178-235): expected value per diagonal. */
int snn_expected(snn_ctx* ctx, int chrom_id, double* out /* [SNN_NDIST] */);

/* ---- a4: getStripe.nulldist (getStripe.py:238-499).  The host owns the sampling plan (it
 * draws with Python's random.Random like the reference); the device answers the two
 * questions the plan needs and then evaluates the planned samples.
 * snn_row_flags: geom = nunits x 6 int64 {i0, i1, wi0, wi1, wj0, wj1}: rows [i0,i1) with a
 * non-zero sum inside their own square block [i0,i1) -> avail (available_cols, :249-273);
 * rows [wi0,wi1) with a non-zero sum over columns [wj0,wj1) -> live (pool, :328-335). */
int snn_row_flags(snn_ctx* ctx, int chrom_id, int nunits, const int64_t* geom,
                  uint8_t* avail /* [nbins] */, uint8_t* live /* [nbins] */);
/* spec = nsamp x 6 int32 {i0, i1, j0, j1, x, shift}: fetched window extent in bins, local
 * row, column offset (:358-360).  out = 4 x SNN_NDIST x nsamp doubles in the order
 * bgleft_up, bgright_up, bgleft_down, bgright_down (row-major [table][distance][sample]). */
int snn_background(snn_ctx* ctx, int chrom_id, int64_t nsamp, const int32_t* spec, double* out);
/* Install the genome-wide tables (all chromosomes concatenated by the host, same layout)
 * used by snn_score's p-values; they are sorted per distance on the device. */
int snn_set_background(snn_ctx* ctx, int64_t ncol, const double* tables);

/* Device-resident form of the two calls above, for the sharded pipeline (each GPU
 * computes the background columns of its own chromosomes, all GPUs need all columns before any
 * p-value): the columns never visit the host.
 * snn_background_dev writes sample k of the call to SAMPLE-MAJOR slot (slot0 + k) of a device
 * buffer: d_buf[(slot0 + k) * 4 * SNN_NDIST + table * SNN_NDIST + distance].  The buffer may be any
 * device allocation of this process (snn_device_alloc, or memory a collective library gathers
 * into: one ncclAllGather of the slot ranges completes it on every GPU).  The call returns after
 * the kernel has finished (the buffer can be handed to another stream).
 * snn_set_background_dev installs the tables from such a buffer: column t of every table is slot
 * slots[t] (host array of ncol int32; NULL = slots 0..ncol-1).  The p-values depend on the multiset
 * of each row only, so the order of the slots is free. */
int snn_background_dev(snn_ctx* ctx, int chrom_id, int64_t nsamp, const int32_t* spec, double* d_buf,
                       int64_t slot0);
int snn_set_background_dev(snn_ctx* ctx, int64_t ncol, const int32_t* slots, const double* d_buf);
/* Plain device memory for such buffers (and tests): allocate / release / copy to the host
 * (synchronises the context's stream first). */
int snn_device_alloc(snn_ctx* ctx, int64_t bytes, void** d_out);
int snn_device_free(snn_ctx* ctx, void* d_ptr);
int snn_device_read(snn_ctx* ctx, void* dst_host, const void* d_src, int64_t bytes);

/* ---- a5-a14: getStripe.extract/search_frame/StripeSearch (getStripe.py:790-1114) for the
 * windows [frame_lo, frame_hi) of one chromosome and ALL maxpixel values at once.
 * gauss_w = the 2*radius+1 weights of scipy's _gaussian_kernel1d(sigma) (computed by the
 * host with numpy so they are the bits the reference would use).  Returns candidates before
 * any RemoveRedundant; if more than `cap` were found returns SNN_ECAP with *n_out = need. */
int snn_extract(snn_ctx* ctx, int chrom_id, int nq, const double* maxpixel_values,
                const double* gauss_w, int radius, int minH, int maxW, int bfilter,
                int gray_variant, int frame_lo, int frame_hi, snn_candidate* out, int64_t cap,
                int64_t* n_out);
/* Per-pass intermediates of one window (tests): any pointer may be NULL.
 * gray/smoothed/mag: float[S*S]; low/edges/vert: uint8[S*S]; nz: int32[400]. */
int snn_debug_stages(snn_ctx* ctx, int chrom_id, double maxpixel_value, const double* gauss_w,
                     int radius, int bfilter, int gray_variant, int frame, int bright, int32_t* S,
                     int32_t* nz, float* gray, float* smoothed, float* mag, uint8_t* low,
                     uint8_t* edges, uint8_t* vert);

/* ---- a16-a18: getStripe.pvalue (:536-606), scoringstripes (:608-788), getMean (:501-534)
 * fused: one pass over each stripe's window.  pos = n x 4 int64 {pos1,pos2,pos3,pos4}.
 * expected = mpmean output of this chromosome (NULL: skip stripiness / O/E outputs).
 * stale = NULL or n x 2 int32 {table (0 down / 1 up), distance}: the background row a stripe
 * that is neither on (x1 == y1) nor under (x2 == y2) the diagonal inherits from the rows
 * processed before it (the reference's loop variables persist, getStripe.py:583-597); the
 * host derives it from the table order.  Without it such stripes get NaN.
 * Rows the reference would crash on (window outside the chromosome, shape mismatch at the
 * last bin) get NaN.  Any output pointer may be NULL.  p-values need snn_set_background. */
int snn_score(snn_ctx* ctx, int chrom_id, int64_t n, const int64_t* pos, const int32_t* stale,
              const double* expected,
              double* pvalue, double* stripiness, double* oe_mean, double* oe_total,
              double* o_mean, double* o_sum);

/* Same for stripes of several resident chromosomes in ONE call (the reference loops over the
 * whole table: getStripe.py:576 `for i in range(len(df))`, :764 `Parallel(...)(... for i in
 * range(nrow))`): chrom_ids = the nchrom chromosomes involved, expected = nchrom x SNN_NDIST
 * (row k = mpmean of chrom_ids[k]; NULL: p-values only), chrom_index[i] = index into chrom_ids
 * of stripe i (NULL: all stripes belong to chrom_ids[0]).  All chromosomes share one bin size. */
int snn_score_multi(snn_ctx* ctx, int nchrom, const int32_t* chrom_ids, const double* expected,
                    int64_t n, const int32_t* chrom_index, const int64_t* pos, const int32_t* stale,
                    double* pvalue, double* stripiness, double* oe_mean, double* oe_total,
                    double* o_mean, double* o_sum);

/* ---- --mask chr:start-end of scoringstripes (getStripe.py:612-639, 701-707): while a mask is set,
 * snn_score / snn_score_multi treat the O/E values of the masked bins [int(start/r), int(end/r)]
 * as NaN for every stripe of chromosome `chrom_id` with
 * start > min(pos1 - 50000, pos3) and start < max(pos2 + 50000, pos4) — columns of the centre,
 * left and right parts and rows of all three, with the reference's index arithmetic (the right
 * part is shifted by one bin; a masked bin equal to int(pos1/r), int(pos2/r) or int(pos4/r)
 * indexes one past the part and raises IndexError in the reference: such stripes get NaN
 * Stripiness and O/E outputs).  p-values and the observed mean/sum are not affected.
 * chrom_id < 0 clears the mask. */
int snn_set_mask(snn_ctx* ctx, int chrom_id, int64_t start_bp, int64_t end_bp);

/* ---- a15: getStripe.RemoveRedundant (:1116-1196) — host code (C++), no device work.
 * group = chromosome index per row, num = window id, pos = n x 4, h/w = the size columns,
 * key = pvalue or Stripiness (unused for by = 0).  by: 0 size, 1 pvalue, 2 score.
 * keep[i] = 1 for surviving rows (original order is kept by the caller). */
int snn_remove_redundant(int64_t n, const int32_t* group, const int64_t* num, const int64_t* pos,
                         const double* h, const double* w, const double* key, int by,
                         uint8_t* keep);

/* ---- measurement: accumulated CUDA-event time per kernel family since the last reset.
 * names[i] is a static string; returns the number of stages written (<= cap). */
int snn_stage_times(snn_ctx* ctx, int cap, const char** names, double* ms, int64_t* launches);
int snn_reset_stage_times(snn_ctx* ctx);
int snn_enable_stage_timing(snn_ctx* ctx, int on);
/* How the stripes of the context's scoring calls were routed since the last reset (reset = 1 clears):
 * out[0] fast kernel (regular stripes, score_fast.cu), out[1] warp kernel, out[2] CTA kernel. */
int snn_score_routing(snn_ctx* ctx, int64_t out[3], int reset);

#ifdef __cplusplus
}
#endif
#endif
