// K9: fused scoring of candidate stripes — getStripe.pvalue (getStripe.py:536-606),
// getStripe.scoringstripes / iterate_idx (:608-788, stats.elementwise_product_sum
// stats.py:184-199) and getStripe.getMean (:501-534) in one pass over each stripe's window.
//
// One CTA (4 warps) per stripe; a warp owns a row of the window and its lanes run along the
// columns (contiguous in the band for rows above the diagonal), so every window byte is
// requested once; per-row results live in shared memory, the background rows are pre-sorted
// (background.cu) so #(bg >= v) is a binary search in L2-resident tables.
//
// Roofline: HBM.  Algorithmic bytes per stripe: 8 * h * (w + 2*bs) window bytes + 48 B out;
// the 12.8 MB background tables are read once per launch (L2-resident afterwards).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "score_warp.cuh"

#define SCORE_THREADS 128
#define SCORE_MAXCOLS 512

struct ScoreParams {
    int64_t binsize, chromsize, nbins;
    int bs;
    int hmax;  // shared-memory capacity in rows
    int64_t bg_ncol;
    // --mask (getStripe.py:612-639): mask_on for this chromosome; start in bp; masked bins [mxs, mxe]
    int mask_on;
    int64_t mask_start, mxs, mxe;
};

__device__ __forceinline__ double qnan() { return __longlong_as_double(0x7ff8000000000000LL); }

// number of entries >= v among the first n (sorted ascending) entries of row
__device__ __forceinline__ int count_ge(const double* __restrict__ row, int n, double v) {
    if (v != v) return 0;
    int lo = 0, hi = n;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (row[mid] < v) lo = mid + 1; else hi = mid;
    }
    return n - lo;
}

// k-th smallest (0-based) of a[0..n) by rank counting; all threads of the CTA call it and get the
// same answer through `slot` (shared).
__device__ double select_kth(const double* a, int n, int k, double* slot) {
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double v = a[i];
        int less = 0, leq = 0;
        for (int j = 0; j < n; ++j) {
            const double u = a[j];
            less += (u < v);
            leq += (u <= v);
        }
        if (less <= k && k < leq) *slot = v;  // every writer writes the same value
    }
    __syncthreads();
    return *slot;
}

// Window accessor: the stripe's window staged in shared memory (the common case) or, for windows
// that do not fit, read through the band / CSR view.
struct Win {
    const double* sm;   // staged window, row-major [h][ncols], or nullptr
    MatView m;
    int64_t R0, cmin;
    int ncols;
    __device__ __forceinline__ double at(int j, int64_t gc) const {
        return sm ? sm[j * ncols + (int)(gc - cmin)] : snn_mat_at(m, R0 + j, gc);
    }
};

__global__ void __launch_bounds__(SCORE_THREADS) k_score(MatView m, const int64_t* __restrict__ pos,
                                                         const int32_t* __restrict__ stale,
                                                         const int32_t* __restrict__ order,
                                                         const double* __restrict__ rexp,
                                                         const double* __restrict__ bg,
                                                         const int32_t* __restrict__ bg_valid, ScoreParams P,
                                                         int wcap, int want_p, int want_s,
                                                         double* __restrict__ out) {
    extern __shared__ double sh[];
    double* cm = sh;               // centre row means (O/E)         [hmax]
    double* lm = cm + P.hmax;      // left                            [hmax]
    double* rm = lm + P.hmax;      // right                           [hmax]
    double* pv = rm + P.hmax;      // per-row p-values / scratch      [hmax]
    double* win = pv + P.hmax;     // staged window                   [wcap]
    int* ridx = (int*)(win + wcap);                          // compacted row index [hmax]
    unsigned char* rdel = (unsigned char*)(ridx + P.hmax);   // row deleted          [hmax]
    __shared__ double rex[SNN_NDIST];   // 1 / (expected[d] + 1e-8)
    __shared__ unsigned char dead[SCORE_MAXCOLS];
    __shared__ unsigned char cmask[SCORE_MAXCOLS];   // --mask: column of its part is masked
    __shared__ double slot;
    __shared__ double red[8];
    __shared__ int redi[4];
    __shared__ int bad_flag;

    const int64_t s = order ? order[blockIdx.x] : blockIdx.x;
    const int64_t p1 = pos[s * 4 + 0], p2 = pos[s * 4 + 1], p3 = pos[s * 4 + 2], p4 = pos[s * 4 + 3];
    const int64_t r = P.binsize;
    const int bs = P.bs;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = SCORE_THREADS / 32;
    double* o = out + s * 6;
    if (threadIdx.x == 0) {
        bad_flag = 0;
        for (int t = 0; t < 6; ++t) o[t] = qnan();
    }
    if (want_s)
        for (int t = threadIdx.x; t < SNN_NDIST; t += SCORE_THREADS) rex[t] = rexp[t];
    __syncthreads();
    // cooler region checks (fake/real cooler raise; we answer NaN)
    if (p1 < 1 || p3 < 1 || p2 < p1 || p4 < p3 || p2 > P.chromsize || p4 > P.chromsize) return;
    // rows of the window: bins [p3 // r, ceil(p4 / r))
    const int64_t R0 = p3 / r, R1 = (p4 + r - 1) / r;
    const int h = (int)(R1 - R0);
    if (h <= 0 || h > P.hmax) return;

    // column ranges of both consumers
    int64_t lmost = p1 - (int64_t)bs * r;
    if (lmost < 1) lmost = 1;
    int64_t rmost = p2 + (int64_t)bs * r;
    if (rmost > P.chromsize) rmost = P.chromsize;
    const int64_t C0 = lmost / r, C1 = (rmost + r - 1) / r;       // p-value window
    const int64_t xsi = p1 / r, xei = p2 / r, ysi = p3 / r, yei = p4 / r;
    int64_t lmb = xsi - bs, rmb = xei + bs;
    if (lmb < 1) lmb = 1;
    if (rmb >= P.nbins) rmb = P.nbins - 1;
    const int64_t cC0 = p1 / r, cC1 = (p2 + r - 1) / r;           // centre: "chr:xs-xe"
    int64_t cmin = C0, cmax = C1;
    if (want_s) {
        if (lmb < cmin) cmin = lmb;
        if (cC0 < cmin) cmin = cC0;
        if (rmb > cmax) cmax = rmb;
        if (cC1 > cmax) cmax = cC1;
    }
    if (!want_p) { cmin = lmb < cC0 ? lmb : cC0; cmax = rmb > cC1 ? rmb : cC1; }
    if (cmin < 0) cmin = 0;
    if (cmax > P.nbins) cmax = P.nbins;
    const int64_t ncw = cmax - cmin;
    Win W;
    W.m = m; W.R0 = R0; W.cmin = cmin; W.ncols = (int)ncw;
    W.sm = (ncw > 0 && (int64_t)h * ncw <= wcap) ? win : nullptr;
    if (W.sm) {
        // upper triangle part (col >= row): a warp per window row, lanes along the band row
        for (int j = warp; j < h; j += nwarp) {
            const int64_t row = R0 + j;
            int64_t c0 = cmin > row ? cmin : row;
            for (int64_t c = c0 + lane; c < cmax; c += 32) win[j * (int)ncw + (int)(c - cmin)] = snn_mat_at(m, row, c);
        }
        // lower triangle part (row > col): a warp per window column, lanes along the band row of that column
        for (int64_t c = cmin + warp; c < cmax; c += nwarp) {
            int64_t i0 = R0 > c + 1 ? R0 : c + 1;
            for (int64_t i = i0 + lane; i < R1; i += 32) win[(int)(i - R0) * (int)ncw + (int)(c - cmin)] = snn_mat_at(m, i, c);
        }
        __syncthreads();
    }

    // ================================ p-value (getStripe.py:552-605) ================================
    if (want_p) {
        const int64_t ncol = C1 - C0;
        int64_t ca, cb, la, lb, ra, rb;
        snn_pyslice(bs, -(int64_t)bs, ncol, &ca, &cb);   // mat[:, bs:-bs]
        snn_pyslice(0, bs, ncol, &la, &lb);              // mat[:, :bs]
        snn_pyslice(-(int64_t)bs, ncol, ncol, &ra, &rb); // mat[:, -bs:]
        if (bs == 0) { ra = 0; rb = ncol; }
        const int64_t x1 = (p1 - 1) / r, x2 = p2 / r, y1 = (p3 - 1) / r, y2 = p4 / r;
        int mode;  // 0: down tables, row j ; 1: up tables, row y2-y1-j-1 ; 2: stale ; 3: undefined
        if (x1 == y1) mode = 0;
        else if (x2 == y2) mode = 1;
        else if (stale && stale[s * 2] >= 0) mode = 2;
        else mode = 3;
        for (int j = threadIdx.x; j < h; j += SCORE_THREADS) {
            double sc = 0.0, sl = 0.0, sr = 0.0;
            for (int64_t c = 0; c < ncol; ++c) {
                const double v = snn_nan0(W.at(j, C0 + c));
                if (c >= ca && c < cb) sc += v;
                if (c >= la && c < lb) sl += v;
                if (c >= ra && c < rb) sr += v;
            }
            const double mc = (cb > ca) ? sc / (double)(cb - ca) : qnan();
            const double ml = (lb > la) ? sl / (double)(lb - la) : qnan();
            const double mr = (rb > ra) ? sr / (double)(rb - ra) : qnan();
            const double dl = mc - ml, dr = mc - mr;
            double p = qnan();
            if (mode != 3) {
                int table, d;
                if (mode == 0) { table = 1; d = j < 399 ? j : 399; }
                else if (mode == 1) {
                    int64_t dd = y2 - y1 - j - 1;
                    if (dd > 399) dd = 399;
                    if (dd < 0) dd += SNN_NDIST;   // numpy negative index
                    d = (int)dd;
                    table = 0;
                } else { table = stale[s * 2] ? 0 : 1; d = stale[s * 2 + 1]; }
                if (d >= 0 && d < SNN_NDIST) {
                    // table order in memory: 0 bgleft_up, 1 bgright_up, 2 bgleft_down, 3 bgright_down
                    const int tl = table == 0 ? 0 : 2, tr = table == 0 ? 1 : 3;
                    const int64_t rl = (int64_t)tl * SNN_NDIST + d, rr = (int64_t)tr * SNN_NDIST + d;
                    const int nl = bg_valid[rl], nr = bg_valid[rr];
                    const double a = (double)count_ge(bg + rl * P.bg_ncol, nl, dl) / (double)nl;
                    const double b = (double)count_ge(bg + rr * P.bg_ncol, nr, dr) / (double)nr;
                    p = a;                // Python max(a, b)
                    if (b > p) p = b;
                    if (p == 0.0) p = 1.0 / (double)P.bg_ncol;
                }
            }
            pv[j] = p;
        }
        __syncthreads();
        // np.median: NaN if any NaN, else mean of the middle pair
        int anynan = 0;
        for (int j = threadIdx.x; j < h; j += SCORE_THREADS) anynan |= (pv[j] != pv[j]);
        anynan = __syncthreads_or(anynan);
        double med;
        if (anynan) med = qnan();
        else if (h & 1) med = select_kth(pv, h, h / 2, &slot);
        else {
            const double a = select_kth(pv, h, h / 2 - 1, &slot);
            const double b = select_kth(pv, h, h / 2, &slot);
            med = (a + b) / 2.0;
        }
        if (threadIdx.x == 0) o[0] = med;
        __syncthreads();
    }
    if (!want_s) return;

    // ============================ stripiness (getStripe.py:661-759) ================================
    const int64_t lC0 = lmb, lC1 = xsi;                   // left:   "chr:lm*r - xsi*r"
    const int64_t rC0 = xei, rC1 = rmb;                   // right:  "chr:xei*r - rm*r"
    const int wc = (int)(cC1 - cC0), wl = (int)(lC1 - lC0), wr = (int)(rC1 - rC0);
    // ---- getMean (getStripe.py:501-534): nanmean / nansum of fetch(rows pos3..pos4, cols pos1..pos2).
    // A function of its own in the reference: it does not depend on the flank geometry below
    // (a stripe in the first bins has no valid left flank, but it has a mean).
    {
        double tot_obs = 0.0;
        int n_obs = 0;
        for (int j = threadIdx.x; j < h; j += SCORE_THREADS) {
            for (int c = 0; c < wc; ++c) {
                const double obs = W.at(j, cC0 + c);
                if (obs == obs) { tot_obs += obs; ++n_obs; }
            }
        }
        for (int off = 16; off > 0; off >>= 1) {
            tot_obs += __shfl_xor_sync(0xffffffffu, tot_obs, off);
            n_obs += __shfl_xor_sync(0xffffffffu, n_obs, off);
        }
        if (lane == 0) { red[4 + warp] = tot_obs; redi[warp] = n_obs; }
        __syncthreads();
        if (threadIdx.x == 0) {
            double b = 0.0;
            int n = 0;
            for (int w = 0; w < nwarp; ++w) { b += red[4 + w]; n += redi[w]; }
            o[4] = n ? b / (double)n : qnan();
            o[5] = b;
        }
        __syncthreads();
    }
    // shapes must agree with expecMatrix (else numpy raises in the reference) and the regions
    // must be valid for cooler (end >= start)
    if (wc != (int)(xei - xsi) || h != (int)(yei - ysi) || wc <= 0 || wl < 0 || wr < 0 ||
        wc + wl + wr > SCORE_MAXCOLS || R0 != ysi)
        return;
    const int ncols = wc + wl + wr;  // column order: centre | left | right
    // ---- --mask (masking(), getStripe.py:620-639, applied at :701-707): a masked bin b in [mxs, mxe]
    // hits index b - start of a part iff 0 <= b - start < end - start + 1; for the centre columns
    // (start xsi, end xei), the rows (ysi, yei) and the left columns (lm, xsi) that range is one
    // longer than the part, so b == xei / yei / xsi raises IndexError in the reference -> NaN here.
    // The right part is called with start xei + 1, end rm: its column x is masked iff xei + 1 + x is.
    bool masked = false, mask_bad = false;
    if (P.mask_on) {
        const int64_t lo = (p1 - 50000 < p3) ? p1 - 50000 : p3, hi = (p2 + 50000 > p4) ? p2 + 50000 : p4;
        masked = P.mask_start > lo && P.mask_start < hi;
    }
    if (masked) {
        const bool in_xei = P.mxs <= xei && xei <= P.mxe;
        const bool in_yei = P.mxs <= yei && yei <= P.mxe;
        const bool in_xsi = P.mxs <= xsi && xsi <= P.mxe;
        // centre cols (h > 0), centre rows (wc > 0), left cols (h > 0): the reference raises
        mask_bad = in_xei || in_yei || in_xsi;
        if (mask_bad) masked = false;    // carry on unmasked for the observed mean/sum, drop the O/E outputs
    }
    if (masked) {
        for (int c = threadIdx.x; c < ncols; c += SCORE_THREADS) {
            int64_t b = c < wc ? xsi + c : (c < wc + wl ? lmb + (c - wc) : xei + 1 + (c - wc - wl));
            cmask[c] = (b >= P.mxs && b <= P.mxe) ? 1 : 0;
        }
    }
    for (int j = threadIdx.x; j < h; j += SCORE_THREADS)
        rdel[j] = (masked && ysi + j >= P.mxs && ysi + j <= P.mxe) ? 2 : 0;   // bit 1 = masked row, bit 0 = deleted
    __syncthreads();
    // ---- pass 1: all-NaN columns of the O/E parts (division keeps NaN of the observed value)
    for (int c = threadIdx.x; c < ncols; c += SCORE_THREADS) {
        int64_t gc = c < wc ? cC0 + c : (c < wc + wl ? lC0 + (c - wc) : rC0 + (c - wc - wl));
        bool all = true;
        if (!(masked && cmask[c])) {
            for (int j = 0; j < h && all; ++j) {
                const double v = W.at(j, gc);
                if (v == v && rdel[j] == 0) all = false;  // observed finite or inf -> O/E not NaN (expected + 1e-8 is finite)
            }
        }
        dead[c] = all ? 1 : 0;
    }
    __syncthreads();
    // rows deleted together with dead columns: index x (column index inside its part) when
    // xs == ys, else shape[0]-1-x  (getStripe.py:713-731); np.delete wraps negatives
    if (threadIdx.x == 0) {
        const bool diag = (p1 == p3);
        bool any = false;
        for (int c = 0; c < ncols; ++c) any |= dead[c] != 0;
        if (any) {
            for (int part = 0; part < 3; ++part) {
                const int w0 = part == 0 ? 0 : (part == 1 ? wc : wc + wl);
                const int wn = part == 0 ? wc : (part == 1 ? wl : wr);
                for (int x = 0; x < wn; ++x) {
                    if (!dead[w0 + x]) continue;
                    int idx = diag ? x : h - 1 - x;
                    if (idx < 0) idx += h;
                    if (idx < 0 || idx >= h) { bad_flag = 1; continue; }  // IndexError in the reference
                    rdel[idx] |= 1;
                }
            }
        }
        redi[1] = any ? 1 : 0;
    }
    __syncthreads();
    if (bad_flag) return;
    if (redi[1]) {
        if (threadIdx.x == 0) {
            int k = 0;
            for (int j = 0; j < h; ++j)
                if (!(rdel[j] & 1)) ridx[k++] = j;
            redi[0] = k;
        }
    } else {
        for (int j = threadIdx.x; j < h; j += SCORE_THREADS) ridx[j] = j;
        if (threadIdx.x == 0) redi[0] = h;
    }
    __syncthreads();
    const int hk = redi[0];
    int lc = 0, ll = 0, lr = 0;  // live columns per part
    for (int c = 0; c < ncols; ++c) {
        if (dead[c]) continue;
        if (c < wc) ++lc; else if (c < wc + wl) ++ll; else ++lr;
    }
    // ---- pass 2: row means of nan0(O/E) over the live columns (thread per row)
    double tot_oe = 0.0;
    for (int k = threadIdx.x; k < hk; k += SCORE_THREADS) {
        const int j = ridx[k];
        const int64_t row = R0 + j;
        double sc = 0.0, sl = 0.0, sr = 0.0;
        const bool rowm = masked && (rdel[j] & 2);
        for (int c = 0; c < ncols; ++c) {
            if (dead[c]) continue;
            if (rowm || (masked && cmask[c])) continue;     // masked O/E value: NaN -> 0
            int64_t gc, ecol;
            if (c < wc) { gc = cC0 + c; ecol = xsi + c; }
            else if (c < wc + wl) { gc = lC0 + (c - wc); ecol = lmb + (c - wc); }
            else { gc = rC0 + (c - wc - wl); ecol = xei + 1 + (c - wc - wl); }  // right: shifted by one
            int64_t dd = ecol - row;
            if (dd < 0) dd = -dd;
            if (dd > 399) dd = 399;
            const double v = snn_nan0(W.at(j, gc) * rex[dd]);
            if (c < wc) sc += v; else if (c < wc + wl) sl += v; else sr += v;
        }
        cm[k] = lc ? sc / (double)lc : qnan();
        lm[k] = ll ? sl / (double)ll : qnan();
        rm[k] = lr ? sr / (double)lr : qnan();
        tot_oe += sc;
    }
    for (int off = 16; off > 0; off >>= 1) tot_oe += __shfl_xor_sync(0xffffffffu, tot_oe, off);
    if (lane == 0) red[warp] = tot_oe;
    __syncthreads();
    if (threadIdx.x == 0 && !mask_bad) {
        double a = 0.0;
        for (int w = 0; w < nwarp; ++w) a += red[w];
        // centerMean = mean(center[rows_rep]) = sum / (h' * w') ; centerTotal = w' * sum  (quirk :747)
        o[2] = (hk > 0 && lc > 0) ? a / ((double)hk * (double)lc) : qnan();
        o[3] = (double)lc * a;
    }
    __syncthreads();
    // ---- pass 3: Sobel-like sums (stats.elementwise_product_sum) and the score
    double acc = 0.0;
    int nacc = 0;
    for (int i = 1 + threadIdx.x; i + 1 < hk; i += SCORE_THREADS) {
        const double l0 = lm[i - 1], l1 = lm[i], l2 = lm[i + 1];
        const double c0 = cm[i - 1], c1 = cm[i], c2 = cm[i + 1];
        const double r0 = rm[i - 1], r1 = rm[i], r2 = rm[i + 1];
        const double gxl = ((-1.0 * l0 + -2.0 * l1) + -1.0 * l2) + ((1.0 * c0 + 2.0 * c1) + 1.0 * c2);
        const double gxr = ((1.0 * c0 + 2.0 * c1) + 1.0 * c2) + ((-1.0 * r0 + -2.0 * r1) + -1.0 * r2);
        double gy = (((1.0 * l0 + 0.0 * l1) + -1.0 * l2) + ((2.0 * c0 + 0.0 * c1) + -2.0 * c2)) +
                    ((1.0 * r0 + 0.0 * r1) + -1.0 * r2);
        if (gy < 0) gy *= -1.0;
        const double gx = (gxl != gxl || gxr != gxr) ? qnan() : fmin(gxl, gxr);  // np.minimum propagates NaN
        const double d = gx - gy;
        if (d == d) { acc += d; ++nacc; }
    }
    for (int off = 16; off > 0; off >>= 1) {
        acc += __shfl_xor_sync(0xffffffffu, acc, off);
        nacc += __shfl_xor_sync(0xffffffffu, nacc, off);
    }
    __syncthreads();
    if (lane == 0) { red[warp] = acc; redi[warp] = nacc; }
    __syncthreads();
    double avgdiff;
    {
        double a = 0.0;
        int n = 0;
        for (int w = 0; w < nwarp; ++w) { a += red[w]; n += redi[w]; }
        avgdiff = n ? a / (double)n : qnan();
    }
    // nanmedian(centerm): compact the non-NaN values into pv
    __syncthreads();
    if (threadIdx.x == 0) {
        int n = 0;
        for (int k = 0; k < hk; ++k)
            if (cm[k] == cm[k]) pv[n++] = cm[k];
        redi[0] = n;
    }
    __syncthreads();
    const int nm = redi[0];
    double med;
    if (nm == 0) med = qnan();
    else if (nm & 1) med = select_kth(pv, nm, nm / 2, &slot);
    else {
        const double a = select_kth(pv, nm, nm / 2 - 1, &slot);
        const double b = select_kth(pv, nm, nm / 2, &slot);
        med = (a + b) / 2.0;
    }
    if (threadIdx.x == 0 && !mask_bad) o[1] = med * avgdiff;
}

// slot y (SF_EXSTRIDE doubles): e[t] = expected -> e[512 + t] = 1 / (expected + 1e-8)   (getStripe.py:650)
// and the mirrored table of the fast kernel, e[2048 + SF_RM + d] = 1 / (expected[min(|d|, 399)] + 1e-8)
__global__ void k_recip_expected(double* __restrict__ ex) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    double* e = ex + (size_t)blockIdx.y * SF_EXSTRIDE;
    if (t < SNN_NDIST) e[512 + t] = 1.0 / (e[t] + .00000001);
    if (t < 2 * SF_RM) {
        int d = t - SF_RM;
        if (d < 0) d = -d;
        if (d > SNN_NDIST - 1) d = SNN_NDIST - 1;
        e[2048 + t] = 1.0 / (e[d] + .00000001);
    }
}

// ---- host side ------------------------------------------------------------------------------------------
// One call scores stripes of any number of resident chromosomes (`slots`): the warp kernel takes
// every stripe whose window is in-band and small (conservative tests on the bp coordinates, no
// divisions), the CTA kernel above takes the rest per chromosome.
static inline size_t up256(size_t v) { return (v + 255) & ~(size_t)255; }

int snn_score_multi_impl(snn_ctx* ctx, const std::vector<ScoreSlot>& slots, int64_t n, const int32_t* slot_of,
                         const int64_t* pos, const int32_t* stale, double* pvalue, double* stripiness,
                         double* oe_mean, double* oe_total, double* o_mean, double* o_sum) {
    const int want_p = pvalue != nullptr;
    bool any_exp = false;
    for (const ScoreSlot& sl : slots) any_exp |= sl.expected != nullptr;
    const int want_s = any_exp && (stripiness || oe_mean || oe_total || o_mean || o_sum);
    if ((!want_p && !want_s) || n <= 0) return SNN_OK;
    const int nslot = (int)slots.size();
    const int64_t r = slots[0].c->binsize;
    for (const ScoreSlot& sl : slots)
        if (sl.c->binsize != r || (want_s && !sl.expected))
            return snn_fail(ctx, SNN_EINVAL, "snn_score: chromosomes of one call share bin size and expected tables");
    const int bs = (int)(50000.0 / (double)r);

    // ---- classification.  Warp classes by height {64, 128, SW_HMAX}; CTA classes as before.
    static const int kWarpH[3] = {64, 128, SW_HMAX};
    struct Cls { int hcap, wcap; };
    static const Cls cls[3] = {{128, 4096}, {512, 12288}, {4800, 0}};
    std::vector<int32_t> worder[3];
    std::vector<std::vector<int32_t>> gorder((size_t)nslot * 3);
    // fast kernel (score_fast.cu): the regular stripes.  SNN_SCORE_FAST (read per call: the tests
    // compare the kernels) = 0 off, 1 always, unset = when it pays: the kernel needs the column-major
    // full band of every chromosome it touches, and building one moves ~8 B * nbins * 2 bw + 64 B per
    // stored pixel through HBM — about what the warp kernel loses on 1 stripe per 150 KB.  A `score`
    // call over 100k stripes builds the bands (and later calls find them resident); the few hundred
    // stripes per chromosome of a `compute` step stay with the warp kernel.
    std::vector<FastRec> frec;
    const char* fenv = getenv("SNN_SCORE_FAST");
    bool fast_on = !(fenv && fenv[0] == '0') && !ctx->force_generic && bs >= 1 && bs <= 64 &&
                   (!want_p || ctx->bg_ncol >= 1);
    if (fast_on && !(fenv && fenv[0] == '1')) {
        std::vector<int64_t> cand((size_t)nslot, 0);
        for (int64_t i = 0; i < n; ++i) {
            const int sl = slot_of ? slot_of[i] : 0;
            if (sl >= 0 && sl < nslot && pos[i * 4 + 1] % r == 0 && pos[i * 4 + 3] % r == 0) ++cand[sl];
        }
        double build_bytes = 0.0, ncand = 0.0;
        for (int sl = 0; sl < nslot; ++sl) {
            const SnnChrom* c = slots[sl].c;
            if (!cand[sl]) continue;
            ncand += (double)cand[sl];
            if (c->fband_state == 0) build_bytes += 16.0 * (double)c->nbins * (double)c->bw + 64.0 * (double)c->nnz;
        }
        fast_on = ncand * 150000.0 >= build_bytes;
    }
    std::vector<char> slot_fast((size_t)nslot, 1);
    for (int sl = 0; sl < nslot; ++sl) {
        const SnnChrom* c = slots[sl].c;
        bool ok = fast_on && c->nbins < (1ll << 30) && c->bw < SF_RM;
        if (ok && want_s)      // the mirrored table must be finite (no nan0 around the O/E products)
            for (int d = 0; d < SNN_NDIST && ok; ++d) {
                const double x = 1.0 / (slots[sl].expected[d] + .00000001);
                ok = x == x && x - x == 0.0;
            }
        slot_fast[sl] = ok;
    }
    const int64_t wmaxbp = (int64_t)std::min(62, SW_MAXCOLS - 2 - 2 * bs) * r;
    for (int64_t i = 0; i < n; ++i) {
        const int sl = slot_of ? slot_of[i] : 0;
        if (sl < 0 || sl >= nslot) return snn_fail(ctx, SNN_EINVAL, "snn_score: chromosome slot out of range");
        const SnnChrom* c = slots[sl].c;
        const int64_t p1 = pos[i * 4], p2 = pos[i * 4 + 1], p3 = pos[i * 4 + 2], p4 = pos[i * 4 + 3];
        const bool valid = p1 >= 1 && p3 >= 1 && p2 >= p1 && p4 >= p3 && p2 <= c->chromsize && p4 <= c->chromsize;
        // --mask: stripes the mask may touch take the CTA kernel (it holds the masking logic)
        const bool mask_hit = want_s && ctx->mask_chrom >= 0 && c->id == ctx->mask_chrom &&
                              ctx->mask_start > std::min(p1 - 50000, p3) && ctx->mask_start < std::max(p2 + 50000, p4);
        int wk = -1;
        if (valid && !mask_hit && !ctx->force_generic && bs <= 64 && wmaxbp >= 0 && (p2 - p1) <= wmaxbp) {
            const int64_t reach = (int64_t)(c->bw - 1 - bs) * r;   // |row - col| < bw for the whole window
            if ((p2 - p3) < reach && (p4 - p1) < reach) {
                for (int k = 0; k < 3 && wk < 0; ++k)
                    if ((p4 - p3) <= (int64_t)(kWarpH[k] - 2) * r) wk = k;
            }
        }
        if (wk >= 0 && slot_fast[sl] && p2 % r == 0 && p4 % r == 0) {
            if (!slots[sl].c->fband_state)          // the column-major full band, built on first use
                if (int rcf = snn_fband_build(ctx, slots[sl].c)) return rcf;
            if (slots[sl].c->fband_state != 1) slot_fast[sl] = 0;
        }
        if (wk >= 0 && slot_fast[sl] && p2 % r == 0 && p4 % r == 0) {
            // regular: the p-value slices and the stripiness parts are the same three column ranges
            // and every shape check of scoringstripes holds (the tests of k_score_warp, evaluated here)
            const int64_t xsi = p1 / r, xei = p2 / r, R0 = p3 / r, R1 = p4 / r;
            const int64_t h = R1 - R0, wcb = xei - xsi;
            int64_t lmost = p1 - (int64_t)bs * r, rmost = p2 + (int64_t)bs * r;
            if (lmost < 1) lmost = 1;
            if (rmost > c->chromsize) rmost = c->chromsize;
            const int64_t C0 = lmost / r, C1 = (rmost + r - 1) / r;
            if (h >= 1 && h <= SF_HMAX && wcb >= 1 && wcb <= 62 && xsi - bs >= 1 && xei + bs <= c->nbins - 1 &&
                C0 == xsi - bs && C1 == xei + bs) {
                FastRec fr;
                fr.xsi = (int32_t)xsi;
                fr.wc = (int32_t)wcb;
                fr.R0 = (int32_t)R0;
                fr.h = (int32_t)h;
                const int64_t x1 = (p1 - 1) / r, y1 = (p3 - 1) / r;
                int mode = 3, aux = 0;
                if (x1 == y1) mode = 0;
                else if (xei == R1) { mode = 1; aux = (int)(R1 - y1); }
                else if (stale && stale[i * 2] >= 0) { mode = 2 | ((stale[i * 2] ? 0 : 1) << 4); aux = stale[i * 2 + 1]; }
                fr.mode = mode | ((p1 == p3) ? 0x100 : 0);
                fr.aux = aux;
                fr.slot = sl;
                fr.idx = (int32_t)i;
                frec.push_back(fr);
                continue;
            }
        }
        if (wk >= 0) { worder[wk].push_back((int32_t)i); continue; }
        int64_t h = 1, w = 1;
        if (p3 >= 1 && p4 >= p3 && p4 <= c->chromsize) h = (p4 + r - 1) / r - p3 / r;
        if (p1 >= 1 && p2 >= p1) w = (p2 + r - 1) / r - p1 / r + 2 * bs + 2;
        int k = 2;
        if (h <= cls[0].hcap && h * w <= cls[0].wcap) k = 0;
        else if (h <= cls[1].hcap && h * w <= cls[1].wcap) k = 1;
        gorder[(size_t)sl * 3 + k].push_back((int32_t)i);
    }

    ctx->score_routed[0] += (int64_t)frec.size();
    for (int k = 0; k < 3; ++k) ctx->score_routed[1] += (int64_t)worder[k].size();
    for (const auto& g : gorder) ctx->score_routed[2] += (int64_t)g.size();

    // ---- device scratch
    const size_t pbytes = up256((size_t)n * 4 * sizeof(int64_t));
    const size_t sbytes = stale ? up256((size_t)n * 2 * sizeof(int32_t)) : 0;
    const size_t ebytes = up256((size_t)nslot * SF_EXSTRIDE * sizeof(double));   // expected | 1/(expected+1e-8) | mirrored
    const size_t cbytes = up256((size_t)nslot * sizeof(ChromDev));
    const size_t orbytes = up256((size_t)n * sizeof(int32_t));
    const size_t slbytes = slot_of ? up256((size_t)n * sizeof(int32_t)) : 0;
    const size_t obytes = up256((size_t)n * 6 * sizeof(double));
    const size_t nfast = frec.size();
    // tallest first: the warps of the fast kernel take the stripes from an atomic counter
    if (nfast) {                                   // counting sort by height (stable; std::stable_sort took 5 ms per 100k)
        std::vector<int32_t> start(SF_HMAX + 2, 0);
        for (const FastRec& f : frec) ++start[SF_HMAX - f.h + 1];
        for (int k = 1; k <= SF_HMAX + 1; ++k) start[k] += start[k - 1];
        std::vector<FastRec> sorted(nfast);
        for (const FastRec& f : frec) sorted[start[SF_HMAX - f.h]++] = f;
        frec.swap(sorted);
    }
    const int fgrid = snn_score_fast_grid((int)nfast);
    const bool any_tall = nfast && frec[0].h > SF_HSM;
    const size_t fbytes = up256(nfast * sizeof(FastRec)) + 256 +
                          (any_tall ? up256((size_t)fgrid * SF_WARPS * 4 * SF_HMAX * sizeof(double)) : 0);
    void* scratch = nullptr;
    int rc = snn_scratch(ctx, pbytes + sbytes + ebytes + cbytes + orbytes + slbytes + obytes + fbytes, &scratch);
    if (rc) return rc;
    char* b = (char*)scratch;
    int64_t* dpos = (int64_t*)b;                      b += pbytes;
    int32_t* dstale = stale ? (int32_t*)b : nullptr;  b += sbytes;
    double* dex = (double*)b;                         b += ebytes;
    ChromDev* dcd = (ChromDev*)b;                     b += cbytes;
    int32_t* dorder = (int32_t*)b;                    b += orbytes;
    int32_t* dslot = slot_of ? (int32_t*)b : nullptr; b += slbytes;
    double* dout = (double*)b;                        b += obytes;
    FastRec* dfrec = (FastRec*)b;
    int* dnext = (int*)((char*)dfrec + up256(nfast * sizeof(FastRec)));
    double* dtall = (double*)((char*)dnext + 256);
    if (nfast)
        SNN_CUDA(cudaMemcpyAsync(dfrec, frec.data(), nfast * sizeof(FastRec), cudaMemcpyHostToDevice, ctx->stream));
    SNN_CUDA(cudaMemcpyAsync(dpos, pos, (size_t)n * 4 * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream));
    if (stale)
        SNN_CUDA(cudaMemcpyAsync(dstale, stale, (size_t)n * 2 * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    if (slot_of)
        SNN_CUDA(cudaMemcpyAsync(dslot, slot_of, (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    std::vector<int32_t> flat;
    flat.reserve((size_t)n);
    size_t wstart[3];
    for (int k = 0; k < 3; ++k) { wstart[k] = flat.size(); flat.insert(flat.end(), worder[k].begin(), worder[k].end()); }
    std::vector<size_t> gstart(gorder.size());
    for (size_t g = 0; g < gorder.size(); ++g) { gstart[g] = flat.size(); flat.insert(flat.end(), gorder[g].begin(), gorder[g].end()); }
    if (!flat.empty())
        SNN_CUDA(cudaMemcpyAsync(dorder, flat.data(), flat.size() * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    std::vector<ChromDev> hcd((size_t)nslot);
    std::vector<double> hex((size_t)nslot * SF_EXSTRIDE, 0.0);
    for (int sl = 0; sl < nslot; ++sl) {
        hcd[sl].m = snn_view(slots[sl].c);
        hcd[sl].chromsize = slots[sl].c->chromsize;
        hcd[sl].nbins = slots[sl].c->nbins;
        hcd[sl].rex = want_s ? dex + (size_t)sl * SF_EXSTRIDE + 512 : nullptr;
        hcd[sl].rexm = want_s ? dex + (size_t)sl * SF_EXSTRIDE + 2048 + SF_RM : nullptr;
        hcd[sl].fb = slots[sl].c->fband ? slots[sl].c->fband + (slots[sl].c->bw - 1) : nullptr;
        hcd[sl].fbT = 2 * (int64_t)slots[sl].c->bw - 2;
        hcd[sl].nanbin = slots[sl].c->nanbin;
        hcd[sl].nanpre = slots[sl].c->nanpre;
        if (want_s) memcpy(hex.data() + (size_t)sl * SF_EXSTRIDE, slots[sl].expected, SNN_NDIST * sizeof(double));
    }
    SNN_CUDA(cudaMemcpyAsync(dcd, hcd.data(), (size_t)nslot * sizeof(ChromDev), cudaMemcpyHostToDevice, ctx->stream));
    int launches = 0;
    if (want_s)
        SNN_CUDA(cudaMemcpyAsync(dex, hex.data(), hex.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    if (nfast) {
        if (int rcp = snn_prepare_score_fast(ctx, dnext)) return rcp;
    }
    // the timed scope holds kernels only: every host-to-device copy of the call is queued above
    SnnScope* scope = new SnnScope(ctx, ST_SCORE, 0);
    struct ScopeEnd { SnnScope*& s; ~ScopeEnd() { delete s; } } scope_end{scope};   // error returns close it too
    if (want_s) {
        k_recip_expected<<<dim3(2 * SF_RM / 256, nslot), 256, 0, ctx->stream>>>(dex);
        SNN_LAUNCH_CHECK("k_recip_expected");
        ++launches;
    }
    // ---- fast kernel: every regular stripe of every chromosome in one launch
    if (nfast) {
        const char* rep = getenv("SNN_SCORE_REPEAT");     // measurement hook: launch the kernel N times (same outputs)
        const int nrep = rep ? atoi(rep) : 1;
        cudaEvent_t ev[18];
        const bool dbg = getenv("SNN_SCORE_DEBUG") != nullptr && nrep <= 16;
        if (dbg) { for (int q = 0; q <= nrep; ++q) cudaEventCreate(&ev[q]); cudaEventRecord(ev[0], ctx->stream); }
        for (int q = 0; q < (nrep < 1 ? 1 : nrep); ++q) {
            if (q) SNN_CUDA(cudaMemsetAsync(dnext, 0, sizeof(int), ctx->stream));
            rc = snn_launch_score_fast(ctx, dcd, dfrec, (int)nfast, bs, want_p, want_s, dnext, dtall, fgrid, dout);
            if (rc) return rc;
            ++launches;
            if (dbg) cudaEventRecord(ev[q + 1], ctx->stream);
        }
        if (dbg) {
            cudaStreamSynchronize(ctx->stream);
            for (int q = 0; q < nrep; ++q) { float ms = 0; cudaEventElapsedTime(&ms, ev[q], ev[q + 1]); fprintf(stderr, "[score debug] launch %d: %.3f ms\n", q, ms); }
            for (int q = 0; q <= nrep; ++q) cudaEventDestroy(ev[q]);
        }
    }
    // ---- warp kernel: one launch per height class, all chromosomes together
    for (int k = 0; k < 3; ++k) {
        if (worder[k].empty()) continue;
        rc = snn_launch_score_warp(ctx, dcd, dslot, dpos, dstale, dorder + wstart[k], (int)worder[k].size(), r, bs,
                                   kWarpH[k], want_p, want_s, dout);
        if (rc) return rc;
        ++launches;
    }
    // ---- CTA kernel for the rest
    const size_t per_row = 4 * sizeof(double) + sizeof(int) + 1;
    bool attr_set = false;
    for (int sl = 0; sl < nslot; ++sl) {
        for (int k = 0; k < 3; ++k) {
            const std::vector<int32_t>& ord = gorder[(size_t)sl * 3 + k];
            if (ord.empty()) continue;
            if (!attr_set) {   // one attribute value for every launch (contexts on several host threads share the function)
                size_t worst = 0;
                for (int q = 0; q < 3; ++q) {
                    size_t v = (size_t)cls[q].hcap * per_row + (size_t)cls[q].wcap * sizeof(double) + 64;
                    if (v > worst) worst = v;
                }
                SNN_CUDA(cudaFuncSetAttribute(k_score, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)worst));
                attr_set = true;
            }
            const SnnChrom* c = slots[sl].c;
            ScoreParams P;
            P.binsize = c->binsize;
            P.chromsize = c->chromsize;
            P.nbins = c->nbins;
            P.bs = bs;
            P.hmax = cls[k].hcap;
            P.bg_ncol = ctx->bg_ncol;
            P.mask_on = (ctx->mask_chrom >= 0 && c->id == ctx->mask_chrom) ? 1 : 0;
            P.mask_start = ctx->mask_start;
            P.mxs = ctx->mask_start / c->binsize;     // int(mask_start / resol), getStripe.py:617-618
            P.mxe = ctx->mask_end / c->binsize;
            const size_t smem = (size_t)cls[k].hcap * per_row + (size_t)cls[k].wcap * sizeof(double) + 64;
            k_score<<<(unsigned)ord.size(), SCORE_THREADS, smem, ctx->stream>>>(
                snn_view(c), dpos, dstale, dorder + gstart[(size_t)sl * 3 + k], hcd[sl].rex, ctx->bg_sorted,
                ctx->bg_valid, P, cls[k].wcap, want_p, want_s, dout);
            SNN_LAUNCH_CHECK("k_score");
            ++launches;
        }
    }
    scope->n = launches;
    delete scope;            // the device-to-host copy of the results is not kernel time
    scope = nullptr;
    std::vector<double> host((size_t)n * 6);
    SNN_CUDA(cudaMemcpyAsync(host.data(), dout, (size_t)n * 6 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    SNN_CUDA(snn_stream_sync(ctx));
    for (int64_t i = 0; i < n; ++i) {
        if (pvalue) pvalue[i] = host[i * 6 + 0];
        if (stripiness) stripiness[i] = host[i * 6 + 1];
        if (oe_mean) oe_mean[i] = host[i * 6 + 2];
        if (oe_total) oe_total[i] = host[i * 6 + 3];
        if (o_mean) o_mean[i] = host[i * 6 + 4];
        if (o_sum) o_sum[i] = host[i * 6 + 5];
    }
    return SNN_OK;
}

int snn_score_impl(snn_ctx* ctx, SnnChrom* c, int64_t n, const int64_t* pos, const int32_t* stale,
                   const double* expected, double* pvalue, double* stripiness, double* oe_mean, double* oe_total,
                   double* o_mean, double* o_sum) {
    std::vector<ScoreSlot> slots(1);
    slots[0].c = c;
    slots[0].expected = expected;
    return snn_score_multi_impl(ctx, slots, n, nullptr, pos, stale, pvalue, stripiness, oe_mean, oe_total, o_mean,
                                o_sum);
}
