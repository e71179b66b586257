// K9, warp-per-stripe version: getStripe.pvalue (getStripe.py:536-606), scoringstripes /
// iterate_idx (:608-788, stats.py:184-199) and getMean (:501-534) fused, for the stripes whose
// window lies inside the diagonal band and is at most SW_MAXCOLS bins wide (everything
// StripeSearch emits, and config 5).  The CTA-per-stripe kernel of score.cu keeps the rest.
//
// Why a warp and not a CTA: a stripe's window is ~1-2 k doubles, the work per stripe is a few
// thousand instructions, and the kernel of score.cu spent its time in __syncthreads and in
// single-thread sections (ncu r01b: 4 % SM busy, 95 us per CTA).  Here nothing is shared
// between warps, every step is a warp-uniform loop or a shuffle/ballot reduction, and a warp
// moves on to its next stripe without waiting for anybody.
//
// Data movement per stripe (the algorithmic bytes of DESIGN.md §3): the window is read ONCE
// from the band — rows above the diagonal by row (lanes along the columns), cells below the
// diagonal by column of the symmetric band (lanes along the rows), both contiguous — into a
// 32-row shared-memory tile with an odd pitch, then each lane owns one window row.
#include "score_warp.cuh"

#define FULL 0xffffffffu

__device__ __forceinline__ double w_qnan() { return __longlong_as_double(0x7ff8000000000000LL); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}
__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}

// #(row[0..n) >= v) for two sorted rows at once (the two dependent-load chains overlap)
__device__ __forceinline__ void count_ge2(const double* __restrict__ ra, int na, double va,
                                          const double* __restrict__ rb, int nb, double vb, int* ca, int* cb) {
    int la = 0, ha = (va == va) ? na : 0;
    int lb = 0, hb = (vb == vb) ? nb : 0;
#pragma unroll 1
    while (la < ha || lb < hb) {
        const bool da = la < ha, db = lb < hb;
        const int ma = (la + ha) >> 1, mb = (lb + hb) >> 1;
        const double xa = da ? __ldg(ra + ma) : 0.0;
        const double xb = db ? __ldg(rb + mb) : 0.0;
        if (da) { if (xa < va) la = ma + 1; else ha = ma; }
        if (db) { if (xb < vb) lb = mb + 1; else hb = mb; }
    }
    *ca = (va == va) ? na - la : 0;
    *cb = (vb == vb) ? nb - lb : 0;
}

// np.median of a[0..n) (n >= 1, no NaN) by rank counting; all lanes return the value.
// The k-th smallest is the element v with #(u < v) <= k < #(u <= v).
__device__ double warp_median(const double* a, int n, int lane) {
    const int khi = n >> 1, klo = (n & 1) ? khi : khi - 1;
    double vlo = 0.0, vhi = 0.0;
    bool flo = false, fhi = false;
    if (n <= 32) {                       // one candidate per lane
        const double va = lane < n ? a[lane] : 0.0;
        int la = 0, qa = 0;
#pragma unroll 4
        for (int j = 0; j < n; ++j) {
            const double u = a[j];
            la += (u < va); qa += (u <= va);
        }
        if (lane < n) {
            if (la <= klo && klo < qa) { vlo = va; flo = true; }
            if (la <= khi && khi < qa) { vhi = va; fhi = true; }
        }
    } else {
#pragma unroll 1
        for (int i0 = 0; i0 < n; i0 += 64) {
            const int ia = i0 + lane, ib = i0 + 32 + lane;
            const double va = ia < n ? a[ia] : 0.0, vb = ib < n ? a[ib] : 0.0;
            int la = 0, qa = 0, lb = 0, qb = 0;
#pragma unroll 2
            for (int j = 0; j < n; ++j) {
                const double u = a[j];
                la += (u < va); qa += (u <= va);
                lb += (u < vb); qb += (u <= vb);
            }
            if (ia < n) {
                if (la <= klo && klo < qa) { vlo = va; flo = true; }
                if (la <= khi && khi < qa) { vhi = va; fhi = true; }
            }
            if (ib < n) {
                if (lb <= klo && klo < qb) { vlo = vb; flo = true; }
                if (lb <= khi && khi < qb) { vhi = vb; fhi = true; }
            }
        }
    }
    const unsigned bl = __ballot_sync(FULL, flo), bh = __ballot_sync(FULL, fhi);
    vlo = __shfl_sync(FULL, vlo, __ffs(bl) - 1);
    vhi = __shfl_sync(FULL, vhi, __ffs(bh) - 1);
    return (n & 1) ? vhi : (vlo + vhi) / 2.0;
}

// np.median of a[0..n) for n > 32: bitonic sort of a padded copy in `buf` (shared, >= 512 doubles,
// the window tile is free by then) instead of the O(n^2 / 32) rank counting above — 45 compare-
// exchange sweeps for 512 values against ~20k comparisons per lane for a 300-row stripe
// (ncu r01p: rank counting was 52 % of the instructions of the tall-stripe launch).
__device__ double warp_median_sort(double* buf, const double* a, int n, int lane) {
    int P = 64;
    while (P < n) P <<= 1;
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    for (int i = lane; i < P; i += 32) buf[i] = i < n ? a[i] : inf;
    __syncwarp();
    for (int k = 2; k <= P; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = lane; t < (P >> 1); t += 32) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int l = i | j;
                const bool up = (i & k) == 0;
                const double x = buf[i], y = buf[l];
                if ((x > y) == up) { buf[i] = y; buf[l] = x; }
            }
            __syncwarp();
        }
    }
    const double hi = buf[n >> 1], lo = buf[(n - 1) >> 1];
    __syncwarp();
    return (n & 1) ? hi : (lo + hi) / 2.0;
}

__device__ __forceinline__ double warp_median_any(double* buf, const double* a, int n, int lane) {
    return n <= 32 ? warp_median(a, n, lane) : warp_median_sort(buf, a, n, lane);
}

__device__ __forceinline__ void cp_async8(double* sdst, const double* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(sdst)), "l"(gsrc));
}

__device__ __forceinline__ uint64_t low_mask64(int n) { return n >= 64 ? ~0ull : ((1ull << n) - 1ull); }

__global__ void __launch_bounds__(SW_WARPS * 32, SW_MINB) k_score_warp(const ChromDev* __restrict__ cds,
                                                              const int32_t* __restrict__ cslot,
                                                              const int64_t* __restrict__ pos,
                                                              const int32_t* __restrict__ stale,
                                                              const int32_t* __restrict__ order, int n,
                                                              const double* __restrict__ bg,
                                                              const int32_t* __restrict__ bg_valid,
                                                              int64_t bg_ncol, int64_t binsize, int bs, int hcap,
                                                              int want_p, int want_s, double* __restrict__ out) {
    extern __shared__ double sh[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* win = sh + (size_t)warp * (SW_WINCAP + 4 * (size_t)hcap);
    double* pv = win + SW_WINCAP;     // per-row p-values, later the compacted centre means
    double* scA = pv + hcap;          // O/E row sums -> row means: centre
    double* slA = scA + hcap;         //                            left
    double* srA = slA + hcap;         //                            right
    const int64_t r = binsize;

    for (int it = blockIdx.x * SW_WARPS + warp; it < n; it += gridDim.x * SW_WARPS) {
        const int64_t s = order ? order[it] : it;
        const ChromDev* cd = cds + (cslot ? cslot[s] : 0);
        const double* __restrict__ band = cd->m.band;
        const int bwm1 = cd->m.bw - 1;
        const int64_t chromsize = cd->chromsize;
        const int nbins = (int)cd->nbins;
        const double* __restrict__ rex = cd->rex;
        const int64_t p1 = pos[s * 4 + 0], p2 = pos[s * 4 + 1], p3 = pos[s * 4 + 2], p4 = pos[s * 4 + 3];
        double res0 = w_qnan(), res1 = w_qnan(), res2 = w_qnan(), res3 = w_qnan(), res4 = w_qnan(), res5 = w_qnan();

        // ---- geometry: ten floor divisions, one per lane, then broadcast
        int64_t lmost = p1 - (int64_t)bs * r;
        if (lmost < 1) lmost = 1;
        int64_t rmost = p2 + (int64_t)bs * r;
        if (rmost > chromsize) rmost = chromsize;
        int64_t num = 0;
        switch (lane) {
            case 0: num = p3; break;
            case 1: num = p4 + r - 1; break;
            case 2: num = lmost; break;
            case 3: num = rmost + r - 1; break;
            case 4: num = p1; break;
            case 5: num = p2; break;
            case 6: num = p4; break;
            case 7: num = p2 + r - 1; break;
            case 8: num = p1 - 1; break;
            case 9: num = p3 - 1; break;
            default: break;
        }
        // coordinates below 2^32 bp (every real genome) take the 32-bit divider (~4x fewer instructions)
        const int qv = ((uint64_t)(num | r) >> 32) == 0 ? (int)((uint32_t)num / (uint32_t)r) : (int)(num / r);
        const int R0 = __shfl_sync(FULL, qv, 0), R1 = __shfl_sync(FULL, qv, 1);
        const int C0 = __shfl_sync(FULL, qv, 2), C1 = __shfl_sync(FULL, qv, 3);
        const int xsi = __shfl_sync(FULL, qv, 4), xei = __shfl_sync(FULL, qv, 5);
        const int yei = __shfl_sync(FULL, qv, 6), cC1 = __shfl_sync(FULL, qv, 7);
        const int x1 = __shfl_sync(FULL, qv, 8), y1 = __shfl_sync(FULL, qv, 9);
        const int ysi = R0, cC0 = xsi, x2 = xei, y2 = yei;
        const int h = R1 - R0;
        int lmb = xsi - bs, rmb = xei + bs;
        if (lmb < 1) lmb = 1;
        if (rmb >= nbins) rmb = nbins - 1;
        int cmin = C0, cmax = C1;
        if (want_s) {
            cmin = min(cmin, min(lmb, cC0));
            cmax = max(cmax, max(rmb, cC1));
        }
        if (!want_p) { cmin = min(lmb, cC0); cmax = max(rmb, cC1); }
        if (cmin < 0) cmin = 0;
        if (cmax > nbins) cmax = nbins;
        const int ncw = cmax - cmin;
        const int pitch = ncw | 1;
        const int RC = min(32, SW_WINCAP / pitch);

        // p-value slices (getStripe.py:561-566) and background row selection (:583-597)
        const int ncol = C1 - C0;
        int64_t ca, cb, la, lb, ra, rb;
        snn_pyslice(bs, -(int64_t)bs, ncol, &ca, &cb);
        snn_pyslice(0, bs, ncol, &la, &lb);
        snn_pyslice(-(int64_t)bs, ncol, ncol, &ra, &rb);
        if (bs == 0) { ra = 0; rb = ncol; }
        int mode;
        if (x1 == y1) mode = 0;
        else if (x2 == y2) mode = 1;
        else if (stale && stale[s * 2] >= 0) mode = 2;
        else mode = 3;

        // stripiness column parts (getStripe.py:684-698) and the shape checks of expecMatrix
        const int wc = cC1 - cC0, wl = xsi - lmb, wr = rmb - xei;
        const bool s_ok = want_s && wc == xei - xsi && h == yei - ysi && wc > 0 && wl >= 0 && wr >= 0 &&
                          wc <= 64 && wl <= 64 && wr <= 64 && R0 == ysi;
        uint64_t aliveC = 0, aliveL = 0, aliveR = 0;
        double tot_obs = 0.0;
        int n_obs = 0;

        #pragma unroll 1
        for (int j0 = 0; j0 < h; j0 += RC) {
            const int nr = min(RC, h - j0);
            const int row0 = R0 + j0;
            __syncwarp();
            // Element (row, col) of the symmetric matrix sits at band[min*(bw-1) + max]: both loops
            // below walk that address linearly.  The copies are cp.async (global -> shared without
            // a register stop): every cell of the tile is in flight before the first one is
            // awaited (ncu r01p: the load-then-store loops exposed one DRAM round trip per row).
            // cells on/above the diagonal: by window row, lanes along the band row
            {
                const int nU = min(nr, cmax - row0);                  // rows that reach the diagonal
                const double* __restrict__ g = band + (int64_t)row0 * bwm1 + cmin + lane;
                double* sdst = win + lane;
                int d = row0 - cmin - lane;                           // row - col of this lane's first cell
                #pragma unroll 4
                for (int rr = 0; rr < nU; ++rr) {
                    for (int cc = 0; cc + lane < ncw; cc += 32)
                        if (d - cc <= 0) cp_async8(sdst + cc, g + cc);
                    g += bwm1;
                    sdst += pitch;
                    ++d;
                }
            }
            // cells below the diagonal: by window column, lanes along that column's band row
            {
                const int nL = min(ncw, row0 + nr - 1 - cmin);        // columns left of the last row
                const double* __restrict__ g = band + (int64_t)cmin * bwm1 + row0 + lane;
                double* sdst = win + lane * pitch;
                int d = row0 + lane - cmin;                           // row - col
                const bool rowok = lane < nr;
                #pragma unroll 4
                for (int cc = 0; cc < nL; ++cc) {
                    if (rowok && d > 0) cp_async8(sdst, g);
                    g += bwm1;
                    ++sdst;
                    --d;
                }
            }
            asm volatile("cp.async.commit_group;");
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncwarp();
            const bool act = lane < nr;
            const int j = j0 + lane;
            const double* wrow = win + (act ? lane : 0) * pitch - cmin;   // indexed by chromosome bin

            if (want_p) {
                double sc = 0.0, sl = 0.0, sr = 0.0;
                #pragma unroll 2
                for (int c = (int)ca; c < (int)cb; ++c) sc += snn_nan0(wrow[C0 + c]);
                #pragma unroll 2
                for (int c = (int)la; c < (int)lb; ++c) sl += snn_nan0(wrow[C0 + c]);
                #pragma unroll 2
                for (int c = (int)ra; c < (int)rb; ++c) sr += snn_nan0(wrow[C0 + c]);
                const double mc = (cb > ca) ? sc / (double)(cb - ca) : w_qnan();
                const double ml = (lb > la) ? sl / (double)(lb - la) : w_qnan();
                const double mr = (rb > ra) ? sr / (double)(rb - ra) : w_qnan();
                const double dl = mc - ml, dr = mc - mr;
                double p = w_qnan();
                int table = 0, d = -1;
                if (mode == 0) { table = 1; d = j < 399 ? j : 399; }
                else if (mode == 1) {
                    int dd = y2 - y1 - j - 1;
                    if (dd > 399) dd = 399;
                    if (dd < 0) dd += SNN_NDIST;
                    d = dd;
                } else if (mode == 2) { table = stale[s * 2] ? 0 : 1; d = stale[s * 2 + 1]; }
                if (act && d >= 0 && d < SNN_NDIST) {
                    const int64_t rl = (int64_t)(table == 0 ? 0 : 2) * SNN_NDIST + d;
                    const int64_t rr_ = (int64_t)(table == 0 ? 1 : 3) * SNN_NDIST + d;
                    const int nl = bg_valid[rl], nr_ = bg_valid[rr_];
                    int cl, cr;
                    count_ge2(bg + rl * bg_ncol, nl, dl, bg + rr_ * bg_ncol, nr_, dr, &cl, &cr);
                    const double a = (double)cl / (double)nl;
                    const double b = (double)cr / (double)nr_;
                    p = a;
                    if (b > p) p = b;
                    if (p == 0.0) p = 1.0 / (double)bg_ncol;
                }
                if (act) pv[j] = p;
            }
            if (s_ok) {
                const int row = row0 + lane;
                double sc = 0.0, sl = 0.0, sr = 0.0;
                #pragma unroll 1
                for (int c = 0; c < wc; ++c) {
                    const double v = wrow[cC0 + c];
                    const bool fin = act && (v == v);
                    if (__ballot_sync(FULL, fin)) aliveC |= 1ull << c;
                    int dd = abs(xsi + c - row);
                    if (dd > 399) dd = 399;
                    sc += snn_nan0(v * __ldg(rex + dd));
                    if (fin) { tot_obs += v; ++n_obs; }
                }
                #pragma unroll 1
                for (int c = 0; c < wl; ++c) {
                    const double v = wrow[lmb + c];
                    if (__ballot_sync(FULL, act && (v == v))) aliveL |= 1ull << c;
                    int dd = abs(lmb + c - row);
                    if (dd > 399) dd = 399;
                    sl += snn_nan0(v * __ldg(rex + dd));
                }
                #pragma unroll 1
                for (int c = 0; c < wr; ++c) {
                    const double v = wrow[xei + c];
                    if (__ballot_sync(FULL, act && (v == v))) aliveR |= 1ull << c;
                    int dd = abs(xei + 1 + c - row);   // the right part's expected is shifted by one
                    if (dd > 399) dd = 399;
                    sr += snn_nan0(v * __ldg(rex + dd));
                }
                if (act) { scA[j] = sc; slA[j] = sl; srA[j] = sr; }
            } else if (want_s && wc > 0) {
                // getMean alone (getStripe.py:501-534): it needs no valid flank geometry
                #pragma unroll 1
                for (int c = 0; c < wc; ++c) {
                    const double v = wrow[cC0 + c];
                    if (act && (v == v)) { tot_obs += v; ++n_obs; }
                }
            }
        }
        __syncwarp();

        // ================================ p-value: np.median over the rows ==========================
        if (want_p) {
            int anynan = 0;
            #pragma unroll 1
            for (int j = lane; j < h; j += 32) anynan |= (pv[j] != pv[j]);
            anynan = __any_sync(FULL, anynan);
            res0 = anynan ? w_qnan() : warp_median_any(win, pv, h, lane);
        }

        // ================================ stripiness ================================================
        if (s_ok) {
            const uint64_t deadC = ~aliveC & low_mask64(wc), deadL = ~aliveL & low_mask64(wl),
                           deadR = ~aliveR & low_mask64(wr);
            const int lc = wc - __popcll(deadC), ll = wl - __popcll(deadL), lr = wr - __popcll(deadR);
            bool bad = false;
            int hk = h;
            if (deadC | deadL | deadR) {
                // rows deleted with the dead columns: index x inside its part when xs == ys, else
                // shape[0]-1-x (getStripe.py:713-731; np.delete wraps negatives, IndexError -> NaN)
                const bool diag = (p1 == p3);
                unsigned rd[SW_HMAX / 32];
#pragma unroll
                for (int t = 0; t < SW_HMAX / 32; ++t) rd[t] = 0u;
                #pragma unroll 1
                for (int part = 0; part < 3; ++part) {
                    uint64_t dm = part == 0 ? deadC : (part == 1 ? deadL : deadR);
                    while (dm) {
                        const int x = __ffsll((long long)dm) - 1;
                        dm &= dm - 1;
                        int idx = diag ? x : h - 1 - x;
                        if (idx < 0) idx += h;
                        if (idx < 0 || idx >= h) { bad = true; continue; }
#pragma unroll
                        for (int t = 0; t < SW_HMAX / 32; ++t)
                            if ((idx >> 5) == t) rd[t] |= 1u << (idx & 31);
                    }
                }
                if (!bad) {
                    // compact the kept rows in place, 32 at a time (order preserved; a chunk's
                    // destinations never reach the rows of later chunks)
                    int base = 0;
                    #pragma unroll 1
                    for (int t = 0; t * 32 < h; ++t) {
                        const int j = t * 32 + lane;
                        const unsigned valid = (h - t * 32 >= 32) ? FULL : ((1u << (h - t * 32)) - 1u);
                        unsigned word = 0u;
#pragma unroll
                        for (int u = 0; u < SW_HMAX / 32; ++u)
                            if (u == t) word = rd[u];
                        const unsigned kept = ~word & valid;
                        const bool mine = (kept >> lane) & 1u;
                        double tc = 0.0, tl = 0.0, tr = 0.0;
                        if (mine) { tc = scA[j]; tl = slA[j]; tr = srA[j]; }
                        __syncwarp();
                        if (mine) {
                            const int k = base + __popc(kept & ((1u << lane) - 1u));
                            scA[k] = tc; slA[k] = tl; srA[k] = tr;
                        }
                        __syncwarp();
                        base += __popc(kept);
                    }
                    hk = base;
                }
            }
            if (!bad) {
                // row means of nan0(O/E) over the live columns
                double tot_oe = 0.0;
                #pragma unroll 1
                for (int k = lane; k < hk; k += 32) {
                    const double sc = scA[k], sl = slA[k], sr = srA[k];
                    tot_oe += sc;
                    scA[k] = lc ? sc / (double)lc : w_qnan();
                    slA[k] = ll ? sl / (double)ll : w_qnan();
                    srA[k] = lr ? sr / (double)lr : w_qnan();
                }
                tot_oe = warp_sum(tot_oe);
                res2 = (hk > 0 && lc > 0) ? tot_oe / ((double)hk * (double)lc) : w_qnan();
                res3 = (double)lc * tot_oe;           // centerTotal quirk (getStripe.py:747)
                __syncwarp();
                // Sobel-like sums (stats.elementwise_product_sum) and the score
                double acc = 0.0;
                int nacc = 0;
                #pragma unroll 1
                for (int i = 1 + lane; i + 1 < hk; i += 32) {
                    const double l0 = slA[i - 1], l1 = slA[i], l2 = slA[i + 1];
                    const double c0 = scA[i - 1], c1 = scA[i], c2 = scA[i + 1];
                    const double r0 = srA[i - 1], r1 = srA[i], r2 = srA[i + 1];
                    const double gxl = ((-1.0 * l0 + -2.0 * l1) + -1.0 * l2) + ((1.0 * c0 + 2.0 * c1) + 1.0 * c2);
                    const double gxr = ((1.0 * c0 + 2.0 * c1) + 1.0 * c2) + ((-1.0 * r0 + -2.0 * r1) + -1.0 * r2);
                    double gy = (((1.0 * l0 + 0.0 * l1) + -1.0 * l2) + ((2.0 * c0 + 0.0 * c1) + -2.0 * c2)) +
                                ((1.0 * r0 + 0.0 * r1) + -1.0 * r2);
                    if (gy < 0) gy *= -1.0;
                    const double gx = (gxl != gxl || gxr != gxr) ? w_qnan() : fmin(gxl, gxr);
                    const double d = gx - gy;
                    if (d == d) { acc += d; ++nacc; }
                }
                acc = warp_sum(acc);
                nacc = warp_sum(nacc);
                const double avgdiff = nacc ? acc / (double)nacc : w_qnan();
                // nanmedian(centre means): compact the non-NaN values into pv
                int nm = 0;
                #pragma unroll 1
                for (int k0 = 0; k0 < hk; k0 += 32) {
                    const int k = k0 + lane;
                    const double v = k < hk ? scA[k] : w_qnan();
                    const bool ok = v == v;
                    const unsigned b = __ballot_sync(FULL, ok);
                    if (ok) pv[nm + __popc(b & ((1u << lane) - 1u))] = v;
                    nm += __popc(b);
                }
                __syncwarp();
                const double med = nm ? warp_median_any(win, pv, nm, lane) : w_qnan();
                res1 = med * avgdiff;
            }
        }
        if (want_s) {      // getMean: independent of the stripiness shape checks
            tot_obs = warp_sum(tot_obs);
            n_obs = warp_sum(n_obs);
            res4 = n_obs ? tot_obs / (double)n_obs : w_qnan();
            res5 = tot_obs;
        }
        if (lane < 6) {
            const double v = lane == 0 ? res0 : lane == 1 ? res1 : lane == 2 ? res2 : lane == 3 ? res3
                           : lane == 4 ? res4 : res5;
            out[s * 6 + lane] = v;
        }
    }
}

int snn_launch_score_warp(snn_ctx* ctx, const ChromDev* cds, const int32_t* cslot, const int64_t* pos,
                          const int32_t* stale, const int32_t* order, int n, int64_t binsize, int bs, int hcap,
                          int want_p, int want_s, double* out) {
    if (n <= 0) return SNN_OK;
    const size_t smem = (size_t)SW_WARPS * (SW_WINCAP + 4 * (size_t)hcap) * sizeof(double);
    static const size_t worst = (size_t)SW_WARPS * (SW_WINCAP + 4 * (size_t)SW_HMAX) * sizeof(double);
    SNN_CUDA(cudaFuncSetAttribute(k_score_warp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)worst));
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int per_sm = (int)((size_t)(227 * 1024) / (smem + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 16) per_sm = 16;
    int grid = (n + SW_WARPS - 1) / SW_WARPS;
    if (grid > sms * per_sm) grid = sms * per_sm;
    k_score_warp<<<grid, SW_WARPS * 32, smem, ctx->stream>>>(cds, cslot, pos, stale, order, n, ctx->bg_sorted,
                                                           ctx->bg_valid, ctx->bg_ncol, binsize, bs, hcap, want_p,
                                                           want_s, out);
    SNN_LAUNCH_CHECK("k_score_warp");
    return SNN_OK;
}
