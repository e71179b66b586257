// K9, fast version: getStripe.pvalue (getStripe.py:536-606), scoringstripes / iterate_idx (:608-788,
// stats.py:184-199) and getMean (:501-534) fused, for REGULAR stripes — the geometry StripeSearch
// emits away from the chromosome ends and `stripenn score` re-reads from its own tables: pos2 and
// pos4 on bin boundaries, both flanks complete, window inside the band.  Everything else (clipped
// flanks, --mask, far pixels, tall stripes) stays with k_score_warp / k_score; the host classifies
// (score.cu) and the two families give identical numbers (tests/test_gpu_score_fast.py).
//
// What is different from k_score_warp (ncu r01q: 5.6 k warp instructions per 55 x 17 stripe, issue
// 27-46 %, 15-29 % of the warps resident because every warp carried an 8 KB window tile):
//  * NO window tile.  The chromosome also lives as a column-major FULL band (band.cu: k_fband_*,
//    pixel (row, col) at fb[col * T + row]); lane = window row, so every window column is ONE coalesced
//    run of rows on both sides of the diagonal and each cell is loaded straight into the register that
//    consumes it.  Shared memory holds only the per-row results (4 doubles per row): 32-64 warps per SM.
//  * for a regular stripe the p-value's column slices (:561-566) and the stripiness parts (:684-698)
//    are the same three column ranges, so ONE pass over the window feeds both: per cell 2 loads
//    (value, mirrored 1/(expected+1e-8) table indexed by col - row), a NaN test, two predicated adds
//    and a multiply.  NaN bookkeeping (dead columns, np.delete of rows) runs only for windows that
//    met a NaN, in a second pass.
//  * the host precomputes the bin geometry (FastRec): no 64-bit divisions in the kernel.
//  * every division by a small integer (column counts, live-column counts, valid background
//    counts) is a Markstein sequence on a correctly rounded reciprocal — five FMA-pipe
//    instructions, correctly rounded like `/` (tools/check_fdiv.c: exhaustive over every count pair up to 26000).
//  * branch-free fixed-depth lower bound in the sorted background rows (both tables interleaved).
//  * medians by a warp quickselect (three-way counts with redux.sync) instead of a bitonic sort.
//
// Roofline: HBM/L2 (window bytes are requested exactly once, sector-aligned runs of h doubles).
// Algorithmic bytes per stripe: 8 * h * (w + 2*bs) + 48.
#include <stdlib.h>

#include "score_warp.cuh"

#define FULL 0xffffffffu

__device__ __forceinline__ double f_qnan() { return __longlong_as_double(0x7ff8000000000000LL); }
__device__ __forceinline__ uint64_t f_low64(int n) { return n >= 64 ? ~0ull : ((1ull << n) - 1ull); }

__device__ __forceinline__ double f_warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}

// a / b for finite a >= 0 and an integer-valued b >= 1 with y = RN(1 / b): two Newton corrections of
// RN(a * y); the last one rounds a value within 2^-104 (relative) of the exact quotient, i.e. the
// correctly rounded quotient (Markstein; tools/check_fdiv.c)
__device__ __forceinline__ double f_div(double a, double b, double y) {
    double q = a * y;
    double r = fma(-b, q, a);
    q = fma(r, y, q);
    r = fma(-b, q, a);
    return fma(r, y, q);
}

// The three column parts of one window row — left flank, centre, right flank — as raw and O/E sums
// in column order per part (the orders of k_score_warp; the full band holds nan0(pixel), so there is
// no NaN test).  The parts advance together, two columns each per iteration: six independent
// accumulation chains and twelve loads in flight per lane.  prow -> pixel (row, column 0) as bytes,
// T8 = column step in bytes; rp -> mirrored 1/(expected+1e-8) entry of column C0 (the right part's
// expected is shifted by one bin, getStripe.py:696-698).
// Every address is base + int * int, forced to ONE IMAD.WIDE through PTX (left to itself the compiler
// merged the bases into the products and expanded full 64-bit multiplications: 10-14 instructions per cell).
__device__ __forceinline__ const double* f_at(const void* base, int k, int step) {
    unsigned long long r;
    asm("mad.wide.s32 %0, %1, %2, %3;" : "=l"(r) : "r"(k), "r"(step), "l"(base));
    return reinterpret_cast<const double*>(r);
}

// window pixels are read exactly once: SF_WIN_CS loads them with the evict-first policy so that they do
// not push the search-tree and expected-value lines out of L1 / L2
#ifndef SF_WIN_LDG      // default: streaming loads (A/B r02A: 0.775 vs 0.79 ms per 100k stripes)
#define F_WIN(p) __ldcs(p)
#else
#define F_WIN(p) __ldg(p)
#endif

template <bool WS>
__device__ __forceinline__ void f_run(const char* __restrict__ prow, int T8, int c, const double* __restrict__ rp, int e,
                                      int n, double& sraw, double& soe) {
    const double* pe = f_at(rp, e, 8);
#pragma unroll 1
    for (; n >= 2; n -= 2, c += 2, pe += 2) {
        const double v0 = F_WIN(f_at(prow, c, T8)), v1 = F_WIN(f_at(prow + T8, c, T8));
        if (WS) {
            const double e0 = __ldg(pe), e1 = __ldg(pe + 1);
            soe += v0 * e0;
            soe += v1 * e1;
        }
        sraw += v0;
        sraw += v1;
    }
    if (n > 0) {
        const double v0 = F_WIN(f_at(prow, c, T8));
        if (WS) soe += v0 * __ldg(pe);
        sraw += v0;
    }
}

template <bool WS>
__device__ __forceinline__ void f_row(const char* __restrict__ prow, int T8, int C0, int xsi, int wc, int bs,
                                      const double* __restrict__ rp, double& rl, double& ol, double& rc, double& oc,
                                      double& rr, double& orr) {
    const int m = bs < wc ? bs : wc;
    // one counter, fixed bases
    const char* l0 = reinterpret_cast<const char*>(f_at(prow, C0, T8));
    const char* c0p = reinterpret_cast<const char*>(f_at(prow, xsi, T8));
    const char* r0 = reinterpret_cast<const char*>(f_at(prow, xsi + wc, T8));
    const char* l1 = l0 + T8;
    const char* c1p = c0p + T8;
    const char* r1 = r0 + T8;
    const double* el = rp;
    const double* ec = rp + bs;
    const double* er = rp + (bs + wc + 1);
    int k = 0;
#pragma unroll 1
    for (; k + 2 <= m; k += 2) {
        const double a0 = F_WIN(f_at(l0, k, T8)), a1 = F_WIN(f_at(l1, k, T8));
        const double c0 = F_WIN(f_at(c0p, k, T8)), c1 = F_WIN(f_at(c1p, k, T8));
        const double b0 = F_WIN(f_at(r0, k, T8)), b1 = F_WIN(f_at(r1, k, T8));
        if (WS) {
            const double* qa = f_at(el, k, 8);
            const double* qc = f_at(ec, k, 8);
            const double* qb = f_at(er, k, 8);
            const double ea0 = __ldg(qa), ea1 = __ldg(qa + 1);
            const double ec0 = __ldg(qc), ec1 = __ldg(qc + 1);
            const double eb0 = __ldg(qb), eb1 = __ldg(qb + 1);
            ol += a0 * ea0;
            oc += c0 * ec0;
            orr += b0 * eb0;
            ol += a1 * ea1;
            oc += c1 * ec1;
            orr += b1 * eb1;
        }
        rl += a0;
        rc += c0;
        rr += b0;
        rl += a1;
        rc += c1;
        rr += b1;
    }
    f_run<WS>(prow, T8, C0 + k, rp, k, bs - k, rl, ol);
    f_run<WS>(prow, T8, xsi + k, rp, bs + k, wc - k, rc, oc);
    f_run<WS>(prow, T8, xsi + wc + k, rp, bs + wc + 1 + k, bs - k, rr, orr);
}

// np.median of a[0..n) (shared memory, n >= 1, no NaN) by quickselect: every round
// takes a pivot from the remaining candidates, counts the candidates below / not above it with two
// redux.sync and keeps one side.  KC = 32-value slots per lane (n <= 32 * KC).
template <int KC>
__device__ double f_median(const double* a, int n, int lane) {
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    double v[KC];
    unsigned cm = 0;
#pragma unroll
    for (int k = 0; k < KC; ++k) {
        const int j = k * 32 + lane;
        const bool ok = j < n;
        v[k] = ok ? a[j] : inf;
        if (ok) cm |= 1u << k;
    }
    const int khi = n >> 1;
    int below = 0;          // values known to be below every candidate
    double xhi = 0.0;
    int lt_hi = 0;          // #(values < xhi)
#pragma unroll 1
    for (int round = 0;; ++round) {
        const unsigned b = __ballot_sync(FULL, cm != 0);       // never 0: the answer is a candidate
        const int rot = (round * 11) & 31;
        const unsigned rb = __funnelshift_r(b, b, rot);
        const int src = (__ffs(rb) - 1 + rot) & 31;
        const int jmine = (__ffs(cm) - 1) * 32 + lane;
        const int jp = __shfl_sync(FULL, jmine, src);
        const double p = a[jp];
        unsigned lt = 0, le = 0;
#pragma unroll
        for (int k = 0; k < KC; ++k) {
            if (v[k] < p) lt |= 1u << k;
            if (v[k] <= p) le |= 1u << k;
        }
        lt &= cm;
        le &= cm;
        const int clt = __reduce_add_sync(FULL, __popc(lt));
        const int cle = __reduce_add_sync(FULL, __popc(le));
        const int kk = khi - below;
        if (kk < clt) cm = lt;
        else if (kk < cle) { xhi = p; lt_hi = below + clt; break; }
        else { cm &= ~le; below += cle; }
    }
    if (n & 1) return xhi;
    double xlo = xhi;
    if (lt_hi > khi - 1) {
        // the lower middle value is the largest value below xhi: maximum over the order-preserving
        // 64-bit keys of the doubles, high words first, then the low words of the lanes that hold it
        double m = -inf;
#pragma unroll
        for (int k = 0; k < KC; ++k)
            if (v[k] < xhi && v[k] > m) m = v[k];
        const uint64_t bits = (uint64_t)__double_as_longlong(m);
        const uint64_t key = bits ^ ((bits >> 63) ? ~0ull : 0x8000000000000000ull);
        const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
        const unsigned hmax = __reduce_max_sync(FULL, hi);
        const unsigned lmax = __reduce_max_sync(FULL, hi == hmax ? lo : 0u);
        const uint64_t kmax = ((uint64_t)hmax << 32) | lmax;
        xlo = __longlong_as_double((long long)(kmax ^ ((kmax >> 63) ? 0x8000000000000000ull : ~0ull)));
    }
    return (xlo + xhi) / 2.0;
}

// Runtime choice of the median's slot count (the register array must be compile-time)
__device__ __forceinline__ double f_median_any(const double* a, int n, int lane) {
    if (n <= 32) return f_median<1>(a, n, lane);
    if (n <= 64) return f_median<2>(a, n, lane);
    if (n <= 128) return f_median<4>(a, n, lane);
    return f_median<SF_HMAX / 32>(a, n, lane);
}

// ONE launch for every regular stripe of a call.  The host sorts the stripes tallest first and the
// warps take them from an atomic counter, so the long stripes start first and the short ones fill
// the tail (per-class launches with a static split left the >128-row class at 11-19 % of the
// warps resident, 425-680 us for 8 % of the stripes: profiles/r02x).  Row results of stripes up to
// SF_HSM rows live in shared memory (4 KB per warp), taller stripes use the warp's slice of `tall`.
template <bool WP, bool WS>
__global__ void __launch_bounds__(SF_WARPS * 32, SF_MINB) k_score_fast(const ChromDev* __restrict__ cds,
                                                            const FastRec* __restrict__ recs, int n,
                                                            int* __restrict__ next,
                                                            const double* __restrict__ bg,
                                                            const int32_t* __restrict__ bg_valid,
                                                            const double* __restrict__ bg_rcp, int bg_ncol,
                                                            int bg_half, const float4* __restrict__ bg_tree,
                                                            int tree_levels, int tree_nodes, int tree_dbg, double pmin,
                                                            int bs,
                                                            double* __restrict__ tall,
                                                            double* __restrict__ out) {
    constexpr int KC = SF_HMAX / 32;
    extern __shared__ double sh[];
    double* rcp = sh;                                  // rcp[k] = 1 / k, k <= 64
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* const sm_rows = sh + 72 + (size_t)warp * 4 * SF_HSM;
    double* const gl_rows = tall + ((size_t)blockIdx.x * SF_WARPS + warp) * 4 * SF_HMAX;
    if (threadIdx.x <= 64) rcp[threadIdx.x] = 1.0 / (double)threadIdx.x;
    __syncthreads();
    const double ybs = rcp[bs], dbs = (double)bs;

    for (;;) {
        int it = 0;
        if (lane == 0) it = atomicAdd(next, 1);
        it = __shfl_sync(FULL, it, 0);
        if (it >= n) break;
        const int4 g0 = __ldg(reinterpret_cast<const int4*>(recs + it));
        const int4 g1 = __ldg(reinterpret_cast<const int4*>(recs + it) + 1);
        const int xsi = g0.x, wc = g0.y, R0 = g0.z, h = g0.w;
        const int mode = g1.x & 15, mtable = (g1.x >> 4) & 1, aux = g1.y;
        const bool diag = (g1.x >> 8) & 1;              // pos1 == pos3
        const ChromDev* cd = cds + g1.z;
        const int64_t s = g1.w;
        const double* __restrict__ fb = cd->fb;
        const int64_t T = cd->fbT;
        const int T8 = (int)T * 8;
        const double* __restrict__ rexm = cd->rexm;
        const int C0 = xsi - bs;
        const double ywc = rcp[wc], dwc = (double)wc;
        const int cap = h <= SF_HSM ? SF_HSM : SF_HMAX;
        double* pv = h <= SF_HSM ? sm_rows : gl_rows;   // per-row p-values, later the compacted centre means
        double* scA = pv + cap;                          // O/E row sums -> row means: centre
        double* slA = scA + cap;                         //                            left
        double* srA = slA + cap;                         //                            right
        double res0 = f_qnan(), res1 = f_qnan(), res2 = f_qnan(), res3 = f_qnan(), res4 = f_qnan(), res5 = f_qnan();
        double tot_obs = 0.0;
        int pnan = 0;
        __syncwarp();

#pragma unroll 1
        for (int j0 = 0; j0 < h; j0 += 32) {
            const int j = j0 + lane;
            const bool act = j < h;
            const int row = R0 + (act ? j : h - 1);        // idle lanes repeat the last row (loads stay in range)
            const char* prow = reinterpret_cast<const char*>(fb + row);
            const double* rp = WS ? rexm + (C0 - row) : nullptr;
            double rl = 0.0, ol = 0.0, rc = 0.0, oc = 0.0, rr = 0.0, orr = 0.0;
            f_row<WS>(prow, T8, C0, xsi, wc, bs, rp, rl, ol, rc, oc, rr, orr);

            if (WP) {
                const double mc = f_div(rc, dwc, ywc), ml = f_div(rl, dbs, ybs), mr = f_div(rr, dbs, ybs);
                const double dl = mc - ml, dr = mc - mr;
                int table = 0, d = -1;
                if (mode == 0) { table = 1; d = j < 399 ? j : 399; }
                else if (mode == 1) {
                    int dd = aux - j - 1;
                    if (dd > 399) dd = 399;
                    if (dd < 0) dd += SNN_NDIST;
                    d = dd;
                } else if (mode == 2) { table = mtable; d = aux; }
                const bool ok = d >= 0 && d < SNN_NDIST;
                const int dq = ok ? d : 0;
                const int rowl = (table == 0 ? 0 : 2) * SNN_NDIST + dq, rowr = rowl + SNN_NDIST;
                const int nl = __ldg(bg_valid + rowl), nr = __ldg(bg_valid + rowr);
                const double yl = __ldg(bg_rcp + rowl), yr = __ldg(bg_rcp + rowr);
                // lower bounds #(row < v).  Scattered loads cost the L1 one sector per lane each (ncu r02v:
                // 58 % of the kernel's L1 sectors were these searches when they were ten halving steps on
                // the float64 row), hence the key tree; the halving steps (probe position pa + step, never
                // past the row; entries past the valid ones are +inf, k_bg_sort) remain as its repeat path
                // and for rows too long for a tree.  Shared-memory copies of the first steps of the hot
                // rows were measured neutral (profiles/r02y) and are gone.
                const int basel = rowl * bg_ncol - 1, baser = rowr * bg_ncol - 1;
                int pa = 0, pb = 0, step = bg_half;
                if (bg_tree) {
                    // 5-ary descent over float32 keys rounded down, against the thresholds rounded up:
                    // one 16-byte load per level and search instead of one 8-byte load per halving
                    // step (10 -> 5 scattered loads for 1000 columns), then ONE float64 entry confirms
                    // the position; a lane whose entry is not below its threshold (an entry within a
                    // float32 ulp above it) repeats the search on the float64 row
                    // tree_dbg (SNN_SCORE_TREE=2, tests): thresholds of +inf send every lane through the repeat
                    const float ul = tree_dbg ? __int_as_float(0x7f800000) : __double2float_ru(dl);
                    const float ur = tree_dbg ? __int_as_float(0x7f800000) : __double2float_ru(dr);
                    const float4* __restrict__ ta = bg_tree + (int64_t)rowl * tree_nodes;
                    const float4* __restrict__ tb = bg_tree + (int64_t)rowr * tree_nodes;
                    int base = 0, cnt = 1;
#pragma unroll 1
                    for (int l = 0; l < tree_levels; ++l) {
                        const float4 ka = __ldg(ta + (base + pa)), kb = __ldg(tb + (base + pb));
                        pa = 5 * pa + (ka.x < ul) + (ka.y < ul) + (ka.z < ul) + (ka.w < ul);
                        pb = 5 * pb + (kb.x < ur) + (kb.y < ur) + (kb.z < ur) + (kb.w < ur);
                        base += cnt;
                        cnt *= 5;
                    }
                    const double xa = __ldg(bg + (basel + max(pa, 1))), xb = __ldg(bg + (baser + max(pb, 1)));
                    const bool bad = (pa > 0 && !(xa < dl)) || (pb > 0 && !(xb < dr));
                    step = 0;
                    if (__any_sync(FULL, bad)) {
                        step = bg_half;
                        if (bad) { pa = 0; pb = 0; }
                        else step = 0;
                    }
                }
#pragma unroll 1
                for (; step > 0; step >>= 1) {
                    const int ia = pa + step, ib = pb + step;
                    const double xa = __ldg(bg + (basel + min(ia, bg_ncol)));
                    const double xb = __ldg(bg + (baser + min(ib, bg_ncol)));
                    if (ia <= bg_ncol && xa < dl) pa = ia;
                    if (ib <= bg_ncol && xb < dr) pb = ib;
                }
                const int cl = (dl == dl) ? nl - pa : 0, cr = (dr == dr) ? nr - pb : 0;
                const double a = f_div((double)cl, (double)nl, yl);
                const double b = f_div((double)cr, (double)nr, yr);
                double pval = a;
                if (b > pval) pval = b;
                if (pval == 0.0) pval = pmin;
                if (!ok) pval = f_qnan();
                if (act) {
                    pv[j] = pval;
                    pnan |= (pval != pval);
                }
            }
            if (WS && act) {
                scA[j] = oc;
                slA[j] = ol;
                srA[j] = orr;
                tot_obs += rc;
            }
        }
        __syncwarp();

        // ================================ p-value: np.median over the rows ==========================
        if (WP) res0 = __any_sync(FULL, pnan) ? f_qnan() : f_median_any(pv, h, lane);

        if (WS) {
            // ---- dead columns (all-NaN inside the stripe's rows) and the finite-cell count, from the
            // NaN-weight bins: pixel (row, col) is NaN exactly when bin row or bin col is (band.cu)
            uint64_t deadC = 0, deadL = 0, deadR = 0;
            int n_obs = h * wc;
            const int32_t* __restrict__ npre = cd->nanpre;
            if (npre) {
                const int nr = __ldg(npre + R0 + h) - __ldg(npre + R0);
                const int nc = __ldg(npre + xsi + wc + bs) - __ldg(npre + C0);
                if (nr | nc) {
                    if (nr == h) {          // no finite cell at all
                        deadC = f_low64(wc);
                        deadL = deadR = f_low64(bs);
                        n_obs = 0;
                    } else {
                        const uint8_t* __restrict__ nb = cd->nanbin;
                        const unsigned l0 = __ballot_sync(FULL, lane < bs && __ldg(nb + C0 + (lane < bs ? lane : 0)));
                        const unsigned l1 = __ballot_sync(FULL, lane + 32 < bs && __ldg(nb + C0 + (lane + 32 < bs ? lane + 32 : 0)));
                        const unsigned c0 = __ballot_sync(FULL, lane < wc && __ldg(nb + xsi + (lane < wc ? lane : 0)));
                        const unsigned c1 = __ballot_sync(FULL, lane + 32 < wc && __ldg(nb + xsi + (lane + 32 < wc ? lane + 32 : 0)));
                        const int xr = xsi + wc;
                        const unsigned r0 = __ballot_sync(FULL, lane < bs && __ldg(nb + xr + (lane < bs ? lane : 0)));
                        const unsigned r1 = __ballot_sync(FULL, lane + 32 < bs && __ldg(nb + xr + (lane + 32 < bs ? lane + 32 : 0)));
                        deadL = ((uint64_t)l1 << 32) | l0;
                        deadC = ((uint64_t)c1 << 32) | c0;
                        deadR = ((uint64_t)r1 << 32) | r0;
                        n_obs = (h - nr) * (wc - __popcll(deadC));
                    }
                }
            }
            const int lc = wc - __popcll(deadC), ll = bs - __popcll(deadL), lr = bs - __popcll(deadR);
            bool bad = false;
            int hk = h;
            if (deadC | deadL | deadR) {
                // rows deleted with the dead columns: index x inside its part when xs == ys, else
                // shape[0]-1-x (getStripe.py:713-731; np.delete wraps negatives, IndexError -> NaN)
                unsigned rd[KC];
#pragma unroll
                for (int t = 0; t < KC; ++t) rd[t] = 0u;
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
                        for (int t = 0; t < KC; ++t)
                            if ((idx >> 5) == t) rd[t] |= 1u << (idx & 31);
                    }
                }
                if (!bad) {
                    int base = 0;
#pragma unroll
                    for (int t = 0; t < KC; ++t) {
                        if (t * 32 < h) {
                            const int j = t * 32 + lane;
                            const unsigned valid = (h - t * 32 >= 32) ? FULL : ((1u << (h - t * 32)) - 1u);
                            const unsigned kept = ~rd[t] & valid;
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
                    }
                    hk = base;
                }
            }
            if (!bad) {
                // row means of nan0(O/E) over the live columns
                const double ylc = rcp[lc], yll = rcp[ll], ylr = rcp[lr];
                const double dlc = (double)lc, dll = (double)ll, dlr = (double)lr;
                double tot_oe = 0.0;
#pragma unroll 1
                for (int k = lane; k < hk; k += 32) {
                    const double sc = scA[k], sl = slA[k], sr = srA[k];
                    tot_oe += sc;
                    scA[k] = lc ? f_div(sc, dlc, ylc) : f_qnan();
                    slA[k] = ll ? f_div(sl, dll, yll) : f_qnan();
                    srA[k] = lr ? f_div(sr, dlr, ylr) : f_qnan();
                }
                tot_oe = f_warp_sum(tot_oe);
                res2 = (hk > 0 && lc > 0) ? tot_oe / ((double)hk * (double)lc) : f_qnan();
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
                    const double gx = (gxl != gxl || gxr != gxr) ? f_qnan() : fmin(gxl, gxr);
                    const double d = gx - gy;
                    if (d == d) { acc += d; ++nacc; }
                }
                acc = f_warp_sum(acc);
                nacc = __reduce_add_sync(FULL, nacc);
                const double avgdiff = nacc ? acc / (double)nacc : f_qnan();
                // nanmedian(centre means): compact the non-NaN values into pv
                int nm = 0;
                __syncwarp();
#pragma unroll 1
                for (int k0 = 0; k0 < hk; k0 += 32) {
                    const int k = k0 + lane;
                    const double v = k < hk ? scA[k] : f_qnan();
                    const bool ok = v == v;
                    const unsigned b = __ballot_sync(FULL, ok);
                    if (ok) pv[nm + __popc(b & ((1u << lane) - 1u))] = v;
                    nm += __popc(b);
                }
                __syncwarp();
                const double med = nm ? f_median_any(pv, nm, lane) : f_qnan();
                res1 = med * avgdiff;
            }
            // getMean (getStripe.py:501-534)
            tot_obs = f_warp_sum(tot_obs);
            res4 = n_obs ? tot_obs / (double)n_obs : f_qnan();
            res5 = tot_obs;
        }
        if (lane < 6) {
            const double v = lane == 0 ? res0 : lane == 1 ? res1 : lane == 2 ? res2 : lane == 3 ? res3
                           : lane == 4 ? res4 : res5;
            out[s * 6 + lane] = v;
        }
    }
}

int snn_launch_score_fast(snn_ctx* ctx, const ChromDev* cds, const FastRec* recs, int n, int bs, int want_p, int want_s,
                          int* next, double* tall, int grid, double* out) {
    if (n <= 0) return SNN_OK;
    const size_t smem = (72 + (size_t)SF_WARPS * 4 * SF_HSM) * sizeof(double);
    int bg_half = 1;
    while (bg_half * 2 <= ctx->bg_ncol) bg_half <<= 1;     // largest power of two <= ncol: positions up to 2*half - 1 >= ncol
    const double pmin = ctx->bg_ncol > 0 ? 1.0 / (double)ctx->bg_ncol : 0.0;
    const char* tenv = getenv("SNN_SCORE_TREE");          // A/B switch and test hook: 0 = float64 rows only
    const float4* tree = (tenv && tenv[0] == '0') ? nullptr : ctx->bg_tree;
    const int tree_dbg = (tenv && tenv[0] == '2') ? 1 : 0;
    int tree_nodes = 0;
    for (int l = 0, c = 1; l < ctx->bg_tree_levels; ++l, c *= 5) tree_nodes += c;
#define SF_LAUNCH(WP, WS)                                                                                          \
    do {                                                                                                           \
        k_score_fast<WP, WS><<<grid, SF_WARPS * 32, smem, ctx->stream>>>(cds, recs, n, next, ctx->bg_sorted,       \
                                                                         ctx->bg_valid, ctx->bg_rcp,               \
                                                                         (int)ctx->bg_ncol, bg_half, tree,         \
                                                                         ctx->bg_tree_levels, tree_nodes,          \
                                                                         tree_dbg, pmin,                           \
                                                                         bs, tall, out);                           \
    } while (0)
    if (want_p && want_s) SF_LAUNCH(true, true);
    else if (want_p) SF_LAUNCH(true, false);
    else SF_LAUNCH(false, true);
#undef SF_LAUNCH
    SNN_LAUNCH_CHECK("k_score_fast");
    return SNN_OK;
}

// Host-side preparation kept out of the timed scope: the work counter and the shared-memory opt-in
// (one constant value per kernel: contexts on several host threads share the functions)
int snn_prepare_score_fast(snn_ctx* ctx, int* next) {
    const int smem = (int)((72 + (size_t)SF_WARPS * 4 * SF_HSM) * sizeof(double));
    SNN_CUDA(cudaFuncSetAttribute(k_score_fast<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    SNN_CUDA(cudaFuncSetAttribute(k_score_fast<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    SNN_CUDA(cudaFuncSetAttribute(k_score_fast<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    SNN_CUDA(cudaMemsetAsync(next, 0, sizeof(int), ctx->stream));
    return SNN_OK;
}

// CTAs of the persistent grid (the host sizes the `tall` scratch from it)
int snn_score_fast_grid(int n) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int grid = (n + SF_WARPS - 1) / SF_WARPS;
    if (grid > sms * SF_MINB) grid = sms * SF_MINB;
    return grid < 1 ? 1 : grid;
}
