// Hysteresis, verticalLine, block scan, segment refinement, grouping, pairing and candidate
// emission of getStripe.StripeSearch, one CTA per brightness pass with
// the whole 400 x 400 window held as bit masks in shared memory (4 masks x 20.8 KB):
//   E  edges  : low mask -> after hysteresis the Canny edge map (skimage canny, ndi.label)
//   C  current: high mask, grown to the connected components of E that contain a high pixel
//   V  vert   : ImageProcessing.verticalLine(edges, 60, 120)            (ImageProcessing.py:61-83)
//   T  testmat: painted candidate segments, refined per row               (getStripe.py:945-978)
// C is dead once the hysteresis has converged (E := C), so T lives in C's storage: three masks =
// 62.4 KB of dynamic shared memory, 11 KB static, 48 registers x 416 threads -> 3 CTAs per SM
// (four masks and 28-byte pending records were 99 KB: 2 CTAs per SM, 26 warps, issue 40 %: profiles/r01p).
// All of this is integer / bit work; the only floating point is the count-weighted mean
// column of the grouping step (float64, sequential, round-half-even).
#include <string.h>

#include <algorithm>

#include "image.cuh"

#define NTHR 416

__device__ __forceinline__ int getbit(const uint32_t* M, int S, int r, int c) {
    if (r < 0 || c < 0 || r >= S || c >= S) return 0;
    return (M[r * SNN_WPR + (c >> 5)] >> (c & 31)) & 1;
}

// bits [s, e) of a 13-word row: apply dst = (dst & ~mask) | (src & mask)
__device__ __forceinline__ void row_assign(uint32_t* dst, const uint32_t* src, int s, int e) {
    if (e <= s) return;
    for (int w = s >> 5; w <= (e - 1) >> 5; ++w) {
        int lo = max(s, w * 32) - w * 32, hi = min(e, w * 32 + 32) - w * 32;  // [lo, hi) within the word
        uint32_t mask = (hi == 32 ? 0xffffffffu : ((1u << hi) - 1u)) & ~((1u << lo) - 1u);
        uint32_t sv = src ? src[w] : 0u;
        dst[w] = (dst[w] & ~mask) | (sv & mask);
    }
}

__device__ __forceinline__ bool row_any(const uint32_t* row, int s, int e) {  // any bit in [s, e)
    if (e <= s) return false;
    for (int w = s >> 5; w <= (e - 1) >> 5; ++w) {
        int lo = max(s, w * 32) - w * 32, hi = min(e, w * 32 + 32) - w * 32;
        uint32_t mask = (hi == 32 ? 0xffffffffu : ((1u << hi) - 1u)) & ~((1u << lo) - 1u);
        if (row[w] & mask) return true;
    }
    return false;
}

// next set bit at position >= p in a 13-word row, or -1
__device__ __forceinline__ int next_set(const uint32_t* row, int p, int S) {
    while (p < S) {
        uint32_t w = row[p >> 5] >> (p & 31);
        if (w) {
            int r = p + __ffs(w) - 1;
            return r < S ? r : -1;
        }
        p = (p | 31) + 1;
    }
    return -1;
}

// next clear bit at position >= p (returns S if none below S)
__device__ __forceinline__ int next_clear(const uint32_t* row, int p, int S) {
    while (p < S) {
        uint32_t w = (~row[p >> 5]) >> (p & 31);
        if (w) {
            int r = p + __ffs(w) - 1;
            return r < S ? r : S;
        }
        p = (p | 31) + 1;
    }
    return S;
}

// r = a + b over 13 words with carry (one 416-bit addition)
__device__ __forceinline__ void add416(const uint32_t (&a)[SNN_WPR], const uint32_t (&b)[SNN_WPR],
                                       uint32_t (&r)[SNN_WPR]) {
    uint32_t c;
    asm("add.cc.u32 %0, %8, %15;\n\t"
        "addc.cc.u32 %1, %9, %16;\n\t"
        "addc.cc.u32 %2, %10, %17;\n\t"
        "addc.cc.u32 %3, %11, %18;\n\t"
        "addc.cc.u32 %4, %12, %19;\n\t"
        "addc.cc.u32 %5, %13, %20;\n\t"
        "addc.cc.u32 %6, %14, %21;\n\t"
        "addc.u32 %7, 0, 0;"
        : "=&r"(r[0]), "=&r"(r[1]), "=&r"(r[2]), "=&r"(r[3]), "=&r"(r[4]), "=&r"(r[5]), "=&r"(r[6]), "=&r"(c)
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(b[0]), "r"(b[1]),
          "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]));
    asm("add.cc.u32 %0, %6, 0xffffffff;\n\t"      // carry flag := c
        "addc.cc.u32 %0, %7, %13;\n\t"
        "addc.cc.u32 %1, %8, %14;\n\t"
        "addc.cc.u32 %2, %9, %15;\n\t"
        "addc.cc.u32 %3, %10, %16;\n\t"
        "addc.cc.u32 %4, %11, %17;\n\t"
        "addc.u32 %5, %12, %18;"
        : "=&r"(r[7]), "=&r"(r[8]), "=&r"(r[9]), "=&r"(r[10]), "=&r"(r[11]), "=&r"(r[12])
        : "r"(c), "r"(a[7]), "r"(a[8]), "r"(a[9]), "r"(a[10]), "r"(a[11]), "r"(a[12]), "r"(b[7]), "r"(b[8]),
          "r"(b[9]), "r"(b[10]), "r"(b[11]), "r"(b[12]));
}

// Flood the seed bits (a subset of m) through the runs of ones of m, both directions, over a
// 13-word row.  Adding a seed to a run of ones ripples a carry to the run's end: the bits that
// differ from m are exactly the part of the run from the seed upwards.  The downward half is
// the same addition on the bit-reversed row.
__device__ __forceinline__ void flood_row(const uint32_t (&m)[SNN_WPR], uint32_t (&g)[SNN_WPR]) {
    uint32_t t[SNN_WPR], mr[SNN_WPR], gr[SNN_WPR];
    add416(m, g, t);
#pragma unroll
    for (int w = 0; w < SNN_WPR; ++w) g[w] = (((t[w] ^ m[w]) & m[w]) | g[w]);
#pragma unroll
    for (int w = 0; w < SNN_WPR; ++w) {
        mr[w] = __brev(m[SNN_WPR - 1 - w]);
        gr[w] = __brev(g[SNN_WPR - 1 - w]);
    }
    add416(mr, gr, t);
#pragma unroll
    for (int w = 0; w < SNN_WPR; ++w) gr[w] = (((t[w] ^ mr[w]) & mr[w]) | gr[w]);
#pragma unroll
    for (int w = 0; w < SNN_WPR; ++w) g[w] = __brev(gr[SNN_WPR - 1 - w]);
}

// The run automaton of ImageProcessing.block (ImageProcessing.py:127-138) fed with gaps instead
// of single rows: `g` zeros in a row.  Zeros fill a 5-slot buffer that ones do not reset; the
// 6th zero closes the run (keeping the longest) and empties the buffer, and so on every 6 zeros.
struct BlockScan {
    int count, buf, last, best, end;
    __device__ __forceinline__ void zeros(int g) {
        if (g <= 0) return;
        if (buf + g <= 5) { buf += g; return; }
        if (count > best) { best = count; end = last; }
        count = 0;
        const int rem = g - (5 - buf) - 1;
        buf = rem % 6;
    }
    // bits of `wd` = rows r0 .. r0+n-1 of the column (bit 0 = row r0)
    __device__ __forceinline__ void word(uint32_t wd, int r0, int n) {
        int pos = 0;
        while (wd) {
            const int b = __ffs(wd) - 1;
            zeros(b - pos);
            const uint32_t rest = wd >> b;                       // starts with a one
            const int n1 = (rest == 0xffffffffu) ? 32 : __ffs(~rest) - 1;   // length of this run of ones
            count += n1;
            last = r0 + b + n1 - 1;
            pos = b + n1;
            wd = (pos >= 32) ? 0u : (wd & ~((1u << pos) - 1u));
        }
        zeros(n - pos);
    }
};

struct SearchParams {
    int nq, minH, maxW;
    int raw_cap;
};

// candidate kept in shared memory until the pass is flushed: 16 bytes (x, right, k fit 16 bits; y and hi
// may hold the 'no row hit' sentinels 1 << 30 and -1)
struct PendCand {
    int y, hi;
    short x, right, k;
    unsigned char ud, pad;
};

#ifdef SEARCH_FOUR_MASKS          // A/B build switch: the 4-mask layout, 2 CTAs per SM
#define SEARCH_MINB 2
#else
#define SEARCH_MINB 3
#endif

__global__ void __launch_bounds__(NTHR, SEARCH_MINB) k_search(const uint32_t* __restrict__ low_all,
                                                 const uint32_t* __restrict__ high_all,
                                                 const FrameInfo* __restrict__ info, SearchParams P,
                                                 int* __restrict__ counter, RawCand* __restrict__ raw,
                                                 uint32_t* __restrict__ dbg_edges, uint32_t* __restrict__ dbg_vert) {
    extern __shared__ uint32_t smem[];
    uint32_t* E = smem;
    uint32_t* C = E + SNN_SMAX * SNN_WPR;
    uint32_t* V = C + SNN_SMAX * SNN_WPR;
#ifdef SEARCH_FOUR_MASKS
    uint32_t* T = V + SNN_SMAX * SNN_WPR;
#else
    uint32_t* T = C;                    // C's storage, after the hysteresis
#endif
    __shared__ uint32_t colV[SNN_WPR], colT[SNN_WPR], colOK[SNN_WPR];   // per-column any-bit / >= 3 ones masks
    __shared__ short cEnd[SNN_SMAX];
    __shared__ unsigned char cUd[SNN_SMAX];
    __shared__ short cnt[SNN_SMAX];
    __shared__ int X[SNN_SMAX + 8];
    __shared__ int nX;
    __shared__ int rmin, rmax;
    __shared__ PendCand pend[SNN_SMAX];
    __shared__ int npend;
    __shared__ int base_sh;
    __shared__ unsigned char dirty[2][SNN_SMAX + 2];   // row r at index r + 1

    const int pass = blockIdx.x;
    const int S = info[pass / (P.nq * SNN_NBRIGHT)].S;
    if (S == 0) return;
    const int tid = threadIdx.x;
    const uint32_t* low = low_all + (int64_t)pass * SNN_SMAX * SNN_WPR;
    const uint32_t* high = high_all + (int64_t)pass * SNN_SMAX * SNN_WPR;
    for (int t = tid; t < S * SNN_WPR; t += NTHR) {
        E[t] = low[t];
        C[t] = high[t];
    }
    if (tid < SNN_WPR) colV[tid] = 0;
    if (tid == 0) npend = 0;
    __syncthreads();

    // ---- hysteresis: grow C inside E to the fixed point (8-connectivity).  A thread owns one
    // row: 3x3 dilation from the rows above/below, then a full horizontal flood inside the row
    // (two 416-bit additions, see flood_row).  A row is recomputed only if itself or a vertical
    // neighbour changed in the previous sweep.
    {
        const int r = tid;
        if (tid < SNN_SMAX + 2) { dirty[0][tid] = (tid >= 1 && tid <= S) ? 1 : 0; dirty[1][tid] = 0; }
        __syncthreads();
        int cur = 0;
        for (;;) {
            bool changed = false;
            uint32_t nw[SNN_WPR];
            if (r < S && (dirty[cur][r] | dirty[cur][r + 1] | dirty[cur][r + 2])) {
                const uint32_t* cu = C + (r - 1) * SNN_WPR;
                const uint32_t* cc = C + r * SNN_WPR;
                const uint32_t* cd = C + (r + 1) * SNN_WPR;
                const bool hasu = r > 0, hasd = r < S - 1;
                uint32_t v[SNN_WPR], er[SNN_WPR], own[SNN_WPR];
#pragma unroll
                for (int w = 0; w < SNN_WPR; ++w) {
                    own[w] = cc[w];
                    er[w] = E[r * SNN_WPR + w];
                    v[w] = own[w] | (hasu ? cu[w] : 0u) | (hasd ? cd[w] : 0u);
                }
#pragma unroll
                for (int w = 0; w < SNN_WPR; ++w) {
                    uint32_t sp = v[w] | (v[w] << 1) | (v[w] >> 1);
                    if (w > 0) sp |= v[w - 1] >> 31;
                    if (w < SNN_WPR - 1) sp |= v[w + 1] << 31;
                    nw[w] = (sp & er[w]) | own[w];
                }
                flood_row(er, nw);
#pragma unroll
                for (int w = 0; w < SNN_WPR; ++w) changed |= (nw[w] != own[w]);
            }
            __syncthreads();  // all reads of C done
            if (changed) {
#pragma unroll
                for (int w = 0; w < SNN_WPR; ++w) C[r * SNN_WPR + w] = nw[w];
            }
            if (r < S) dirty[cur ^ 1][r + 1] = changed ? 1 : 0;
            cur ^= 1;
            if (!__syncthreads_or(changed)) break;
        }
    }
    // E := edges
    for (int t = tid; t < S * SNN_WPR; t += NTHR) E[t] = C[t];
    __syncthreads();
    if (dbg_edges)
        for (int t = tid; t < S * SNN_WPR; t += NTHR) dbg_edges[(int64_t)pass * SNN_SMAX * SNN_WPR + t] = E[t];

    // ---- verticalLine: hit(x, y) = Fx > 0 && 3 Fy^2 < Fx^2 ; vert[x][y-1] = hit[x][y].
    // vert[r][c] can only be set where column c has an edge pixel in rows r-1..r+1; for those
    // columns the 3x3 neighbourhood (columns c..c+2) is cut out of 64-bit windows of the rows.
    if (tid < S) {
        const int r = tid;
        const bool hasu = r > 0, hasd = r < S - 1;
        const uint32_t* eu = E + (r - 1) * SNN_WPR;
        const uint32_t* ec = E + r * SNN_WPR;
        const uint32_t* ed = E + (r + 1) * SNN_WPR;
        uint32_t a_lo = hasu ? eu[0] : 0u, b_lo = ec[0], c_lo = hasd ? ed[0] : 0u;
#pragma unroll
        for (int w = 0; w < SNN_WPR; ++w) {
            uint32_t a_hi = 0u, b_hi = 0u, c_hi = 0u;
            if (w + 1 < SNN_WPR) {
                a_hi = hasu ? eu[w + 1] : 0u;
                b_hi = ec[w + 1];
                c_hi = hasd ? ed[w + 1] : 0u;
            }
            uint32_t d = a_lo | b_lo | c_lo;
            uint32_t out = 0;
            if (d) {
                const uint64_t A = ((uint64_t)a_hi << 32) | a_lo, B = ((uint64_t)b_hi << 32) | b_lo,
                               Cc = ((uint64_t)c_hi << 32) | c_lo;
                while (d) {
                    const int bit = __ffs(d) - 1;
                    d &= d - 1;
                    if (w * 32 + bit > S - 2) break;          // bits ascend: nothing further qualifies
                    const int ta = (int)(A >> bit) & 7, tb = (int)(B >> bit) & 7, tc = (int)(Cc >> bit) & 7;
                    // [1 2 1] sums: columns c / c+2 over the three rows, rows r-1 / r+1 over the three columns
                    const int col0 = (ta & 1) + 2 * (tb & 1) + (tc & 1);
                    const int col2 = (ta >> 2) + 2 * (tb >> 2) + (tc >> 2);
                    const int rowu = (ta & 1) + 2 * ((ta >> 1) & 1) + (ta >> 2);
                    const int rowd = (tc & 1) + 2 * ((tc >> 1) & 1) + (tc >> 2);
                    const int fx = col0 - col2, fy = rowd - rowu;
                    if (fx > 0 && 3 * fy * fy < fx * fx) out |= 1u << bit;
                }
            }
            V[r * SNN_WPR + w] = out;
            if (out) atomicOr(&colV[w], out);
            a_lo = a_hi; b_lo = b_hi; c_lo = c_hi;
        }
    }
    __syncthreads();
    if (dbg_vert)
        for (int t = tid; t < S * SNN_WPR; t += NTHR) dbg_vert[(int64_t)pass * SNN_SMAX * SNN_WPR + t] = V[t];

    // ---- block(vert, c) for every column + candidate test (ImageProcessing.py:100-195,
    // getStripe.py:924-940).  A warp owns 32 columns: per block of 32 rows every lane loads one
    // row's word of (vert[:, c-1] | vert[:, c] | vert[:, c+1]), 32 ballots transpose it so that
    // lane b holds column b's 32 rows, and the run automaton consumes it gap by gap.
    {
        const int w = tid >> 5, lane = tid & 31;        // word index of this warp's columns
        const int c = tid;
        if (w * 32 < S) {
            const uint32_t near = colV[w] | (w > 0 ? colV[w - 1] >> 31 : 0u) |
                                  (w + 1 < SNN_WPR ? colV[w + 1] << 31 : 0u);
            BlockScan bsn;
            bsn.count = bsn.buf = bsn.last = bsn.best = bsn.end = 0;
            if (near | (near << 1) | (near >> 1)) {      // some column of this warp sees a one
                for (int r0 = 0; r0 < S; r0 += 32) {
                    const int rr = r0 + lane;
                    uint32_t v3 = 0u;
                    if (rr < S) {
                        const uint32_t* vr = V + rr * SNN_WPR;
                        const uint32_t v = vr[w];
                        v3 = v | (v << 1) | (v >> 1);
                        if (w > 0) v3 |= vr[w - 1] >> 31;
                        if (w + 1 < SNN_WPR) v3 |= vr[w + 1] << 31;
                    }
                    uint32_t mine = 0u;
                    if (__any_sync(0xffffffffu, v3 != 0u)) {
#pragma unroll
                        for (int b = 0; b < 32; ++b) {
                            const uint32_t bal = __ballot_sync(0xffffffffu, (v3 >> b) & 1u);
                            if (lane == b) mine = bal;
                        }
                    }
                    bsn.word(mine, r0, min(32, S - r0));
                }
            } else {
                bsn.zeros(S);
            }
            if (bsn.count > bsn.best) { bsn.best = bsn.count; bsn.end = bsn.last; }
            if (c < S) {
                int best = bsn.best, end = bsn.end;
                if (end < c) end = end - best + 1;
                int ud = 0;
                if (best > P.minH) {
                    int lo = min(c, end), hi = max(c, end);
                    bool has = false;
                    for (int i = max(lo, 0); i <= hi && i < S; ++i)
                        if (getbit(V, S, i, c)) { has = true; break; }
                    if (has) ud = (end > c) ? 2 : 1;
                }
                cEnd[c] = (short)end;
                cUd[c] = (unsigned char)ud;
            }
        }
    }
    __syncthreads();

    for (int ud = 1; ud <= 2; ++ud) {
        for (int t = tid; t < S * SNN_WPR; t += NTHR) T[t] = 0;
        if (tid < SNN_WPR) { colT[tid] = 0; colOK[tid] = 0; }
        if (tid == 0) nX = 0;
        __syncthreads();
        // paint the candidate segments (half-open row ranges)
        if (tid < S && cUd[tid] == ud) {
            const int c = tid, e = cEnd[c];
            int st = (ud == 1) ? e : c, en = (ud == 1) ? c : e;
            if (st < 0) st = 0;
            for (int r = st; r < en; ++r) atomicOr(&T[r * SNN_WPR + (c >> 5)], 1u << (c & 31));
        }
        __syncthreads();
        // per-row refinement of every horizontal run of ones (getStripe.py:957-978)
        if (tid < S) {
            const int r = tid;
            uint32_t* tr = T + r * SNN_WPR;
            uint32_t orig[SNN_WPR];
            uint32_t anyrow = 0;
            for (int w = 0; w < SNN_WPR; ++w) { orig[w] = tr[w]; anyrow |= orig[w]; }
            if (anyrow) {
                int p = 0;
                for (;;) {
                    int s = next_set(orig, p, S);
                    if (s < 0) break;
                    int e = next_clear(orig, s, S) - 1;  // inclusive end of the run
                    if (row_any(E + r * SNN_WPR, s, e + 1)) {
                        row_assign(tr, V + r * SNN_WPR, s, e);  // testmat[r, s:e] = vert[r, s:e]
                    } else {
                        row_assign(tr, nullptr, s, e);          // testmat[r, s:e] = 0
                        int sum = s + e;
                        int med = sum >> 1;
                        if ((sum & 1) && (med & 1)) med += 1;   // np.round((s+e)/2): half to even
                        tr[med >> 5] |= 1u << (med & 31);
                    }
                    p = e + 1;
                }
                for (int w = 0; w < SNN_WPR; ++w)
                    if (tr[w]) atomicOr(&colT[w], tr[w]);
            }
        }
        __syncthreads();
        // ones per column: the same ballot transpose, popcount instead of the automaton
        {
            const int w = tid >> 5, lane = tid & 31;
            if (w * 32 < S) {
                int n = 0;
                if (colT[w]) {
                    for (int r0 = 0; r0 < S; r0 += 32) {
                        const int rr = r0 + lane;
                        const uint32_t tw = rr < S ? T[rr * SNN_WPR + w] : 0u;
                        if (__any_sync(0xffffffffu, tw != 0u)) {
#pragma unroll
                            for (int b = 0; b < 32; ++b) {
                                const uint32_t bal = __ballot_sync(0xffffffffu, (tw >> b) & 1u);
                                if (lane == b) n += __popc(bal);
                            }
                        }
                    }
                }
                if (tid < S) cnt[tid] = (short)n;
                const uint32_t okm = __ballot_sync(0xffffffffu, tid < S && n >= 3);
                if (lane == 0) colOK[w] = okm;
            }
        }
        __syncthreads();
        // grouping state machine (getStripe.py:996-1030), sequential
        if (tid == 0) {
            int n = 0;
            // columns with >= 3 ones, ascending; walk consecutive pairs
            int prev = -1, prevlen = 0;
            bool open_ = false;
            bool have_ln = false;
            int gcol[SNN_SMAX];
            int glen[SNN_SMAX];
            int gn = 0;
            bool unit_len = false;  // the single-column case stores ln = [1.0]
            int nx = 0;
            auto flush = [&](void) {
                // fr = [a / tot] ; round(sum(c * fr))  in float64, sequential, half-even
                double total = 0.0;
                if (unit_len) total = 1.0;
                else {
                    long long ti = 0;
                    for (int k = 0; k < gn; ++k) ti += glen[k];
                    total = (double)ti;
                }
                double acc = 0.0;
                for (int k = 0; k < gn; ++k) {
                    double fr = (unit_len ? 1.0 : (double)glen[k]) / total;
                    double term = (double)gcol[k] * fr;
                    acc = (k == 0) ? term : acc + term;
                }
                double rv = rint(acc);
                X[nx++] = (int)rv;
            };
            for (int wq = 0; wq < SNN_WPR; ++wq) {      // columns with >= 3 ones, ascending
                uint32_t okm = colOK[wq];
                while (okm) {
                    const int c = wq * 32 + __ffs(okm) - 1;
                    okm &= okm - 1;
                    ++n;
                    if (prev >= 0) {
                        // one step of the loop `for k in range(n-1)` with cur = prev, nxt = c
                        const bool adjacent = (c - prev) == 1;
                        if (adjacent && open_) {
                            gcol[gn] = c; glen[gn] = cnt[c]; ++gn;
                        } else if (adjacent) {
                            gcol[0] = prev; glen[0] = prevlen; gcol[1] = c; glen[1] = cnt[c]; gn = 2;
                            unit_len = false; open_ = true; have_ln = true;
                        } else if (!open_) {
                            gcol[0] = prev; glen[0] = 1; gn = 1; unit_len = true; have_ln = true;
                            flush();
                        } else {
                            flush();
                            gcol[0] = prev; glen[0] = prevlen; gn = 1; unit_len = false; open_ = false;
                            have_ln = true;
                        }
                    }
                    prev = c;
                    prevlen = cnt[c];
            
                }
            }
            if (!have_ln) X[nx++] = 0;
            else flush();
            // sorted(set(...))
            for (int a = 1; a < nx; ++a) {
                int v = X[a], b = a - 1;
                while (b >= 0 && X[b] > v) { X[b + 1] = X[b]; --b; }
                X[b + 1] = v;
            }
            int u = 0;
            for (int a = 0; a < nx; ++a)
                if (a == 0 || X[a] != X[u - 1]) X[u++] = X[a];
            nX = u;
        }
        __syncthreads();
        // pairing (getStripe.py:1034-1078)
        const int nx = nX;
        int kout = 0;
        for (int k = 0; k + 1 < nx; ++k) {
            const int n = X[k], m = X[k + 1];
            const int gap = m - n;
            if (!(gap > 1 && gap <= P.maxW)) continue;  // uniform across the CTA
            if (tid == 0) { rmin = 1 << 30; rmax = -1; }
            __syncthreads();
            if (tid < S) {
                const uint32_t* tr = T + tid * SNN_WPR;
                bool hit = row_any(tr, max(0, n - 1), min(n + 2, S)) || row_any(tr, max(0, m - 1), min(m + 2, S));
                if (hit) { atomicMin(&rmin, tid); atomicMax(&rmax, tid); }
            }
            __syncthreads();
            if (tid == 0) {
                const int right = (gap > 4) ? m - 2 : m;
                int lo = rmin, hi = rmax;
                if (ud == 1) hi = right; else lo = n;
                if (npend < SNN_SMAX) {
                    PendCand pc;
                    pc.ud = (unsigned char)ud; pc.pad = 0; pc.k = (short)kout;
                    pc.x = (short)n; pc.y = lo; pc.right = (short)right; pc.hi = hi;
                    pend[npend++] = pc;
                }
            }
            ++kout;
            __syncthreads();
        }
        __syncthreads();
    }
    // ---- flush the candidates of this pass
    if (tid == 0) base_sh = npend ? atomicAdd(&counter[0], npend) : 0;
    __syncthreads();
    const int base = base_sh;
    for (int t = tid; t < npend; t += NTHR) {
        if (base + t < P.raw_cap) {
            const PendCand pc = pend[t];
            RawCand rc;
            rc.pass = pass; rc.ud = pc.ud; rc.k = pc.k;
            rc.x = pc.x; rc.y = pc.y; rc.right = pc.right; rc.hi = pc.hi;
            raw[base + t] = rc;
        } else counter[1] = 1;
    }
}

// ---- finish: bp coordinates and the window sum of each candidate (getStripe.py:1087-1095) --------
// The raw candidates of the current batch are entries [counter[2], counter[0]) of `raw` (k_search
// appends, k_advance moves counter[2] up after this kernel): no host round trip between batches.
__global__ void k_finish(const RawCand* __restrict__ raw, const int* __restrict__ counter, int cap,
                         const FrameInfo* __restrict__ info, const int32_t* __restrict__ nz_all,
                         const double* __restrict__ Dc_all, int nq, int64_t binsize, int64_t chromsize,
                         snn_candidate* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int hi = min(counter[0], cap);
    // grid-stride over the batch's candidates [counter[2], counter[0]): one batch may append up to
    // raw_cap of them, more than the launch has warps
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int warp = counter[2] + ((blockIdx.x * blockDim.x + threadIdx.x) >> 5); warp < hi; warp += nwarps) {
    const RawCand rc = raw[warp];
    const int f = rc.pass / (nq * SNN_NBRIGHT);
    const FrameInfo fi = info[f];
    const int32_t* nz = nz_all + (int64_t)f * SNN_SMAX;
    const double* Dc = Dc_all + (int64_t)f * SNN_SMAX * SNN_SMAX;
    const int w = rc.right - rc.x + 1, h = rc.hi - rc.y + 1;
    double s = 0.0;
    if (h > 0 && w > 0) {
        for (int t = lane; t < h * w; t += 32) {
            int r = rc.y + t / w, c = rc.x + t % w;
            if (r >= 0 && r < fi.S && c >= 0 && c < fi.S) s += Dc[r * SNN_SMAX + c];
        }
    }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if (lane == 0) {
        snn_candidate c;
        c.q = (rc.pass / SNN_NBRIGHT) % nq;
        c.frame = fi.idx;
        c.bright = rc.pass % SNN_NBRIGHT;
        c.ud = rc.ud;
        c.k = rc.k;
        c.x = rc.x; c.y = rc.y; c.w = w; c.h = h;
        c.start = fi.start; c.end = fi.end; c.S = fi.S;
        auto sa = [&](int i) -> int64_t {  // start_array[i]
            int ii = min(max(i, 0), fi.S - 1);
            return ((int64_t)fi.start + nz[ii]) * binsize + 1;
        };
        auto ea = [&](int i) -> int64_t {  // end_array[i]; only the window's last bin is clipped
            int ii = min(max(i, 0), fi.S - 1);
            int64_t g = (int64_t)fi.start + nz[ii];
            int64_t e = (g + 1) * binsize;
            if (g == fi.end && e >= chromsize) e = chromsize;
            return e;
        };
        c.pos1 = sa(rc.x);
        c.pos2 = ea(rc.right);
        c.pos3 = sa(rc.y);
        c.pos4 = ea(rc.hi);
        c.total = s;
        out[warp] = c;
    }
    }
}

__global__ void k_advance(int* counter) { counter[2] = counter[0]; }

// ---- host side ------------------------------------------------------------------------------------------
static size_t align_up(size_t v) { return (v + 255) & ~(size_t)255; }

static int make_plan(snn_ctx* ctx, ExtractPlan* pl, int nq, const double* mp, const double* gw, int radius,
                     int minH, int maxW, int bfilter, int variant, int fb, int total_frames, bool dbg,
                     uint32_t** dbgE, uint32_t** dbgV, float** dbgM) {
    pl->nq = nq;
    pl->minH = minH;
    pl->maxW = maxW;
    pl->fb = fb;
    GrayParams& g = pl->gray;
    g.nq = nq;
    g.bf = bfilter;
    g.variant = variant;
    // np.ones((bf,bf)) / (bf*bf)
    g.kv = 1.0 / (double)(bfilter * bfilter);
    {   // what the constant red channel becomes: sequential sum of bf*bf taps, clip, float32
        volatile double s = 0.0;
        for (int t = 0; t < bfilter * bfilter; ++t) {
            volatile double term = g.kv * 1.0;
            s = s + term;
        }
        double cl = s < 0.0 ? 0.0 : (s > 1.0 ? 1.0 : s);
        g.r32 = (float)cl;
    }
    {   // np.arange(0.5, 1.01, 0.1): start + i * delta with delta = (start + step) - start
        volatile double first = 0.5, second = 0.5 + 0.1;
        volatile double delta = second - first;
        for (int i = 0; i < SNN_NBRIGHT; ++i) {
            volatile double t = (double)i * delta;
            g.bright[i] = first + t;
            volatile double num = 1.0 - 0.0, den = g.bright[i] - 0.0;
            g.kb[i] = num / den;
        }
    }
    for (int i = 0; i < nq; ++i) g.mp[i] = mp[i];
    pl->gauss.radius = radius;
    for (int i = 0; i <= radius; ++i) pl->gauss.w[i] = gw[i];

    const size_t npass = (size_t)fb * nq * SNN_NBRIGHT;
    const size_t img = (size_t)SNN_SMAX * SNN_SMAX;
    pl->batch_cap = (int)std::min<size_t>(npass * 24 + 4096, (size_t)16 << 20);   // candidates of one batch
    pl->raw_cap = (int)std::min<size_t>((size_t)total_frames * nq * SNN_NBRIGHT * 24 + 4096, (size_t)16 << 20);
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes); return o; };
    size_t o_info = take(fb * sizeof(FrameInfo));
    size_t o_nz = take((size_t)fb * SNN_SMAX * sizeof(int32_t));
    size_t o_flags = take((size_t)fb * SNN_SMAX * sizeof(int));
    size_t o_Dc = take((size_t)fb * img * sizeof(double));
    size_t o_bleed = take((size_t)fb * img * sizeof(float));
    size_t o_A = take(npass * img * sizeof(float));
    size_t o_B = take(npass * img * sizeof(float));
    size_t o_C = take(npass * img * sizeof(float));
    size_t o_low = take(npass * SNN_SMAX * SNN_WPR * sizeof(uint32_t));
    size_t o_high = take(npass * SNN_SMAX * SNN_WPR * sizeof(uint32_t));
    size_t o_cnt = take(64);
    size_t o_raw = take((size_t)pl->raw_cap * sizeof(RawCand));
    size_t o_cand = take((size_t)pl->raw_cap * sizeof(snn_candidate));
    size_t o_dE = 0, o_dV = 0, o_dM = 0;
    if (dbg) {
        o_dE = take(npass * SNN_SMAX * SNN_WPR * sizeof(uint32_t));
        o_dV = take(npass * SNN_SMAX * SNN_WPR * sizeof(uint32_t));
        o_dM = take(npass * img * sizeof(float));
    }
    void* base = nullptr;
    int rc = snn_scratch(ctx, off, &base);
    if (rc) return rc;
    char* b = (char*)base;
    pl->info = (FrameInfo*)(b + o_info);
    pl->nz = (int32_t*)(b + o_nz);
    pl->flags = (int*)(b + o_flags);
    pl->Dc = (double*)(b + o_Dc);
    pl->bleed = (float*)(b + o_bleed);
    pl->bufA = (float*)(b + o_A);
    pl->bufB = (float*)(b + o_B);
    pl->bufC = (float*)(b + o_C);
    pl->low = (uint32_t*)(b + o_low);
    pl->high = (uint32_t*)(b + o_high);
    pl->counter = (int*)(b + o_cnt);
    pl->raw = (RawCand*)(b + o_raw);
    pl->cand = (snn_candidate*)(b + o_cand);
    rc = snn_fused_make_tmap(ctx, pl, npass);
    if (rc) pl->tmap_ok = false;      // the unfused kernels take over (snn_fused_supported)
    if (dbg) {
        *dbgE = (uint32_t*)(b + o_dE);
        *dbgV = (uint32_t*)(b + o_dV);
        *dbgM = (float*)(b + o_dM);
    }
    return SNN_OK;
}

#ifdef SEARCH_FOUR_MASKS
static const size_t kSearchSmem = 4 * (size_t)SNN_SMAX * SNN_WPR * sizeof(uint32_t);
#else
static const size_t kSearchSmem = 3 * (size_t)SNN_SMAX * SNN_WPR * sizeof(uint32_t);
#endif

// Enqueue one batch of windows: image chain, search (appends raw candidates), finish (turns the
// batch's raw candidates into records while its compacted windows are still in the scratch
// buffers), advance.  Nothing here waits for the device.
static int run_batch(snn_ctx* ctx, SnnChrom* c, const ExtractPlan& pl, int frame0, int nf, uint32_t* dbgE,
                     uint32_t* dbgV, float* dbgM) {
    const int npass = nf * pl.nq * SNN_NBRIGHT;
    int rc = snn_image_chain(ctx, c, pl, frame0, nf, dbgM);
    if (rc) return rc;
    {
        SnnScope scope(ctx, ST_SEARCH, 1);
        SearchParams sp;
        sp.nq = pl.nq;
        sp.minH = pl.minH;
        sp.maxW = pl.maxW;
        sp.raw_cap = pl.raw_cap;
        k_search<<<npass, NTHR, kSearchSmem, ctx->stream>>>(pl.low, pl.high, pl.info, sp, pl.counter, pl.raw, dbgE,
                                                            dbgV);
        SNN_LAUNCH_CHECK("k_search");
    }
    {
        SnnScope scope(ctx, ST_FINISH, 2);
        const int blocks = (pl.batch_cap * 32 + 255) / 256;
        k_finish<<<blocks, 256, 0, ctx->stream>>>(pl.raw, pl.counter, pl.raw_cap, pl.info, pl.nz, pl.Dc, pl.nq,
                                                  c->binsize, c->chromsize, pl.cand);
        SNN_LAUNCH_CHECK("k_finish");
        k_advance<<<1, 1, 0, ctx->stream>>>(pl.counter);
        SNN_LAUNCH_CHECK("k_advance");
    }
    return SNN_OK;
}

// After the last batch: how many candidates were found (one synchronisation per chromosome).
static int fetch_count(snn_ctx* ctx, const ExtractPlan& pl, int* n_found) {
    int cnt[2] = {0, 0};
    SNN_CUDA(cudaMemcpyAsync(cnt, pl.counter, sizeof cnt, cudaMemcpyDeviceToHost, ctx->stream));
    SNN_CUDA(snn_stream_sync(ctx));
    if (cnt[1] || cnt[0] > pl.raw_cap)
        return snn_fail(ctx, SNN_ENOMEM, "snn_extract: more than %d raw candidates in one call", pl.raw_cap);
    *n_found = cnt[0];
    return SNN_OK;
}

static int ensure_search_attr(snn_ctx* ctx) {
    SNN_CUDA(cudaFuncSetAttribute(k_search, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSearchSmem));
    return SNN_OK;
}

int snn_extract_impl(snn_ctx* ctx, SnnChrom* c, int nq, const double* mp, const double* gw, int radius, int minH,
                     int maxW, int bfilter, int gray_variant, int frame_lo, int frame_hi, snn_candidate* out,
                     int64_t cap, int64_t* n_out) {
    *n_out = 0;
    int rc = ensure_search_attr(ctx);
    if (rc) return rc;
    const int total = frame_hi - frame_lo;
    if (total <= 0) return SNN_OK;
    // windows per batch: keep the per-pass intermediates (~2 MB each) around 3 GB
    int fb = 1536 / (nq * SNN_NBRIGHT);
    if (fb < 1) fb = 1;
    if (fb > total) fb = total;
    ExtractPlan pl;
    rc = make_plan(ctx, &pl, nq, mp, gw, radius, minH, maxW, bfilter, gray_variant, fb, total, false, nullptr,
                   nullptr, nullptr);
    if (rc) return rc;
    SNN_CUDA(cudaMemsetAsync(pl.counter, 0, 64, ctx->stream));
    for (int f0 = frame_lo; f0 < frame_hi; f0 += fb) {
        int nf = std::min(fb, frame_hi - f0);
        rc = run_batch(ctx, c, pl, f0, nf, nullptr, nullptr, nullptr);
        if (rc) return rc;
    }
    int found = 0;
    rc = fetch_count(ctx, pl, &found);
    if (rc) return rc;
    std::vector<snn_candidate> all((size_t)found);
    if (found > 0) {
        SNN_CUDA(cudaMemcpyAsync(all.data(), pl.cand, (size_t)found * sizeof(snn_candidate), cudaMemcpyDeviceToHost,
                                 ctx->stream));
        SNN_CUDA(snn_stream_sync(ctx));
    }
    // the reference's emission order: maxpixel, window, brightness pass, up/down, pair index
    std::sort(all.begin(), all.end(), [](const snn_candidate& a, const snn_candidate& b) {
        if (a.q != b.q) return a.q < b.q;
        if (a.frame != b.frame) return a.frame < b.frame;
        if (a.bright != b.bright) return a.bright < b.bright;
        if (a.ud != b.ud) return a.ud < b.ud;
        return a.k < b.k;
    });
    *n_out = (int64_t)all.size();
    if ((int64_t)all.size() > cap)
        return snn_fail(ctx, SNN_ECAP, "snn_extract: %lld candidates, caller's buffer holds %lld",
                        (long long)all.size(), (long long)cap);
    if (!all.empty()) memcpy(out, all.data(), all.size() * sizeof(snn_candidate));
    return SNN_OK;
}

static void unpack_bits(const std::vector<uint32_t>& words, int S, uint8_t* out) {
    for (int r = 0; r < S; ++r)
        for (int c = 0; c < S; ++c) out[r * S + c] = (words[r * SNN_WPR + (c >> 5)] >> (c & 31)) & 1;
}

int snn_debug_stages_impl(snn_ctx* ctx, SnnChrom* c, double mp, const double* gw, int radius, int bfilter,
                          int gray_variant, int frame, int bright, int32_t* S_out, int32_t* nz, float* gray,
                          float* smoothed, float* mag, uint8_t* low, uint8_t* edges, uint8_t* vert) {
    int rc = ensure_search_attr(ctx);
    if (rc) return rc;
    ExtractPlan pl;
    uint32_t *dE = nullptr, *dV = nullptr;
    float* dM = nullptr;
    rc = make_plan(ctx, &pl, 1, &mp, gw, radius, 10, 8, bfilter, gray_variant, 1, 1, true, &dE, &dV, &dM);
    if (rc) return rc;
    SNN_CUDA(cudaMemsetAsync(pl.counter, 0, 64, ctx->stream));
    rc = run_batch(ctx, c, pl, frame, 1, dE, dV, dM);
    if (rc) return rc;
    FrameInfo fi;
    SNN_CUDA(cudaMemcpyAsync(&fi, pl.info, sizeof fi, cudaMemcpyDeviceToHost, ctx->stream));
    SNN_CUDA(snn_stream_sync(ctx));
    const int S = fi.S;
    *S_out = S;
    if (nz) SNN_CUDA(cudaMemcpy(nz, pl.nz, SNN_SMAX * sizeof(int32_t), cudaMemcpyDeviceToHost));
    if (S == 0) return SNN_OK;
    const size_t img = (size_t)SNN_SMAX * SNN_SMAX;
    std::vector<float> tmp(img);
    auto grab_f = [&](const float* dev, float* dst) -> int {
        SNN_CUDA(cudaMemcpy(tmp.data(), dev + (size_t)bright * img, img * sizeof(float), cudaMemcpyDeviceToHost));
        for (int r = 0; r < S; ++r) memcpy(dst + (size_t)r * S, tmp.data() + (size_t)r * SNN_SMAX, S * sizeof(float));
        return SNN_OK;
    };
    if (gray && (rc = grab_f(pl.bufA, gray))) return rc;
    if (smoothed && (rc = grab_f(pl.bufC, smoothed))) return rc;
    if (mag && (rc = grab_f(dM, mag))) return rc;
    std::vector<uint32_t> words((size_t)SNN_SMAX * SNN_WPR);
    auto grab_b = [&](const uint32_t* dev, uint8_t* dst) -> int {
        SNN_CUDA(cudaMemcpy(words.data(), dev + (size_t)bright * SNN_SMAX * SNN_WPR, words.size() * 4,
                            cudaMemcpyDeviceToHost));
        unpack_bits(words, S, dst);
        return SNN_OK;
    };
    if (low && (rc = grab_b(pl.low, low))) return rc;
    if (edges && (rc = grab_b(dE, edges))) return rc;
    if (vert && (rc = grab_b(dV, vert))) return rc;
    return SNN_OK;
}
