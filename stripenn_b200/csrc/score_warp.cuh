// Shared declarations of the two scoring kernels (score.cu: CTA per stripe, any window;
// score_warp.cu: warp per stripe, in-band windows up to SW_MAXCOLS bins wide).
#pragma once
#include "snn_internal.cuh"

#define SW_WARPS 4         // stripes in flight per CTA (one per warp)
#ifndef SW_WINCAP
#define SW_WINCAP 1056     // doubles of window tile per warp: 32 rows x pitch 33
#endif
#ifndef SW_MINB
#define SW_MINB 6          // CTAs per SM the register allocation aims at
#endif
#define SW_MAXCOLS 128     // widest window the warp kernel takes (pitch 129 -> 8-row tiles)
#define SW_HMAX 320        // tallest stripe the warp kernel takes

// Per-chromosome constants of a scoring launch (device array, one entry per chromosome slot)
struct ChromDev {
    MatView m;
    int64_t chromsize, nbins;
    const double* rex;     // 1 / (expected[d] + 1e-8), SNN_NDIST entries (null without stripiness)
    // fast kernel (score_fast.cu): the column-major full band, pre-offset so that pixel (row, col) is
    // fb[col * fbT + row]; rexm[d] = rex[min(|d|, 399)] for |d| < SF_RM (null without stripiness)
    const double* fb;
    int64_t fbT;
    const double* rexm;
    const uint8_t* nanbin;   // weight of the bin is NaN; null when the chromosome has no weights
    const int32_t* nanpre;   // exclusive prefix counts of nanbin
};

// ---- fast kernel: "regular" stripes (the geometry StripeSearch emits away from the chromosome ends:
// pos2 and pos4 on bin boundaries, full flanks) whose window lies inside the band -----------------------
#define SF_WARPS 8
#ifndef SF_MINB
#define SF_MINB 4          // CTAs per SM the register allocation aims at (64 registers)
#endif
#define SF_RM 1024         // half-width of the mirrored 1/(expected+1e-8) table
#define SF_EXSTRIDE 4096   // doubles per chromosome slot: expected[512] | rex[512] | pad | rexm[2*SF_RM] from 2048
struct FastRec {           // host-side geometry of one stripe, all in bins
    int32_t xsi;           // first centre column = int(pos1 / r)
    int32_t wc;            // centre width = int(pos2 / r) - xsi
    int32_t R0;            // first row = int(pos3 / r)
    int32_t h;             // rows
    int32_t mode;          // bits 0-3: background row rule (getStripe.py:583-597) 0 down, 1 up, 2 stale, 3 none;
                           // bit 4: the stale table; bit 8: pos1 == pos3
    int32_t aux;           // mode 1: y2 - y1; mode 2: the stale distance
    int32_t slot;          // chromosome slot
    int32_t idx;           // output row
};
#define SF_HMAX 320         // tallest stripe the fast kernel takes
#define SF_HSM 128          // row results of stripes up to this height stay in shared memory
int snn_launch_score_fast(snn_ctx* ctx, const ChromDev* cds, const FastRec* recs, int n, int bs, int want_p, int want_s,
                          int* next, double* tall, int grid, double* out);
int snn_score_fast_grid(int n);
int snn_prepare_score_fast(snn_ctx* ctx, int* next);

int snn_launch_score_warp(snn_ctx* ctx, const ChromDev* cds, const int32_t* cslot, const int64_t* pos,
                          const int32_t* stale, const int32_t* order, int n, int64_t binsize, int bs, int hcap,
                          int want_p, int want_s, double* out);

struct ScoreSlot {
    SnnChrom* c;
    const double* expected;   // host, SNN_NDIST entries, or null (p-values only)
};
int snn_score_multi_impl(snn_ctx* ctx, const std::vector<ScoreSlot>& slots, int64_t n, const int32_t* slot_of,
                         const int64_t* pos, const int32_t* stale, double* pvalue, double* stripiness,
                         double* oe_mean, double* oe_total, double* o_mean, double* o_sum);
