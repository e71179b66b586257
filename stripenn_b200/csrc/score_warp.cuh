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
};

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
