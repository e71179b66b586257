// Specialised (default-parameter) versions of the image-chain kernels of image.cu.  Same
// arithmetic, same operation order, bit-identical outputs (tests compare both paths with the
// oracle); what changes is how often each value is converted and re-read:
//   k_gray6        one CTA computes saturate(D, M) once per tile and emits all 6 brightness
//                  passes from shared memory (2 float64 divisions per pixel instead of 12)
//   k_gauss_v8/h8  every thread produces 8 consecutive outputs along the filter axis from
//                  8 + 2r inputs held in registers: each float32 input is converted to float64
//                  once per thread (≈3 conversions per output instead of 17)
//   k_sobel_nms2   derivative images staged once per tile (float32 subtraction is exact
//                  here, see the comment), [1 2 1] passes from float64 copies in shared memory
// ncu (profiles/r01a_*): the generic kernels were bound by the conversion (XU) pipe at 64-68 %
// and by issue slots at 84 %, not by FP64 or HBM.
#include <stdlib.h>

#include "image.cuh"

__device__ __forceinline__ int f_reflect101(int i, int n) {
    if (i < 0) i = -i;
    if (i >= n) i = 2 * n - 2 - i;
    return i;
}

// Correctly rounded float64 quotient a / b from y = RN(1 / b) in three instructions instead of the
// ~30 of a float64 division (Markstein: q0 = RN(a y) is faithful, r = a - q0 b is exact in an FMA,
// RN(q0 + r y) = RN(a / b)); checked against true division on 2e9 random and gray-shaped cases by
// tools/check_division_identities.c.  Operands must be finite and away from over/underflow.
__device__ __forceinline__ double div_by_recip(double a, double b, double y) {
    const double q0 = __dmul_rn(a, y);
    const double r = __fma_rn(-q0, b, a);
    return __fma_rn(r, y, q0);
}

// ---- saturate once, six brightness passes (bf = 3) -------------------------------------------
#define GT 34  // 32 + bf - 1
__global__ void __launch_bounds__(256) k_gray6(const double* __restrict__ Dc_all, const FrameInfo* __restrict__ info,
                                               GrayParams P, float* __restrict__ gray_all) {
    __shared__ double G[GT * GT];
    __shared__ double Pk[GT * GT];
    const int fq = blockIdx.z;  // frame * nq + q
    const int f = fq / P.nq, q = fq - f * P.nq;
    const int S = info[f].S;
    const int ty0 = blockIdx.y * 32, tx0 = blockIdx.x * 32;
    if (ty0 >= S || tx0 >= S) return;
    const double M = P.mp[q];
    const double* Dc = Dc_all + (int64_t)f * SNN_SMAX * SNN_SMAX;
    const int tid = threadIdx.y * 32 + threadIdx.x;
    // maxpixel values of real data are ordinary positive numbers; anything else (NaN of an empty
    // chromosome, 0, huge) keeps the true divisions
    const bool fastdiv = M > 1e-100 && M < 1e100;
    const double yM = 1.0 / M, y255 = 1.0 / 255.0;
    for (int t = tid; t < GT * GT; t += 256) {
        const int yy = t / GT, xx = t - yy * GT;
        int r = f_reflect101(ty0 + yy - 1, S), c = f_reflect101(tx0 + xx - 1, S);
        double g = 0.0;
        if (r >= 0 && r < S && c >= 0 && c < S) {
            const double D = Dc[r * SNN_SMAX + c];
            const double num = 255.0 * (M - D);
            if (fastdiv && fabs(D) < 1e100) {
                double blue = div_by_recip(num, M, yM);     // (255.0 * (M - D)) / M, getStripe.py:889
                if (blue < 0.0) blue = 0.0;
                g = div_by_recip(blue, 255.0, y255);
            } else {
                double blue = num / M;
                if (blue < 0.0) blue = 0.0;
                g = blue / 255.0;
            }
            g = fmin(fmax(g, 0.0), 1.0);
        }
        G[t] = g;
    }
    const float CR = 0.299f, CG = 0.587f, CB = 0.114f;
    const int x = threadIdx.x, y0 = threadIdx.y * 4;
    const float rterm = P.r32 * CR;                   // the red channel's share (constant image)
    for (int bi = 0; bi < SNN_NBRIGHT; ++bi) {
        const double b = P.bright[bi];
        const double k = P.kb[bi];                    // (1.0 - 0.0) / (b - 0.0), divided on the host
        __syncthreads();  // G ready / previous pass done with Pk
#pragma unroll
        for (int u = 0; u < (GT * GT + 255) / 256; ++u) {
            const int t = tid + 256 * u;
            if (t < GT * GT) {
                const double g = G[t];                    // never NaN: the tile is clipped with fmin/fmax above
                // ImageProcessing.py:17-23: v <= 0 -> 0 ; v > b -> 1 ; else k * v
                const double a = (g > b) ? 1.0 : k * g;   // g = 0 gives k * 0 = 0
                Pk[t] = P.kv * a;                         // the tap product of cv.filter2D
            }
        }
        __syncthreads();
        if (tx0 + x < S) {
            double p[6][3];
#pragma unroll
            for (int i = 0; i < 6; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j) p[i][j] = Pk[(y0 + i) * GT + x + j];
            float* gray = gray_all + ((int64_t)fq * SNN_NBRIGHT + bi) * SNN_SMAX * SNN_SMAX +
                          (int64_t)(ty0 + y0) * SNN_SMAX + tx0 + x;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                // cv.filter2D: accumulator from 0, taps row-major; 0.0 + p is p for p >= 0
                double acc = p[i][0];
                acc = acc + p[i][1];
                acc = acc + p[i][2];
#pragma unroll
                for (int ky = 1; ky < 3; ++ky)
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) acc = acc + p[i + ky][kx];
                if (acc > 1.0) acc = 1.0;             // np.clip(blur, 0, 1): the sum is never negative or NaN
                const float g = (float)acc;
                float out;
                if (P.variant == SNN_GRAY_IPP) {
                    const float t = rterm + (g * CG);
                    out = __fmaf_rn(g, CB, t);
                } else if (P.variant == SNN_GRAY_FMA2) {
                    const float t = __fmaf_rn(g, CG, rterm);
                    out = __fmaf_rn(g, CB, t);
                } else {
                    out = (rterm + (g * CG)) + (g * CB);
                }
                if (ty0 + y0 + i < S) gray[i * SNN_SMAX] = out;
            }
        }
    }
}

// ---- Gaussian, 8 outputs per thread ---------------------------------------------------------------
#define NOUT 8

// axis 0: thread = (column, chunk of 8 rows); loads are coalesced across the warp's columns
template <int R>
__global__ void __launch_bounds__(256) k_gauss_v8(const float* __restrict__ in_all, const FrameInfo* __restrict__ info,
                                                  int passes_per_frame, GaussParams G, float* __restrict__ out_all) {
    const int pass = blockIdx.z;
    const int S = info[pass / passes_per_frame].S;
    const int j = blockIdx.x * 32 + threadIdx.x;
    const int i0 = (blockIdx.y * 8 + threadIdx.y) * NOUT;
    if (j >= S || i0 >= S) return;
    const float* in = in_all + (int64_t)pass * SNN_SMAX * SNN_SMAX + j;
    float* out = out_all + (int64_t)pass * SNN_SMAX * SNN_SMAX + j;
    double x[NOUT + 2 * R];
#pragma unroll
    for (int t = 0; t < NOUT + 2 * R; ++t) {
        const int r = i0 - R + t;
        x[t] = (r >= 0 && r < S) ? (double)in[r * SNN_SMAX] : 0.0;
    }
    double w[R + 1];
#pragma unroll
    for (int t = 0; t <= R; ++t) w[t] = G.w[t];
#pragma unroll
    for (int o = 0; o < NOUT; ++o) {
        if (i0 + o >= S) break;
        double t = x[o + R] * w[R];
#pragma unroll
        for (int k = R; k >= 1; --k) t = t + (x[o + R - k] + x[o + R + k]) * w[R - k];
        out[(i0 + o) * SNN_SMAX] = (float)t;
    }
}

// axis 1 + bleed-over division: thread = (row, chunk of 8 columns); float4 loads
template <int R>
__global__ void __launch_bounds__(256) k_gauss_h8(const float* __restrict__ in_all, const FrameInfo* __restrict__ info,
                                                  int passes_per_frame, GaussParams G,
                                                  const float* __restrict__ bleed_all, float* __restrict__ out_all) {
    static_assert(R % 4 == 0 || R == 6 || R == 10, "R must keep 4-float alignment handling simple");
    const int pass = blockIdx.z;
    const int f = pass / passes_per_frame;
    const int S = info[f].S;
    const int nchunk = (S + NOUT - 1) / NOUT;
    const int t = blockIdx.x * 256 + threadIdx.x;
    const int i = t / nchunk;
    if (i >= S) return;
    const int j0 = (t - i * nchunk) * NOUT;
    const float* in = in_all + (int64_t)pass * SNN_SMAX * SNN_SMAX + i * SNN_SMAX;
    const float* bleed = bleed_all + (int64_t)f * SNN_SMAX * SNN_SMAX + i * SNN_SMAX;
    float* out = out_all + (int64_t)pass * SNN_SMAX * SNN_SMAX + i * SNN_SMAX;
    // inputs j0-R .. j0+NOUT+R-1 ; load as aligned float4 groups covering [j0 - RA, j0 + NOUT + RA)
    constexpr int RA = (R + 3) / 4 * 4;
    double x[NOUT + 2 * RA];
#pragma unroll
    for (int g = 0; g < (NOUT + 2 * RA) / 4; ++g) {
        const int c = j0 - RA + 4 * g;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c >= 0 && c < SNN_SMAX) v = *reinterpret_cast<const float4*>(in + c);
        x[4 * g + 0] = (c + 0 >= 0 && c + 0 < S) ? (double)v.x : 0.0;
        x[4 * g + 1] = (c + 1 >= 0 && c + 1 < S) ? (double)v.y : 0.0;
        x[4 * g + 2] = (c + 2 >= 0 && c + 2 < S) ? (double)v.z : 0.0;
        x[4 * g + 3] = (c + 3 >= 0 && c + 3 < S) ? (double)v.w : 0.0;
    }
    double w[R + 1];
#pragma unroll
    for (int k = 0; k <= R; ++k) w[k] = G.w[k];
    float res[NOUT];
#pragma unroll
    for (int o = 0; o < NOUT; ++o) {
        double acc = x[o + RA] * w[R];
#pragma unroll
        for (int k = R; k >= 1; --k) acc = acc + (x[o + RA - k] + x[o + RA + k]) * w[R - k];
        res[o] = (float)acc;
    }
    const float4 b0 = *reinterpret_cast<const float4*>(bleed + j0);
    const float4 b1 = *reinterpret_cast<const float4*>(bleed + j0 + 4);
    float4 o0, o1;
    o0.x = res[0] / b0.x; o0.y = res[1] / b0.y; o0.z = res[2] / b0.z; o0.w = res[3] / b0.w;
    o1.x = res[4] / b1.x; o1.y = res[5] / b1.y; o1.z = res[6] / b1.z; o1.w = res[7] / b1.w;
    // columns >= S inside the last chunk hold don't-care values (never read: every consumer masks
    // with S); the row pitch (400) keeps the stores inside the row
    *reinterpret_cast<float4*>(out + j0) = o0;
    *reinterpret_cast<float4*>(out + j0 + 4) = o1;
}

// ---- Gaussian, marching versions -------------------------------------------------------------------
// ncu on k_gauss_v8/h8 (profiles/r01c): 61 / 94 instructions per output for 25 float64 operations,
// load latency exposed (long-scoreboard 5.9 stalled warps per issue) because every thread loads
// its 24 inputs and then computes.  Here a thread walks along the filter axis with the 2R+1
// inputs of the current output in registers (rotating slots, the loop is unrolled by 2R+1 so no
// value is ever moved), each input is loaded and widened ONCE, and the loads of the next 2R+1
// outputs are in flight while the current ones are computed.  What is left per output is the
// 25 float64 operations of scipy's correlate1d plus ~8 others: the FP64 pipe is the bound.
// ---- one output of scipy's correlate1d, fused form with an exactness check ------------------------
// The reference accumulates  t = x_c w_R ;  t = t + (x_-k + x_+k) * w_(R-k)  (k = R..1)  in float64,
// one rounding per multiply and per add, then stores float32.  The FP64 pipe is what bounds these
// kernels (profiles/r01p: 81 %), so the common path uses 8 pair sums + 1 multiply + 8 FMAs (17
// float64 instructions instead of 25).  With non-negative inputs (gray >= 0.299, zero padding,
// positive weights: no cancellation) both forms are within 9u of the exact sum T relative to T
// (u = 2^-53; standard gamma_9 bound: the first product passes 9 roundings in the literal form, 9
// in the fused one), so they differ by < 18u T <= 36 ulps of t however t falls in its binade.  Two
// float64 values round to different float32 only if a rounding TIE (dropped 29 mantissa bits =
// 0x10000000) lies between them: when the fused t is farther than tieK = 64 ulps from every tie the
// literal t rounds to the same float32; otherwise (2.4e-7 of the outputs) the literal form is
// evaluated.  A negative input (sign bits are OR-ed into `sgn` by the callers, never seen in this
// pipeline) voids the bound and forces the literal form (tieK = 2^28 matches every t).
// SNN_GAUSS_TIE_K=268435456 forces the literal form everywhere (tests), 0 disables the check (A/B).
#define GAUSS_TIE_ALWAYS 0x10000000u
template <int R>
__device__ __forceinline__ double gauss_fused(const double* x, const int u, const GaussParams& G) {
    constexpr int W = 2 * R + 1;
    double t = x[(u + 1 + R) % W] * G.w[R];
#pragma unroll
    for (int k = R; k >= 1; --k) t = __fma_rn(x[(u + 1 + R - k) % W] + x[(u + 1 + R + k) % W], G.w[R - k], t);
    return t;
}
__device__ __forceinline__ bool gauss_near_tie(const double t, const unsigned tieK) {
    const unsigned d = ((unsigned)__double2loint(t) & 0x1fffffffu) - 0x10000000u + tieK;
    return d <= 2u * tieK;
}
// The literal form, out of line and straight from memory (cold: ~1 output in 4 million): p points
// at the centre input, `stride` apart along the filter axis, inputs [lo, hi] (relative to the
// centre) exist, the others are the zero padding of mode 'constant'.
__device__ __forceinline__ float gauss_literal(const float* p, int stride, int lo, int hi, int R, const GaussParams& G) {
    double t = (double)p[0] * G.w[R];
#pragma unroll 1
    for (int k = R; k >= 1; --k) {
        const double a = (-k >= lo) ? (double)p[-k * stride] : 0.0;
        const double b = (k <= hi) ? (double)p[k * stride] : 0.0;
        t = t + (a + b) * G.w[R - k];
    }
    return (float)t;
}

#define GV_SEG 102   // outputs per thread along axis 0 (6 groups of 17 at the default radius)
#ifndef GVM_MINB
#define GVM_MINB 5   // min CTAs per SM asked of the compiler: 96 registers, no spill (6: 80 registers + a 48-byte spill; A/B in DESIGN.md)
#endif

template <int R>
__global__ void __launch_bounds__(128, GVM_MINB) k_gauss_vm(const float* __restrict__ in_all, const FrameInfo* __restrict__ info,
                                                  int passes_per_frame, GaussParams G, unsigned tieK,
                                                  float* __restrict__ out_all) {
    constexpr int W = 2 * R + 1;
    const int pass = blockIdx.z;
    const int S = info[pass / passes_per_frame].S;
    const int j = blockIdx.x * 32 + threadIdx.x;
    const int i0 = (blockIdx.y * 4 + threadIdx.y) * GV_SEG;
    if (j >= S || i0 >= S) return;
    const int i1 = min(i0 + GV_SEG, S);
    const float* in = in_all + (int64_t)pass * SNN_SMAX * SNN_SMAX + j;
    float* out = out_all + (int64_t)pass * SNN_SMAX * SNN_SMAX + j;
    double x[W];   // at output row i (slot u = (i - i0) % W): row i - R + k lives in slot (u + 1 + k) % W
    x[0] = 0.0;
    int sgn = 0;   // OR of the input bit patterns: the sign bit says "a negative input was seen"
#pragma unroll
    for (int k = 0; k < 2 * R; ++k) {
        const int r = i0 - R + k;
        const float v = (r >= 0 && r < S) ? in[r * SNN_SMAX] : 0.f;
        sgn |= __float_as_int(v);
        x[k + 1] = (double)v;
    }
    float nx[W];
#pragma unroll
    for (int u = 0; u < W; ++u) {
        const int r = i0 + R + u;
        nx[u] = (r < S) ? in[r * SNN_SMAX] : 0.f;
    }
    const float* pin = in + (int64_t)(i0 + W + R) * SNN_SMAX;   // first row of the next group's inputs
    float* pout = out + (int64_t)i0 * SNN_SMAX;
    for (int base = i0; base < i1; base += W) {
        float cur[W];
#pragma unroll
        for (int u = 0; u < W; ++u) cur[u] = nx[u];
        if (base + W < i1) {
            if (base + 2 * W + R <= S) {          // every row of the next group exists: plain loads
#pragma unroll
                for (int u = 0; u < W; ++u) nx[u] = pin[u * SNN_SMAX];
            } else {
#pragma unroll
                for (int u = 0; u < W; ++u) nx[u] = (base + W + R + u < S) ? pin[u * SNN_SMAX] : 0.f;
            }
        }
        const int left = i1 - base;                 // outputs still wanted (tail group: stores predicated)
#pragma unroll
        for (int u = 0; u < W; ++u) sgn |= __float_as_int(cur[u]);
        const unsigned kk = sgn < 0 ? GAUSS_TIE_ALWAYS : tieK;
        unsigned redo = 0;                          // outputs of this group that sit next to a rounding tie
#pragma unroll
        for (int u = 0; u < W; ++u) {
            x[u] = (double)cur[u];
            const double t = gauss_fused<R>(x, u, G);
            redo |= gauss_near_tie(t, kk) ? (1u << u) : 0u;
            if (u < left) pout[u * SNN_SMAX] = (float)t;
        }
        // no branch inside the unrolled group (the 17 accumulation chains interleave freely); the
        // rare near-tie outputs are overwritten afterwards, recomputed literally from memory
        if (__builtin_expect(redo != 0u, 0)) {
#pragma unroll 1
            for (int u = 0; u < W; ++u)
                if (((redo >> u) & 1u) && u < left) {
                    const int i = base + u;
                    pout[u * SNN_SMAX] = gauss_literal(in + (int64_t)i * SNN_SMAX, SNN_SMAX, -i, S - 1 - i, R, G);
                }
        }
        pin += W * SNN_SMAX;
        pout += W * SNN_SMAX;
    }
}

// axis 1 + bleed-over division: a CTA (4 warps) owns a tile of 32 rows x 8(2R+1) columns.  It
// stages the tile plus R halo columns in shared memory (coalesced loads, zeros outside the
// image), then lane = row and warp = 2(2R+1)-column segment: the same walk as above along the
// row out of shared memory (odd pitch: the 32 lanes hit 32 banks) into a second tile, and a
// coalesced pass divides by the bleed-over image and stores.  37 KB per CTA: 6 CTAs per SM.
#define GH_ROWS 32
#ifndef GH_WARPS
#define GH_WARPS 4
#endif
template <int R>
__global__ void __launch_bounds__(GH_WARPS * 32) k_gauss_hm(const float* __restrict__ in_all,
                                                            const FrameInfo* __restrict__ info, int passes_per_frame,
                                                            GaussParams G, unsigned tieK,
                                                            const float* __restrict__ bleed_all,
                                                            float* __restrict__ out_all) {
    constexpr int W = 2 * R + 1;
    constexpr int TC = 2 * W * GH_WARPS;          // tile columns
    constexpr int PIN = (TC + 2 * R) | 1;         // input tile pitch (tile column t at index t + R)
    constexpr int POUT = TC | 1;
    __shared__ float tin[GH_ROWS * PIN];
    __shared__ float tout[GH_ROWS * POUT];
    const int pass = blockIdx.z;
    const int f = pass / passes_per_frame;
    const int S = info[f].S;
    const int r0 = blockIdx.y * GH_ROWS, ct0 = blockIdx.x * TC;
    if (r0 >= S || ct0 >= S) return;
    const int nr = min(GH_ROWS, S - r0);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float* in = in_all + (int64_t)pass * SNN_SMAX * SNN_SMAX + (int64_t)r0 * SNN_SMAX;
    // phase 1: rows -> tin with cp.async (all copies of a thread in flight at once; a source
    // size of 0 zero-fills the cells outside the image).  Warp = row, lanes along the columns.
    {
        constexpr int NCOL = TC + 2 * R;
        constexpr int NIT = (NCOL + 31) / 32;
        const unsigned tin_s = (unsigned)__cvta_generic_to_shared(tin);
#pragma unroll 2
        for (int rr = warp; rr < GH_ROWS; rr += GH_WARPS) {
            const bool rowok = rr < nr;
            const float* grow = in + rr * SNN_SMAX + (ct0 - R);
            const unsigned sdst = tin_s + (unsigned)(rr * PIN) * 4u;
#pragma unroll
            for (int it = 0; it < NIT; ++it) {
                const int p = lane + 32 * it;
                if (p < NCOL) {
                    const int c = ct0 - R + p;
                    const bool ok = rowok && c >= 0 && c < S;
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(sdst + (unsigned)p * 4u),
                                 "l"(ok ? grow + p : in), "r"(ok ? 4 : 0));
                }
            }
        }
        asm volatile("cp.async.commit_group;");
        asm volatile("cp.async.wait_group 0;");
    }
    __syncthreads();
    // phase 2: lane = row, warp = segment of 2W columns
    {
        const float* pin = tin + lane * PIN + warp * 2 * W;   // pin[k] = tile column (segment start - R + k)
        float* pout = tout + lane * POUT + warp * 2 * W;
        double x[W];
        x[0] = 0.0;
        int sgn = 0;
#pragma unroll
        for (int k = 0; k < 2 * R; ++k) {
            const float v = pin[k];
            sgn |= __float_as_int(v);
            x[k + 1] = (double)v;
        }
#pragma unroll
        for (int g = 0; g < 2; ++g) {
            float cur[W];
#pragma unroll
            for (int u = 0; u < W; ++u) {
                cur[u] = pin[g * W + u + 2 * R];
                sgn |= __float_as_int(cur[u]);
            }
            const unsigned kk = sgn < 0 ? GAUSS_TIE_ALWAYS : tieK;
            unsigned redo = 0;
#pragma unroll
            for (int u = 0; u < W; ++u) {
                x[u] = (double)cur[u];
                const double t = gauss_fused<R>(x, u, G);
                redo |= gauss_near_tie(t, kk) ? (1u << u) : 0u;
                pout[g * W + u] = (float)t;
            }
            // the tile already holds zeros outside the image: every input "exists"
            if (__builtin_expect(redo != 0u, 0)) {
#pragma unroll 1
                for (int u = 0; u < W; ++u)
                    if ((redo >> u) & 1u) pout[g * W + u] = gauss_literal(pin + g * W + u + R, 1, -R, R, R, G);
            }
        }
    }
    __syncthreads();
    // phase 3: divide by the bleed-over image, coalesced stores.  (Both exact reciprocal forms of
    // this division were measured and dropped — tools/check_division_identities.c, identities 2 and
    // 3: the float64 product moves work onto the FP64/conversion pipes the kernel is bound by, 35.1
    // vs 30.1 ms per step; the float32 Markstein form is neutral within noise and costs a second
    // per-window image.)
    const float* bleed = bleed_all + (int64_t)f * SNN_SMAX * SNN_SMAX + (int64_t)r0 * SNN_SMAX + ct0;
    float* out = out_all + (int64_t)pass * SNN_SMAX * SNN_SMAX + (int64_t)r0 * SNN_SMAX + ct0;
    const int nc = min(TC, S - ct0);
    constexpr int NIT3 = (TC + 31) / 32;
    // four rows at a time: all their bleed-over loads are in flight before the first division
#pragma unroll 1
    for (int rb = warp; rb < nr; rb += 4 * GH_WARPS) {
        float bv[4][NIT3];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int rr = rb + k * GH_WARPS;
#pragma unroll
            for (int it = 0; it < NIT3; ++it) {
                const int t = lane + 32 * it;
                bv[k][it] = (rr < nr && t < nc) ? bleed[rr * SNN_SMAX + t] : 1.f;
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int rr = rb + k * GH_WARPS;
#pragma unroll
            for (int it = 0; it < NIT3; ++it) {
                const int t = lane + 32 * it;
                if (rr < nr && t < nc) out[rr * SNN_SMAX + t] = tout[rr * POUT + t] / bv[k][it];
            }
        }
    }
}

// ---- axis 1 + bleed-over division, direct-store version (SNN_GAUSS_HD=1; A/B candidate) -----------
// k_gauss_hm spends more than half of its instructions outside the accumulation (profiles/r01s:
// phase 3 = second shared-memory tile, bleed-over loads and division 35 %, staging 18 %): 548 M
// warp instructions per launch against 253 M for the same arithmetic along axis 0.  Here the
// thread that walks along a row keeps its outputs in registers, divides them by the bleed-over
// values of its own row (five float4 loads per 20 outputs, issued before the arithmetic) and
// stores float4s straight to global memory — no second tile, no third phase, one barrier.  The
// window ring has HD_N = 20 slots (>= 2R+1, a multiple of 4) so that an unrolled group yields 20
// outputs = 5 aligned float4s; input k of the thread's segment lives in slot k % 20.
#define HD_ROWS 32
#define HD_WARPS 5
#define HD_N 20
#define HD_SEG 40                      // outputs per thread: 2 groups; 5 warps x 40 = 200 columns per CTA
#ifndef HD_MINB
#define HD_MINB 4
#endif
template <int R>
__device__ __forceinline__ double gauss_fused_ring(const double* x, const int u, const GaussParams& G) {
    // centre = input u + R of the group -> slot (u + R) % HD_N; pairs k away on both sides
    double t = x[(u + R) % HD_N] * G.w[R];
#pragma unroll
    for (int k = R; k >= 1; --k)
        t = __fma_rn(x[(u + R - k + HD_N) % HD_N] + x[(u + R + k) % HD_N], G.w[R - k], t);
    return t;
}

template <int R>
__global__ void __launch_bounds__(HD_WARPS * 32, HD_MINB) k_gauss_hd(const float* __restrict__ in_all,
                                                            const FrameInfo* __restrict__ info, int passes_per_frame,
                                                            GaussParams G, unsigned tieK,
                                                            const float* __restrict__ bleed_all,
                                                            float* __restrict__ out_all) {
    static_assert(2 * R + 1 <= HD_N, "ring too small");
    constexpr int TC = HD_SEG * HD_WARPS;         // tile columns
    constexpr int NCOL = TC + 2 * R;
    constexpr int PIN = NCOL | 1;                 // odd pitch: lane = row hits 32 banks
    __shared__ float tin[HD_ROWS * PIN];
    const int pass = blockIdx.z;
    const int f = pass / passes_per_frame;
    const int S = info[f].S;
    const int r0 = blockIdx.y * HD_ROWS, ct0 = blockIdx.x * TC;
    if (r0 >= S || ct0 >= S) return;
    const int nr = min(HD_ROWS, S - r0);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float* in = in_all + (int64_t)pass * SNN_SMAX * SNN_SMAX + (int64_t)r0 * SNN_SMAX;
    {   // staging: warp = row, lanes along the columns; zero fill outside the image
        constexpr int NIT = (NCOL + 31) / 32;
        const unsigned tin_s = (unsigned)__cvta_generic_to_shared(tin);
#pragma unroll 2
        for (int rr = warp; rr < HD_ROWS; rr += HD_WARPS) {
            const bool rowok = rr < nr;
            const float* grow = in + rr * SNN_SMAX + (ct0 - R);
            const unsigned sdst = tin_s + (unsigned)(rr * PIN) * 4u;
#pragma unroll
            for (int it = 0; it < NIT; ++it) {
                const int p = lane + 32 * it;
                if (p < NCOL) {
                    const int c = ct0 - R + p;
                    const bool ok = rowok && c >= 0 && c < S;
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(sdst + (unsigned)p * 4u),
                                 "l"(ok ? grow + p : in), "r"(ok ? 4 : 0));
                }
            }
        }
        asm volatile("cp.async.commit_group;");
        asm volatile("cp.async.wait_group 0;");
    }
    __syncthreads();
    const int cs = ct0 + warp * HD_SEG;           // first output column of this thread (multiple of 4)
    if (lane >= nr || cs >= S) return;
    const float* pin = tin + lane * PIN + warp * HD_SEG;      // pin[k] = image column cs - R + k
    const int64_t rowoff = (int64_t)(r0 + lane) * SNN_SMAX + cs;
    const float* bleed = bleed_all + (int64_t)f * SNN_SMAX * SNN_SMAX + rowoff;
    float* out = out_all + (int64_t)pass * SNN_SMAX * SNN_SMAX + rowoff;
    double x[HD_N];
    int sgn = 0;
#pragma unroll
    for (int k = 0; k < 2 * R; ++k) {
        const float v = pin[k];
        sgn |= __float_as_int(v);
        x[k] = (double)v;
    }
#pragma unroll
    for (int k = 2 * R; k < HD_N; ++k) x[k] = 0.0;
#pragma unroll 1
    for (int g = 0; g < HD_SEG / HD_N; ++g) {
        if (cs + g * HD_N >= S) break;            // the rest of the segment lies right of the image
        float4 bl[HD_N / 4];
#pragma unroll
        for (int m = 0; m < HD_N / 4; ++m) bl[m] = *reinterpret_cast<const float4*>(bleed + g * HD_N + 4 * m);
        float cur[HD_N];
#pragma unroll
        for (int u = 0; u < HD_N; ++u) {
            cur[u] = pin[g * HD_N + u + 2 * R];
            sgn |= __float_as_int(cur[u]);
        }
        const unsigned kk = sgn < 0 ? GAUSS_TIE_ALWAYS : tieK;
        unsigned redo = 0;
        float o[HD_N];
#pragma unroll
        for (int u = 0; u < HD_N; ++u) {
            x[(u + 2 * R) % HD_N] = (double)cur[u];
            const double t = gauss_fused_ring<R>(x, u, G);
            redo |= gauss_near_tie(t, kk) ? (1u << u) : 0u;
            o[u] = (float)t;
        }
        if (__builtin_expect(redo != 0u, 0)) {    // next to a rounding tie: the literal form decides
#pragma unroll
            for (int u = 0; u < HD_N; ++u)
                if ((redo >> u) & 1u) o[u] = gauss_literal(pin + g * HD_N + u + R, 1, -R, R, R, G);
        }
        // columns >= S of the last group hold don't-care values (never read; inside the row pitch)
#pragma unroll
        for (int m = 0; m < HD_N / 4; ++m) {
            float4 q;
            q.x = o[4 * m] / bl[m].x;
            q.y = o[4 * m + 1] / bl[m].y;
            q.z = o[4 * m + 2] / bl[m].z;
            q.w = o[4 * m + 3] / bl[m].w;
            *reinterpret_cast<float4*>(out + g * HD_N + 4 * m) = q;
        }
    }
}

// ---- Sobel + magnitude + NMS ------------------------------------------------------------------------
__device__ __forceinline__ int f_symm(int i, int n) {
    if (i < 0) i = -i - 1;
    if (i >= n) i = 2 * n - 1 - i;
    return i;
}

// ndi.sobel's derivative pass computes (x[i-1] - x[i+1]) * -1 in float64 and stores float32.
// For float32 operands that equals the float32 subtraction: when the exponents are within 2^29
// the float64 difference is exact (one rounding either way); otherwise the smaller operand is
// below 2^-29 of the larger, far from any float32 rounding boundary.  So the derivative images
// are built with __fsub_rn; the [1 2 1] pass keeps the reference's float64 accumulation.
__global__ void __launch_bounds__(256) k_sobel_nms2(const float* __restrict__ sm_all, const FrameInfo* __restrict__ info,
                                                    int passes_per_frame, uint32_t* __restrict__ low_all,
                                                    uint32_t* __restrict__ high_all, float* __restrict__ mag_dbg) {
    __shared__ float sm[36][37];     // smoothed, tile + halo 2 (reflected)
    __shared__ double di[34][37];    // d/d(axis 0) at rows r-1..r+32, cols c-2..c+33 (cols 0..35)
    __shared__ double dj[36][35];    // d/d(axis 1) at rows r-2..r+33 (rows 0..35), cols c-1..c+32
    __shared__ float si[34][35], sj[34][35], mg[34][35];
    const int pass = blockIdx.z;
    const int S = info[pass / passes_per_frame].S;
    const int ty0 = blockIdx.y * 32, tx0 = blockIdx.x * 32;
    if (ty0 >= S) return;
    uint32_t* low = low_all + (int64_t)pass * SNN_SMAX * SNN_WPR;
    uint32_t* high = high_all + (int64_t)pass * SNN_SMAX * SNN_WPR;
    const int tid = threadIdx.y * 32 + threadIdx.x;
    if (tx0 >= S) {
        if (tid < 32 && ty0 + tid < S) {
            low[(ty0 + tid) * SNN_WPR + blockIdx.x] = 0u;
            high[(ty0 + tid) * SNN_WPR + blockIdx.x] = 0u;
        }
        return;
    }
    const float* smo = sm_all + (int64_t)pass * SNN_SMAX * SNN_SMAX;
    for (int t = tid; t < 36 * 36; t += 256) {
        const int yy = t / 36, xx = t - yy * 36;
        const int r = f_symm(ty0 + yy - 2, S), c = f_symm(tx0 + xx - 2, S);
        float v = 0.f;
        if (r >= 0 && r < S && c >= 0 && c < S) v = smo[r * SNN_SMAX + c];
        sm[yy][xx] = v;
    }
    __syncthreads();
    // derivative images.  di[a][b]: image row ty0-1+a, image col tx0-2+b  (sm row a+1, col b)
    for (int t = tid; t < 34 * 36; t += 256) {
        const int a = t / 36, b = t - a * 36;
        di[a][b] = (double)(-__fsub_rn(sm[a][b], sm[a + 2][b]));
    }
    // dj[a][b]: image row ty0-2+a, image col tx0-1+b  (sm row a, col b+1)
    for (int t = tid; t < 36 * 34; t += 256) {
        const int a = t / 34, b = t - a * 34;
        dj[a][b] = (double)(-__fsub_rn(sm[a][b], sm[a][b + 2]));
    }
    __syncthreads();
    for (int t = tid; t < 34 * 34; t += 256) {
        const int yy = t / 34, xx = t - yy * 34;
        const int r = ty0 + yy - 1, c = tx0 + xx - 1;
        float is = 0.f, js = 0.f, m = 0.f;
        if (r >= 0 && r < S && c >= 0 && c < S) {
            // [1 2 1] along axis 1 of the axis-0 derivative; neighbours outside the image reflect
            // onto the pixel's own column (the derivative image is what scipy pads)
            const int bc = xx + 1;  // di column of image col c
            const int bl = (c - 1 < 0) ? bc : bc - 1, br = (c + 1 >= S) ? bc : bc + 1;
            is = (float)(di[yy][bc] * 2.0 + (di[yy][bl] + di[yy][br]) * 1.0);
            const int ar = yy + 1;  // dj row of image row r
            const int au = (r - 1 < 0) ? ar : ar - 1, ad = (r + 1 >= S) ? ar : ar + 1;
            js = (float)(dj[ar][xx] * 2.0 + (dj[au][xx] + dj[ad][xx]) * 1.0);
            float mm = is * is;
            mm = mm + js * js;
            m = sqrtf(mm);
        }
        si[yy][xx] = is;
        sj[yy][xx] = js;
        mg[yy][xx] = m;
    }
    __syncthreads();
    const int x = threadIdx.x;
#pragma unroll
    for (int yi = 0; yi < 4; ++yi) {
        const int y = threadIdx.y + 8 * yi;
        const int r = ty0 + y, c = tx0 + x;
        bool keep = false, strong = false;
        if (r < S && c < S) {
            const int Y = y + 1, X = x + 1;
            const float m = mg[Y][X];
            if (mag_dbg) mag_dbg[(int64_t)pass * SNN_SMAX * SNN_SMAX + r * SNN_SMAX + c] = m;
            if (r >= 1 && r <= S - 2 && c >= 1 && c <= S - 2 && m >= 0.1f) {
                const float is = si[Y][X], js = sj[Y][X];
                const bool up = is >= 0.f, down = is <= 0.f, right = js >= 0.f, left = js <= 0.f;
                const bool cond1 = (up && right) || (down && left);
                const bool cond2 = (down && right) || (up && left);
                const float ai = fabsf(is), aj = fabsf(js);
                float w = 0.f, n11 = 0.f, n12 = 0.f, n21 = 0.f, n22 = 0.f;
                bool ok = true;
                if (cond1) {
                    if (ai > aj) {
                        w = aj / ai;
                        n11 = mg[Y + 1][X]; n12 = mg[Y + 1][X + 1]; n21 = mg[Y - 1][X]; n22 = mg[Y - 1][X - 1];
                    } else {
                        w = ai / aj;
                        n11 = mg[Y][X + 1]; n12 = mg[Y + 1][X + 1]; n21 = mg[Y][X - 1]; n22 = mg[Y - 1][X - 1];
                    }
                } else if (cond2) {
                    if (ai < aj) {
                        w = ai / aj;
                        n11 = mg[Y][X + 1]; n12 = mg[Y - 1][X + 1]; n21 = mg[Y][X - 1]; n22 = mg[Y + 1][X - 1];
                    } else {
                        w = aj / ai;
                        n11 = mg[Y - 1][X]; n12 = mg[Y - 1][X + 1]; n21 = mg[Y + 1][X]; n22 = mg[Y + 1][X - 1];
                    }
                } else {
                    ok = false;
                }
                if (ok) {
                    const double omw = 1.0 - (double)w;
                    const double plus = (double)(n12 * w) + (double)n11 * omw;
                    const double minus = (double)(n22 * w) + (double)n21 * omw;
                    const double md = (double)m;
                    keep = (plus <= md) && (minus <= md);
                    strong = keep && (m >= 0.2f);
                }
            }
        }
        const unsigned lw = __ballot_sync(0xffffffffu, keep);
        const unsigned hw = __ballot_sync(0xffffffffu, strong);
        if (x == 0 && r < S) {
            low[r * SNN_WPR + blockIdx.x] = lw;
            high[r * SNN_WPR + blockIdx.x] = hw;
        }
    }
}

// ---- Sobel + magnitude + NMS, register-marching version ----------------------------------------------
// A warp owns a 32-column strip (28 useful columns + 2 halo lanes each side) and marches down
// the rows keeping s(rho-2..rho), dj(rho-2..rho), mag(rho-3..rho-1) in registers; horizontal
// neighbours come from warp shuffles.  No shared-memory staging of the images, one conversion
// per derivative value.  Rows / columns outside the image are LOADED reflected (scipy
// 'reflect'), which makes every derivative at a reflected position equal to the derivative at
// its source position — exactly how ndi.sobel pads the derivative image in its second pass.
#define SN_ROWS 64
#define SN_WARPS 15
template <bool DBG>
__global__ void __launch_bounds__(SN_WARPS * 32, 2) k_sobel_nms3(const float* __restrict__ sm_all,
                                                              const FrameInfo* __restrict__ info,
                                                              int passes_per_frame, uint32_t* __restrict__ low_all,
                                                              uint32_t* __restrict__ high_all,
                                                              float* __restrict__ mag_dbg) {
    __shared__ uint32_t lowS[SN_ROWS][SNN_WPR], highS[SN_ROWS][SNN_WPR];
    const int pass = blockIdx.y;
    const int S = info[pass / passes_per_frame].S;
    const int r0 = blockIdx.x * SN_ROWS;
    if (r0 >= S) return;
    const int r1 = min(r0 + SN_ROWS, S);
    for (int t = threadIdx.x; t < SN_ROWS * SNN_WPR; t += SN_WARPS * 32) {
        (&lowS[0][0])[t] = 0u;
        (&highS[0][0])[t] = 0u;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cbase = warp * 28;
    if (cbase < S) {    // S >= 11 here (smaller windows are skipped upstream)
        const int c = cbase + lane - 2;
        const int cc = min(max(f_symm(c, S), 0), S - 1);   // lanes beyond the image read a valid (unused) column
        const float* colp = sm_all + (int64_t)pass * SNN_SMAX * SNN_SMAX + cc;
        const unsigned FULL = 0xffffffffu;
        const bool lanemine = lane >= 2 && lane < 30 && c >= 1 && c <= S - 2;   // NMS columns of this lane
        const bool lanedbg = lane >= 2 && lane < 30 && c < S;
        const int rlo = max(r0, 1), rhi = min(r1, S - 1);                       // NMS rows [rlo, rhi)
        const int word = cbase >> 5, off = cbase & 31;
        float s0 = 0.f, s1 = 0.f;
        double dj0 = 0.0, dj1 = 0.0;
        float mA = 0.f, mB = 0.f, isB = 0.f, jsB = 0.f;
        float mAl = 0.f, mAr = 0.f, mBl = 0.f, mBr = 0.f;   // horizontal neighbours, carried down the rows
        // two rows are in flight while a row is processed
        float nxt = colp[f_symm(r0 - 2, S) * SNN_SMAX];
        float nxt2 = colp[max(f_symm(r0 - 1, S), 0) * SNN_SMAX];
#pragma unroll 3
        for (int rho = r0 - 2; rho <= r1 + 1; ++rho) {
            const float s2 = nxt;
            nxt = nxt2;
            {
                int rn = rho + 2;                       // scipy 'reflect' row of the iteration after next
                rn = rn < 0 ? -rn - 1 : rn;
                rn = rn >= S ? 2 * S - 1 - rn : rn;
                rn = min(max(rn, 0), S - 1);
                nxt2 = colp[rn * SNN_SMAX];
            }
            // axis-1 derivative of row rho
            const float sl = __shfl_up_sync(FULL, s2, 1), sr = __shfl_down_sync(FULL, s2, 1);
            const double djn = (double)(-__fsub_rn(sl, sr));
            // axis-0 derivative of row rho-1 (float32, exact — see above), then the [1 2 1] passes;
            // the neighbours travel as float32 and are widened after the shuffle
            const float di1f = -__fsub_rn(s0, s2);
            const double di1 = (double)di1f;
            const double dil = (double)__shfl_up_sync(FULL, di1f, 1), dir = (double)__shfl_down_sync(FULL, di1f, 1);
            const float is1 = (float)(di1 * 2.0 + (dil + dir) * 1.0);
            const float js1 = (float)(dj1 * 2.0 + (dj0 + djn) * 1.0);
            float mm = is1 * is1;
            mm = mm + js1 * js1;
            const float m1 = sqrtf(mm);
            // NMS of row rho-2 from mag rows rho-3 (mA), rho-2 (mB), rho-1 (m1)
            const float m1l = __shfl_up_sync(FULL, m1, 1), m1r = __shfl_down_sync(FULL, m1, 1);
            const int r = rho - 2;
            bool keep = false, strong = false;
            if (DBG && lanedbg && rho - 1 >= r0 && rho - 1 < r1)
                mag_dbg[(int64_t)pass * SNN_SMAX * SNN_SMAX + (rho - 1) * SNN_SMAX + c] = m1;
            const bool rowok = r >= rlo && r < rhi;     // warp-uniform
            if (rowok && lanemine && mB >= 0.1f) {
                const float is = isB, js = jsB, m = mB;
                const bool up = is >= 0.f, down = is <= 0.f, right = js >= 0.f, left = js <= 0.f;
                const bool c1 = (up && right) || (down && left);
                const bool c2 = !c1 && ((down && right) || (up && left));
                const float ai = fabsf(is), aj = fabsf(js);
                if (c1 || c2) {
                    // the four branches of _nonmaximum_suppression_bilinear differ only in WHICH
                    // neighbours they read: the weight is always smaller/larger gradient component
                    // (cond1: |i|>|j| -> |j|/|i| else |i|/|j| ; cond2: |i|<|j| -> |i|/|j| else |j|/|i|)
                    const bool idom = c1 ? (ai > aj) : !(ai < aj);
                    const float w = (idom ? aj : ai) / (idom ? ai : aj);
                    const float n12 = c1 ? m1r : mAr, n22 = c1 ? mAl : m1l;
                    const float n11 = idom ? (c1 ? m1 : mA) : mBr;
                    const float n21 = idom ? (c1 ? mA : m1) : mBl;
                    const double omw = 1.0 - (double)w;
                    const double plus = (double)(n12 * w) + (double)n11 * omw;
                    const double minus = (double)(n22 * w) + (double)n21 * omw;
                    const double md = (double)m;
                    keep = (plus <= md) && (minus <= md);
                    strong = keep && (m >= 0.2f);
                }
            }
            if (rowok) {
                const unsigned lw = (__ballot_sync(FULL, keep) >> 2) & 0x0fffffffu;
                if (lw) {
                    const unsigned hw = (__ballot_sync(FULL, strong) >> 2) & 0x0fffffffu;
                    if (lane == 0) {
                        atomicOr(&lowS[r - r0][word], lw << off);
                        if (off > 4 && word + 1 < SNN_WPR) atomicOr(&lowS[r - r0][word + 1], lw >> (32 - off));
                        if (hw) {
                            atomicOr(&highS[r - r0][word], hw << off);
                            if (off > 4 && word + 1 < SNN_WPR) atomicOr(&highS[r - r0][word + 1], hw >> (32 - off));
                        }
                    }
                }
            }
            s0 = s1; s1 = s2;
            dj0 = dj1; dj1 = djn;
            mA = mB; mB = m1;
            mAl = mBl; mAr = mBr; mBl = m1l; mBr = m1r;
            isB = is1; jsB = js1;
        }
    }
    __syncthreads();
    uint32_t* low = low_all + (int64_t)pass * SNN_SMAX * SNN_WPR + (int64_t)r0 * SNN_WPR;
    uint32_t* high = high_all + (int64_t)pass * SNN_SMAX * SNN_WPR + (int64_t)r0 * SNN_WPR;
    for (int t = threadIdx.x; t < (r1 - r0) * SNN_WPR; t += SN_WARPS * 32) {
        low[t] = (&lowS[0][0])[t];
        high[t] = (&highS[0][0])[t];
    }
}

// ---- Sobel + magnitude + NMS, two columns per lane ---------------------------------------------------
// Same walk as k_sobel_nms3, but a lane owns two ADJACENT columns (a warp covers 64 columns, 60
// useful + 2 halo each side): half of every horizontal neighbour is already in the lane's own
// registers, so a row costs 6 shuffles for 60 pixels instead of 6 for 28, the per-row index
// arithmetic is shared by twice as many pixels, and the two columns give the scheduler two
// independent dependency chains.  Bits go to the shared-memory masks with atomicOr from the
// (few) lanes that keep a pixel.
#define SN4_COLS 60
#define SN4_WARPS 7
#ifndef SN4_MINB
#define SN4_MINB 4
#endif

// skimage _nonmaximum_suppression_bilinear for one pixel (see k_sobel_nms3 for the branch merge)
__device__ __forceinline__ void nms_pixel(float is, float js, float m, float mAl, float mA, float mAr, float mBl,
                                          float mBr, float m1l, float m1, float m1r, bool* keep, bool* strong) {
    const bool up = is >= 0.f, down = is <= 0.f, right = js >= 0.f, left = js <= 0.f;
    const bool c1 = (up && right) || (down && left);
    const bool c2 = !c1 && ((down && right) || (up && left));
    *keep = false;
    *strong = false;
    if (c1 || c2) {
        const float ai = fabsf(is), aj = fabsf(js);
        const bool idom = c1 ? (ai > aj) : !(ai < aj);
        const float w = (idom ? aj : ai) / (idom ? ai : aj);
        const float n12 = c1 ? m1r : mAr, n22 = c1 ? mAl : m1l;
        const float n11 = idom ? (c1 ? m1 : mA) : mBr;
        const float n21 = idom ? (c1 ? mA : m1) : mBl;
        const double omw = 1.0 - (double)w;
        const double plus = (double)(n12 * w) + (double)n11 * omw;
        const double minus = (double)(n22 * w) + (double)n21 * omw;
        const double md = (double)m;
        *keep = (plus <= md) && (minus <= md);
        *strong = *keep && (m >= 0.2f);
    }
}

template <bool DBG>
__global__ void __launch_bounds__(SN4_WARPS * 32, SN4_MINB) k_sobel_nms4(const float* __restrict__ sm_all,
                                                                  const FrameInfo* __restrict__ info,
                                                                  int passes_per_frame, uint32_t* __restrict__ low_all,
                                                                  uint32_t* __restrict__ high_all,
                                                                  float* __restrict__ mag_dbg) {
    const int pass = blockIdx.y;
    const int S = info[pass / passes_per_frame].S;
    const int r0 = blockIdx.x * SN_ROWS;
    if (r0 >= S) return;
    const int r1 = min(r0 + SN_ROWS, S);
    // the masks are zeroed by the launcher; every warp adds its bits with global atomics (only
    // the ~2 % of pixels that survive the suppression) and no warp ever waits for another
    uint32_t* low = low_all + (int64_t)pass * SNN_SMAX * SNN_WPR;
    uint32_t* high = high_all + (int64_t)pass * SNN_SMAX * SNN_WPR;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cbase = warp * SN4_COLS;
    if (cbase < S) {    // S >= 11 here (smaller windows are skipped upstream)
        const unsigned FULL = 0xffffffffu;
        const int ca = cbase + 2 * lane - 2, cb = ca + 1;
        const float* base = sm_all + (int64_t)pass * SNN_SMAX * SNN_SMAX;
        const float* pa = base + min(max(f_symm(ca, S), 0), S - 1);
        const float* pb = base + min(max(f_symm(cb, S), 0), S - 1);
        const bool useful = lane >= 1 && lane <= 30;
        const bool mine_a = useful && ca >= 1 && ca <= S - 2, mine_b = useful && cb >= 1 && cb <= S - 2;
        const int rlo = max(r0, 1), rhi = min(r1, S - 1);       // NMS rows [rlo, rhi)
        // bit positions of the two columns in the 13-word rows (ca is even: both in one word)
        const int wword = max(ca, 0) >> 5, wshift = max(ca, 0) & 31;
        float s0a = 0.f, s1a = 0.f, s0b = 0.f, s1b = 0.f;
        double dj0a = 0.0, dj1a = 0.0, dj0b = 0.0, dj1b = 0.0;
        float mAa = 0.f, mBa = 0.f, isBa = 0.f, jsBa = 0.f, mAb = 0.f, mBb = 0.f, isBb = 0.f, jsBb = 0.f;
        float mAl = 0.f, mBl = 0.f, mAr = 0.f, mBr = 0.f;   // column ca-1 / cb+1 of the two previous rows
        int rn = f_symm(r0 - 2, S);
        float nxa = pa[rn * SNN_SMAX], nxb = pb[rn * SNN_SMAX];
        rn = max(f_symm(r0 - 1, S), 0);
        float nx2a = pa[rn * SNN_SMAX], nx2b = pb[rn * SNN_SMAX];
#pragma unroll 3
        for (int rho = r0 - 2; rho <= r1 + 1; ++rho) {
            const float s2a = nxa, s2b = nxb;
            nxa = nx2a;
            nxb = nx2b;
            {
                int q = rho + 2;                        // scipy 'reflect' row of the iteration after next
                q = q < 0 ? -q - 1 : q;
                q = q >= S ? 2 * S - 1 - q : q;
                q = min(max(q, 0), S - 1);
                nx2a = pa[q * SNN_SMAX];
                nx2b = pb[q * SNN_SMAX];
            }
            // axis-1 derivative of row rho: -(s[c-1] - s[c+1])
            const float sla = __shfl_up_sync(FULL, s2b, 1), srb = __shfl_down_sync(FULL, s2a, 1);
            const double djna = (double)(-__fsub_rn(sla, s2b));
            const double djnb = (double)(-__fsub_rn(s2a, srb));
            // axis-0 derivative of row rho-1, then the [1 2 1] passes (float64 accumulation)
            const float dia = -__fsub_rn(s0a, s2a), dib = -__fsub_rn(s0b, s2b);
            const float dila = __shfl_up_sync(FULL, dib, 1), dirb = __shfl_down_sync(FULL, dia, 1);
            const double dda = (double)dia, ddb = (double)dib;
            const float is1a = (float)(dda * 2.0 + ((double)dila + ddb) * 1.0);
            const float is1b = (float)(ddb * 2.0 + (dda + (double)dirb) * 1.0);
            const float js1a = (float)(dj1a * 2.0 + (dj0a + djna) * 1.0);
            const float js1b = (float)(dj1b * 2.0 + (dj0b + djnb) * 1.0);
            float mma = is1a * is1a, mmb = is1b * is1b;
            mma = mma + js1a * js1a;
            mmb = mmb + js1b * js1b;
            const float m1a = sqrtf(mma), m1b = sqrtf(mmb);
            const float m1l = __shfl_up_sync(FULL, m1b, 1), m1r = __shfl_down_sync(FULL, m1a, 1);
            const int r = rho - 2;
            if (DBG && useful && rho - 1 >= r0 && rho - 1 < r1) {
                float* md = mag_dbg + (int64_t)pass * SNN_SMAX * SNN_SMAX + (rho - 1) * SNN_SMAX;
                if (ca >= 0 && ca < S) md[ca] = m1a;
                if (cb >= 0 && cb < S) md[cb] = m1b;
            }
            if (r >= rlo && r < rhi) {      // warp-uniform
                bool ka = false, ha = false, kb = false, hb = false;
                if (mine_a && mBa >= 0.1f)
                    nms_pixel(isBa, jsBa, mBa, mAl, mAa, mAb, mBl, mBb, m1l, m1a, m1b, &ka, &ha);
                if (mine_b && mBb >= 0.1f)
                    nms_pixel(isBb, jsBb, mBb, mAa, mAb, mAr, mBa, mBr, m1a, m1b, m1r, &kb, &hb);
                if (ka | kb) {
                    atomicOr(&low[r * SNN_WPR + wword], ((ka ? 1u : 0u) | (kb ? 2u : 0u)) << wshift);
                    if (ha | hb) atomicOr(&high[r * SNN_WPR + wword], ((ha ? 1u : 0u) | (hb ? 2u : 0u)) << wshift);
                }
            }
            s0a = s1a; s1a = s2a; s0b = s1b; s1b = s2b;
            dj0a = dj1a; dj1a = djna; dj0b = dj1b; dj1b = djnb;
            mAl = mBl; mBl = m1l; mAr = mBr; mBr = m1r;
            mAa = mBa; mBa = m1a; mAb = mBb; mBb = m1b;
            isBa = is1a; jsBa = js1a; isBb = is1b; jsBb = js1b;
        }
    }
}

// ---- launchers (image.cu picks them when the parameters match) ----------------------------------------
int snn_launch_gray6(snn_ctx* ctx, const ExtractPlan& pl, int nf) {
    k_gray6<<<dim3(13, 13, nf * pl.nq), dim3(32, 8), 0, ctx->stream>>>(pl.Dc, pl.info, pl.gray, pl.bufA);
    SNN_LAUNCH_CHECK("k_gray6");
    return SNN_OK;
}

bool snn_gauss_fast_supported(int radius) { return radius == 8 || radius == 10 || radius == 6; }

template <int R>
static int launch_gauss_march(snn_ctx* ctx, const ExtractPlan& pl, int npass, int ppf, bool axis1_only) {
    constexpr int TC = 2 * (2 * R + 1) * GH_WARPS;
    const int nseg = (SNN_SMAX + GV_SEG - 1) / GV_SEG;
    const char* tk = getenv("SNN_GAUSS_TIE_K");     // read per launch: the tests switch it
    const unsigned tieK = tk ? (unsigned)strtoul(tk, nullptr, 0) : 64u;
    if (!axis1_only)
        k_gauss_vm<R><<<dim3(13, (nseg + 3) / 4, npass), dim3(32, 4), 0, ctx->stream>>>(pl.bufA, pl.info, ppf, pl.gauss,
                                                                                       tieK, pl.bufB);
    static_assert(SNN_SMAX % (HD_SEG * HD_WARPS) == 0, "k_gauss_hd tiles must cover the row pitch exactly");
    const char* hd = getenv("SNN_GAUSS_HD");        // A/B switch: direct-store axis-1 kernel (default sigma only)
    if constexpr (2 * R + 1 <= HD_N && 2 * R + 1 > HD_N - 4) {
        if (hd && hd[0] == '1') {
            k_gauss_hd<R><<<dim3(SNN_SMAX / (HD_SEG * HD_WARPS), (SNN_SMAX + HD_ROWS - 1) / HD_ROWS, npass),
                            HD_WARPS * 32, 0, ctx->stream>>>(pl.bufB, pl.info, ppf, pl.gauss, tieK, pl.bleed, pl.bufC);
            SNN_LAUNCH_CHECK("k_gauss_hd");
            return SNN_OK;
        }
    }
    (void)hd;
    k_gauss_hm<R><<<dim3((SNN_SMAX + TC - 1) / TC, (SNN_SMAX + GH_ROWS - 1) / GH_ROWS, npass), GH_WARPS * 32, 0,
                        ctx->stream>>>(pl.bufB, pl.info, ppf, pl.gauss, tieK, pl.bleed, pl.bufC);
    SNN_LAUNCH_CHECK("k_gauss_march");
    return SNN_OK;
}

int snn_launch_gauss_fast(snn_ctx* ctx, const ExtractPlan& pl, int npass, int ppf, bool axis1_only) {
    static const bool old8 = getenv("SNN_GAUSS8") != nullptr;   // A/B switch for the 8-outputs-per-thread kernels
    if (!old8 || axis1_only) {
        switch (pl.gauss.radius) {
            case 8: return launch_gauss_march<8>(ctx, pl, npass, ppf, axis1_only);
            case 10: return launch_gauss_march<10>(ctx, pl, npass, ppf, axis1_only);
            default: return launch_gauss_march<6>(ctx, pl, npass, ppf, axis1_only);
        }
    }
    const dim3 gv(13, (SNN_SMAX + 8 * NOUT - 1) / (8 * NOUT), npass), bv(32, 8);
    const int per_pass = SNN_SMAX * ((SNN_SMAX + NOUT - 1) / NOUT);
    const dim3 gh((per_pass + 255) / 256, 1, npass), bh(256);
    switch (pl.gauss.radius) {
        case 8:
            k_gauss_v8<8><<<gv, bv, 0, ctx->stream>>>(pl.bufA, pl.info, ppf, pl.gauss, pl.bufB);
            k_gauss_h8<8><<<gh, bh, 0, ctx->stream>>>(pl.bufB, pl.info, ppf, pl.gauss, pl.bleed, pl.bufC);
            break;
        case 10:
            k_gauss_v8<10><<<gv, bv, 0, ctx->stream>>>(pl.bufA, pl.info, ppf, pl.gauss, pl.bufB);
            k_gauss_h8<10><<<gh, bh, 0, ctx->stream>>>(pl.bufB, pl.info, ppf, pl.gauss, pl.bleed, pl.bufC);
            break;
        default:
            k_gauss_v8<6><<<gv, bv, 0, ctx->stream>>>(pl.bufA, pl.info, ppf, pl.gauss, pl.bufB);
            k_gauss_h8<6><<<gh, bh, 0, ctx->stream>>>(pl.bufB, pl.info, ppf, pl.gauss, pl.bleed, pl.bufC);
            break;
    }
    SNN_LAUNCH_CHECK("k_gauss_fast");
    return SNN_OK;
}

int snn_launch_sobel_nms2(snn_ctx* ctx, const ExtractPlan& pl, int npass, int ppf, float* mag_dbg) {
    const dim3 grid((SNN_SMAX + SN_ROWS - 1) / SN_ROWS, npass);
    static const bool old3 = getenv("SNN_NMS3") != nullptr;   // A/B switch: one column per lane
    if (!old3) {
        const size_t mask_bytes = (size_t)npass * SNN_SMAX * SNN_WPR * sizeof(uint32_t);
        SNN_CUDA(cudaMemsetAsync(pl.low, 0, mask_bytes, ctx->stream));
        SNN_CUDA(cudaMemsetAsync(pl.high, 0, mask_bytes, ctx->stream));
        if (mag_dbg)
            k_sobel_nms4<true><<<grid, SN4_WARPS * 32, 0, ctx->stream>>>(pl.bufC, pl.info, ppf, pl.low, pl.high, mag_dbg);
        else
            k_sobel_nms4<false><<<grid, SN4_WARPS * 32, 0, ctx->stream>>>(pl.bufC, pl.info, ppf, pl.low, pl.high, mag_dbg);
        SNN_LAUNCH_CHECK("k_sobel_nms4");
        return SNN_OK;
    }
    if (mag_dbg)
        k_sobel_nms3<true><<<grid, SN_WARPS * 32, 0, ctx->stream>>>(pl.bufC, pl.info, ppf, pl.low, pl.high, mag_dbg);
    else
        k_sobel_nms3<false><<<grid, SN_WARPS * 32, 0, ctx->stream>>>(pl.bufC, pl.info, ppf, pl.low, pl.high, mag_dbg);
    SNN_LAUNCH_CHECK("k_sobel_nms3");
    return SNN_OK;
}
