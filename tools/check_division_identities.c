#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <omp.h>

/* CPU check of the two exact-division identities the image kernels rely on (plain IEEE-754
 * arithmetic, so a CPU check is as good as a GPU one):
 *  (1) float64, Markstein: y = RN(1/b), q0 = RN(a*y), r = fma(-q0, b, a), q = fma(r, y, q0)  ==  RN(a/b).
 *      k_gray6: (255*(M-D))/M and blue/255 in 3 instructions instead of a ~30-instruction float64 division.
 *  (2) float32 through float64: (float)((double)x * RN64(1/(double)b))  ==  x / b in float32.
 *  (3) float32 Markstein: y = RN32(1/b), q0 = x*y, r = fmaf(-q0, b, x), q = fmaf(r, y, q0)  ==  x / b.
 *      k_gauss_hm: smoothed / bleed on the (idle) FP32 pipe, no MUFU.RCP.
 *  (4) float32 division by the interior bleed-over constant b0 = 1 + 2^-23: RN(x / b0) is x's bit
 *      pattern minus 1, or minus 2 when the mantissa field f is 0 or >= 0x400002.  EXHAUSTIVE over
 *      every positive float with biased exponent 2..254 (2.1e9 values).  Measured in k_gauss_hm phase 3 and not kept.
 *   gcc -O2 -march=native -fopenmp -o /tmp/divchk/chk tools/check_division_identities.c -lm && /tmp/divchk/chk 2000000000
 */
static inline uint64_t rng(uint64_t* s) {
    uint64_t x = *s;
    x ^= x << 13; x ^= x >> 7; x ^= x << 17;
    return *s = x;
}
static inline double markstein(double a, double b, double y) {
    double q0 = a * y;
    double r = fma(-q0, b, a);
    return fma(r, y, q0);
}
static inline double u01(uint64_t* s) { return (double)(rng(s) >> 11) * (1.0 / 9007199254740992.0); }

int main(int argc, char** argv) {
    long long n = argc > 1 ? atoll(argv[1]) : 200000000LL;
    long long bad1 = 0, bad2 = 0, bad3 = 0, bad4 = 0, bad5 = 0;
    const double y255 = 1.0 / 255.0;
#pragma omp parallel reduction(+ : bad1, bad2, bad3, bad4, bad5)
    {
        uint64_t s = 0x9E3779B97F4A7C15ull * (uint64_t)(1 + omp_get_thread_num());
        long long per = n / omp_get_num_threads();
        for (long long i = 0; i < per; ++i) {
            /* (1a) the k_gray6 shape: M a quantile (integer-ish or balanced float), D a pixel value */
            double M = (rng(&s) & 1) ? (double)(1 + (rng(&s) % 5000)) : 0.01 + 3000.0 * u01(&s);
            double D = (rng(&s) & 1) ? (double)(rng(&s) % 20000) : 4000.0 * u01(&s) * u01(&s);
            double a = 255.0 * (M - D);
            double y = 1.0 / M;
            if (markstein(a, M, y) != a / M) ++bad1;
            double blue = a / M;
            if (blue < 0.0) blue = 0.0;
            if (markstein(blue, 255.0, y255) != blue / 255.0) ++bad2;
            /* (1b) arbitrary normal doubles (random bit patterns, exponents kept moderate) */
            union { uint64_t u; double d; } A, B;
            A.u = (rng(&s) & 0x800FFFFFFFFFFFFFull) | ((uint64_t)(1023 - 200 + (rng(&s) % 400)) << 52);
            B.u = (rng(&s) & 0x800FFFFFFFFFFFFFull) | ((uint64_t)(1023 - 200 + (rng(&s) % 400)) << 52);
            if (markstein(A.d, B.d, 1.0 / B.d) != A.d / B.d) ++bad3;
            /* (2) float32 division through a float64 reciprocal */
            union { uint32_t u; float f; } X, Bf;
            X.u = ((uint32_t)rng(&s) & 0x007FFFFFu) | ((uint32_t)(127 - 40 + (rng(&s) % 80)) << 23);
            Bf.u = ((uint32_t)rng(&s) & 0x007FFFFFu) | ((uint32_t)(127 - 40 + (rng(&s) % 80)) << 23);
            double rd = 1.0 / (double)Bf.f;
            volatile float q = X.f / Bf.f;
            if ((float)((double)X.f * rd) != q) ++bad4;
            {
                volatile float yv = 1.0f / Bf.f;
                float y = yv;
                float q0 = X.f * y;
                float r = fmaf(-q0, Bf.f, X.f);
                if (fmaf(r, y, q0) != q) ++bad5;
                /* the k_gauss_hm shape: x in [0.1, 1.1), b in [0.2, 1.01) */
                union { uint32_t u; float f; } x2, b2;
                x2.f = 0.1f + (float)u01(&s);
                b2.f = 0.2f + 0.81f * (float)u01(&s);
                x2.u ^= (uint32_t)rng(&s) & 0xFF; b2.u ^= (uint32_t)rng(&s) & 0xFF;
                volatile float y2v = 1.0f / b2.f;
                float y2 = y2v;
                volatile float qq = x2.f / b2.f;
                float q02 = x2.f * y2;
                float r2 = fmaf(-q02, b2.f, x2.f);
                if (fmaf(r2, y2, q02) != qq) ++bad5;
            }
        }
    }
    long long bad6 = 0, n6 = 0;
    {
        union { uint32_t u; float f; } b0;
        b0.u = 0x3f800001u;
        const float B0 = b0.f;
#pragma omp parallel for reduction(+ : bad6, n6) schedule(static)
        for (long long bits = 0x01000000LL; bits <= 0x7f7fffffLL; ++bits) {
            union { uint32_t u; float f; } x, q, w;
            x.u = (uint32_t)bits;
            volatile float qv = x.f / B0;
            q.f = qv;
            const uint32_t fm = x.u & 0x7fffffu;
            w.u = x.u - ((fm - 1u >= 0x400001u) ? 2u : 1u);
            if (w.u != q.u) ++bad6;
            ++n6;
        }
    }
    printf("division by 1+2^-23 as an integer step: %lld floats checked, %lld bad\n", n6, bad6);
    if (bad6) return 1;
    printf("cases %lld  markstein(gray M): %lld bad  markstein(/255): %lld bad  markstein(random): %lld bad  f32-via-f64: %lld bad  f32 markstein: %lld bad\n",
           n, bad1, bad2, bad3, bad4, bad5);
    return (bad1 | bad2 | bad3 | bad4 | bad5) ? 1 : 0;
}
