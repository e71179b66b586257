#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <omp.h>

/* CPU check of f_div (stripenn_b200/csrc/score_fast.cu): a / b for an integer-valued b with y = RN(1/b),
 *     q = RN(a*y); r = fma(-b, q, a); q = fma(r, y, q); r = fma(-b, q, a); q = fma(r, y, q)   ==   RN(a / b)
 * (Markstein's final correction applied to a quotient that is already within 2^-104 of the exact one).
 * Plain IEEE-754 arithmetic, so the CPU answers for the GPU.
 *  (1) EXHAUSTIVE: every count / valid pair 0 <= c <= n <= NMAX (the p-value fractions, getStripe.py:599-603);
 *  (2) random float64 sums (moderate exponents, random mantissas) over every column count 1..64
 *      (the row means of pvalue / scoringstripes, getStripe.py:568-573, 733-742);
 *  (3) random float64 over random integers up to 25600 (the largest background table).
 *   gcc -O2 -march=native -fopenmp -o /tmp/divchk/fdiv tools/check_fdiv.c -lm && /tmp/divchk/fdiv 6000 400000000
 */
static inline uint64_t rng(uint64_t* s) {
    uint64_t x = *s;
    x ^= x << 13; x ^= x >> 7; x ^= x << 17;
    return *s = x;
}
static inline double f_div(double a, double b, double y) {
    double q = a * y;
    double r = fma(-b, q, a);
    q = fma(r, y, q);
    r = fma(-b, q, a);
    return fma(r, y, q);
}

int main(int argc, char** argv) {
    int nmax = argc > 1 ? atoi(argv[1]) : 6000;
    long long n = argc > 2 ? atoll(argv[2]) : 200000000LL;
    long long bad1 = 0, bad2 = 0, bad3 = 0, cnt1 = 0;
#pragma omp parallel for schedule(dynamic, 16) reduction(+ : bad1, cnt1)
    for (int b = 1; b <= nmax; ++b) {
        const double y = 1.0 / (double)b;
        for (int c = 0; c <= b; ++c) {
            volatile double want = (double)c / (double)b;
            if (f_div((double)c, (double)b, y) != want) ++bad1;
            ++cnt1;
        }
    }
#pragma omp parallel reduction(+ : bad2, bad3)
    {
        uint64_t s = 0x9E3779B97F4A7C15ull * (uint64_t)(1 + omp_get_thread_num());
        long long per = n / omp_get_num_threads();
        for (long long i = 0; i < per; ++i) {
            union { uint64_t u; double d; } A;
            A.u = (rng(&s) & 0x000FFFFFFFFFFFFFull) | ((uint64_t)(1023 - 300 + (rng(&s) % 600)) << 52);
            const double b = (double)(1 + (rng(&s) % 64));
            volatile double w2 = A.d / b;
            if (f_div(A.d, b, 1.0 / b) != w2) ++bad2;
            const double b3 = (double)(1 + (rng(&s) % 25600));
            volatile double w3 = A.d / b3;
            if (f_div(A.d, b3, 1.0 / b3) != w3) ++bad3;
        }
    }
    printf("{\"exhaustive_pairs\": %lld, \"nmax\": %d, \"bad_pairs\": %lld, \"random\": %lld, \"bad_small_divisor\": %lld, "
           "\"bad_large_divisor\": %lld}\n", cnt1, nmax, bad1, n, bad2, bad3);
    return (bad1 | bad2 | bad3) ? 1 : 0;
}
