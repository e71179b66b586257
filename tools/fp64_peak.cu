// Pipe-throughput micro-benchmark for the B200 image chain ("the image
// chain is FP64-bound; the peak is to be measured by a micro-benchmark in the build").
// MEASURED_PEAKS.json has HBM and bf16 figures only, so the FP64 roof the image-chain kernels are
// reported against comes from here: DFMA / DADD / DMUL rates, the float<->double conversion rates
// (F2F, the second pipe those kernels lean on), IEEE sqrtf / float division, and DFMA + F2F mixes
// that show whether the two pipes overlap.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/fp64_peak tools/fp64_peak.cu
//   tools/bin/fp64_peak [out.json]
//
// Every kernel runs CHAINS independent dependency chains per thread, ITERS iterations, on
// 148 x 8 CTAs of 256 threads (all warps resident), timed with CUDA events (best of 5).
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define CHAINS 8
#define ITERS 4096
#define CHECK(x)                                                                        \
    do {                                                                                \
        cudaError_t e = (x);                                                            \
        if (e != cudaSuccess) {                                                         \
            fprintf(stderr, "%s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e)); \
            exit(1);                                                                    \
        }                                                                               \
    } while (0)

__global__ void k_dfma(double* out, double a, double b) {
    double x[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) x[c] = a * (threadIdx.x + c);
#pragma unroll 1
    for (int i = 0; i < ITERS; ++i)
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) x[c] = __fma_rn(x[c], a, b);
    double s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += x[c];
    if (s == 12345.678) out[threadIdx.x] = s;
}
__global__ void k_dadd(double* out, double a, double b) {
    double x[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) x[c] = a * (threadIdx.x + c);
#pragma unroll 1
    for (int i = 0; i < ITERS; ++i)
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) x[c] = __dadd_rn(x[c], b);
    double s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += x[c];
    if (s == 12345.678) out[threadIdx.x] = s;
}
__global__ void k_dmul(double* out, double a, double b) {
    double x[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) x[c] = a * (threadIdx.x + c + 1);
#pragma unroll 1
    for (int i = 0; i < ITERS; ++i)
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) x[c] = __dmul_rn(x[c], b);
    double s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += x[c];
    if (s == 12345.678) out[threadIdx.x] = s;
}
// one widening and one narrowing conversion per chain step (dependent)
__global__ void k_cvt_pair(double* out, float a) {
    float f[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) f[c] = a * (threadIdx.x + c + 1);
#pragma unroll 1
    for (int i = 0; i < ITERS; ++i)
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) {
            double d;
            asm volatile("cvt.f64.f32 %0, %1;" : "=d"(d) : "f"(f[c]));
            asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(f[c]) : "d"(d));
        }
    float s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += f[c];
    if (s == 12345.678f) out[threadIdx.x] = s;
}
// widening only: the float is advanced with an integer add (ALU pipe), the double's high word is
// folded into an integer (ALU pipe)
__global__ void k_cvt_widen(double* out, float a) {
    int f[CHAINS];
    int acc = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) f[c] = __float_as_int(a * (threadIdx.x + c + 1));
#pragma unroll 1
    for (int i = 0; i < ITERS; ++i)
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) {
            double d;
            asm volatile("cvt.f64.f32 %0, %1;" : "=d"(d) : "f"(__int_as_float(f[c])));
            acc ^= __double2hiint(d);
            f[c] += 1;
        }
    if (acc == 0x12345678) out[threadIdx.x] = acc;
}
__global__ void k_cvt_narrow(double* out, double a) {
    int hi[CHAINS];
    int acc = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) hi[c] = __double2hiint(a * (threadIdx.x + c + 1));
#pragma unroll 1
    for (int i = 0; i < ITERS; ++i)
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) {
            float f;
            asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(f) : "d"(__hiloint2double(hi[c], 0x12345678 + i)));
            acc ^= __float_as_int(f);
            hi[c] += 1;
        }
    if (acc == 0x12345678) out[threadIdx.x] = acc;
}
// float32 -> float64 with integer instructions only (normal, non-zero inputs): hi = (b >> 3) +
// 0x38000000 with the sign moved up, lo = b << 29
__device__ __forceinline__ double widen_int(int b) {
    const unsigned u = (unsigned)b;
    const unsigned hi = ((u & 0x7fffffffu) >> 3) + 0x38000000u + (u & 0x80000000u);
    return __hiloint2double((int)hi, (int)(u << 29));
}
__global__ void k_widen_int(double* out, float a) {
    int f[CHAINS];
    int acc = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) f[c] = __float_as_int(a * (threadIdx.x + c + 1));
#pragma unroll 1
    for (int i = 0; i < ITERS; ++i)
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) {
            const double d = widen_int(f[c]);
            acc ^= __double2hiint(d) ^ __double2loint(d);
            f[c] += 1;
        }
    if (acc == 0x12345678) out[threadIdx.x] = acc;
}
__global__ void k_sqrtf(double* out, float a) {
    float f[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) f[c] = a * (threadIdx.x + c + 1);
#pragma unroll 1
    for (int i = 0; i < ITERS; ++i)
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) f[c] = __fsqrt_rn(f[c]) + a;
    float s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += f[c];
    if (s == 12345.678f) out[threadIdx.x] = s;
}
__global__ void k_divf(double* out, float a) {
    float f[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) f[c] = a * (threadIdx.x + c + 1);
#pragma unroll 1
    for (int i = 0; i < ITERS; ++i)
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) f[c] = __fdiv_rn(a, f[c]) + a;
    float s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += f[c];
    if (s == 12345.678f) out[threadIdx.x] = s;
}
// NF DFMAs per widening conversion, the conversion feeding the chain (the Gaussian's shape)
template <int NF>
__global__ void k_mix(double* out, double a, double b, float fa) {
    double x[CHAINS];
    int f[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) {
        x[c] = a * (threadIdx.x + c);
        f[c] = __float_as_int(fa * (threadIdx.x + c + 1));
    }
#pragma unroll 1
    for (int i = 0; i < ITERS; ++i)
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) {
            double d;
            asm volatile("cvt.f64.f32 %0, %1;" : "=d"(d) : "f"(__int_as_float(f[c])));
            f[c] += 1;
            x[c] = __fma_rn(x[c], a, d);
#pragma unroll
            for (int k = 1; k < NF; ++k) x[c] = __fma_rn(x[c], a, b);
        }
    double s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += x[c];
    if (s == 12345.678) out[threadIdx.x] = s;
}
// the same mix with the integer widening
template <int NF>
__global__ void k_mix_int(double* out, double a, double b, float fa) {
    double x[CHAINS];
    int f[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) {
        x[c] = a * (threadIdx.x + c);
        f[c] = __float_as_int(fa * (threadIdx.x + c + 1));
    }
#pragma unroll 1
    for (int i = 0; i < ITERS; ++i)
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) {
            const double d = widen_int(f[c]);
            f[c] += 1;
            x[c] = __fma_rn(x[c], a, d);
#pragma unroll
            for (int k = 1; k < NF; ++k) x[c] = __fma_rn(x[c], a, b);
        }
    double s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += x[c];
    if (s == 12345.678) out[threadIdx.x] = s;
}
// shared-memory float loads feeding DFMAs (LDS + F2F + DFMA per step): axis-1 Gaussian's shape
__global__ void k_lds_mix(double* out, double a, float fa) {
    __shared__ float sh[256 + 64];
    for (int t = threadIdx.x; t < 256 + 64; t += 256) sh[t] = fa * (t + 1);
    __syncthreads();
    double x[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) x[c] = a * (threadIdx.x + c);
    int p = threadIdx.x;
#pragma unroll 1
    for (int i = 0; i < ITERS; ++i) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) x[c] = __fma_rn((double)sh[p + c + (i & 31)], a, x[c]);
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += x[c];
    if (s == 12345.678) out[threadIdx.x] = s;
}

template <typename F>
static double time_ms(F launch) {
    cudaEvent_t e0, e1;
    CHECK(cudaEventCreate(&e0));
    CHECK(cudaEventCreate(&e1));
    launch();
    CHECK(cudaDeviceSynchronize());
    double best = 1e30;
    for (int r = 0; r < 5; ++r) {
        CHECK(cudaEventRecord(e0));
        launch();
        CHECK(cudaEventRecord(e1));
        CHECK(cudaEventSynchronize(e1));
        float ms;
        CHECK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    CHECK(cudaGetLastError());
    return best;
}

int main(int argc, char** argv) {
    int dev = 0, sms = 0, khz = 0;
    CHECK(cudaSetDevice(dev));
    CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    CHECK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev));
    cudaDeviceProp prop;
    CHECK(cudaGetDeviceProperties(&prop, dev));
    double* out;
    CHECK(cudaMalloc(&out, 4096));
    const int grid = sms * 8, block = 256;
    const double nthread = (double)grid * block;
    const double steps = nthread * CHAINS * ITERS;   // chain steps per launch
    struct Row { const char* name; double ms; double ops_per_step; double flops_per_op; };
    Row rows[16];
    int n = 0;
    rows[n++] = {"dfma", time_ms([&] { k_dfma<<<grid, block>>>(out, 1.0000001, 1e-9); }), 1, 2};
    rows[n++] = {"dadd", time_ms([&] { k_dadd<<<grid, block>>>(out, 1.0000001, 1e-9); }), 1, 1};
    rows[n++] = {"dmul", time_ms([&] { k_dmul<<<grid, block>>>(out, 1.0000001, 1.0000001); }), 1, 1};
    rows[n++] = {"cvt_f32_f64_pair", time_ms([&] { k_cvt_pair<<<grid, block>>>(out, 1.25f); }), 2, 0};
    rows[n++] = {"cvt_widen", time_ms([&] { k_cvt_widen<<<grid, block>>>(out, 1.25f); }), 1, 0};
    rows[n++] = {"cvt_narrow", time_ms([&] { k_cvt_narrow<<<grid, block>>>(out, 1.25); }), 1, 0};
    rows[n++] = {"widen_int_emulated", time_ms([&] { k_widen_int<<<grid, block>>>(out, 1.25f); }), 1, 0};
    rows[n++] = {"sqrtf_rn", time_ms([&] { k_sqrtf<<<grid, block>>>(out, 1.25f); }), 1, 0};
    rows[n++] = {"divf_rn", time_ms([&] { k_divf<<<grid, block>>>(out, 1.25f); }), 1, 0};
    rows[n++] = {"mix_1cvt_2dfma", time_ms([&] { k_mix<2><<<grid, block>>>(out, 1.0000001, 1e-9, 1.25f); }), 2, 2};
    rows[n++] = {"mix_1cvt_4dfma", time_ms([&] { k_mix<4><<<grid, block>>>(out, 1.0000001, 1e-9, 1.25f); }), 4, 2};
    rows[n++] = {"mix_1cvt_8dfma", time_ms([&] { k_mix<8><<<grid, block>>>(out, 1.0000001, 1e-9, 1.25f); }), 8, 2};
    rows[n++] = {"mixint_1widen_2dfma", time_ms([&] { k_mix_int<2><<<grid, block>>>(out, 1.0000001, 1e-9, 1.25f); }), 2, 2};
    rows[n++] = {"mixint_1widen_4dfma", time_ms([&] { k_mix_int<4><<<grid, block>>>(out, 1.0000001, 1e-9, 1.25f); }), 4, 2};
    rows[n++] = {"lds_cvt_dfma", time_ms([&] { k_lds_mix<<<grid, block>>>(out, 1.0000001, 1.25f); }), 1, 2};

    FILE* f = argc > 1 ? fopen(argv[1], "w") : stdout;
    if (!f) f = stdout;
    fprintf(f, "{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz_attr\": %d, \"grid\": %d, \"block\": %d, \"chains\": %d, "
               "\"iters\": %d, \"rows\": {",
            prop.name, sms, khz, grid, block, CHAINS, ITERS);
    for (int i = 0; i < n; ++i) {
        const double ops = steps * rows[i].ops_per_step;           // counted operations (DFMAs in the mixes)
        const double gops = ops / (rows[i].ms * 1e-3) / 1e9;
        const double per_clk_sm = ops / (rows[i].ms * 1e-3) / ((double)khz * 1e3) / sms;
        fprintf(f, "%s\"%s\": {\"ms\": %.4f, \"gops\": %.1f, \"ops_per_clk_per_sm_at_attr_clock\": %.2f, \"tflops\": %.2f}",
                i ? ", " : "", rows[i].name, rows[i].ms, gops, per_clk_sm, gops * rows[i].flops_per_op / 1e3);
    }
    fprintf(f, "}}\n");
    if (f != stdout) fclose(f);
    return 0;
}
