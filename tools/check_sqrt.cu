// Exhaustive check (run on the GPU) of the lean float32 square root used by k_gh_nms
// (image_fused.cu: fz_sqrt): rsqrt.approx.ftz + two multiplications + two FMAs must equal the IEEE
// result __fsqrt_rn for EVERY float32 in [2^-100, 2^100] (the range fz_sqrt uses it on; zero is a
// select, everything else takes sqrtf).  Prints the number of mismatches (must be 0).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o tools/bin/check_sqrt tools/check_sqrt.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>

__device__ __forceinline__ float lean_sqrt(float x) {
    const float xc = fmaxf(x, 7.888609052e-31f);
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(xc));
    const float sq = __fmul_rn(x, r), h = __fmul_rn(0.5f, r);
    const float e = __fmaf_rn(-sq, sq, x);
    return __fmaf_rn(e, h, sq);
}

__global__ void k_check(uint32_t lo, uint32_t hi, unsigned long long* bad, uint32_t* first_bad) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    unsigned long long mine = 0;
    for (uint64_t b = (uint64_t)lo + blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; b <= hi; b += stride) {
        const float x = __uint_as_float((uint32_t)b);
        const float a = lean_sqrt(x), w = __fsqrt_rn(x);
        if (__float_as_uint(a) != __float_as_uint(w)) {
            ++mine;
            atomicMin(first_bad, (uint32_t)b);
        }
    }
    if (mine) atomicAdd(bad, mine);
}

int main() {
    unsigned long long* bad;
    uint32_t* first;
    cudaMalloc(&bad, 8);
    cudaMalloc(&first, 4);
    cudaMemset(bad, 0, 8);
    cudaMemset(first, 0xff, 4);
    const float lo = 7.888609052e-31f, hi = 1.267650600e30f;      // 2^-100, 2^100
    uint32_t blo, bhi;
    memcpy(&blo, &lo, 4);
    memcpy(&bhi, &hi, 4);
    k_check<<<148 * 16, 256>>>(blo, bhi, bad, first);
    k_check<<<1, 32>>>(0u, 0u, bad, first);                 // zero goes through the clamp
    unsigned long long nb = 0;
    uint32_t fb = 0;
    cudaMemcpy(&nb, bad, 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(&fb, first, 4, cudaMemcpyDeviceToHost);
    cudaError_t e = cudaDeviceSynchronize();
    printf("{\"checked_from\": \"0x%08x\", \"checked_to\": \"0x%08x\", \"values\": %llu, \"mismatches\": %llu, \"first_bad\": \"0x%08x\", \"cuda\": \"%s\"}\n",
           blo, bhi, (unsigned long long)bhi - blo + 1, nb, fb, cudaGetErrorString(e));
    return nb ? 1 : 0;
}
