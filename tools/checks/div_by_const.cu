// Exhaustive check (GPU): Markstein's reciprocal-multiply + one FMA correction equals the IEEE division
// __fdiv_rn((float)v, (float)M) for every v < 2^24 and M = 1..16 (the refocus finalize divides integer sums by M).
//   nvcc -arch=sm_100a -o /tmp/div_by_const tools/checks/div_by_const.cu && /tmp/div_by_const
#include <cstdio>
#include <cuda_runtime.h>

__global__ void check(int M, unsigned long long *bad) {
    const float fM = (float)M, inv = __frcp_rn(fM);
    for (unsigned v = blockIdx.x * blockDim.x + threadIdx.x; v < (1u << 24); v += gridDim.x * blockDim.x) {
        const float a = (float)v;
        const float q = __fmul_rn(a, inv);
        const float r = __fmaf_rn(-fM, q, a);
        const float qq = __fmaf_rn(r, inv, q);
        if (qq != __fdiv_rn(a, fM)) atomicAdd(bad, 1ull);
    }
}

int main() {
    unsigned long long *bad, h = 0, total = 0;
    cudaMalloc(&bad, 8);
    for (int M = 1; M <= 16; ++M) {
        cudaMemset(bad, 0, 8);
        check<<<1184, 256>>>(M, bad);
        cudaMemcpy(&h, bad, 8, cudaMemcpyDeviceToHost);
        printf("M=%2d mismatches=%llu\n", M, h);
        total += h;
    }
    printf("total mismatches: %llu\n", total);
    return total != 0;
}
