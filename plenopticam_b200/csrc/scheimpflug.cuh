// Scheimpflug composite from a refocus stack ("next" row) — replaces the per-pixel Python loop of
// LfpScheimpflug.scheimpflug_from_stack (lfp_refocuser/lfp_scheimpflug.py:68-114):
//   out[y,x,:] = stack[a_map[y,x]][y,x,:],   a_map = a_y[y] | a_x[x] | (a_x[x] + a_y[y]) / 2   (integer mean)
// Pure gather: one read of the selected pixels, one write.
#pragma once
#include "common.cuh"

namespace pcb {

__global__ void __launch_bounds__(256)
scheimpflug_kernel(const float *__restrict__ stack, int N, int S, int E, int C, const int32_t *__restrict__ a_y,
                   const int32_t *__restrict__ a_x, int mode, float *__restrict__ out) {
    const int y = blockIdx.y;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    const int x = e / C;
    int a = mode == 0 ? a_y[y] : (mode == 1 ? a_x[x] : (a_x[x] + a_y[y]) / 2);
    a = min(max(a, 0), N - 1);
    out[(size_t)y * E + e] = __ldg(stack + ((size_t)a * S + y) * E + e);
}

}  // namespace pcb
