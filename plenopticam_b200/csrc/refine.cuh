// (c') Sub-pixel "refocus refinement" (opt_refi): replaces the up-sample / shift / sum / down-sample chain of
// lfp_refocuser/lfp_shiftandsum.py:93-135 + misc/data_proc.py:57-92.
//
// The chain is linear and separable (see lfp_refocuser/refinement.py):
//     R_a = (1/M) * sum_{j,i active}  A_{a(c-j)} . V_ji . B_{a(c-i)}^T
// with A_s, B_s narrow-band operators prepared on the host in float64.  Per slice two banded passes:
//   horizontal  H[j][r][x][c] = sum_{i active} sum_m  Bx[ix[i]][g(x)][m][q(x)] * vp[j,i,r, first+m, c]
//   vertical    out[y][x][c]  = (1/M) sum_{j active} sum_m  Ay[iy[j]][g(y)][m][q(y)] * H[j][first+m][x][c]
// Outputs come in groups of 4 that share one sample window, so each loaded sample feeds 4 outputs x 3 channels
// (12 FMAs per 3 shared-memory loads + one 16-byte weight load).
#pragma once
#include "common.cuh"

namespace pcb {

constexpr int RF_GROUP = 4;
constexpr int RFH_THREADS = 160;

// grid (S, M): one CTA per (image row r, view-row j); thread = one group of 4 output pixels, all 3 channels.
template <typename TIn>
__global__ void __launch_bounds__(RFH_THREADS)
refine_rows_kernel(const TIn *__restrict__ vp, int M, int S, int T, const float *__restrict__ bx_vals,
                   const int32_t *__restrict__ bx_first, int Kg, const int32_t *__restrict__ ix_n,
                   const uint8_t *__restrict__ view_zeroed, float *__restrict__ H) {
    extern __shared__ __align__(16) float s_rows[];      // [nact][T*3]
    __shared__ int s_act[128];
    __shared__ int s_nact;
    const int r = blockIdx.x, j = blockIdx.y;
    const int E = T * 3;
    if (threadIdx.x == 0) {
        int n = 0;
        for (int i = 0; i < M; ++i)
            if (!view_zeroed[j * M + i]) s_act[n++] = i;
        s_nact = n;
    }
    __syncthreads();
    const int nact = s_nact;
    if (nact == 0) return;
    for (int ii = 0; ii < nact; ++ii) {
        const TIn *src = vp + ((size_t)(j * M + s_act[ii]) * S + r) * E;
        float *dst = s_rows + (size_t)ii * E;
        for (int e = threadIdx.x; e < E; e += RFH_THREADS) dst[e] = ld_as_float(src + e);
    }
    __syncthreads();

    const int G = (T + RF_GROUP - 1) / RF_GROUP;
    for (int g = threadIdx.x; g < G; g += RFH_THREADS) {
        float acc[RF_GROUP][3];
#pragma unroll
        for (int q = 0; q < RF_GROUP; ++q) acc[q][0] = acc[q][1] = acc[q][2] = 0.f;
        for (int ii = 0; ii < nact; ++ii) {
            const int sidx = ix_n[s_act[ii]];
            const int f = bx_first[(size_t)sidx * G + g];
            const float4 *w = reinterpret_cast<const float4 *>(bx_vals) + ((size_t)sidx * G + g) * Kg;
            const float *row = s_rows + (size_t)ii * E;
#pragma unroll 4
            for (int m = 0; m < Kg; ++m) {
                const float4 w4 = __ldg(w + m);
                const float *px = row + min(f + m, T - 1) * 3;
                const float v0 = px[0], v1 = px[1], v2 = px[2];
                acc[0][0] = fmaf(w4.x, v0, acc[0][0]); acc[0][1] = fmaf(w4.x, v1, acc[0][1]); acc[0][2] = fmaf(w4.x, v2, acc[0][2]);
                acc[1][0] = fmaf(w4.y, v0, acc[1][0]); acc[1][1] = fmaf(w4.y, v1, acc[1][1]); acc[1][2] = fmaf(w4.y, v2, acc[1][2]);
                acc[2][0] = fmaf(w4.z, v0, acc[2][0]); acc[2][1] = fmaf(w4.z, v1, acc[2][1]); acc[2][2] = fmaf(w4.z, v2, acc[2][2]);
                acc[3][0] = fmaf(w4.w, v0, acc[3][0]); acc[3][1] = fmaf(w4.w, v1, acc[3][1]); acc[3][2] = fmaf(w4.w, v2, acc[3][2]);
            }
        }
        float *out = H + ((size_t)j * S + r) * E;
#pragma unroll
        for (int q = 0; q < RF_GROUP; ++q) {
            const int x = g * RF_GROUP + q;
            if (x < T) { out[x * 3 + 0] = acc[q][0]; out[x * 3 + 1] = acc[q][1]; out[x * 3 + 2] = acc[q][2]; }
        }
    }
}

// grid (Gy, ceil(E/256)): thread = one (x, c) element of a group of 4 output rows.
__global__ void __launch_bounds__(256)
refine_cols_kernel(const float *__restrict__ H, int M, int S, int E, const float *__restrict__ ay_vals,
                   const int32_t *__restrict__ ay_first, int Kg, const int32_t *__restrict__ iy_n,
                   const uint8_t *__restrict__ view_zeroed, float *__restrict__ out_n, pcb_minmax_t *mm) {
    __shared__ int s_rowact[128];
    const int gy = blockIdx.x;
    const int Gy = gridDim.x;
    const int e = blockIdx.y * blockDim.x + threadIdx.x;
    for (int j = threadIdx.x; j < M; j += blockDim.x) {
        int any = 0;
        for (int i = 0; i < M; ++i) any |= !view_zeroed[j * M + i];
        s_rowact[j] = any;
    }
    __syncthreads();
    float mn = INFINITY, mx = -INFINITY;
    if (e < E) {
        float acc[RF_GROUP] = {0.f, 0.f, 0.f, 0.f};
        for (int j = 0; j < M; ++j) {
            if (!s_rowact[j]) continue;
            const int sidx = iy_n[j];
            const int f = ay_first[(size_t)sidx * Gy + gy];
            const float4 *w = reinterpret_cast<const float4 *>(ay_vals) + ((size_t)sidx * Gy + gy) * Kg;
            const float *col = H + (size_t)j * S * E + e;
#pragma unroll 4
            for (int m = 0; m < Kg; ++m) {
                const float4 w4 = __ldg(w + m);                      // same address for the whole CTA: broadcast
                const float v = __ldg(col + (size_t)min(f + m, S - 1) * E);
                acc[0] = fmaf(w4.x, v, acc[0]);
                acc[1] = fmaf(w4.y, v, acc[1]);
                acc[2] = fmaf(w4.z, v, acc[2]);
                acc[3] = fmaf(w4.w, v, acc[3]);
            }
        }
        const float invM = 1.0f / (float)M;
#pragma unroll
        for (int q = 0; q < RF_GROUP; ++q) {
            const int y = gy * RF_GROUP + q;
            if (y < S) {
                const float v = acc[q] * invM;
                out_n[(size_t)y * E + e] = v;
                mn = fminf(mn, v);
                mx = fmaxf(mx, v);
            }
        }
    }
    if (mm != nullptr) block_minmax_commit(mn, mx, mm);
}

template <typename TIn>
static int launch_refine(const void *vp, int M, int S, int T, int N, const float *ay_vals, const int32_t *ay_first,
                         int Kgy, const float *bx_vals, const int32_t *bx_first, int Kgx, const int32_t *iy,
                         const int32_t *ix, const uint8_t *view_zeroed, float *tmp, float *out, pcb_minmax_t *mm,
                         cudaStream_t st) {
    const int E = T * 3;
    const size_t smem = (size_t)M * E * sizeof(float);
    if (smem > 200 * 1024) return fail(PCB_ERR_UNSUPPORTED, "%s: view rows do not fit shared memory", "pcb_refocus_refi");
    PCB_CUDA(cudaFuncSetAttribute(refine_rows_kernel<TIn>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int Gy = (S + RF_GROUP - 1) / RF_GROUP;
    for (int n = 0; n < N; ++n) {
        refine_rows_kernel<TIn><<<dim3(S, M), RFH_THREADS, smem, st>>>((const TIn *)vp, M, S, T, bx_vals, bx_first, Kgx,
                                                                      ix + (size_t)n * M, view_zeroed, tmp);
        PCB_LAUNCH_OK();
        refine_cols_kernel<<<dim3(Gy, (E + 255) / 256), 256, 0, st>>>(tmp, M, S, E, ay_vals, ay_first, Kgy,
                                                                     iy + (size_t)n * M, view_zeroed,
                                                                     out + (size_t)n * S * E, mm);
        PCB_LAUNCH_OK();
    }
    return PCB_OK;
}

}  // namespace pcb
