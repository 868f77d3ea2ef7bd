// (c) Integer shift-and-sum refocus stack.
// Replaces lfp_refocuser/lfp_shiftandsum.py:77-139 (opt_refi off) with circular_view_aperture
// (lfp_extractor/lfp_viewpoints.py:229-250) and the /M scaling (:89) fused:
//
//   out[n,y,x,c] = (1/M) * sum_{j,i active} vp[j,i, clamp(y + a_n(c0-j), 0, S-1), clamp(x + a_n(c0-i), 0, T-1), c]
//
// uint16/uint8 light fields are accumulated exactly in 32-bit integers (order independent, so the result
// is the correctly rounded quotient), float32 ones in float32 in the reference's j-major / i-minor order.
#pragma once
#include "common.cuh"

namespace pcb {

template <typename T> struct AccOf { typedef uint32_t type; };
template <> struct AccOf<float> { typedef float type; };

__device__ __forceinline__ float finish_acc(uint32_t acc, int M) { return (float)((double)acc / (double)M); }
__device__ __forceinline__ float finish_acc(float acc, int M) { return acc / (float)M; }

constexpr int RF_THREADS = 256;
constexpr int RF_EPT = 8;   // elements per thread: one CTA covers 2048 consecutive (x,c) elements of a row

// ZERO: samples outside the light field count as 0 instead of replicating the edge - the boundary rule of
// LfpShiftAndSum.refo_from_scratch (lfp_refocuser/lfp_shiftandsum.py:141-221: "if index exceeds image boundaries, leave
// pixel empty"), which also sums every view (no circular aperture) and crops the result symmetrically, so that its
// output pixel (h, j) gathers exactly the shifts a (c0 - v), a (c0 - u) of refo_from_vp.
template <typename T, int CT, bool ZERO = false>
__global__ void __launch_bounds__(RF_THREADS)
refocus_int_kernel(const T *__restrict__ vp, const int32_t *__restrict__ a_list,
                   const uint8_t *__restrict__ view_zeroed, float *__restrict__ out,
                   int M, int S, int Tn, int Crt, pcb_minmax_t *mm) {
    typedef typename AccOf<T>::type Acc;
    const int C = CT ? CT : Crt;
    const int y = blockIdx.x, n = blockIdx.y;
    const int e0 = blockIdx.z * (RF_THREADS * RF_EPT);
    const int ne = Tn * C;
    const int a = a_list[n];
    const int c0 = M / 2;

    int xk[RF_EPT], ck[RF_EPT];
    Acc acc[RF_EPT];
#pragma unroll
    for (int k = 0; k < RF_EPT; ++k) {
        int e = e0 + threadIdx.x + k * RF_THREADS;
        e = e < ne ? e : ne - 1;
        xk[k] = e / C;
        ck[k] = e - xk[k] * C;
        acc[k] = 0;
    }

    const size_t view_stride = (size_t)S * Tn * C;
    for (int j = 0; j < M; ++j) {
        int r = y + a * (c0 - j);
        if (ZERO && (r < 0 || r > S - 1)) continue;
        r = max(0, min(S - 1, r));
        for (int i = 0; i < M; ++i) {
            if (view_zeroed[j * M + i]) continue;
            const int dx = a * (c0 - i);
            const T *rowp = vp + (size_t)(j * M + i) * view_stride + (size_t)r * Tn * C;
#pragma unroll
            for (int k = 0; k < RF_EPT; ++k) {
                const int xr = xk[k] + dx;
                const int xs = max(0, min(Tn - 1, xr));
                const Acc v = (Acc)__ldg(rowp + xs * C + ck[k]);
                acc[k] += (ZERO && xr != xs) ? (Acc)0 : v;
            }
        }
    }

    float *orow = out + ((size_t)n * S + y) * ne;
    float mn = INFINITY, mx = -INFINITY;
#pragma unroll
    for (int k = 0; k < RF_EPT; ++k) {
        const int e = e0 + threadIdx.x + k * RF_THREADS;
        if (e < ne) {
            const float v = finish_acc(acc[k], M);
            orow[e] = v;
            mn = fminf(mn, v);
            mx = fmaxf(mx, v);
        }
    }
    if (mm != nullptr) block_minmax_commit(mn, mx, mm);   // stack min/max for Normalizer's falsy-bound fallback
}

template <typename T>
static int launch_refocus_int(const void *vp, int M, int S, int Tn, int C, const int32_t *a_list, int N,
                              const uint8_t *view_zeroed, float *out, pcb_minmax_t *mm, cudaStream_t st,
                              bool zero_boundary = false) {
    const int ne = Tn * C;
    dim3 grid(S, N, (ne + RF_THREADS * RF_EPT - 1) / (RF_THREADS * RF_EPT)), block(RF_THREADS);
    if (zero_boundary) {
        if (C == 3)
            refocus_int_kernel<T, 3, true><<<grid, block, 0, st>>>((const T *)vp, a_list, view_zeroed, out, M, S, Tn, C, mm);
        else
            refocus_int_kernel<T, 0, true><<<grid, block, 0, st>>>((const T *)vp, a_list, view_zeroed, out, M, S, Tn, C, mm);
        PCB_LAUNCH_OK();
        return PCB_OK;
    }
    if (C == 3)
        refocus_int_kernel<T, 3><<<grid, block, 0, st>>>((const T *)vp, a_list, view_zeroed, out, M, S, Tn, C, mm);
    else if (C == 1)
        refocus_int_kernel<T, 1><<<grid, block, 0, st>>>((const T *)vp, a_list, view_zeroed, out, M, S, Tn, C, mm);
    else
        refocus_int_kernel<T, 0><<<grid, block, 0, st>>>((const T *)vp, a_list, view_zeroed, out, M, S, Tn, C, mm);
    PCB_LAUNCH_OK();
    return PCB_OK;
}

}  // namespace pcb
