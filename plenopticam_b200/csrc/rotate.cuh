// (a') Cubic-spline image rotation.
// Replaces LfpRotator._rotate_img (lfp_aligner/lfp_rotator.py:129-145):
//   scipy.ndimage.rotate(img, deg, reshape=False, order=3)   (mode='constant', cval=0, prefilter=True)
//
// scipy = spline_filter (cubic B-spline coefficients, IIR with pole z1 = sqrt(3)-2, mirror boundary) followed
// by a 4x4-tap B-spline evaluation at the rotated coordinate.  The two-sided IIR's impulse response is
// h[k] = sqrt(3) * z1^|k|, i.e. it decays by 0.268 per sample: truncating at |k| <= PF_K = 14 leaves a
// relative error of 1e-8, below float32 resolution, and turns the only serial step of the path into a
// separable FIR over the mirror-extended signal that parallelises over every pixel.
#pragma once
#include "common.cuh"

namespace pcb {

constexpr int PF_K = 14;
struct PfTaps { float h[PF_K + 1]; };     // h[0..K], passed by value (kernel-parameter constant bank)

__device__ __forceinline__ int mirror_idx(int i, int n) {
    // whole-sample symmetric extension, period 2n-2 (scipy 'mirror')
    if (n == 1) return 0;
    const int period = 2 * n - 2;
    i = i % period;
    if (i < 0) i += period;
    return i < n ? i : period - i;
}

// horizontal pass: out[y, x, c] = sum_k h[|k|] in[y, mirror(x+k), c]
template <typename TIn>
__global__ void __launch_bounds__(256)
prefilter_rows_kernel(const TIn *__restrict__ in, float *__restrict__ out, int H, int W, int C, PfTaps taps) {
    extern __shared__ float s_row[];   // (tile_w + 2K) * C
    const int y = blockIdx.y;
    const int tile_w = blockDim.x;     // pixels per CTA
    const int x0 = blockIdx.x * tile_w;
    const int span = (tile_w + 2 * PF_K) * C;
    const TIn *src = in + (size_t)y * W * C;
    for (int idx = threadIdx.x; idx < span; idx += blockDim.x) {
        const int px = idx / C, c = idx - px * C;
        const int x = mirror_idx(x0 - PF_K + px, W);
        s_row[idx] = ld_as_float(src + (size_t)x * C + c);
    }
    __syncthreads();
    const int x = x0 + threadIdx.x;
    if (x < W) {
        for (int c = 0; c < C; ++c) {
            const float *p = s_row + (threadIdx.x + PF_K) * C + c;
            float acc = taps.h[0] * p[0];
#pragma unroll
            for (int k = 1; k <= PF_K; ++k) acc = fmaf(taps.h[k], p[k * C] + p[-k * C], acc);
            out[((size_t)y * W + x) * C + c] = acc;
        }
    }
}

// vertical pass on the row-filtered image; threads run along the contiguous (x, c) axis, each CTA walks a
// strip of rows keeping a sliding register window of 2K+1 samples.
constexpr int PF_ROWS = 32;   // output rows per CTA
__global__ void __launch_bounds__(256)
prefilter_cols_kernel(const float *__restrict__ in, float *__restrict__ out, int H, int W, int C, PfTaps taps) {
    const size_t rowlen = (size_t)W * C;
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= rowlen) return;
    const int y0 = blockIdx.y * PF_ROWS;
    float win[2 * PF_K + 1];
#pragma unroll
    for (int k = 0; k < 2 * PF_K; ++k) win[k + 1] = __ldg(in + (size_t)mirror_idx(y0 - PF_K + k, H) * rowlen + e);
    const int y1 = min(y0 + PF_ROWS, H);
    for (int y = y0; y < y1; ++y) {
#pragma unroll
        for (int k = 0; k < 2 * PF_K; ++k) win[k] = win[k + 1];
        win[2 * PF_K] = __ldg(in + (size_t)mirror_idx(y + PF_K, H) * rowlen + e);
        float acc = taps.h[0] * win[PF_K];
#pragma unroll
        for (int k = 1; k <= PF_K; ++k) acc = fmaf(taps.h[k], win[PF_K + k] + win[PF_K - k], acc);
        out[(size_t)y * rowlen + e] = acc;
    }
}

// out[y,x,c] = sum_{p,q} By[p] Bx[q] coef[mirror(fy-1+p), mirror(fx-1+q), c]  if (Y,X) inside, else 0
__global__ void __launch_bounds__(256)
rotate_sample_kernel(const float *__restrict__ coef, float *__restrict__ out, int H, int W, int C,
                     double cs, double sn, double cy, double cx) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (x >= W) return;
    const double dy = (double)y - cy, dx = (double)x - cx;
    const double Y = cs * dy + sn * dx + cy;
    const double X = -sn * dy + cs * dx + cx;
    float *o = out + ((size_t)y * W + x) * C;
    if (!(Y >= 0.0 && Y <= (double)(H - 1) && X >= 0.0 && X <= (double)(W - 1))) {
        for (int c = 0; c < C; ++c) o[c] = 0.f;
        return;
    }
    const double fyd = floor(Y), fxd = floor(X);
    const float ty = (float)(Y - fyd), tx = (float)(X - fxd);
    const int iy = (int)fyd, ix = (int)fxd;
    float wy[4], wx[4];
    {
        const float t = ty, t2 = t * t, t3 = t2 * t, u = 1.f - t;
        wy[0] = u * u * u * (1.f / 6.f);
        wy[1] = (3.f * t3 - 6.f * t2 + 4.f) * (1.f / 6.f);
        wy[2] = (-3.f * t3 + 3.f * t2 + 3.f * t + 1.f) * (1.f / 6.f);
        wy[3] = t3 * (1.f / 6.f);
    }
    {
        const float t = tx, t2 = t * t, t3 = t2 * t, u = 1.f - t;
        wx[0] = u * u * u * (1.f / 6.f);
        wx[1] = (3.f * t3 - 6.f * t2 + 4.f) * (1.f / 6.f);
        wx[2] = (-3.f * t3 + 3.f * t2 + 3.f * t + 1.f) * (1.f / 6.f);
        wx[3] = t3 * (1.f / 6.f);
    }
    int ry[4], rx[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        ry[p] = mirror_idx(iy - 1 + p, H);
        rx[p] = mirror_idx(ix - 1 + p, W);
    }
    for (int c = 0; c < C; ++c) {
        float acc = 0.f;
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            const float *rowp = coef + (size_t)ry[p] * W * C + c;
            float racc = 0.f;
#pragma unroll
            for (int q = 0; q < 4; ++q) racc = fmaf(wx[q], __ldg(rowp + (size_t)rx[q] * C), racc);
            acc = fmaf(wy[p], racc, acc);
        }
        o[c] = acc;
    }
}

static PfTaps prefilter_taps() {
    PfTaps t;
    const double z1 = sqrt(3.0) - 2.0;
    double g = sqrt(3.0);   // -6 z1 / (1 - z1^2)
    for (int k = 0; k <= PF_K; ++k) { t.h[k] = (float)g; g *= z1; }
    return t;
}

template <typename TIn>
static int launch_prefilter(const void *img, int H, int W, int C, float *tmp, float *coef, cudaStream_t st) {
    const PfTaps taps = prefilter_taps();
    {
        const int tile_w = 256;
        dim3 grid((W + tile_w - 1) / tile_w, H), block(tile_w);
        const size_t smem = (size_t)(tile_w + 2 * PF_K) * C * sizeof(float);
        prefilter_rows_kernel<TIn><<<grid, block, smem, st>>>((const TIn *)img, tmp, H, W, C, taps);
        PCB_LAUNCH_OK();
    }
    {
        const size_t rowlen = (size_t)W * C;
        dim3 grid((unsigned)((rowlen + 255) / 256), (H + PF_ROWS - 1) / PF_ROWS), block(256);
        prefilter_cols_kernel<<<grid, block, 0, st>>>(tmp, coef, H, W, C, taps);
        PCB_LAUNCH_OK();
    }
    return PCB_OK;
}

}  // namespace pcb
