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

// horizontal pass: out[y, x, c] = sum_k h[|k|] in[y, mirror(x+k), c]   (any channel count; one output per thread)
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

// The same for C == 3 with a register window: a thread owns PR_P adjacent pixels of one channel and reads their
// PR_P + 2K samples once (4 shared-memory reads per output instead of 29).  PR_P * 3 = 27 is odd, so the lane
// stride is conflict free; results go back through shared memory so that the row leaves in contiguous runs.
constexpr int PR_P = 9;
constexpr int PR_G = 85;                       // pixel groups per CTA: 3 channels x 85 groups = 255 threads
constexpr int PR_TW = PR_P * PR_G;             // 765 pixels per CTA
template <typename TIn>
__global__ void __launch_bounds__(256)
prefilter_rows3_kernel(const TIn *__restrict__ in, float *__restrict__ out, int H, int W, PfTaps taps) {
    __shared__ float s_in[(PR_TW + 2 * PF_K) * 3];
    __shared__ float s_out[PR_TW * 3];
    const int y = blockIdx.y;
    const int x0 = blockIdx.x * PR_TW;
    const int npx = min(PR_TW, W - x0);
    const TIn *src = in + (size_t)y * W * 3;
    const bool interior = x0 >= PF_K && x0 + PR_TW + PF_K <= W;
    if (interior) {
        const TIn *p = src + (size_t)(x0 - PF_K) * 3;
        for (int idx = threadIdx.x; idx < (PR_TW + 2 * PF_K) * 3; idx += 256) s_in[idx] = ld_as_float(p + idx);
    } else {
        for (int idx = threadIdx.x; idx < (PR_TW + 2 * PF_K) * 3; idx += 256) {
            const int px = idx / 3, c = idx - px * 3;
            s_in[idx] = ld_as_float(src + (size_t)mirror_idx(x0 - PF_K + px, W) * 3 + c);
        }
    }
    __syncthreads();
    if (threadIdx.x < 3 * PR_G) {
        const int c = threadIdx.x / PR_G, g = threadIdx.x - c * PR_G;
        const float *p = s_in + g * (PR_P * 3) + c;
        float w[PR_P + 2 * PF_K];
#pragma unroll
        for (int m = 0; m < PR_P + 2 * PF_K; ++m) w[m] = p[m * 3];
        float *o = s_out + g * (PR_P * 3) + c;
#pragma unroll
        for (int j = 0; j < PR_P; ++j) {
            float acc = taps.h[0] * w[j + PF_K];
#pragma unroll
            for (int k = 1; k <= PF_K; ++k) acc = fmaf(taps.h[k], w[j + PF_K + k] + w[j + PF_K - k], acc);
            o[j * 3] = acc;
        }
    }
    __syncthreads();
    float *dst = out + ((size_t)y * W + x0) * 3;
    for (int idx = threadIdx.x; idx < npx * 3; idx += 256) dst[idx] = s_out[idx];
}

// vertical pass on the row-filtered image; threads run along the contiguous (x, c) axis.  A CTA owns a strip of PF_ROWS
// rows: every thread requests the PF_ROWS + 2K samples of its column up front (all loads in flight, static register
// indices: no sliding-window moves) and then evaluates the PF_ROWS outputs from registers.
constexpr int PF_ROWS = 16;   // output rows per CTA
__global__ void __launch_bounds__(256)
prefilter_cols_kernel(const float *__restrict__ in, float *__restrict__ out, int H, int W, int C, PfTaps taps) {
    const size_t rowlen = (size_t)W * C;
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= rowlen) return;
    const int y0 = blockIdx.y * PF_ROWS;
    float win[PF_ROWS + 2 * PF_K];
    if (y0 >= PF_K && y0 + PF_ROWS + PF_K <= H) {             // interior strip: no mirroring
        const float *p = in + (size_t)(y0 - PF_K) * rowlen + e;
#pragma unroll
        for (int k = 0; k < PF_ROWS + 2 * PF_K; ++k) win[k] = __ldg(p + (size_t)k * rowlen);
    } else {
#pragma unroll
        for (int k = 0; k < PF_ROWS + 2 * PF_K; ++k) win[k] = __ldg(in + (size_t)mirror_idx(y0 - PF_K + k, H) * rowlen + e);
    }
#pragma unroll
    for (int r = 0; r < PF_ROWS; ++r) {
        float acc = taps.h[0] * win[r + PF_K];
#pragma unroll
        for (int k = 1; k <= PF_K; ++k) acc = fmaf(taps.h[k], win[r + PF_K + k] + win[r + PF_K - k], acc);
        if (y0 + r < H) out[(size_t)(y0 + r) * rowlen + e] = acc;
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
    if (iy >= 1 && iy + 2 < H && ix >= 1 && ix + 2 < W) {      // interior: no mirroring (almost every pixel)
#pragma unroll
        for (int p = 0; p < 4; ++p) { ry[p] = iy - 1 + p; rx[p] = ix - 1 + p; }
    } else {
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            ry[p] = mirror_idx(iy - 1 + p, H);
            rx[p] = mirror_idx(ix - 1 + p, W);
        }
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
    if (C == 3) {
        dim3 grid((W + PR_TW - 1) / PR_TW, H), block(256);
        prefilter_rows3_kernel<TIn><<<grid, block, 0, st>>>((const TIn *)img, tmp, H, W, taps);
        PCB_LAUNCH_OK();
    } else {
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
