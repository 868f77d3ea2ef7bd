// De-vignetting by white-image division — first of the "next" rows: it sits immediately before the
// resampler in LfpAligner (lfp_aligner/top_level.py:51-57) and is purely element-wise / streaming.
// Replaces LfpDevignetter.__init__ / main / wht_img_divide / _estimate_noise_level
// (lfp_aligner/lfp_devignetter.py:33-57,59-93,160-176):
//
//   wgray = rgb2gry(wht)  (3-channel white image; BT.709 stand-in)  or  wht
//   wn    = wgray / percentile(wgray, 99.9)                                    (pcb_percentile_f64 + normalize)
//   out   = lfp / wn  where wn != 0, else 0                                    (float64 divisions: bit-exact)
//   noise = std(wn - convolve2d(wn, gauss(int(pitch), sigma=1), 'same'))       (separable, zero padded)
//
// All arithmetic is float64 with explicit _rn intrinsics (no FMA contraction) wherever the reference's numpy
// expression is a plain IEEE operation, so `lfp_img` / `wht_img` are bit-identical to the reference; the noise
// level (a decision scalar) is accumulated in one pass (sum, sum of squares) and agrees to ~1e-12 relative.
#pragma once
#include "common.cuh"

namespace pcb {

template <typename T> __device__ __forceinline__ double ld_as_double(const T *p) { return (double)__ldg(p); }

// wgray[i] = C == 3 ? ((r*0.2126 + g*0.7152) + b*0.0722) : v       (numpy: (img[..., :3] * w).sum(-1))
template <typename T>
__global__ void __launch_bounds__(256)
devig_white_kernel(const T *__restrict__ wht, int64_t npx, int C, double *__restrict__ wgray) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < npx; i += stride) {
        double v;
        if (C == 3) {
            const double r = ld_as_double(wht + 3 * i), g = ld_as_double(wht + 3 * i + 1), b = ld_as_double(wht + 3 * i + 2);
            v = __dadd_rn(__dadd_rn(__dmul_rn(r, 0.2126), __dmul_rn(g, 0.7152)), __dmul_rn(b, 0.0722));
        } else {
            v = ld_as_double(wht + i * C);
        }
        wgray[i] = v;
    }
}

__global__ void __launch_bounds__(256)
devig_normalize_kernel(double *__restrict__ w, int64_t n, const double *__restrict__ p) {
    const double d = *p;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) w[i] = __ddiv_rn(w[i], d);
}

// out[i, c] = wn[i] != 0 ? lfp[i, c] / wn[i] : 0
template <typename TIn, typename TOut>
__global__ void __launch_bounds__(256)
devig_divide_kernel(const TIn *__restrict__ lfp, int64_t npx, int C, const double *__restrict__ wn,
                    TOut *__restrict__ out) {
    const int64_t n = npx * C;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += stride) {
        const double w = __ldg(wn + e / C);
        const double v = ld_as_double(lfp + e);
        out[e] = (TOut)(w != 0.0 ? __ddiv_rn(v, w) : 0.0);
    }
}

// ---- noise level: separable Gaussian (zero padded 'same'), then sum / sum of squares of (w - flt) -------------
constexpr int DV_MAXG = 64;
struct DevigGauss { double g[DV_MAXG]; int len; };

__global__ void __launch_bounds__(256)
devig_blur_rows_kernel(const double *__restrict__ w, int H, int W, DevigGauss gk, double *__restrict__ tmp) {
    const int y = blockIdx.y;
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= W) return;
    const int half = gk.len / 2;
    const double *row = w + (size_t)y * W;
    double acc = 0.0;
    for (int k = 0; k < gk.len; ++k) {
        const int xx = x + half - k;                     // convolution: flipped kernel (symmetric anyway)
        if (xx >= 0 && xx < W) acc += gk.g[k] * __ldg(row + xx);
    }
    tmp[(size_t)y * W + x] = acc;
}

__global__ void __launch_bounds__(256)
devig_blur_cols_stats_kernel(const double *__restrict__ w, const double *__restrict__ tmp, int H, int W, DevigGauss gk,
                             double *__restrict__ stats /* [0] sum, [1] sum of squares */) {
    __shared__ double s_a[8], s_b[8];
    const int y = blockIdx.y;
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    double d = 0.0;
    if (x < W) {
        const int half = gk.len / 2;
        double acc = 0.0;
        for (int k = 0; k < gk.len; ++k) {
            const int yy = y + half - k;
            if (yy >= 0 && yy < H) acc += gk.g[k] * __ldg(tmp + (size_t)yy * W + x);
        }
        d = __ldg(w + (size_t)y * W + x) - acc;
    }
    double a = d, b = d * d;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { s_a[warp] = a; s_b[warp] = b; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double ta = 0.0, tb = 0.0;
        for (int k = 0; k < (int)(blockDim.x >> 5); ++k) { ta += s_a[k]; tb += s_b[k]; }
        atomicAdd(stats, ta);
        atomicAdd(stats + 1, tb);
    }
}

__global__ void devig_stats_finish_kernel(double *stats, double n) {
    const double mean = stats[0] / n;
    const double var = stats[1] / n - mean * mean;
    stats[2] = sqrt(var > 0.0 ? var : 0.0);
}

static inline int devig_grid(int64_t n) {
    int64_t b = (n + 256 * 4 - 1) / (256 * 4);
    return (int)(b < 1 ? 1 : (b > 148 * 16 ? 148 * 16 : b));
}

}  // namespace pcb
