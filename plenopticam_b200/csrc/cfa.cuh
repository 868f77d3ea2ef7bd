// Bayer white balance with highlight protection and soft clipping — "next" row f1: the element-wise
// stage of CfaProcessor that sits before the demosaic / resampler.  Replaces CfaProcessor.safe_bayer_awb =
// _reshape_bayer + set_gains + _correct_bayer_highlights + apply_awb + clamp + soft_clipping
// (lfp_aligner/cfa_processor.py:96-114,142-169,171-203,233-255).
//
// One thread owns one Bayer quad (2 x 2 samples: the reference's 4-channel view of the mosaic):
//   m      = min of the quad (first minimum wins, like argmin),   mb = m * gain[argmin]
//   s      = min(m * mean(white quad), 1)^2                       (weight of the "saturated" estimate)
//   q_i   <- max(q_i, (mb (1 - s) + q_i * max(gain) * s) / gain_i)   where q_i > 0.99 / white_i
//   out_i  = max(q_i * gain_i, 0);   out_i = softclip(out_i / L, 7) * L   (L = 2^-exposure bias, optional)
// The reference's dtype flow is kept operation by operation (float32 samples and white image, float64 gains): the
// explicit _rn intrinsics stop the compiler from contracting a*b+c into an FMA, so everything before the soft
// clipping is bit-identical to numpy; exp / log of the soft clipping are CUDA's (<= 1 ulp from numpy's).
#pragma once
#include "common.cuh"

namespace pcb {

struct CfaArgs {
    const float *bay;      // (H, W) float32 mosaic
    const float *wht;      // (H, W) float32 white image or nullptr
    int H, W;              // even
    double g[4];           // gains of quad positions (0,0) (0,1) (1,0) (1,1)
    double fact;           // max gain
    int soft_clip;
    double max_lum, b, log1pb;   // 2^-exp, exp(7), log(1 + exp(7)) as the host's libm rounds them
    double *out;           // (H, W) float64
};

template <bool HAS_WHT>
__global__ void __launch_bounds__(256)
cfa_safe_awb_kernel(CfaArgs p) {
    const int qx = blockIdx.x * blockDim.x + threadIdx.x;
    const int qy = blockIdx.y;
    if (qx >= p.W / 2) return;
    const size_t o0 = (size_t)(2 * qy) * p.W + 2 * qx, o1 = o0 + p.W;
    const float2 r0 = __ldg(reinterpret_cast<const float2 *>(p.bay + o0));
    const float2 r1 = __ldg(reinterpret_cast<const float2 *>(p.bay + o1));
    float q[4] = {r0.x, r0.y, r1.x, r1.y};
    float m = q[0];
    int mi = 0;
#pragma unroll
    for (int i = 1; i < 4; ++i)
        if (q[i] < m) { m = q[i]; mi = i; }
    const double mb = __dmul_rn((double)m, p.g[mi]);
    double res[4];
    if (HAS_WHT) {
        const float2 w0 = __ldg(reinterpret_cast<const float2 *>(p.wht + o0));
        const float2 w1 = __ldg(reinterpret_cast<const float2 *>(p.wht + o1));
        const float w[4] = {w0.x, w0.y, w1.x, w1.y};
        const float mean = __fdiv_rn(__fadd_rn(__fadd_rn(__fadd_rn(w[0], w[1]), w[2]), w[3]), 4.f);
        float s = __fmul_rn(m, mean);
        if (s > 1.f) s = 1.f;
        s = __fmul_rn(s, s);
        const double om = (double)__fsub_rn(1.f, s), sd = (double)s;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float ref = w[i] != 0.f ? __fdiv_rn(0.99f, w[i]) : INFINITY;
            if (!isfinite(ref)) ref = 0.f;
            float v = q[i];
            if (v > ref) {
                const double inm = __ddiv_rn(__dadd_rn(__dmul_rn(mb, om), __dmul_rn(__dmul_rn((double)v, p.fact), sd)), p.g[i]);
                v = (float)fmax((double)v, inm);
            }
            res[i] = __dmul_rn((double)v, p.g[i]);
        }
    } else {
        double s = (double)m;
        if (s > 1.0) s = 1.0;
        s = __dmul_rn(s, s);
        const double om = __dsub_rn(1.0, s);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float v = q[i];
            if ((double)v > 0.99) {
                const double inm = __ddiv_rn(__dadd_rn(__dmul_rn(mb, om), __dmul_rn(__dmul_rn((double)v, p.fact), s)), p.g[i]);
                v = (float)fmax((double)v, inm);
            }
            res[i] = __dmul_rn((double)v, p.g[i]);
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        double v = res[i];
        if (v < 0.0) v = 0.0;
        if (p.soft_clip) {
            const double e = exp(__dmul_rn(-7.0, __ddiv_rn(v, p.max_lum)));
            const double den = __dadd_rn(1.0, __dmul_rn(p.b, e));
            v = __dmul_rn(__ddiv_rn(log(__ddiv_rn(__dadd_rn(1.0, p.b), den)), p.log1pb), p.max_lum);
        }
        res[i] = v;
    }
    *reinterpret_cast<double2 *>(p.out + o0) = make_double2(res[0], res[1]);
    *reinterpret_cast<double2 *>(p.out + o1) = make_double2(res[2], res[3]);
}

}  // namespace pcb
