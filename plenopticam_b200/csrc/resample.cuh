// (a) Micro-image resampling around calibrated MLA centroids.
// Replaces lfp_aligner/lfp_local_resampler.py:44-148 (method 'linear') and the uint16 quantisation of
// lfp_aligner/lfp_resampler.py:60 (misc/normalizer.py:43-46,69-78).
//
// Closed form:  P[ly,lx,v,u,ch] = bilin(raw, oy+v+wy, ox+u+wx)   (oy,ox,wy,wx: pcb_lens_t)
//   rec: aligned[ly*M+v, lx*M+u] = P[ly,lx,v,u]
//   hex: aligned[ly*M+v, k*M+u]  = P[ly,x0_k,v,u] + (P[ly,x1_k,v,u] - P[ly,x0_k,v,u]) * w_k
//
// Thread mapping: one thread owns one output COLUMN (k,u,ch) of one lens row and walks the M+1 raw rows
// top to bottom, so each horizontally-interpolated row is reused for the two output rows it feeds
// (2 loads per patch per output instead of 4) and every store instruction of a warp is one contiguous
// run in the aligned image.
#pragma once
#include "common.cuh"

namespace pcb {

struct ResampleArgs {
    const void *raw;
    int H, W, C;
    const pcb_lens_t *lens;
    int LY, LX, M;
    const pcb_hexcol_t *hexcol;   // nullptr for rec
    int Xs, hex_odd, flip;
    int ncols;                    // Xs*M*C elements per aligned row
};

enum { RS_MINMAX = 0, RS_F32 = 1, RS_U16 = 2 };

template <typename TIn, int MODE>
__global__ void __launch_bounds__(256)
resample_kernel(ResampleArgs p, float *__restrict__ out_f32, uint16_t *__restrict__ out_u16,
                pcb_minmax_t *mm_rw, const pcb_minmax_t *mm_ro) {
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    const int ly = blockIdx.y;
    const bool active = col < p.ncols;
    const int M = p.M, C = p.C;
    float mn = INFINITY, mx = -INFINITY;

    float q_mn = 0.f, q_mx = 0.f;
    bool q_scale = false;
    if (MODE == RS_U16) {
        q_mn = key_to_float(mm_ro->kmin);
        q_mx = key_to_float(mm_ro->kmax);
        q_scale = (q_mx != q_mn) && (q_mx != 0.f);
    }

    if (active) {
        const int px = col / C;
        const int ch = col - px * C;
        const int k = px / M;
        const int u = px - k * M;
        const int us = p.flip ? (M - 1 - u) : u;

        int x0 = k, x1 = k;
        float w = 0.f;
        if (p.hexcol != nullptr) {
            const pcb_hexcol_t hc = p.hexcol[((ly + p.hex_odd) & 1) * p.Xs + k];
            x0 = hc.x0; x1 = hc.x1; w = hc.w;
        }
        const pcb_lens_t L0 = p.lens[ly * p.LX + x0];
        const bool two = (w != 0.f);
        pcb_lens_t L1 = L0;
        if (two) L1 = p.lens[ly * p.LX + x1];
        const bool ok0 = L0.oy != PCB_LENS_INVALID;
        const bool ok1 = two && (L1.oy != PCB_LENS_INVALID);

        const size_t stride = (size_t)p.W * C;
        const TIn *r0 = (const TIn *)p.raw + ((size_t)(ok0 ? L0.oy : 0) * p.W + (ok0 ? L0.ox : 0) + us) * C + ch;
        const TIn *r1 = (const TIn *)p.raw + ((size_t)(ok1 ? L1.oy : 0) * p.W + (ok1 ? L1.ox : 0) + us) * C + ch;

        float h0p = 0.f, h1p = 0.f;
        if (ok0) { float a = ld_as_float(r0), b = ld_as_float(r0 + C); h0p = fmaf(b - a, L0.wx, a); }
        if (ok1) { float a = ld_as_float(r1), b = ld_as_float(r1 + C); h1p = fmaf(b - a, L1.wx, a); }

        const size_t out_row0 = (size_t)ly * M;
#pragma unroll 4
        for (int r = 1; r <= M; ++r) {
            float h0 = 0.f, h1 = 0.f;
            if (ok0) {
                const TIn *q = r0 + r * stride;
                float a = ld_as_float(q), b = ld_as_float(q + C);
                h0 = fmaf(b - a, L0.wx, a);
            }
            if (ok1) {
                const TIn *q = r1 + r * stride;
                float a = ld_as_float(q), b = ld_as_float(q + C);
                h1 = fmaf(b - a, L1.wx, a);
            }
            float val = fmaf(h0 - h0p, L0.wy, h0p);
            if (two) {
                const float v1 = fmaf(h1 - h1p, L1.wy, h1p);
                val = fmaf(v1 - val, w, val);
            }
            h0p = h0; h1p = h1;
            const int v = p.flip ? (M - r) : (r - 1);
            const size_t o = (out_row0 + v) * (size_t)p.ncols + col;
            if (MODE == RS_F32) out_f32[o] = val;
            if (MODE == RS_U16) out_u16[o] = quantize_u16(val, q_mn, q_mx, q_scale);
            mn = fminf(mn, val);
            mx = fmaxf(mx, val);
        }
    }
    if (MODE != RS_U16 && mm_rw != nullptr) block_minmax_commit(mn, mx, mm_rw);
}

template <int MODE>
static int launch_resample(const ResampleArgs &a, int raw_dtype, float *out_f32, uint16_t *out_u16,
                           pcb_minmax_t *mm_rw, const pcb_minmax_t *mm_ro, cudaStream_t st) {
    dim3 block(256);
    dim3 grid((a.ncols + 255) / 256, a.LY);
    switch (raw_dtype) {
    case PCB_U16: resample_kernel<uint16_t, MODE><<<grid, block, 0, st>>>(a, out_f32, out_u16, mm_rw, mm_ro); break;
    case PCB_F32: resample_kernel<float, MODE><<<grid, block, 0, st>>>(a, out_f32, out_u16, mm_rw, mm_ro); break;
    case PCB_U8:  resample_kernel<uint8_t, MODE><<<grid, block, 0, st>>>(a, out_f32, out_u16, mm_rw, mm_ro); break;
    default: return fail(PCB_ERR_INVALID, "%s: unsupported raw dtype %lld", "pcb_resample", raw_dtype);
    }
    PCB_LAUNCH_OK();
    return PCB_OK;
}

}  // namespace pcb
