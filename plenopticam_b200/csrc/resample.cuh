// (a) Micro-image resampling around calibrated MLA centroids.
// Replaces lfp_aligner/lfp_local_resampler.py:44-148 (method 'linear') and the uint16 quantisation of
// lfp_aligner/lfp_resampler.py:60 (misc/normalizer.py:43-46,69-78).
//
// Closed form:  P[ly,lx,v,u,ch] = bilin(raw, oy+v+wy, ox+u+wx)   (oy,ox,wy,wx: pcb_lens_t)
//   rec: aligned[ly*M+v, lx*M+u] = P[ly,lx,v,u]
//   hex: aligned[ly*M+v, k*M+u]  = P[ly,x0_k,v,u] + (P[ly,x1_k,v,u] - P[ly,x0_k,v,u]) * w_k
//
// One CTA owns a tile of KT output lens columns of one lens row.  It finds the bounding box of the raw
// pixels its source lenses touch, stages that box ONCE in shared memory as float32 (one int->float conversion
// per raw sample instead of one per use; every raw sample is used ~4.6 times), and then each thread walks one
// output COLUMN (k,u,ch) down its M+1 source rows, so each horizontally interpolated row feeds the two output
// rows it belongs to and every store instruction of a warp is one contiguous run of the aligned image.
// A tile whose box does not fit the shared-memory budget (pathological centroid scatter) takes the same
// arithmetic straight from global memory.
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
    int KT;                       // output lens columns per CTA
    int cap_floats;               // shared-memory tile capacity
};

enum { RS_MINMAX = 0, RS_F32 = 1, RS_U16 = 2 };
constexpr int RS_THREADS = 256;
constexpr int RS_MAX_SRC = 64;    // source lenses per CTA kept in shared memory

template <int MODE>
struct ResampleOut {
    float *f32;
    uint16_t *u16;
    float q_mn, q_mx;
    bool q_scale;
    __device__ __forceinline__ void put(size_t o, float val) const {
        if (MODE == RS_F32) f32[o] = val;
        if (MODE == RS_U16) u16[o] = quantize_u16(val, q_mn, q_mx, q_scale);
    }
};

// One output column from a row-addressable source: src(r) points at the sample (row r, column u, ch) of the
// patch; src(r)[C] is its right neighbour.
template <int MODE, typename Ld0, typename Ld1>
__device__ __forceinline__ void resample_column(const ResampleArgs &p, const ResampleOut<MODE> &out, int ly, int col,
                                                bool ok0, bool ok1, bool two, float wx0, float wy0, float wx1,
                                                float wy1, float w, Ld0 ld0, Ld1 ld1, float &mn, float &mx) {
    const int M = p.M;
    float h0p = 0.f, h1p = 0.f;
    if (ok0) { float a, b; ld0(0, a, b); h0p = fmaf(b - a, wx0, a); }
    if (ok1) { float a, b; ld1(0, a, b); h1p = fmaf(b - a, wx1, a); }
    const size_t out_row0 = (size_t)ly * M;
#pragma unroll 5
    for (int r = 1; r <= M; ++r) {
        float h0 = 0.f, h1 = 0.f;
        if (ok0) { float a, b; ld0(r, a, b); h0 = fmaf(b - a, wx0, a); }
        if (ok1) { float a, b; ld1(r, a, b); h1 = fmaf(b - a, wx1, a); }
        float val = fmaf(h0 - h0p, wy0, h0p);
        if (two) {
            const float v1 = fmaf(h1 - h1p, wy1, h1p);
            val = fmaf(v1 - val, w, val);
        }
        h0p = h0; h1p = h1;
        const int v = p.flip ? (M - r) : (r - 1);
        out.put((out_row0 + v) * (size_t)p.ncols + col, val);
        mn = fminf(mn, val);
        mx = fmaxf(mx, val);
    }
}

template <typename TIn, int MODE, int CT>
__global__ void __launch_bounds__(RS_THREADS, 4)
resample_tile_kernel(ResampleArgs p, float *__restrict__ out_f32, uint16_t *__restrict__ out_u16,
                     pcb_minmax_t *mm_rw, const pcb_minmax_t *mm_ro) {
    extern __shared__ __align__(16) float tile[];
    __shared__ pcb_lens_t s_lens[RS_MAX_SRC];
    __shared__ int s_box[4];      // ymin, ymax, xmin, xmax over the valid source lenses (oy / ox)
    __shared__ int s_src[2];      // first / last source lens column

    const int C = CT ? CT : p.C;
    const int M = p.M;
    const int ly = blockIdx.y;
    const int k0 = blockIdx.x * p.KT;
    const int kn = min(p.KT, p.Xs - k0);
    const int tid = threadIdx.x;
    const pcb_hexcol_t *hc = p.hexcol ? p.hexcol + ((ly + p.hex_odd) & 1) * p.Xs : nullptr;

    ResampleOut<MODE> out;
    out.f32 = out_f32; out.u16 = out_u16; out.q_mn = 0.f; out.q_mx = 0.f; out.q_scale = false;
    if (MODE == RS_U16) {
        out.q_mn = key_to_float(mm_ro->kmin);
        out.q_mx = key_to_float(mm_ro->kmax);
        out.q_scale = (out.q_mx != out.q_mn) && (out.q_mx != 0.f);
    }

    // ---- source lens range of this tile (hexcol x0/x1 are non-decreasing in k)
    if (tid == 0) {
        int xa = k0, xb = k0 + kn - 1;
        if (hc) {
            xa = hc[k0].x0;
            xb = hc[k0 + kn - 1].x1;
            if (hc[k0 + kn - 1].w == 0.f) xb = hc[k0 + kn - 1].x0;
        }
        s_src[0] = xa;
        s_src[1] = xb;
        s_box[0] = INT32_MAX; s_box[1] = INT32_MIN; s_box[2] = INT32_MAX; s_box[3] = INT32_MIN;
    }
    __syncthreads();
    const int xa = s_src[0];
    const int nsrc = s_src[1] - xa + 1;
    const bool lens_in_smem = nsrc <= RS_MAX_SRC;
    if (lens_in_smem && tid < nsrc) {
        const pcb_lens_t L = p.lens[ly * p.LX + xa + tid];
        s_lens[tid] = L;
        if (L.oy != PCB_LENS_INVALID) {
            atomicMin(&s_box[0], L.oy); atomicMax(&s_box[1], L.oy);
            atomicMin(&s_box[2], L.ox); atomicMax(&s_box[3], L.ox);
        }
    }
    __syncthreads();
    const int ymin = s_box[0], xmin = s_box[2];
    const int R = s_box[1] - ymin + M + 1;               // staged rows
    const int Wt = (s_box[3] - xmin + M + 1) * C;        // staged elements per row
    const bool any_valid = s_box[1] != INT32_MIN;
    const bool use_tile = lens_in_smem && any_valid && (long long)R * Wt <= p.cap_floats;

    if (use_tile) {
        // stage: warps take rows, lanes run along the contiguous (x, ch) axis
        const int warp = tid >> 5, lane = tid & 31;
        for (int rr = warp; rr < R; rr += RS_THREADS / 32) {
            const TIn *src = (const TIn *)p.raw + ((size_t)(ymin + rr) * p.W + xmin) * C;
            float *dst = tile + rr * Wt;
            for (int e = lane; e < Wt; e += 32) dst[e] = ld_as_float(src + e);
        }
    }
    __syncthreads();

    float mn = INFINITY, mx = -INFINITY;
    const int ncols_tile = kn * M * C;
    for (int lc = tid; lc < ncols_tile; lc += RS_THREADS) {
        const int px = lc / C;
        const int ch = lc - px * C;
        const int kl = px / M;
        const int u = px - kl * M;
        const int k = k0 + kl;
        const int us = p.flip ? (M - 1 - u) : u;
        const int col = k0 * M * C + lc;

        int x0 = k, x1 = k;
        float w = 0.f;
        if (hc) { const pcb_hexcol_t h = hc[k]; x0 = h.x0; x1 = h.x1; w = h.w; }
        const bool two = (w != 0.f);
        const pcb_lens_t L0 = lens_in_smem ? s_lens[x0 - xa] : p.lens[ly * p.LX + x0];
        pcb_lens_t L1 = L0;
        if (two) L1 = lens_in_smem ? s_lens[x1 - xa] : p.lens[ly * p.LX + x1];
        const bool ok0 = L0.oy != PCB_LENS_INVALID;
        const bool ok1 = two && (L1.oy != PCB_LENS_INVALID);

        if (use_tile) {
            const float *t0 = tile + (ok0 ? (L0.oy - ymin) * Wt + (L0.ox - xmin + us) * C + ch : 0);
            const float *t1 = tile + (ok1 ? (L1.oy - ymin) * Wt + (L1.ox - xmin + us) * C + ch : 0);
            auto ld0 = [&](int r, float &a, float &b) { a = t0[r * Wt]; b = t0[r * Wt + C]; };
            auto ld1 = [&](int r, float &a, float &b) { a = t1[r * Wt]; b = t1[r * Wt + C]; };
            resample_column<MODE>(p, out, ly, col, ok0, ok1, two, L0.wx, L0.wy, L1.wx, L1.wy, w, ld0, ld1, mn, mx);
        } else {
            const size_t stride = (size_t)p.W * C;
            const TIn *g0 = (const TIn *)p.raw + ((size_t)(ok0 ? L0.oy : 0) * p.W + (ok0 ? L0.ox : 0) + us) * C + ch;
            const TIn *g1 = (const TIn *)p.raw + ((size_t)(ok1 ? L1.oy : 0) * p.W + (ok1 ? L1.ox : 0) + us) * C + ch;
            auto ld0 = [&](int r, float &a, float &b) { a = ld_as_float(g0 + r * stride); b = ld_as_float(g0 + r * stride + C); };
            auto ld1 = [&](int r, float &a, float &b) { a = ld_as_float(g1 + r * stride); b = ld_as_float(g1 + r * stride + C); };
            resample_column<MODE>(p, out, ly, col, ok0, ok1, two, L0.wx, L0.wy, L1.wx, L1.wy, w, ld0, ld1, mn, mx);
        }
    }
    if (MODE != RS_U16 && mm_rw != nullptr) block_minmax_commit(mn, mx, mm_rw);
}

// Tile geometry chosen on the host from the nominal lens pitch; the kernel guards the real box per CTA.
static void resample_geometry(ResampleArgs &a, size_t *smem_bytes) {
    const int cap_bytes = 56 * 1024;                       // 4 CTAs of 256 threads per SM
    const double pitch = (double)a.W / (double)a.LX;      // nominal lens pitch in raw pixels
    const double src_per_out = (double)a.LX / (double)a.Xs;
    const int rows = a.M + 1 + 4;                          // M+1 rows plus centroid scatter allowance
    int KT = 16;
    for (; KT > 1; --KT) {
        const double cols_px = KT * src_per_out * pitch + pitch + a.M + 4;
        if ((double)rows * cols_px * a.C * 4.0 <= cap_bytes) break;
    }
    a.KT = KT;
    a.cap_floats = cap_bytes / 4;
    *smem_bytes = cap_bytes;
}

template <typename TIn, int MODE>
static int launch_resample_t(ResampleArgs a, float *out_f32, uint16_t *out_u16, pcb_minmax_t *mm_rw,
                             const pcb_minmax_t *mm_ro, cudaStream_t st) {
    size_t smem;
    resample_geometry(a, &smem);
    dim3 block(RS_THREADS), grid((a.Xs + a.KT - 1) / a.KT, a.LY);
    if (a.C == 3) {
        PCB_CUDA(cudaFuncSetAttribute(resample_tile_kernel<TIn, MODE, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        resample_tile_kernel<TIn, MODE, 3><<<grid, block, smem, st>>>(a, out_f32, out_u16, mm_rw, mm_ro);
    } else if (a.C == 1) {
        PCB_CUDA(cudaFuncSetAttribute(resample_tile_kernel<TIn, MODE, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        resample_tile_kernel<TIn, MODE, 1><<<grid, block, smem, st>>>(a, out_f32, out_u16, mm_rw, mm_ro);
    } else {
        PCB_CUDA(cudaFuncSetAttribute(resample_tile_kernel<TIn, MODE, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        resample_tile_kernel<TIn, MODE, 0><<<grid, block, smem, st>>>(a, out_f32, out_u16, mm_rw, mm_ro);
    }
    PCB_LAUNCH_OK();
    return PCB_OK;
}

template <int MODE>
static int launch_resample(const ResampleArgs &a, int raw_dtype, float *out_f32, uint16_t *out_u16,
                           pcb_minmax_t *mm_rw, const pcb_minmax_t *mm_ro, cudaStream_t st) {
    switch (raw_dtype) {
    case PCB_U16: return launch_resample_t<uint16_t, MODE>(a, out_f32, out_u16, mm_rw, mm_ro, st);
    case PCB_F32: return launch_resample_t<float, MODE>(a, out_f32, out_u16, mm_rw, mm_ro, st);
    case PCB_U8:  return launch_resample_t<uint8_t, MODE>(a, out_f32, out_u16, mm_rw, mm_ro, st);
    }
    return fail(PCB_ERR_INVALID, "%s: unsupported raw dtype %lld", "pcb_resample", raw_dtype);
}

}  // namespace pcb
