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
    int skip;                     // min/max pass of resample_col: tiles that cannot move the bounds are skipped
};

enum { RS_MINMAX = 0, RS_F32 = 1, RS_U16 = 2 };
constexpr int RS_THREADS = 256;
constexpr int RS_MAX_SRC = 64;    // source lenses per CTA kept in shared memory

template <int MODE>
struct ResampleOut {
    float *f32;
    uint16_t *u16;
    QuantU16 qz;
    __device__ __forceinline__ void put(size_t o, float val) const {
        if (MODE == RS_F32) f32[o] = val;
        if (MODE == RS_U16) u16[o] = qz(val);
    }
};

template <typename T> __device__ __forceinline__ float as_float(T v) { return (float)v; }

// One output column.  s0 / s1 point at the sample (row 0, column u, ch) of the two source patches inside a
// row-addressable source (the shared-memory tile or the raw image); the right neighbour is C elements on, the
// next row `stride` elements on.
template <int MODE, typename TS>
__device__ __forceinline__ void resample_column(const ResampleArgs &p, const ResampleOut<MODE> &out, int ly, int col,
                                                int C, bool ok0, bool ok1, bool two, float wx0, float wy0, float wx1,
                                                float wy1, float w, const TS *s0, const TS *s1, size_t stride,
                                                float &mn, float &mx) {
    const int M = p.M;
    float h0p = 0.f, h1p = 0.f;
    if (ok0) { const float a = as_float(s0[0]), b = as_float(s0[C]); h0p = fmaf(b - a, wx0, a); }
    if (ok1) { const float a = as_float(s1[0]), b = as_float(s1[C]); h1p = fmaf(b - a, wx1, a); }
    // output element of patch row v = 0 and the step to the next written row (reversed when flipping)
    size_t o = ((size_t)ly * M + (p.flip ? M - 1 : 0)) * (size_t)p.ncols + col;
    const long long ostep = p.flip ? -(long long)p.ncols : (long long)p.ncols;
#pragma unroll 5
    for (int r = 1; r <= M; ++r) {
        s0 += stride;
        s1 += stride;
        float h0 = 0.f, h1 = 0.f;
        if (ok0) { const float a = as_float(s0[0]), b = as_float(s0[C]); h0 = fmaf(b - a, wx0, a); }
        if (ok1) { const float a = as_float(s1[0]), b = as_float(s1[C]); h1 = fmaf(b - a, wx1, a); }
        float val = fmaf(h0 - h0p, wy0, h0p);
        const float v1 = fmaf(h1 - h1p, wy1, h1p);
        val = two ? fmaf(v1 - val, w, val) : val;
        h0p = h0; h1p = h1;
        out.put(o, val);
        o += ostep;
        mn = fminf(mn, val);
        mx = fmaxf(mx, val);
    }
}

// Stage `R` rows of `Wt` elements starting at raw element (ymin*W + xmin)*C into the float tile.
// uint16 rows go through aligned 32-bit loads (two samples per load); RS_STAGE_UNROLL loads stay in flight.
constexpr int RS_STAGE_UNROLL = 12;
template <typename TIn>
__device__ __forceinline__ void stage_tile(const ResampleArgs &p, float *tile, int ymin, int xmin, int R, int Wt, int C) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t total_elems = (size_t)p.H * p.W * C;
    for (int rr = warp; rr < R; rr += RS_THREADS / 32) {
        const size_t gE = ((size_t)(ymin + rr) * p.W + xmin) * C;
        const TIn *src = (const TIn *)p.raw + gE;
        float *dst = tile + rr * Wt;
        bool done = false;
        if constexpr (sizeof(TIn) == 2) {
            if ((reinterpret_cast<uintptr_t>(p.raw) & 3) == 0) {
                const int g = (int)(gE & 1);
                const uint32_t *wsrc = reinterpret_cast<const uint32_t *>(src - g);
                int nw = (Wt + g + 1) >> 1;
                if (gE - g + 2 * (size_t)nw > total_elems) --nw;          // never read past the image
                for (int base = lane; base < nw; base += 32 * RS_STAGE_UNROLL) {
                    uint32_t v[RS_STAGE_UNROLL];
#pragma unroll
                    for (int u = 0; u < RS_STAGE_UNROLL; ++u) {
                        const int m = base + u * 32;
                        if (m < nw) v[u] = __ldg(wsrc + m);
                    }
#pragma unroll
                    for (int u = 0; u < RS_STAGE_UNROLL; ++u) {
                        const int m = base + u * 32;
                        if (m < nw) {
                            const int e = 2 * m - g;
                            const float f0 = (float)(v[u] & 0xffffu), f1 = (float)(v[u] >> 16);
                            if (g == 0 && !(Wt & 1) && e + 1 < Wt) {
                                *reinterpret_cast<float2 *>(dst + e) = make_float2(f0, f1);   // 8-byte aligned pair
                            } else {
                                if (e >= 0) dst[e] = f0;
                                if (e + 1 < Wt) dst[e + 1] = f1;
                            }
                        }
                    }
                }
                for (int e = 2 * nw - g + lane; e < Wt; e += 32) dst[e] = ld_as_float(src + e);   // tail sample
                done = true;
            }
        }
        if (!done) {
            for (int base = lane; base < Wt; base += 32 * RS_STAGE_UNROLL) {
                float v[RS_STAGE_UNROLL];
#pragma unroll
                for (int u = 0; u < RS_STAGE_UNROLL; ++u) {
                    const int e = base + u * 32;
                    if (e < Wt) v[u] = ld_as_float(src + e);
                }
#pragma unroll
                for (int u = 0; u < RS_STAGE_UNROLL; ++u) {
                    const int e = base + u * 32;
                    if (e < Wt) dst[e] = v[u];
                }
            }
        }
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
    out.f32 = out_f32; out.u16 = out_u16;
    out.qz.init(0.f, 0.f);
    if (MODE == RS_U16) out.qz.init(key_to_float(mm_ro->kmin), key_to_float(mm_ro->kmax));

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

    if (use_tile) stage_tile<TIn>(p, tile, ymin, xmin, R, Wt, C);   // warps take rows, lanes run along (x, ch)
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
            resample_column<MODE, float>(p, out, ly, col, C, ok0, ok1, two, L0.wx, L0.wy, L1.wx, L1.wy, w, t0, t1,
                                         (size_t)Wt, mn, mx);
        } else {
            const size_t stride = (size_t)p.W * C;
            const TIn *g0 = (const TIn *)p.raw + ((size_t)(ok0 ? L0.oy : 0) * p.W + (ok0 ? L0.ox : 0) + us) * C + ch;
            const TIn *g1 = (const TIn *)p.raw + ((size_t)(ok1 ? L1.oy : 0) * p.W + (ok1 ? L1.ox : 0) + us) * C + ch;
            resample_column<MODE, TIn>(p, out, ly, col, C, ok0, ok1, two, L0.wx, L0.wy, L1.wx, L1.wy, w, g0, g1, stride,
                                       mn, mx);
        }
    }
    if (MODE != RS_U16 && mm_rw != nullptr) block_minmax_commit(mn, mx, mm_rw);
}

// Tile geometry chosen on the host from the nominal lens pitch; the kernel guards the real box per CTA.
static void resample_geometry(ResampleArgs &a, size_t *smem_bytes) {
    const int cap_bytes = 54 * 1024;                       // 4 CTAs of 256 threads per SM (with the static smem)
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
