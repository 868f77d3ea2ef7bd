// (a) Micro-image resampling, lean column kernel for uint16 RGB exposures (the pipeline's case).
//
// Same decomposition as resample.cuh (one CTA = one lens row x RC_KT output lens columns, the bounding box of the
// raw pixels it needs staged once in shared memory as float32, one thread per output COLUMN (k,u,ch) walking
// down the patch rows so that every store of a warp is one contiguous run), rebuilt to cut the ~45-65 issue
// slots per output sample of resample_tile_kernel to ~15-25:
//
//   * the bilinear interpolation of both source lenses and the hex lerp between them are folded into EIGHT
//     per-thread weights, constant over the column:
//         out[v] = sum_{l in 0,1} sum_{dy,dx in 0,1}  W[l][dy][dx] * raw[o_l + (v+dy, u+dx)]
//     W[0][dy][dx] = (1-w) by_0[dy] bx_0[dx],  W[1][dy][dx] = w by_1[dy] bx_1[dx]   (b = {1-frac, frac}).
//     Per output row: 4 new shared-memory samples (the other 4 are last row's) and 8 FMAs; an invalid lens or a
//     rec grid (single source lens) is just a zero weight set — no branches in the row loop;
//   * the tile has a compile-time row stride, so every LDS of the (fully unrolled, M known) row loop has an
//     immediate offset: no address arithmetic in the loop;
//   * uint16 -> float32 while staging uses the 2^23 magic-number trick (LOP3 + FADD at full rate) instead of I2F.
//
// Float order differs from the separable form by a few ulp (parity bar: 1e-4 of the range; measured ~1e-6).
// A CTA whose box exceeds the tile (pathological centroid scatter) reads the raw image directly.
#pragma once
#include "common.cuh"
#include "resample.cuh"

namespace pcb {

constexpr int RC_THREADS = 320;
constexpr int RC_KT = 14;            // output lens columns per CTA: 14 * 15 * 3 = 630 columns ~ 2 per thread
constexpr int RC_TS = 672;           // tile row stride in floats
constexpr int RC_ROWS = 19;          // tile rows (M + 1 + scatter allowance for M = 15)
constexpr int RC_MAX_SRC = 32;       // source lenses per CTA
constexpr float RC_SKIP_EPS = 4e-6f; // >> the rounding of 8 products of 3 factors + 7 FMAs (~16 * 2^-24 = 1e-6)

__device__ __forceinline__ float u16_to_float_magic(uint32_t v16) {      // v16 < 65536
    return __uint_as_float(0x4B000000u | v16) - 8388608.0f;
}

// Stage rows [ymin, ymin+R) of the uint16 raw image into the float tile.  Returns through `shift` (0/1) the tile
// column of raw element xmin*3: rows are copied from the 32-bit word that contains their first element, so when
// every row starts on the same half of a word (W*3 even and a 4-byte aligned image) the copy is whole words:
// one LDG.32, two PRMT + FADD conversions and one 8-byte STS per two samples, no per-sample guards.
// TRACK: also fold every staged sample into (smn, smx) - the min/max pass uses them to skip tiles that cannot move the
// global bounds.
template <bool TRACK>
__device__ __forceinline__ int rc_stage(const uint16_t *__restrict__ raw, size_t total_elems, int W, float *tile,
                                        int ymin, int xmin, int R, int Wt, float &smn, float &smx) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t gE0 = ((size_t)ymin * W + xmin) * 3;
    const bool words = ((reinterpret_cast<uintptr_t>(raw) & 3) == 0) && ((W * 3) & 1) == 0;
    if (words) {
        const int g = (int)(gE0 & 1);
        const int nw = (Wt + g + 1) >> 1;                                 // words per row (may cover one extra sample)
        for (int rr = warp; rr < R; rr += RC_THREADS / 32) {
            const size_t gE = gE0 + (size_t)rr * W * 3 - g;               // even
            const uint32_t *wsrc = reinterpret_cast<const uint32_t *>(raw + gE);
            float2 *dst = reinterpret_cast<float2 *>(tile + rr * RC_TS);
            const int nwr = (gE + 2 * (size_t)nw > total_elems) ? nw - 1 : nw;      // never read past the image
            constexpr int U = 12;
            for (int base = lane; base < nwr; base += 32 * U) {
                uint32_t v[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int mm_ = base + u * 32;
                    if (mm_ < nwr) v[u] = __ldg(wsrc + mm_);
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int mm_ = base + u * 32;
                    if (mm_ < nwr) {
                        const float2 f2 = make_float2(__uint_as_float(__byte_perm(v[u], 0x4B000000u, 0x7610)) - 8388608.0f,
                                                      __uint_as_float(__byte_perm(v[u], 0x4B000000u, 0x7632)) - 8388608.0f);
                        dst[mm_] = f2;
                        if (TRACK) {
                            smn = fminf(smn, fminf(f2.x, f2.y));
                            smx = fmaxf(smx, fmaxf(f2.x, f2.y));
                        }
                    }
                }
            }
            if (nwr < nw && lane == 0) {                                  // last word of the image: its low half only
                const float f = (float)__ldg(raw + gE + 2 * (size_t)nwr);
                dst[nwr] = make_float2(f, f);
                if (TRACK) { smn = fminf(smn, f); smx = fmaxf(smx, f); }
            }
        }
        return g;
    }
    for (int rr = warp; rr < R; rr += RC_THREADS / 32) {
        const uint16_t *src = raw + gE0 + (size_t)rr * W * 3;
        float *dst = tile + rr * RC_TS;
        for (int e = lane; e < Wt; e += 32) {
            const float f = (float)__ldg(src + e);
            dst[e] = f;
            if (TRACK) { smn = fminf(smn, f); smx = fmaxf(smx, f); }
        }
    }
    return 0;
}

template <int MODE>
__device__ __forceinline__ void rc_emit(const ResampleOut<MODE> &out, size_t o, float val, float &mn, float &mx) {
    out.put(o, val);
    if (MODE != RS_U16) { mn = fminf(mn, val); mx = fmaxf(mx, val); }
}

// The output columns of one tile: thread `tid` of `nthreads` walks columns tid, tid + nthreads, ... down the patch rows.
// Shared by the one-tile-per-CTA kernel below and the persistent pipelined kernel of resample_pipe.cuh.
// a staged sample as float: float32 tiles as they are, uint16 tiles (raw samples read in place) through the 2^23 trick
__device__ __forceinline__ float rc_ld(const float *q) { return *q; }
__device__ __forceinline__ float rc_ld(const uint16_t *q) { return u16_to_float_magic((uint32_t)*q); }

// TRAW: element type of the raw exposure (the direct global path of a tile that was not staged reads it).
// TE: element type of the staged tile (float32, or uint16_t = the raw samples as fetched, converted on load).
template <int MODE, int MT, int TS = RC_TS, typename TRAW = uint16_t, typename TE = float>
__device__ __forceinline__ void rc_columns(const ResampleArgs &p, const ResampleOut<MODE> &out, const TE *tile, int shift,
                                           const pcb_lens_t *s_lens, int xa, bool lens_in_smem, bool use_tile, int ymin,
                                           int xmin, int ly, int k0, int kn, const pcb_hexcol_t *hc, int tid, int nthreads,
                                           float &mn, float &mx) {
    const int M = MT ? MT : p.M;
    const TRAW *raw = reinterpret_cast<const TRAW *>(p.raw);
    const int ncols_tile = kn * M * 3;
    for (int lc = tid; lc < ncols_tile; lc += nthreads) {
        const int px = lc / 3;
        const int ch = lc - px * 3;
        const int kl = px / M;
        const int u = px - kl * M;
        const int k = k0 + kl;
        const int us = p.flip ? (M - 1 - u) : u;
        const int col = k0 * M * 3 + lc;

        int x0 = k, x1 = k;
        float w = 0.f;
        if (hc) { const pcb_hexcol_t h = hc[k]; x0 = h.x0; x1 = h.x1; w = h.w; }
        const bool two = (w != 0.f);
        const pcb_lens_t L0 = lens_in_smem ? s_lens[x0 - xa] : p.lens[ly * p.LX + x0];
        pcb_lens_t L1 = L0;
        if (two) L1 = lens_in_smem ? s_lens[x1 - xa] : p.lens[ly * p.LX + x1];
        const bool ok0 = L0.oy != PCB_LENS_INVALID;
        const bool ok1 = two && (L1.oy != PCB_LENS_INVALID);
        // eight weights: [lens][dy][dx]
        const float h0 = ok0 ? (two ? 1.f - w : 1.f) : 0.f, h1 = ok1 ? w : 0.f;
        const float a00 = h0 * (1.f - L0.wy), a01 = h0 * L0.wy, a10 = h1 * (1.f - L1.wy), a11 = h1 * L1.wy;
        const float W000 = a00 * (1.f - L0.wx), W001 = a00 * L0.wx, W010 = a01 * (1.f - L0.wx), W011 = a01 * L0.wx;
        const float W100 = a10 * (1.f - L1.wx), W101 = a10 * L1.wx, W110 = a11 * (1.f - L1.wx), W111 = a11 * L1.wx;

        size_t o = ((size_t)ly * M + (p.flip ? M - 1 : 0)) * (size_t)p.ncols + col;
        const long long ostep = p.flip ? -(long long)p.ncols : (long long)p.ncols;

        if (use_tile) {
            const TE *t0 = tile + shift + (ok0 ? (L0.oy - ymin) * TS + (L0.ox - xmin + us) * 3 + ch : 0);
            const TE *t1 = ok1 ? tile + shift + (L1.oy - ymin) * TS + (L1.ox - xmin + us) * 3 + ch : t0;   // finite x 0
            float pa0 = rc_ld(t0), pb0 = rc_ld(t0 + 3), pa1 = rc_ld(t1), pb1 = rc_ld(t1 + 3);
            if (MT) {
#pragma unroll
                for (int r = 1; r <= (MT ? MT : 1); ++r) {
                    const float a0 = rc_ld(t0 + r * TS), b0 = rc_ld(t0 + r * TS + 3), a1 = rc_ld(t1 + r * TS), b1 = rc_ld(t1 + r * TS + 3);
                    float val = W000 * pa0;
                    val = fmaf(W001, pb0, val); val = fmaf(W010, a0, val); val = fmaf(W011, b0, val);
                    val = fmaf(W100, pa1, val); val = fmaf(W101, pb1, val); val = fmaf(W110, a1, val);
                    val = fmaf(W111, b1, val);
                    rc_emit<MODE>(out, o, val, mn, mx);
                    o += ostep;
                    pa0 = a0; pb0 = b0; pa1 = a1; pb1 = b1;
                }
            } else {
                for (int r = 1; r <= M; ++r) {
                    t0 += TS; t1 += TS;
                    const float a0 = rc_ld(t0), b0 = rc_ld(t0 + 3), a1 = rc_ld(t1), b1 = rc_ld(t1 + 3);
                    float val = W000 * pa0;
                    val = fmaf(W001, pb0, val); val = fmaf(W010, a0, val); val = fmaf(W011, b0, val);
                    val = fmaf(W100, pa1, val); val = fmaf(W101, pb1, val); val = fmaf(W110, a1, val);
                    val = fmaf(W111, b1, val);
                    rc_emit<MODE>(out, o, val, mn, mx);
                    o += ostep;
                    pa0 = a0; pb0 = b0; pa1 = a1; pb1 = b1;
                }
            }
        } else {
            const size_t stride = (size_t)p.W * 3;
            const TRAW *g0 = raw + ((size_t)(ok0 ? L0.oy : 0) * p.W + (ok0 ? L0.ox : 0) + us) * 3 + ch;
            const TRAW *g1 = ok1 ? raw + ((size_t)L1.oy * p.W + L1.ox + us) * 3 + ch : g0;
            float pa0 = (float)__ldg(g0), pb0 = (float)__ldg(g0 + 3), pa1 = (float)__ldg(g1), pb1 = (float)__ldg(g1 + 3);
            for (int r = 1; r <= M; ++r) {
                g0 += stride; g1 += stride;
                const float a0 = (float)__ldg(g0), b0 = (float)__ldg(g0 + 3), a1 = (float)__ldg(g1), b1 = (float)__ldg(g1 + 3);
                float val = W000 * pa0;
                val = fmaf(W001, pb0, val); val = fmaf(W010, a0, val); val = fmaf(W011, b0, val);
                val = fmaf(W100, pa1, val); val = fmaf(W101, pb1, val); val = fmaf(W110, a1, val);
                val = fmaf(W111, b1, val);
                rc_emit<MODE>(out, o, val, mn, mx);
                o += ostep;
                pa0 = a0; pb0 = b0; pa1 = a1; pb1 = b1;
            }
        }
    }
}

// MT: micro-image size known at compile time (row loop fully unrolled, immediate LDS offsets) or 0 = runtime.
template <int MODE, int MT>
__global__ void __launch_bounds__(RC_THREADS, 4)
resample_col_kernel(ResampleArgs p, float *__restrict__ out_f32, uint16_t *__restrict__ out_u16,
                    pcb_minmax_t *mm_rw, const pcb_minmax_t *mm_ro) {
    extern __shared__ __align__(16) float tile[];          // [RC_ROWS][RC_TS]
    __shared__ pcb_lens_t s_lens[RC_MAX_SRC];
    __shared__ int s_box[4];      // ymin, ymax, xmin, xmax over the valid source lenses (oy / ox)
    __shared__ int s_src[2];      // first / last source lens column
    __shared__ int s_smm[2];      // min / max of the staged samples (bit patterns of non-negative floats order like ints)
    __shared__ int s_ninv;        // invalid source lenses of this CTA
    __shared__ float s_g[2];      // min/max pass: the global bounds as lane 0 saw them at the start of this CTA

    const int M = MT ? MT : p.M;
    const int ly = blockIdx.y;
    const int k0 = blockIdx.x * RC_KT;
    const int kn = min(RC_KT, p.Xs - k0);
    const int tid = threadIdx.x;
    const pcb_hexcol_t *hc = p.hexcol ? p.hexcol + ((ly + p.hex_odd) & 1) * p.Xs : nullptr;

    ResampleOut<MODE> out;
    out.f32 = out_f32; out.u16 = out_u16;
    out.qz.init(0.f, 0.f);
    if (MODE == RS_U16) out.qz.init(key_to_float(mm_ro->kmin), key_to_float(mm_ro->kmax));
    // Min/max pass: the running global bounds as they are NOW (stale is fine: they only ever widen).  A tile whose staged
    // raw samples lie strictly inside them cannot move them - every output is a convex combination of staged samples,
    // up to float rounding (a few 2^-24, covered by RC_SKIP_EPS) - so its interpolation is skipped altogether.  The
    // final min / max are exactly those of the full computation.  (NaN from freshly reset keys compares false: no skip.)
    // The bounds are read ONCE per CTA (lane 0 of warp 0) and handed to everybody through shared memory: they change
    // under the kernel's feet, and a skip decision taken by some warps only would leave block_minmax_commit reading
    // partial results that were never written.
    float g_mn = 0.f, g_mx = 0.f;
    const bool may_skip = MODE == RS_MINMAX && mm_rw != nullptr && p.skip;

    // Tables of the tile, by warp 0, in ONE global-memory latency: the two hex-column entries that bound the source
    // lenses and, speculatively, the 32 lenses starting just below where the first one must be (x0(k) = floor(k LX /
    // (Xs - 1) + parity / 2), so a float estimate is off by one at most); the exact window is then picked out of the
    // speculative one with shuffles.  One barrier instead of two dependent load-and-barrier rounds.
    if (tid < 32) {
        const int lane = tid;
        if (may_skip) {
            uint32_t kmn = 0u, kmx = 0u;
            if (lane == 0) {
                kmn = *reinterpret_cast<volatile uint32_t *>(&mm_rw->kmin);
                kmx = *reinterpret_cast<volatile uint32_t *>(&mm_rw->kmax);
            }
            g_mn = key_to_float(__shfl_sync(0xffffffffu, kmn, 0));
            g_mx = key_to_float(__shfl_sync(0xffffffffu, kmx, 0));
        }
        int xa = k0, xb = k0 + kn - 1;
        int xa_est = k0;
        pcb_hexcol_t h;
        h.x0 = h.x1 = 0; h.w = 0.f; h.reserved = 0;
        if (hc) {
            xa_est = max((int)floorf((float)k0 * (float)p.LX / (float)max(p.Xs - 1, 1)) - 1, 0);
            if (lane < 2) h = hc[lane == 0 ? k0 : k0 + kn - 1];
        }
        const int xs = min(xa_est + lane, p.LX - 1);
        pcb_lens_t Ls = p.lens[ly * p.LX + xs];                      // in flight together with the hex entries
        if (hc) {
            xa = __shfl_sync(0xffffffffu, h.x0, 0);
            const int bx0 = __shfl_sync(0xffffffffu, h.x0, 1), bx1 = __shfl_sync(0xffffffffu, h.x1, 1);
            const float bw = __shfl_sync(0xffffffffu, h.w, 1);
            xb = bw == 0.f ? bx0 : bx1;
        }
        const int nsrc_w = xb - xa + 1;
        // lane l owns source lens xa + l: held by lane (xa + l - xa_est) of the speculative load when that is in range
        const int from = xa + lane - xa_est;
        pcb_lens_t L;
        L.oy = __shfl_sync(0xffffffffu, Ls.oy, from & 31);
        L.ox = __shfl_sync(0xffffffffu, Ls.ox, from & 31);
        L.wy = __shfl_sync(0xffffffffu, Ls.wy, from & 31);
        L.wx = __shfl_sync(0xffffffffu, Ls.wx, from & 31);
        const bool mine = lane < nsrc_w && nsrc_w <= RC_MAX_SRC;
        if (mine && (from < 0 || from > 31 || xa_est + from > p.LX - 1)) L = p.lens[ly * p.LX + xa + lane];
        const bool valid = mine && L.oy != PCB_LENS_INVALID;
        int b0 = valid ? L.oy : INT32_MAX, b1 = valid ? L.oy : INT32_MIN;
        int b2 = valid ? L.ox : INT32_MAX, b3 = valid ? L.ox : INT32_MIN;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            b0 = min(b0, __shfl_xor_sync(0xffffffffu, b0, o)); b1 = max(b1, __shfl_xor_sync(0xffffffffu, b1, o));
            b2 = min(b2, __shfl_xor_sync(0xffffffffu, b2, o)); b3 = max(b3, __shfl_xor_sync(0xffffffffu, b3, o));
        }
        const unsigned inv = __ballot_sync(0xffffffffu, mine && !valid);
        if (mine) s_lens[lane] = L;
        if (lane == 0) {
            s_src[0] = xa; s_src[1] = xb;
            s_box[0] = b0; s_box[1] = b1; s_box[2] = b2; s_box[3] = b3;
            s_smm[0] = 0x7f800000; s_smm[1] = 0;
            s_ninv = __popc(inv);
            s_g[0] = g_mn; s_g[1] = g_mx;
        }
    }
    __syncthreads();
    if (may_skip) { g_mn = s_g[0]; g_mx = s_g[1]; }
    const int xa = s_src[0];
    const int nsrc = s_src[1] - xa + 1;
    const bool lens_in_smem = nsrc <= RC_MAX_SRC;
    const int ymin = s_box[0], xmin = s_box[2];
    const int R = s_box[1] - ymin + M + 1;               // staged rows
    const int Wt = (s_box[3] - xmin + M + 1) * 3;        // staged elements per row
    const bool any_valid = s_box[1] != INT32_MIN;
    const bool use_tile = lens_in_smem && any_valid && R <= RC_ROWS && Wt + 2 <= RC_TS;
    const uint16_t *raw = reinterpret_cast<const uint16_t *>(p.raw);
    int shift = 0;
    float smn = INFINITY, smx = 0.f;
    if (use_tile) {
        if (may_skip) shift = rc_stage<true>(raw, (size_t)p.H * p.W * 3, p.W, tile, ymin, xmin, R, Wt, smn, smx);
        else shift = rc_stage<false>(raw, (size_t)p.H * p.W * 3, p.W, tile, ymin, xmin, R, Wt, smn, smx);
    }
    if (may_skip && use_tile) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            smn = fminf(smn, __shfl_xor_sync(0xffffffffu, smn, o));
            smx = fmaxf(smx, __shfl_xor_sync(0xffffffffu, smx, o));
        }
        if ((tid & 31) == 0) {
            atomicMin(&s_smm[0], __float_as_int(smn));
            atomicMax(&s_smm[1], __float_as_int(smx));
        }
    }
    __syncthreads();
    if (may_skip && use_tile && s_ninv == 0) {
        const float tmn = __int_as_float(s_smm[0]), tmx = __int_as_float(s_smm[1]);
        if (tmn * (1.f - RC_SKIP_EPS) > g_mn && tmx * (1.f + RC_SKIP_EPS) < g_mx) return;      // uniform over the CTA
    }

    float mn = INFINITY, mx = -INFINITY;
    rc_columns<MODE, MT>(p, out, tile, shift, s_lens, xa, lens_in_smem, use_tile, ymin, xmin, ly, k0, kn, hc, tid, RC_THREADS,
                         mn, mx);
    if (MODE != RS_U16 && mm_rw != nullptr) block_minmax_commit(mn, mx, mm_rw);
}

// uint16, C == 3, M + 1 + 3 <= RC_ROWS: the lean kernel applies.
static bool resample_col_applies(const ResampleArgs &a, int raw_dtype) {
    if (getenv("PCB_RESAMPLE_KERNEL") && !strcmp(getenv("PCB_RESAMPLE_KERNEL"), "tile")) return false;
    return raw_dtype == PCB_U16 && a.C == 3 && a.M + 2 <= RC_ROWS && a.M >= 1;
}

template <int MODE>
static int launch_resample_col(const ResampleArgs &a_in, float *out_f32, uint16_t *out_u16, pcb_minmax_t *mm_rw,
                               const pcb_minmax_t *mm_ro, cudaStream_t st) {
    ResampleArgs a = a_in;
    a.skip = (getenv("PCB_RESAMPLE_NOSKIP") && atoi(getenv("PCB_RESAMPLE_NOSKIP"))) ? 0 : 1;
    const size_t smem = (size_t)RC_ROWS * RC_TS * sizeof(float);
    dim3 block(RC_THREADS), grid((a.Xs + RC_KT - 1) / RC_KT, a.LY);
    if (a.M == 15) {
        PCB_CUDA(cudaFuncSetAttribute(resample_col_kernel<MODE, 15>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        resample_col_kernel<MODE, 15><<<grid, block, smem, st>>>(a, out_f32, out_u16, mm_rw, mm_ro);
    } else {
        PCB_CUDA(cudaFuncSetAttribute(resample_col_kernel<MODE, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        resample_col_kernel<MODE, 0><<<grid, block, smem, st>>>(a, out_f32, out_u16, mm_rw, mm_ro);
    }
    PCB_LAUNCH_OK();
    return PCB_OK;
}

}  // namespace pcb
