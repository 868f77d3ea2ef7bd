// (c') Sub-pixel "refocus refinement" (opt_refi): replaces the up-sample / shift / sum / down-sample chain of
// lfp_refocuser/lfp_shiftandsum.py:93-135 + misc/data_proc.py:57-92.
//
// The chain is linear and separable (see lfp_refocuser/refinement.py):
//     R_a = (1/M) * sum_{j,i active}  A_{a(c-j)} . V_ji . B_{a(c-i)}^T ,      A_s = D_s . P
// P (samples -> cubic B-spline coefficients, ~29 taps) does not depend on the slice; D_s is a band of <= 5 taps.
//
//   once per light field   coef[j,i] = Py . V_ji . Px^T                            (refi_prefilter_{x,y}_kernel)
//   per batch of slices    H[n][j][r][x]  = sum_{i active} sum_m Dx[s(n,i)][m][x] * coef[j,i][r][fx+m]   (refine_h_kernel)
//                          out[n][y][x]   = (1/M) sum_{j active} sum_m Dy[s(n,j)][g(y)][m][q(y)] * H[n][j][fy+m][x]
//
// refine_h_kernel is where the work is (177 views x K taps per output and slice).  One CTA owns view-row j, RH_ROWS
// image rows and a chunk of RH_XC pixels; it stages, once, the coefficient windows every slice of the batch can
// reach (per view: chunk + the offsets its own shifts produce) and then walks all slices.  Lanes run along x, so
// the K weights + window start of a thread are coalesced loads, fetched once per (slice, view) and used for all
// RH_ROWS rows from registers; the samples come from shared memory at a 3-word lane stride (conflict free).
#pragma once
#include "common.cuh"

namespace pcb {

constexpr int RF_GROUP = 8;          // output rows that share one sample window in the vertical pass
constexpr int RF_MAXM = 32;          // view columns per row supported by the control block

// ---- prefilter -------------------------------------------------------------------------------------------------
constexpr int RP_THREADS = 320;

// grid (S, M*M): one CTA per (image row r, view); tmp[view][r][x][c] = sum_m px[m][x] * V[view][r][pf[x]+m][c]
template <typename TIn>
__global__ void __launch_bounds__(RP_THREADS)
refi_prefilter_x_kernel(const TIn *__restrict__ vp, int S, int T, const float *__restrict__ px_vals,
                        const int32_t *__restrict__ px_first, int Kp, const uint8_t *__restrict__ view_zeroed,
                        float *__restrict__ tmp, int r_off = 0) {
    extern __shared__ __align__(16) float s_row[];        // [T*3]
    const int r = r_off + blockIdx.x, view = blockIdx.y;
    if (view_zeroed[view]) return;
    const int E = T * 3;
    const TIn *src = vp + ((size_t)view * S + r) * E;
    for (int e = threadIdx.x; e < E; e += RP_THREADS) s_row[e] = ld_as_float(src + e);
    __syncthreads();
    float *dst = tmp + ((size_t)view * S + r) * E;
    for (int x = threadIdx.x; x < T; x += RP_THREADS) {
        const float *px = s_row + px_first[x] * 3;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll 4
        for (int m = 0; m < Kp; ++m) {
            const float w = __ldg(px_vals + (size_t)m * T + x);
            a0 = fmaf(w, px[m * 3 + 0], a0);
            a1 = fmaf(w, px[m * 3 + 1], a1);
            a2 = fmaf(w, px[m * 3 + 2], a2);
        }
        dst[x * 3 + 0] = a0; dst[x * 3 + 1] = a1; dst[x * 3 + 2] = a2;
    }
}

// grid (ceil(E/256), Gy, M*M): thread = one (x, c) element of a group of RP_GY output rows that share one window
//   coef[view][y][e] = sum_m py[g(y)][m][q(y)] * tmp[view][pf[g]+m][e]        (weights uniform over the CTA)
constexpr int RP_GY = 8;
__global__ void __launch_bounds__(256)
refi_prefilter_y_kernel(const float *__restrict__ tmp, int S, int E, const float *__restrict__ py_vals,
                        const int32_t *__restrict__ py_first, int Kg, const uint8_t *__restrict__ view_zeroed,
                        float *__restrict__ coef, int gy0 = 0) {
    const int gy = gy0 + blockIdx.y, view = blockIdx.z;
    if (view_zeroed[view]) return;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    const float4 *w = reinterpret_cast<const float4 *>(py_vals) + (size_t)gy * Kg * 2;
    const int f = py_first[gy];
    const float *col = tmp + (size_t)view * S * E + e;
    float acc[RP_GY];
#pragma unroll
    for (int q = 0; q < RP_GY; ++q) acc[q] = 0.f;
#pragma unroll 4
    for (int m = 0; m < Kg; ++m) {
        const float4 wa = __ldg(w + 2 * m), wb = __ldg(w + 2 * m + 1);
        const float v = __ldg(col + (size_t)min(f + m, S - 1) * E);
        acc[0] = fmaf(wa.x, v, acc[0]); acc[1] = fmaf(wa.y, v, acc[1]);
        acc[2] = fmaf(wa.z, v, acc[2]); acc[3] = fmaf(wa.w, v, acc[3]);
        acc[4] = fmaf(wb.x, v, acc[4]); acc[5] = fmaf(wb.y, v, acc[5]);
        acc[6] = fmaf(wb.z, v, acc[6]); acc[7] = fmaf(wb.w, v, acc[7]);
    }
#pragma unroll
    for (int q = 0; q < RP_GY; ++q) {
        const int y = gy * RP_GY + q;
        if (y < S) coef[((size_t)view * S + y) * E + e] = acc[q];
    }
}

// ---- prefilter, Toeplitz interior ---------------------------------------------------------------------------------
// Away from the image border the rows of P are one and the same symmetric filter (the cubic B-spline prefilter
// sqrt(3) z1^|k|, 29 taps at 1e-8): the host hands over those taps and the range of rows that equal them EXACTLY (as
// float32), and the kernels below run that range with the taps in the kernel-parameter bank and a register window,
// in the same summation order as the table kernels above (bit-identical results); border rows keep the tables.
constexpr int RPT_K = 29;                      // taps of the interior filter this fast path is built for
struct RpTaps { float h[RPT_K]; int c, lo, hi; };   // row x in [lo, hi] uses samples x - c ... x - c + RPT_K - 1

constexpr int RPX_P = 9;                       // adjacent pixels of one channel per thread (27-float lane stride)
template <typename TIn>
__global__ void __launch_bounds__(RP_THREADS)
refi_prefilter_x3_kernel(const TIn *__restrict__ vp, int S, int T, const float *__restrict__ px_vals,
                         const int32_t *__restrict__ px_first, int Kp, RpTaps tp,
                         const uint8_t *__restrict__ view_zeroed, float *__restrict__ tmp, int r_off = 0) {
    extern __shared__ __align__(16) float s_buf[];        // [T*3] samples, [T*3] results
    const int r = r_off + blockIdx.x, view = blockIdx.y;
    if (view_zeroed[view]) return;
    const int E = T * 3;
    float *s_row = s_buf, *s_out = s_buf + E;
    const TIn *src = vp + ((size_t)view * S + r) * E;
    for (int e = threadIdx.x; e < E; e += RP_THREADS) s_row[e] = ld_as_float(src + e);
    __syncthreads();
    // tasks: (channel, group of RPX_P pixels) for the groups that lie inside [lo, hi], then one task per border
    // (pixel, channel) — a border pixel costs about as much as half a group, so it gets a thread of its own
    const int g_lo = (tp.lo + RPX_P - 1) / RPX_P;
    const int g_hi = tp.hi >= RPX_P - 1 ? (tp.hi - (RPX_P - 1)) / RPX_P + 1 : 0;      // one past the last inner group
    const int ngi = max(g_hi - g_lo, 0);
    const int xl = ngi ? g_lo * RPX_P : T;                   // border pixels: [0, xl) and [xr, T)
    const int xr = ngi ? g_hi * RPX_P : T;
    const int nbx = xl + (T - xr);
    for (int task = threadIdx.x; task < 3 * (ngi + nbx); task += RP_THREADS) {
        if (task < 3 * ngi) {
            const int c = task / ngi, g = g_lo + (task - c * ngi);
            const int xa = g * RPX_P;
            const float *p = s_row + (xa - tp.c) * 3 + c;
            // every sample is read once and goes to the (up to RPX_P) outputs it contributes to; for a fixed output the
            // taps still arrive in ascending order
            float acc[RPX_P];
#pragma unroll
            for (int j = 0; j < RPX_P; ++j) acc[j] = 0.f;
#pragma unroll
            for (int m = 0; m < RPX_P + RPT_K - 1; ++m) {
                const float v = p[m * 3];
#pragma unroll
                for (int j = 0; j < RPX_P; ++j)
                    if (m - j >= 0 && m - j < RPT_K) acc[j] = fmaf(tp.h[m - j], v, acc[j]);
            }
#pragma unroll
            for (int j = 0; j < RPX_P; ++j) s_out[(xa + j) * 3 + c] = acc[j];
        } else {
            const int bt = task - 3 * ngi;
            const int c = bt / nbx, bx = bt - c * nbx;
            const int x = bx < xl ? bx : xr + (bx - xl);
            const float *px = s_row + px_first[x] * 3 + c;
            float acc = 0.f;
            for (int m = 0; m < Kp; ++m) acc = fmaf(__ldg(px_vals + (size_t)m * T + x), px[m * 3], acc);
            s_out[x * 3 + c] = acc;
        }
    }
    __syncthreads();
    float *dst = tmp + ((size_t)view * S + r) * E;
    for (int e = threadIdx.x; e < E; e += RP_THREADS) dst[e] = s_out[e];
}

// grid (ceil(E/256), strips, M*M): rows y0 + blockIdx.y * RPY_R ... of the interior; thread = one (x, c) element.
constexpr int RPY_R = 16;
__global__ void __launch_bounds__(256)
refi_prefilter_yt_kernel(const float *__restrict__ tmp, int S, int E, RpTaps tp, int y_first,
                         const uint8_t *__restrict__ view_zeroed, float *__restrict__ coef) {
    const int view = blockIdx.z;
    if (view_zeroed[view]) return;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    const int y0 = y_first + blockIdx.y * RPY_R;
    const float *col = tmp + ((size_t)view * S + (y0 - tp.c)) * E + e;
    float acc[RPY_R];
#pragma unroll
    for (int j = 0; j < RPY_R; ++j) acc[j] = 0.f;
#pragma unroll
    for (int m0 = 0; m0 < RPY_R + RPT_K - 1; m0 += 11) {            // 44 samples in four batches of 11 loads in flight
        float v[11];
#pragma unroll
        for (int u = 0; u < 11; ++u)
            if (m0 + u < RPY_R + RPT_K - 1) v[u] = __ldg(col + (size_t)(m0 + u) * E);
#pragma unroll
        for (int u = 0; u < 11; ++u) {
            const int m = m0 + u;
#pragma unroll
            for (int j = 0; j < RPY_R; ++j)
                if (m < RPY_R + RPT_K - 1 && m - j >= 0 && m - j < RPT_K) acc[j] = fmaf(tp.h[m - j], v[u], acc[j]);
        }
    }
    float *o = coef + ((size_t)view * S + y0) * E + e;
#pragma unroll
    for (int j = 0; j < RPY_R; ++j) o[(size_t)j * E] = acc[j];
}

// ---- horizontal pass ---------------------------------------------------------------------------------------------
constexpr int RH_XC = 160;           // output pixels per CTA
constexpr int RH_ROWS = 3;           // image rows per CTA: the weights of a thread are reused for all of them
constexpr int RH_SG = 4;             // thread groups per CTA, each RH_XC wide, sharing the staged windows and taking
                                     //   every RH_SG-th slice: 4x the warps in flight for the same shared memory
constexpr int RH_THREADS = RH_XC * RH_SG;
constexpr int RH_MAX_IX = 1024;      // slice x view operator indices kept in shared memory per launch

struct RefineHCtl {                  // per-launch control block, passed by value
    int32_t lo[RF_MAXM];             // per view column: smallest / largest coefficient offset (relative to the output
    int32_t hi[RF_MAXM];             //   pixel) any slice of the batch reads
    uint32_t active[RF_MAXM];        // bit i of active[j]: view (j, i) is not masked
};

template <int K>
__global__ void __launch_bounds__(RH_THREADS)
refine_h_kernel(const float *__restrict__ coef, int M, int S, int T, int nb, const float *__restrict__ dx_vals,
                const int32_t *__restrict__ dx_first, const int32_t *__restrict__ ix, RefineHCtl ctl,
                float *__restrict__ H) {
    extern __shared__ __align__(16) float s_win[];         // [view][row][width_i * 3]
    __shared__ int s_off[RF_MAXM];                         // float offset of view i's block in s_win
    __shared__ int s_x0[RF_MAXM];                          // first staged pixel of view i
    __shared__ int s_wE[RF_MAXM];                          // staged floats per row of view i
    __shared__ int s_act[RF_MAXM];                         // the unmasked view columns of this view-row
    __shared__ int2 s_view[RF_MAXM];                       // per ACTIVE view: {float offset of pixel 0 in s_win, row pitch}
    __shared__ int s_nact;
    __shared__ int s_ix[RH_MAX_IX];                        // ix[n][i] of the batch
    const int j = blockIdx.z;
    const int r0 = blockIdx.y * RH_ROWS;
    const int xc0 = blockIdx.x * RH_XC;
    const int tid = threadIdx.x;
    const int sg = tid / RH_XC, tx = tid - sg * RH_XC;
    const uint32_t act = ctl.active[j];
    if (act == 0u) return;
    const int nrow = min(RH_ROWS, S - r0);
    const int xcn = min(RH_XC, T - xc0);

    if (tid == 0) {
        int off = 0, na = 0;
        for (int i = 0; i < M; ++i) {
            const int a = max(xc0 + ctl.lo[i], 0), b = min(xc0 + xcn - 1 + ctl.hi[i], T - 1);
            s_x0[i] = a;
            s_off[i] = off;
            s_wE[i] = (b - a + 1) * 3;
            if ((act >> i) & 1u) {
                s_view[na] = make_int2(off - a * 3, (b - a + 1) * 3);
                off += (b - a + 1) * 3 * RH_ROWS;
                s_act[na++] = i;
            }
        }
        s_nact = na;
    }
    __syncthreads();
    // operator index of every (slice, ACTIVE view) of the batch, compacted so that the slice loop indexes it by ii
    const int nact = s_nact;
    for (int e = tid; e < nb * nact; e += RH_THREADS) {
        const int n = e / nact, ii = e - n * nact;
        s_ix[e] = __ldg(ix + n * M + s_act[ii]);
    }
    __syncthreads();
    // ---- stage the coefficient windows (rows beyond S replicate nothing: they are never read)
    for (int i = 0; i < M; ++i) {
        if (!((act >> i) & 1u)) continue;
        const int a = s_x0[i];
        const int wE = (min(xc0 + xcn - 1 + ctl.hi[i], T - 1) - a + 1) * 3;
        float *dst = s_win + s_off[i];
        for (int rr = 0; rr < nrow; ++rr) {
            const float *src = coef + (((size_t)(j * M + i) * S + r0 + rr) * T + a) * 3;
            for (int e = tid; e < wE; e += RH_THREADS) dst[rr * wE + e] = __ldg(src + e);
        }
    }
    __syncthreads();
    if (tx >= xcn) return;
    const int x = xc0 + tx;

    for (int n = sg; n < nb; n += RH_SG) {                 // the RH_SG thread groups of a CTA take slices round-robin
        float acc[RH_ROWS][3];
#pragma unroll
        for (int rr = 0; rr < RH_ROWS; ++rr) acc[rr][0] = acc[rr][1] = acc[rr][2] = 0.f;
        const int *ixn = s_ix + n * nact;
        for (int ii = 0; ii < nact; ++ii) {
            const int sidx = ixn[ii];                                      // uniform, from shared memory
            const int2 vw = s_view[ii];
            const int f = __ldg(dx_first + (size_t)sidx * T + x);
            float w[K];
#pragma unroll
            for (int m = 0; m < K; ++m) w[m] = __ldg(dx_vals + ((size_t)sidx * K + m) * T + x);
            const int wE = vw.y;
            const float *px = s_win + vw.x + f * 3;
#pragma unroll
            for (int rr = 0; rr < RH_ROWS; ++rr) {
#pragma unroll
                for (int m = 0; m < K; ++m) {
                    acc[rr][0] = fmaf(w[m], px[rr * wE + m * 3 + 0], acc[rr][0]);
                    acc[rr][1] = fmaf(w[m], px[rr * wE + m * 3 + 1], acc[rr][1]);
                    acc[rr][2] = fmaf(w[m], px[rr * wE + m * 3 + 2], acc[rr][2]);
                }
            }
        }
#pragma unroll
        for (int rr = 0; rr < RH_ROWS; ++rr) {
            if (rr < nrow) {
                float *o = H + ((((size_t)n * M + j) * S + r0 + rr) * T + x) * 3;
                o[0] = acc[rr][0]; o[1] = acc[rr][1]; o[2] = acc[rr][2];
            }
        }
    }
}

// ---- horizontal pass, slice quads ----------------------------------------------------------------------------------
// refine_h_kernel issues one shared-memory load per FMA (every tap of every slice fetches its own sample), so it runs at
// the LDS rate, a quarter of the FMA rate.  Four CONSECUTIVE slices read almost the same samples of a view - their shifts
// a (c - i) differ by |c - i| up-sampled pixels, i.e. by less than half a coarse pixel - so this form walks the slices in
// quads: per view and output pixel the host prepares ONE window [fu, fu + W) that covers the bands of all four slices
// (W = K + 0..3) and the four bands padded with zeros to that window, as float4 {slice 0..3} per tap
// (lfp_refocuser/refinement.py::quad_tables).  A thread then loads each sample of the window once and feeds 4 slices x 3
// channels from it, for RQ_ROWS image rows whose weights it shares: per tap 1 coalesced LDG.128 + 3 RQ_ROWS LDS for
// 12 RQ_ROWS FMAs.  Adding the zero taps changes nothing in the sums (x + 0 * s = x), and the taps of one slice still
// arrive in ascending sample order, view by view: the results equal refine_h_kernel's bit for bit.
constexpr int RQ = 4;                // slices per quad
constexpr int RQ_XC = 128;           // output pixels per CTA
constexpr int RQ_ROWS = 6;           // image rows per CTA (and per thread)
constexpr int RQ_SG = 4;             // thread groups per CTA; they take the quads round-robin
constexpr int RQ_THREADS = RQ_XC * RQ_SG;
constexpr int RQ_MAXW = 8;

struct RefineQCtl {                  // per-launch control block, passed by value
    int32_t lo[RF_MAXM];             // per view column: smallest / largest coefficient offset (relative to the output
    int32_t hi[RF_MAXM];             //   pixel) any quad window touches
    uint32_t active[RF_MAXM];        // bit i of active[j]: view (j, i) is not masked
    int32_t woff[RF_MAXM];           // float4 offset of view column i's block [G][W_i][T] in the weight table
    int32_t W[RF_MAXM];              // window length of view column i
};

constexpr int RQ_PIX = 20;           // floats per staged pixel: RQ_ROWS x 3 channels + 2 pad.  5 float4 = an ODD number of
                                     //   16-byte units, so the LDS.128 of a quarter-warp (lane stride 80 B) hit 8 distinct
                                     //   bank groups: conflict free
static_assert(RQ_ROWS * 3 <= RQ_PIX && RQ_PIX % 4 == 0 && (RQ_PIX / 4) % 2 == 1, "pixel record layout");

// One view: W taps; per tap one float4 of weights {slice 0..3} and the pixel record of the window sample (6 rows x 3
// channels in 5 LDS.128 off ONE base register with immediate offsets).
template <int W>
__device__ __forceinline__ void rq_view(float (&acc)[RQ_ROWS][RQ][3], const float4 *__restrict__ wrec, int T,
                                        const float4 *__restrict__ px) {
    float4 wt = __ldg(wrec);
#pragma unroll
    for (int t = 0; t < W; ++t) {
        const float4 wn = t + 1 < W ? __ldg(wrec + (size_t)(t + 1) * T) : wt;     // next tap's weights in flight
        const float4 v0 = px[t * 5 + 0], v1 = px[t * 5 + 1], v2 = px[t * 5 + 2], v3 = px[t * 5 + 3], v4 = px[t * 5 + 4];
        const float s[RQ_ROWS][3] = {{v0.x, v0.y, v0.z}, {v0.w, v1.x, v1.y}, {v1.z, v1.w, v2.x},
                                     {v2.y, v2.z, v2.w}, {v3.x, v3.y, v3.z}, {v3.w, v4.x, v4.y}};
#pragma unroll
        for (int rr = 0; rr < RQ_ROWS; ++rr) {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                acc[rr][0][c] = fmaf(wt.x, s[rr][c], acc[rr][0][c]);
                acc[rr][1][c] = fmaf(wt.y, s[rr][c], acc[rr][1][c]);
                acc[rr][2][c] = fmaf(wt.z, s[rr][c], acc[rr][2][c]);
                acc[rr][3][c] = fmaf(wt.w, s[rr][c], acc[rr][3][c]);
            }
        }
        wt = wn;
    }
}

__global__ void __launch_bounds__(RQ_THREADS, 1)
refine_hq_kernel(const float *__restrict__ coef, int M, int S, int T, int nb, int G, const float4 *__restrict__ wq,
                 const int32_t *__restrict__ fu, RefineQCtl ctl, float *__restrict__ H, int rb_off) {
    extern __shared__ __align__(16) float s_win[];         // [active view][pixel of the window][RQ_PIX]
    __shared__ int s_act[RF_MAXM];                         // the unmasked view columns of this view-row
    __shared__ int2 s_view[RF_MAXM];                       // per ACTIVE view: {float offset of pixel 0's record, pixels}
    __shared__ int2 s_tab[RF_MAXM];                        // per ACTIVE view: {float4 offset of its weight block, W}
    __shared__ int s_nact;
    const int j = blockIdx.z;
    const int r0 = (rb_off + blockIdx.y) * RQ_ROWS;
    const int xc0 = blockIdx.x * RQ_XC;
    const int tid = threadIdx.x;
    const int sg = tid / RQ_XC, tx = tid - sg * RQ_XC;
    const uint32_t act = ctl.active[j];
    if (act == 0u) return;
    const int nrow = min(RQ_ROWS, S - r0);
    const int xcn = min(RQ_XC, T - xc0);

    if (tid == 0) {
        int off = 0, na = 0;
        for (int i = 0; i < M; ++i) {
            if (!((act >> i) & 1u)) continue;
            const int a = max(xc0 + ctl.lo[i], 0), b = min(xc0 + xcn - 1 + ctl.hi[i], T - 1);
            s_view[na] = make_int2(off - a * RQ_PIX, b - a + 1);
            s_tab[na] = make_int2(ctl.woff[i], ctl.W[i]);
            off += (b - a + 1) * RQ_PIX;
            s_act[na++] = i;
        }
        s_nact = na;
    }
    __syncthreads();
    const int nact = s_nact;
    // ---- stage the coefficient windows, transposed to pixel records [x][row][channel]: one (view, row) pair per warp
    //      and turn, 4-byte cp.async so that every copy of the CTA is in flight before the first is waited for; rows
    //      beyond S are zero (their sums are never stored)
    {
        const int warp = tid >> 5, lane = tid & 31;
        for (int pr = warp; pr < nact * RQ_ROWS; pr += RQ_THREADS / 32) {
            const int ii = pr / RQ_ROWS, rr = pr - ii * RQ_ROWS;
            const int i = s_act[ii];
            const int2 vw = s_view[ii];
            const int a = max(xc0 + ctl.lo[i], 0);
            float *dst = s_win + vw.x + a * RQ_PIX + rr * 3;
            const int nE = vw.y * 3;
            if (rr < nrow) {
                const float *src = coef + (((size_t)(j * M + i) * S + r0 + rr) * T + a) * 3;
                for (int e = lane; e < nE; e += 32) {
                    const int p = e / 3, c = e - p * 3;
                    const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst + p * RQ_PIX + c);
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(src + e) : "memory");
                }
            } else {
                for (int e = lane; e < nE; e += 32) {
                    const int p = e / 3, c = e - p * 3;
                    dst[p * RQ_PIX + c] = 0.f;
                }
            }
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
    }
    __syncthreads();
    if (tx >= xcn) return;
    const int x = xc0 + tx;

    for (int g = sg; g < G; g += RQ_SG) {                  // the RQ_SG thread groups take the quads round-robin
        float acc[RQ_ROWS][RQ][3];
#pragma unroll
        for (int rr = 0; rr < RQ_ROWS; ++rr)
#pragma unroll
            for (int q = 0; q < RQ; ++q) acc[rr][q][0] = acc[rr][q][1] = acc[rr][q][2] = 0.f;
        for (int ii = 0; ii < nact; ++ii) {
            const int2 vw = s_view[ii], tb = s_tab[ii];
            const int i = s_act[ii];
            const int f = __ldg(fu + ((size_t)g * M + i) * T + x);
            const float4 *wrec = wq + (size_t)tb.x + ((size_t)g * tb.y) * T + x;
            const float4 *px = reinterpret_cast<const float4 *>(s_win + vw.x + f * RQ_PIX);
            switch (tb.y) {                                // uniform over the CTA
            case 1: rq_view<1>(acc, wrec, T, px); break;
            case 2: rq_view<2>(acc, wrec, T, px); break;
            case 3: rq_view<3>(acc, wrec, T, px); break;
            case 4: rq_view<4>(acc, wrec, T, px); break;
            case 5: rq_view<5>(acc, wrec, T, px); break;
            case 6: rq_view<6>(acc, wrec, T, px); break;
            case 7: rq_view<7>(acc, wrec, T, px); break;
            default: rq_view<8>(acc, wrec, T, px); break;
            }
        }
#pragma unroll
        for (int q = 0; q < RQ; ++q) {
            const int n = g * RQ + q;
            if (n >= nb) break;
#pragma unroll
            for (int rr = 0; rr < RQ_ROWS; ++rr) {
                if (rr < nrow) {
                    float *o = H + ((((size_t)n * M + j) * S + r0 + rr) * T + x) * 3;
                    o[0] = acc[rr][q][0]; o[1] = acc[rr][q][1]; o[2] = acc[rr][q][2];
                }
            }
        }
    }
}

// ---- vertical pass -----------------------------------------------------------------------------------------------
// out[n][y][e] = (1/M) sum_{j active} sum_m Dy[s(n,j)][g(y)][m][q(y)] * H[n][j][first + m][e]
// grid (Gy, ceil(E/256), nb): thread = one (x, c) element of a group of RF_GROUP = 8 output rows of slice n that share
// one window of KG sample rows per view row (KG = K + spread of the band starts inside a group, compile time; the host
// tables keep every window inside the image, so there is no clamping).  The pass is bound by the latency of the H loads
// (2.9 GB read once from DRAM), so: wide groups (13 rows loaded for 8 outputs instead of 9 for 4), the weights of the
// whole CTA staged once in shared memory (they are the same for every thread: broadcast LDS.128 instead of global
// loads), and the samples of the next view row requested before the current ones are used.
// The finished values go to every destination of `dst` (pcb_fanout_t): the local stack and, in a slice-sharded job, the
// peer-mapped stacks of the other GPUs - the exchange is these stores, issued as soon as a tile is done, so the NVLink
// traffic overlaps the arithmetic of the CTAs that are still running.
template <int KG>
__global__ void __launch_bounds__(256)
refine_v_kernel(const float *__restrict__ H, int M, int S, int E, const float *__restrict__ dy_vals,
                const int32_t *__restrict__ dy_first, const int32_t *__restrict__ iy, RefineHCtl ctl, pcb_fanout_t dst,
                int Gy, int g_off) {
    __shared__ __align__(16) float s_w[RF_MAXM * KG * RF_GROUP];      // [active view row][tap][output row of the group]
    __shared__ int s_j[RF_MAXM];           // active view rows
    __shared__ int s_f[RF_MAXM];           // first sample row of this group for each of them
    __shared__ int s_n;
    const int gy = g_off + blockIdx.x, n = blockIdx.z;
    const int e = blockIdx.y * blockDim.x + threadIdx.x;
    if (threadIdx.x < 32) {                // M <= 32: one lane per view row, compacted with a ballot
        const int j = threadIdx.x;
        const bool on = j < M && ctl.active[j] != 0u;
        const unsigned m = __ballot_sync(0xffffffffu, on);
        if (on) {
            const int k = __popc(m & ((1u << j) - 1u));
            const int sidx = __ldg(iy + (size_t)n * M + j);
            s_j[k] = j;
            s_f[k] = __ldg(dy_first + (size_t)sidx * Gy + gy);
            // the weights of this (slice, group): KG * RF_GROUP consecutive floats per view row
            const float *src = dy_vals + ((size_t)sidx * Gy + gy) * (KG * RF_GROUP);
            for (int t = 0; t < KG * RF_GROUP; ++t) s_w[k * KG * RF_GROUP + t] = __ldg(src + t);
        }
        if (j == 0) s_n = __popc(m);
    }
    __syncthreads();
    const int na = s_n;
    float mn = INFINITY, mx = -INFINITY;
    if (e < E) {
        float acc[RF_GROUP];
#pragma unroll
        for (int q = 0; q < RF_GROUP; ++q) acc[q] = 0.f;
        const size_t sE = (size_t)E;
        const float *base = H + ((size_t)n * M) * S * sE + e;
        float v[KG], vn[KG];
        if (na > 0) {
            const float *col = base + ((size_t)s_j[0] * S + s_f[0]) * sE;
#pragma unroll
            for (int m = 0; m < KG; ++m) v[m] = __ldg(col + m * sE);
        }
        for (int k = 0; k < na; ++k) {
            if (k + 1 < na) {
                const float *col = base + ((size_t)s_j[k + 1] * S + s_f[k + 1]) * sE;
#pragma unroll
                for (int m = 0; m < KG; ++m) vn[m] = __ldg(col + m * sE);
            }
            const float4 *w = reinterpret_cast<const float4 *>(s_w + k * KG * RF_GROUP);
#pragma unroll
            for (int m = 0; m < KG; ++m) {
                const float4 wa = w[2 * m], wb = w[2 * m + 1];       // same address for the whole CTA: broadcast
                acc[0] = fmaf(wa.x, v[m], acc[0]); acc[1] = fmaf(wa.y, v[m], acc[1]);
                acc[2] = fmaf(wa.z, v[m], acc[2]); acc[3] = fmaf(wa.w, v[m], acc[3]);
                acc[4] = fmaf(wb.x, v[m], acc[4]); acc[5] = fmaf(wb.y, v[m], acc[5]);
                acc[6] = fmaf(wb.z, v[m], acc[6]); acc[7] = fmaf(wb.w, v[m], acc[7]);
            }
#pragma unroll
            for (int m = 0; m < KG; ++m) v[m] = vn[m];
        }
        const float invM = 1.0f / (float)M;
        const size_t o_n = (size_t)n * S * E + e;
#pragma unroll
        for (int q = 0; q < RF_GROUP; ++q) {
            const int y = gy * RF_GROUP + q;
            if (y < S) {
                const float val = acc[q] * invM;
#pragma unroll
                for (int d = 0; d < PCB_MAX_PEERS; ++d)
                    if (d < dst.n) dst.out[d][o_n + (size_t)y * E] = val;
                mn = fminf(mn, val);
                mx = fmaxf(mx, val);
            }
        }
    }
    if (dst.mm[0] != nullptr) block_minmax_commit_fanout(mn, mx, dst);
}

// ---- the summed UP-SAMPLED image of one slice (slow path) -----------------------------------------------------------
// The reference writes, per refined slice, the up-sampled sum itself as upscale_*.png (lfp_shiftandsum.py:126-132).  With
// coef = B-spline coefficients of the views, Up_ji = Ey . coef_ji . Ex^T (4 taps per fine sample and axis) and
//     U_a[Y, X] = (1/M) sum_{(j,i) active} Up_ji[clamp(Y + a (c0 - j)), clamp(X + a (c0 - i))]
// separates into a horizontal pass per view row (G_j, S x T M) and a vertical pass over the view rows.  732 MB per slice
// at Illum size: offered for exporters that want the file, not part of the timed path.
__global__ void __launch_bounds__(256)
refi_up_h_kernel(const float *__restrict__ coef, int M, int S, int T, int TM, int a, const int32_t *__restrict__ fx,
                 const float4 *__restrict__ wx, RefineHCtl ctl, float *__restrict__ G) {
    const int X = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y, j = blockIdx.z;
    const uint32_t act = ctl.active[j];
    if (act == 0u || X >= TM) return;
    const int c0 = M / 2;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    for (int i = 0; i < M; ++i) {
        if (!((act >> i) & 1u)) continue;
        const int Xs = min(max(X + a * (c0 - i), 0), TM - 1);
        const int f = __ldg(fx + Xs);
        const float4 w = __ldg(wx + Xs);
        const float *row = coef + (((size_t)(j * M + i) * S + r) * T + f) * 3;
        const float wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int o = min(f + q, T - 1) - f;                 // a padded tap (weight 0) never leaves the row
            a0 = fmaf(wv[q], __ldg(row + o * 3 + 0), a0);
            a1 = fmaf(wv[q], __ldg(row + o * 3 + 1), a1);
            a2 = fmaf(wv[q], __ldg(row + o * 3 + 2), a2);
        }
    }
    float *o = G + (((size_t)j * S + r) * TM + X) * 3;
    o[0] = a0; o[1] = a1; o[2] = a2;
}

__global__ void __launch_bounds__(256)
refi_up_v_kernel(const float *__restrict__ G, int M, int S, int SM, int E, int a, const int32_t *__restrict__ fy,
                 const float4 *__restrict__ wy, RefineHCtl ctl, float *__restrict__ U) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x, Y = blockIdx.y;
    if (e >= E) return;
    const int c0 = M / 2;
    float acc = 0.f;
    for (int j = 0; j < M; ++j) {
        if (ctl.active[j] == 0u) continue;
        const int Ys = min(max(Y + a * (c0 - j), 0), SM - 1);
        const int f = __ldg(fy + Ys);
        const float4 w = __ldg(wy + Ys);
        const float *col = G + ((size_t)j * S + f) * E + e;
        const float wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int p = 0; p < 4; ++p) acc = fmaf(wv[p], __ldg(col + (size_t)(min(f + p, S - 1) - f) * E), acc);
    }
    U[(size_t)Y * E + e] = acc * (1.0f / (float)M);
}

static int launch_refi_upscaled(const float *coef, int M, int S, int T, int a, const int32_t *fx, const float *wx,
                                const int32_t *fy, const float *wy, const uint8_t *view_zeroed_host, float *G, float *U,
                                cudaStream_t st) {
    RefineHCtl ctl;
    memset(&ctl, 0, sizeof(ctl));
    for (int j = 0; j < M; ++j)
        for (int i = 0; i < M; ++i)
            if (!view_zeroed_host[j * M + i]) ctl.active[j] |= 1u << i;
    const int TM = T * M, SM = S * M, E = TM * 3;
    refi_up_h_kernel<<<dim3((TM + 255) / 256, S, M), 256, 0, st>>>(coef, M, S, T, TM, a, fx,
                                                                   reinterpret_cast<const float4 *>(wx), ctl, G);
    PCB_LAUNCH_OK();
    refi_up_v_kernel<<<dim3((E + 255) / 256, SM), 256, 0, st>>>(G, M, S, SM, E, a, fy, reinterpret_cast<const float4 *>(wy),
                                                               ctl, U);
    PCB_LAUNCH_OK();
    return PCB_OK;
}

// ---- launchers ---------------------------------------------------------------------------------------------------
// tx / ty: interior Toeplitz description of the two axes (lo > hi: none)
// rows: optional restriction (row-sharded stacks): the x pass runs for rows [tmp_lo, tmp_hi) only and the y pass produces
// coefficient rows [h_lo, h_hi) (whole strips / groups that intersect it).
template <typename TIn>
static int launch_refi_prefilter(const void *vp, int M, int S, int T, const float *py_vals, const int32_t *py_first,
                                 int Kpy, const float *px_vals, const int32_t *px_first, int Kpx,
                                 const uint8_t *view_zeroed, float *tmp, float *coef, cudaStream_t st,
                                 const RpTaps *tx = nullptr, const RpTaps *ty = nullptr,
                                 const pcb_refi_rows_t *rows = nullptr) {
    const int E = T * 3;
    const size_t smem = (size_t)E * sizeof(float);
    if (2 * smem > 200 * 1024) return fail(PCB_ERR_UNSUPPORTED, "%s: image row does not fit shared memory", "pcb_refi_prefilter");
    const int xr0 = rows ? max(rows->tmp_lo, 0) : 0, xr1 = rows ? min(rows->tmp_hi, S) : S;
    const int ca = rows ? max(rows->h_lo, 0) : 0, cb = rows ? min(rows->h_hi, S) : S;
    if (xr1 <= xr0 || cb <= ca) return fail(PCB_ERR_INVALID, "%s: empty row range", "pcb_refi_prefilter");
    if (tx && tx->hi - tx->lo + 1 >= 2 * RPX_P) {
        PCB_CUDA(cudaFuncSetAttribute(refi_prefilter_x3_kernel<TIn>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * smem)));
        refi_prefilter_x3_kernel<TIn><<<dim3(xr1 - xr0, M * M), RP_THREADS, 2 * smem, st>>>((const TIn *)vp, S, T, px_vals,
                                                                                          px_first, Kpx, *tx, view_zeroed,
                                                                                          tmp, xr0);
    } else {
        PCB_CUDA(cudaFuncSetAttribute(refi_prefilter_x_kernel<TIn>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        refi_prefilter_x_kernel<TIn><<<dim3(xr1 - xr0, M * M), RP_THREADS, smem, st>>>((const TIn *)vp, S, T, px_vals, px_first,
                                                                                     Kpx, view_zeroed, tmp, xr0);
    }
    PCB_LAUNCH_OK();
    const int G = (S + RP_GY - 1) / RP_GY;
    const dim3 gx((E + 255) / 256);
    const int ga = ca / RP_GY, gb = (cb + RP_GY - 1) / RP_GY;            // table groups that intersect [ca, cb)
    // interior rows in strips of RPY_R, aligned to the row groups of the table kernel; the groups before / after keep it
    int ya = 0, strips = 0;
    if (ty && ty->hi >= ty->lo) {
        ya = ((ty->lo + RP_GY - 1) / RP_GY) * RP_GY;
        strips = (ty->hi + 1 - ya) / RPY_R;
        if (strips < 1) strips = 0;
    }
    if (strips > 0) {
        const int s_lo = ca > ya ? (ca - ya) / RPY_R : 0;
        const int s_hi = min(strips, cb > ya ? (cb - ya + RPY_R - 1) / RPY_R : 0);
        if (s_hi > s_lo) {
            refi_prefilter_yt_kernel<<<dim3(gx.x, s_hi - s_lo, M * M), 256, 0, st>>>(tmp, S, E, *ty, ya + s_lo * RPY_R,
                                                                                    view_zeroed, coef);
            PCB_LAUNCH_OK();
        }
        const int g_before = ya / RP_GY, g_after0 = (ya + strips * RPY_R) / RP_GY;
        const int b0 = ga, b1 = min(gb, g_before);                           // border groups before the interior
        if (b1 > b0) {
            refi_prefilter_y_kernel<<<dim3(gx.x, b1 - b0, M * M), 256, 0, st>>>(tmp, S, E, py_vals, py_first, Kpy,
                                                                               view_zeroed, coef, b0);
            PCB_LAUNCH_OK();
        }
        const int a0 = max(ga, g_after0), a1 = min(gb, G);                   // ... and after it
        if (a1 > a0) {
            refi_prefilter_y_kernel<<<dim3(gx.x, a1 - a0, M * M), 256, 0, st>>>(tmp, S, E, py_vals, py_first, Kpy,
                                                                               view_zeroed, coef, a0);
            PCB_LAUNCH_OK();
        }
    } else {
        refi_prefilter_y_kernel<<<dim3(gx.x, min(gb, G) - ga, M * M), 256, 0, st>>>(tmp, S, E, py_vals, py_first, Kpy,
                                                                                  view_zeroed, coef, ga);
        PCB_LAUNCH_OK();
    }
    return PCB_OK;
}

template <int K>
static int launch_refine_h(const float *coef, int M, int S, int T, int nb, const float *dx_vals, const int32_t *dx_first,
                           const int32_t *ix, const RefineHCtl &ctl, size_t smem, float *H, cudaStream_t st) {
    PCB_CUDA(cudaFuncSetAttribute(refine_h_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((T + RH_XC - 1) / RH_XC, (S + RH_ROWS - 1) / RH_ROWS, M);
    refine_h_kernel<K><<<grid, RH_THREADS, smem, st>>>(coef, M, S, T, nb, dx_vals, dx_first, ix, ctl, H);
    PCB_LAUNCH_OK();
    return PCB_OK;
}

// Horizontal pass in slice quads; returns PCB_ERR_UNSUPPORTED (without touching g_err's meaning for the caller) when the
// tables do not fit this kernel, so that launch_refine falls back to refine_h_kernel.
static int launch_refine_hq(const float *coef, int M, int S, int T, int nb, const pcb_refi_quads_t &qd,
                            const uint8_t *view_zeroed_host, float *H, cudaStream_t st,
                            const pcb_refi_rows_t *rows = nullptr) {
    if (M > RF_MAXM || qd.G * RQ < nb || qd.wq == nullptr || qd.fu == nullptr) return PCB_ERR_UNSUPPORTED;
    RefineQCtl ctl;
    memset(&ctl, 0, sizeof(ctl));
    size_t max_px = 0;
    for (int j = 0; j < M; ++j) {
        size_t px = 0;
        for (int i = 0; i < M; ++i) {
            if (view_zeroed_host[j * M + i]) continue;
            if (qd.W[i] < 1 || qd.W[i] > RQ_MAXW || qd.hi[i] < qd.lo[i]) return PCB_ERR_UNSUPPORTED;
            ctl.active[j] |= 1u << i;
            px += (size_t)RQ_XC + (size_t)(qd.hi[i] - qd.lo[i]);
        }
        if (px > max_px) max_px = px;
    }
    for (int i = 0; i < M; ++i) { ctl.lo[i] = qd.lo[i]; ctl.hi[i] = qd.hi[i]; ctl.woff[i] = qd.woff[i]; ctl.W[i] = qd.W[i]; }
    const size_t smem = max_px * RQ_PIX * sizeof(float);
    if (smem > 220 * 1024 || (S + RQ_ROWS - 1) / RQ_ROWS > 65535) return PCB_ERR_UNSUPPORTED;
    PCB_CUDA(cudaFuncSetAttribute(refine_hq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int nblk = (S + RQ_ROWS - 1) / RQ_ROWS;
    const int b0 = rows ? max(rows->h_lo, 0) / RQ_ROWS : 0;
    const int b1 = rows ? min((min(rows->h_hi, S) + RQ_ROWS - 1) / RQ_ROWS, nblk) : nblk;
    if (b1 <= b0) return fail(PCB_ERR_INVALID, "%s: empty row range", "pcb_refocus_refi");
    dim3 grid((T + RQ_XC - 1) / RQ_XC, b1 - b0, M);
    refine_hq_kernel<<<grid, RQ_THREADS, smem, st>>>(coef, M, S, T, nb, qd.G, reinterpret_cast<const float4 *>(qd.wq), qd.fu,
                                                     ctl, H, b0);
    PCB_LAUNCH_OK();
    return PCB_OK;
}

static int launch_refine_v(const float *H, int M, int S, int T, int nb, const float *dy_vals, const int32_t *dy_first,
                           int Kgy, const int32_t *iy, const uint8_t *view_zeroed_host, const pcb_fanout_t &dst,
                           cudaStream_t st, const pcb_refi_rows_t *rows = nullptr) {
    const int E = T * 3;
    RefineHCtl ctl;
    memset(&ctl, 0, sizeof(ctl));
    for (int j = 0; j < M; ++j)
        for (int i = 0; i < M; ++i)
            if (!view_zeroed_host[j * M + i]) ctl.active[j] |= 1u << i;
    const int Gy = (S + RF_GROUP - 1) / RF_GROUP;
    int g0 = 0, g1 = Gy;
    if (rows) {
        if (rows->out_lo % RF_GROUP != 0 || (rows->out_hi % RF_GROUP != 0 && rows->out_hi < S))
            return fail(PCB_ERR_INVALID, "%s: output row range must start / end on multiples of %lld", "pcb_refocus_refi", RF_GROUP);
        g0 = max(rows->out_lo, 0) / RF_GROUP;
        g1 = min((min(rows->out_hi, S) + RF_GROUP - 1) / RF_GROUP, Gy);
        if (g1 <= g0) return fail(PCB_ERR_INVALID, "%s: empty row range", "pcb_refocus_refi");
    }
    const dim3 grid(g1 - g0, (E + 255) / 256, nb);
    if (Kgy > S) return fail(PCB_ERR_INVALID, "%s: vertical window of %lld rows exceeds the image", "pcb_refocus_refi", Kgy);
#define PCB_RV(K) case K: refine_v_kernel<K><<<grid, 256, 0, st>>>(H, M, S, E, dy_vals, dy_first, iy, ctl, dst, Gy, g0); break;
    switch (Kgy) {
        PCB_RV(1) PCB_RV(2) PCB_RV(3) PCB_RV(4) PCB_RV(5) PCB_RV(6) PCB_RV(7) PCB_RV(8) PCB_RV(9) PCB_RV(10)
        PCB_RV(11) PCB_RV(12) PCB_RV(13) PCB_RV(14) PCB_RV(15) PCB_RV(16) PCB_RV(17) PCB_RV(18) PCB_RV(19) PCB_RV(20)
    default: return fail(PCB_ERR_UNSUPPORTED, "%s: vertical window of %lld rows (1..20 supported)", "pcb_refocus_refi", Kgy);
    }
#undef PCB_RV
    PCB_LAUNCH_OK();
    return PCB_OK;
}

static int launch_refine(const float *coef, int M, int S, int T, int nb, const float *dy_vals, const int32_t *dy_first,
                         int Kgy, const float *dx_vals, const int32_t *dx_first, int Kx, const int32_t *iy,
                         const int32_t *ix, const int32_t *xlo_host, const int32_t *xhi_host,
                         const uint8_t *view_zeroed_host, float *H, const pcb_fanout_t &dst, cudaStream_t st,
                         const pcb_refi_quads_t *quads = nullptr, const pcb_refi_rows_t *rows = nullptr) {
    if (quads != nullptr && !(getenv("PCB_REFINE_KERNEL") && !strcmp(getenv("PCB_REFINE_KERNEL"), "taps"))) {
        const int rq = launch_refine_hq(coef, M, S, T, nb, *quads, view_zeroed_host, H, st, rows);
        if (rq == PCB_OK)
            return launch_refine_v(H, M, S, T, nb, dy_vals, dy_first, Kgy, iy, view_zeroed_host, dst, st, rows);
        if (rq != PCB_ERR_UNSUPPORTED) return rq;
    }
    const int E = T * 3;
    RefineHCtl ctl;
    memset(&ctl, 0, sizeof(ctl));
    size_t max_px = 0;
    for (int j = 0; j < M; ++j) {
        size_t px = 0;
        for (int i = 0; i < M; ++i) {
            if (view_zeroed_host[j * M + i]) continue;
            ctl.active[j] |= 1u << i;
            px += (size_t)RH_XC + (size_t)(xhi_host[i] - xlo_host[i]);
        }
        if (px > max_px) max_px = px;
    }
    for (int i = 0; i < M; ++i) { ctl.lo[i] = xlo_host[i]; ctl.hi[i] = xhi_host[i]; }
    const size_t smem = max_px * 3 * RH_ROWS * sizeof(float);
    if (nb * M > RH_MAX_IX)
        return fail(PCB_ERR_INVALID, "%s: at most %lld slice x view entries per call", "pcb_refocus_refi", RH_MAX_IX);
    if (smem > 200 * 1024)
        return fail(PCB_ERR_UNSUPPORTED, "%s: coefficient windows of one view row do not fit shared memory", "pcb_refocus_refi");
    int rc;
    switch (Kx) {
    case 1: rc = launch_refine_h<1>(coef, M, S, T, nb, dx_vals, dx_first, ix, ctl, smem, H, st); break;
    case 2: rc = launch_refine_h<2>(coef, M, S, T, nb, dx_vals, dx_first, ix, ctl, smem, H, st); break;
    case 3: rc = launch_refine_h<3>(coef, M, S, T, nb, dx_vals, dx_first, ix, ctl, smem, H, st); break;
    case 4: rc = launch_refine_h<4>(coef, M, S, T, nb, dx_vals, dx_first, ix, ctl, smem, H, st); break;
    case 5: rc = launch_refine_h<5>(coef, M, S, T, nb, dx_vals, dx_first, ix, ctl, smem, H, st); break;
    case 6: rc = launch_refine_h<6>(coef, M, S, T, nb, dx_vals, dx_first, ix, ctl, smem, H, st); break;
    case 7: rc = launch_refine_h<7>(coef, M, S, T, nb, dx_vals, dx_first, ix, ctl, smem, H, st); break;
    case 8: rc = launch_refine_h<8>(coef, M, S, T, nb, dx_vals, dx_first, ix, ctl, smem, H, st); break;
    case 9: rc = launch_refine_h<9>(coef, M, S, T, nb, dx_vals, dx_first, ix, ctl, smem, H, st); break;
    case 10: rc = launch_refine_h<10>(coef, M, S, T, nb, dx_vals, dx_first, ix, ctl, smem, H, st); break;
    case 11: rc = launch_refine_h<11>(coef, M, S, T, nb, dx_vals, dx_first, ix, ctl, smem, H, st); break;
    case 12: rc = launch_refine_h<12>(coef, M, S, T, nb, dx_vals, dx_first, ix, ctl, smem, H, st); break;
    default: return fail(PCB_ERR_UNSUPPORTED, "%s: band of %lld taps (1..12 supported)", "pcb_refocus_refi", Kx);
    }
    if (rc) return rc;
    // (the per-slice kernel has no row restriction: it fills every row of H, the vertical pass takes what it needs)
    return launch_refine_v(H, M, S, T, nb, dy_vals, dy_first, Kgy, iy, view_zeroed_host, dst, st, rows);
}

}  // namespace pcb
