// (c) Integer shift-and-sum, row-stationary, for uint16 RGB light fields (C == 3, M <= 16).
//
// Same decomposition as refocus_rows.cuh (one CTA owns the image row r of view-row j and serves every slice
// from shared memory, partial rows are folded into the stack accumulator with integer REDs), but the inner
// loop is rebuilt around what bounds it: shared-memory wavefronts and issue slots.
//
//   * A staged pixel is ONE 64-bit word  R | G << 21 | B << 42.  A view row contributes <= 16 views, and
//     16 * 65535 < 2^20, so the three 21-bit fields never carry into each other: one 64-bit integer add
//     (IADD3 + IADD3.X) accumulates a whole pixel, exactly.  Refocus shifts are whole pixels, so every
//     shifted read is an aligned LDS.64 — no second, element-shifted copy of the row and no hi/lo split.
//     Per pixel and view: 1 IMAD (address), 1 LDS.64, 2 IADD3   (was: 1.5 LDS.32 + 1.5 LEA + 0.75 IADD3 +
//     1.5 IMAD.HI on two row copies, padded to groups of four views).
//   * Row offsets are  a * (c0 - i)  — computed, not tabulated: no table reads in the loop.
//   * The stack accumulator is planar per pixel,  accRG[n][y][x] = R | G << 32  (one 64-bit RED; the low
//     field never carries: totals < 2^24) and accB[n][y][x] (32-bit RED): consecutive lanes hit consecutive
//     addresses, 12 full sectors per 32 pixels.
//   * CTA order is r-major (blockIdx.x = j) as before, so the RED window stays within +-(M/2)|a| rows.
//
// Rows whose source is clamped in y are served from a small edge table written by the r = 0 and r = S-1
// CTAs and added by the finalize pass, which also applies the 1/M scaling and reduces the stack min/max.
#pragma once
#include "common.cuh"
#include "refocus_rows.cuh"

namespace pcb {

constexpr int PX_MAXV = 16;          // views per view-row
constexpr int PX_STAGE_UNROLL = 4;   // pixels in flight per thread and row while staging (x 2 rows x 3 loads)

struct PxCtl {                       // per-launch control block, passed by value (kernel-parameter bank)
    int32_t a[RR_MAX_SLICES];
    uint8_t first[16];               // the unmasked views of view-row j are the columns first[j] .. first[j]+nact[j]-1
    uint8_t nact[16];                // (a circular aperture always leaves one contiguous run per row)
    uint32_t row_active;             // bit j set -> view-row j has at least one unmasked view
};

struct PxArgs {
    const uint16_t *vp;
    unsigned long long *accRG;   // [N][S][Tp]
    uint32_t *accB;              // [N][S][Tp]
    unsigned long long *edgeRG;  // [N][M][2][Tp]
    uint32_t *edgeB;             // [N][M][2][Tp]
    int M, S, T, N;
    int Tp;        // accumulator row pitch in pixels (multiple of 4)
    int pad;       // replicated pixels on each side of a staged row (max |shift|, capped at T)
    int chunk;     // output pixels per CTA
    int nxc;       // x chunks
    int rowPx;     // staged pixels per row  (threads*PPT + 2*pad)
    int nsl, n0;   // slices of this launch, first slice
};

__device__ __forceinline__ unsigned long long pack_px(uint32_t r, uint32_t g, uint32_t b) {
    const uint32_t lo = r | (g << 21);
    const uint32_t hi = (g >> 11) | (b << 10);
    return ((unsigned long long)hi << 32) | lo;
}

// Stage one pair of view rows.  FULL: every pixel slot of the unrolled tile exists (no per-slot guards, so all
// 6 * PX_STAGE_UNROLL loads of a thread are in flight before the first is consumed).
template <bool FULL>
__device__ __forceinline__ void px_stage_tile(const uint16_t *srcA, const uint16_t *srcB, unsigned long long *dA,
                                              unsigned long long *dB, bool two, int base, int THREADS, int need,
                                              int xbase, int Tm1, uint64_t pol) {
    uint16_t va[PX_STAGE_UNROLL][3], vb[PX_STAGE_UNROLL][3];
#pragma unroll
    for (int u = 0; u < PX_STAGE_UNROLL; ++u) {
        const int idx = base + u * THREADS;
        if (FULL || idx < need) {
            const int x = min(max(xbase + idx, 0), Tm1) * 3;
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
                va[u][ch] = ld_stream_u16(srcA + x + ch, pol);
                vb[u][ch] = ld_stream_u16(srcB + x + ch, pol);
            }
        }
    }
#pragma unroll
    for (int u = 0; u < PX_STAGE_UNROLL; ++u) {
        const int idx = base + u * THREADS;
        if (FULL || idx < need) {
            dA[idx] = pack_px(va[u][0], va[u][1], va[u][2]);
            if (two) dB[idx] = pack_px(vb[u][0], vb[u][1], vb[u][2]);
        }
    }
}

// All slices of one CTA with the number of active views known at compile time: the NACT shifted LDS.64 of a
// slice are independent and issued back to back.  Address of a read = (per-thread base) + (warp-uniform offset).
template <int NACT, int PPT, bool CLAMP>
__device__ __forceinline__ void px_slices(const PxArgs &p, const PxCtl &ctl, const char *sbase, int THREADS, int j,
                                          int r, int x0, int tid) {
    const int c0 = p.M / 2;
    const int rowBytes = p.rowPx * 8;
    const bool is_top = (r == 0), is_bot = (r == p.S - 1);
    bool valid[PPT];
#pragma unroll
    for (int k = 0; k < PPT; ++k) {
        const int t = tid + k * THREADS;
        valid[k] = t < p.chunk && x0 + t < p.T;
    }
    unsigned long long *rg_t = p.accRG + x0 + tid;       // 32-bit row offsets are added per slice (plan: < 2^31)
    uint32_t *b_t = p.accB + x0 + tid;
    const int dj = c0 - j;
    const int d0 = c0 - (int)ctl.first[j];
    for (int al = 0; al < p.nsl; ++al) {
        const int a = ctl.a[al];
        unsigned long long v[NACT][PPT];
        // view ii = column first+ii is read at  ii*rowBytes + a*(c0 - first - ii)*8 : linear in ii
        const char *q0 = sbase + a * d0 * 8;
        const int qstep = rowBytes - a * 8;
#pragma unroll
        for (int ii = 0; ii < NACT; ++ii) {
            const char *q = q0 + ii * qstep;
            if (CLAMP) {
                const int s = max(-p.T, min(p.T, a * (d0 - ii)));
                q = sbase + ii * rowBytes + s * 8;
            }
#pragma unroll
            for (int k = 0; k < PPT; ++k)
                v[ii][k] = *reinterpret_cast<const unsigned long long *>(q + k * THREADS * 8);
        }
        unsigned long long tot[PPT];
#pragma unroll
        for (int k = 0; k < PPT; ++k) {
            unsigned long long s0 = v[0][k], s1 = 0ull;
#pragma unroll
            for (int ii = 1; ii < NACT; ++ii) {
                if (ii & 1) s1 += v[ii][k]; else s0 += v[ii][k];
            }
            tot[k] = s0 + s1;
        }
        const int n = p.n0 + al;
        const int y = r - a * dj;
        const bool in_range = (y >= 0 && y < p.S);
        const uint32_t ro = (uint32_t)(n * p.S + (in_range ? y : 0)) * (uint32_t)p.Tp;
#pragma unroll
        for (int k = 0; k < PPT; ++k) {
            if (!valid[k]) continue;
            const uint32_t lo = (uint32_t)tot[k], hi = (uint32_t)(tot[k] >> 32);
            const uint32_t R = lo & 0x1FFFFFu;
            const uint32_t G = __funnelshift_r(lo, hi, 21) & 0x1FFFFFu;
            const uint32_t B = hi >> 10;
            const unsigned long long rg = ((unsigned long long)G << 32) | R;
            if (in_range) {
                atomicAdd(rg_t + ro + k * THREADS, rg);
                atomicAdd(b_t + ro + k * THREADS, B);
            }
            if (is_top || is_bot) {
                const size_t e = (size_t)(n * p.M + j) * 2 * p.Tp + x0 + tid + k * THREADS;
                if (is_top) { p.edgeRG[e] = rg; p.edgeB[e] = B; }
                if (is_bot) { p.edgeRG[e + p.Tp] = rg; p.edgeB[e + p.Tp] = B; }
            }
        }
    }
}

template <int PPT, bool CLAMP>
__global__ void __launch_bounds__(512, 2)
refocus_px_kernel(PxArgs p, PxCtl ctl) {
    const int THREADS = blockDim.x;
    extern __shared__ __align__(16) unsigned long long srow[];   // [nact][rowPx]

    const int j = blockIdx.x;           // fastest: the M view-rows of one image row r run together
    const int r = blockIdx.y;
    const int x0 = blockIdx.z * p.chunk;
    const int tid = threadIdx.x;
    const int nact = ctl.nact[j];
    if (nact == 0) return;

    // ---- stage: every active view row, edge-replicated by `pad` pixels, as packed 64-bit pixels
    {
        const size_t view_stride = (size_t)p.S * p.T * 3;
        const uint16_t *row0 = p.vp + ((size_t)j * p.M * p.S + r) * (size_t)p.T * 3;
        const uint64_t pol = l2_evict_first_policy();
        const int need = p.rowPx;
        const int xbase = x0 - p.pad;
        const int tile = THREADS * PX_STAGE_UNROLL;
        for (int ii = 0; ii < nact; ii += 2) {
            const bool two = ii + 1 < nact;
            const uint16_t *srcA = row0 + (size_t)(ctl.first[j] + ii) * view_stride;
            const uint16_t *srcB = srcA + (two ? view_stride : 0);
            unsigned long long *dA = srow + (size_t)ii * need;
            unsigned long long *dB = dA + need;
            int t0 = 0;
            for (; t0 + tile <= need; t0 += tile)
                px_stage_tile<true>(srcA, srcB, dA, dB, two, t0 + tid, THREADS, need, xbase, p.T - 1, pol);
            if (t0 < need)
                px_stage_tile<false>(srcA, srcB, dA, dB, two, t0 + tid, THREADS, need, xbase, p.T - 1, pol);
        }
    }
    __syncthreads();

    const char *sbase = reinterpret_cast<const char *>(srow) + (size_t)(tid + p.pad) * 8;
#define PCB_PX_NACT(K) case K: px_slices<K, PPT, CLAMP>(p, ctl, sbase, THREADS, j, r, x0, tid); break;
    switch (nact) {
        PCB_PX_NACT(1) PCB_PX_NACT(2) PCB_PX_NACT(3) PCB_PX_NACT(4) PCB_PX_NACT(5) PCB_PX_NACT(6) PCB_PX_NACT(7)
        PCB_PX_NACT(8) PCB_PX_NACT(9) PCB_PX_NACT(10) PCB_PX_NACT(11) PCB_PX_NACT(12) PCB_PX_NACT(13)
        PCB_PX_NACT(14) PCB_PX_NACT(15) PCB_PX_NACT(16)
    }
#undef PCB_PX_NACT
}

// The edge tables, made cumulative over the view rows in the order "clamped first".  Output row y of slice a takes the
// view rows whose source row y + a (c0 - j) leaves the image from the edge table; for one side (top / bottom) those rows
// are always a prefix or a suffix of j (the source row is monotone in j), so with running sums along j stored in place
// the finalize pass needs ONE table row per side instead of one per clamped view row (a chain of dependent L2 loads that
// cost 30 % of that pass).  Rows without an unmasked view carry the sum on.  Integer sums: exact, same bits.
__global__ void __launch_bounds__(256)
edge_cumsum_kernel(unsigned long long *__restrict__ edgeRG, uint32_t *__restrict__ edgeB, PxCtl ctl, int M, int Tp, int n0) {
    const int x = blockIdx.x * 256 + threadIdx.x;
    const int sl = blockIdx.y >> 1, side = blockIdx.y & 1;
    const int a = ctl.a[sl];
    if (x >= Tp || a == 0) return;                         // a == 0: no source row ever leaves the image
    const int n = n0 + sl;
    const bool desc = (a > 0) == (side == 0);              // top of a > 0 and bottom of a < 0 clamp the LARGE j first
    // all loads first (a load behind a store to the same array would wait for it: M round trips instead of one)
    unsigned long long vrg[16];
    uint32_t vb[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const int j = desc ? M - 1 - k : k;
        vrg[k] = 0ull; vb[k] = 0u;
        if (k < M && ((ctl.row_active >> j) & 1u)) {
            const size_t e = (size_t)((n * M + j) * 2 + side) * Tp + x;
            vrg[k] = edgeRG[e]; vb[k] = edgeB[e];
        }
    }
    unsigned long long crg = 0ull;
    uint32_t cb = 0u;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        if (k < M) {
            const int j = desc ? M - 1 - k : k;
            const size_t e = (size_t)((n * M + j) * 2 + side) * Tp + x;
            crg += vrg[k]; cb += vb[k];
            edgeRG[e] = crg;
            edgeB[e] = cb;
        }
    }
}

static int launch_edge_cumsum(unsigned long long *edgeRG, uint32_t *edgeB, const PxCtl &ctl, int M, int Tp, int n0, int nloc,
                              cudaStream_t st) {
    edge_cumsum_kernel<<<dim3((Tp + 255) / 256, nloc * 2), 256, 0, st>>>(edgeRG, edgeB, ctl, M, Tp, n0);
    PCB_LAUNCH_OK();
    return PCB_OK;
}

// out[n,y,x,:] = (acc[n,y,x] + sum_{j: source row clamped} edge[n,j,side,x]) / M ;  fused min/max.
// The three floats of a pixel are transposed through shared memory (warp-locally) so that the row leaves in
// 128-byte runs.
constexpr int PXF_THREADS = 256;
constexpr int PXF_IT = 3;          // row segments whose accumulator loads are in flight together

// POST fuses the refocuser's colour management into this epilogue (LfpRefocuser.shift_and_sum,
// lfp_refocuser/top_level.py:68-70): 0 = linear stack; 1 = clip((x - lo) / (hi - lo)) + sRGB with the bounds in
// lohi[2], the min / max of the LINEAR values still go to *mm; 2 = fix-up pass, launched after every POST = 1 launch of
// the stack has finished: Normalizer replaces a bound that is exactly 0 by the data min / max (misc/normalizer.py:
// 40-41), known only then — the pass returns at once unless that changes a bound, else it rewrites the stack.
template <bool EXACT24, int POST = 0>
__global__ void __launch_bounds__(PXF_THREADS)
refocus_px_finalize_kernel(unsigned long long *accRG, uint32_t *accB,
                           const unsigned long long *__restrict__ edgeRG, const uint32_t *__restrict__ edgeB,
                           PxCtl ctl, float *__restrict__ out, int M, int S, int T, int Tp, int n0, int nsl,
                           pcb_minmax_t *mm, const double *__restrict__ lohi = nullptr, int leave_zero = 0) {
    double lo = 0.0, inv = 1.0;
    bool scale = false;
    // leave_zero (PCB_REFOCUS_LEAVE_ZEROED): the LAST pass that reads an accumulator word stores zero behind it, so the
    // next exposure starts from a clean accumulator without a memset.  POST 0: this pass.  POST 1: this pass, unless the
    // upper bound is exactly 0 - only then can the fix-up pass need the sums again (a lower bound of 0 implies a data
    // minimum of 0: the values are non-negative and slice 0 belongs to the stack), and then the fix-up pass zeroes (when
    // it does not run either, every sum is 0 already).
    bool zero_acc = leave_zero != 0;
    if (POST) {
        lo = lohi[0];
        double hi = lohi[1];
        if (POST == 1) zero_acc = zero_acc && hi != 0.0;
        if (POST == 2) {
            const double dmin = (double)key_to_float(mm->kmin), dmax = (double)key_to_float(mm->kmax);
            if (!((lo == 0.0 && dmin != 0.0) || (hi == 0.0 && dmax != 0.0))) return;
            if (lo == 0.0) lo = dmin;
            if (hi == 0.0) hi = dmax;
        }
        scale = (hi != lo) && (hi != 0.0);
        inv = scale ? 1.0 / (hi - lo) : 1.0;
    }
    __shared__ float s_tile[PXF_THREADS * 3];
    DivByConst divM;
    divM.init((float)M);
    float mn = INFINITY, mx = -INFINITY;
    // a CTA walks a few (slice, row) pairs (grid-strided), so that the prologue above and the block reduction below are
    // paid once per CTA, not once per row; the fix-up pass runs a small grid: the launch that usually has nothing to do
    // costs microseconds
    for (int row = blockIdx.x; row < S * nsl; row += (int)gridDim.x) {
    const int sl = row / S;
    const int y = row - sl * S, n = n0 + sl;
    // the view rows whose source row leaves the image at the top / at the bottom are a prefix or a suffix of j: one row
    // of the CUMULATIVE edge table (edge_cumsum_kernel) per side, that of the least clamped j.  Every warp works it out
    // for itself (lane = view row), no shared memory, no barrier.
    int offT = -1, offB = -1;
    {
        const int a = ctl.a[sl], lj = threadIdx.x & 31;
        const int ys = y + a * (M / 2 - lj);
        const unsigned mt = __ballot_sync(0xffffffffu, lj < M && ys < 0);
        const unsigned mb = __ballot_sync(0xffffffffu, lj < M && ys > S - 1);
        if (mt) offT = ((n * M + (a > 0 ? __ffs(mt) - 1 : 31 - __clz(mt))) * 2 + 0) * Tp;
        if (mb) offB = ((n * M + (a > 0 ? 31 - __clz(mb) : __ffs(mb) - 1)) * 2 + 1) * Tp;
    }
    const size_t arow = ((size_t)n * S + y) * Tp;
    float *orow = out + ((size_t)n * S + y) * (size_t)T * 3;
    // the accumulator words of PXF_IT row segments are requested up front (the kernel is bound by the latency of
    // these loads, not by their bandwidth); each segment then goes through the shared-memory transposition
    for (int xb0 = 0; xb0 < T; xb0 += PXF_THREADS * PXF_IT) {
        unsigned long long rg[PXF_IT];
        uint32_t b[PXF_IT];
#pragma unroll
        for (int k = 0; k < PXF_IT; ++k) {
            const int x = xb0 + k * PXF_THREADS + threadIdx.x;
            rg[k] = 0ull; b[k] = 0u;
            if (x < T) { rg[k] = accRG[arow + x]; b[k] = accB[arow + x]; }
        }
        // ... and so are the (at most two) edge-table rows: every load of the row is in flight before the first sum
        unsigned long long et[PXF_IT], eb[PXF_IT];
        uint32_t ft[PXF_IT], fb[PXF_IT];
#pragma unroll
        for (int k = 0; k < PXF_IT; ++k) { et[k] = eb[k] = 0ull; ft[k] = fb[k] = 0u; }
        if (offT >= 0) {
#pragma unroll
            for (int k = 0; k < PXF_IT; ++k) {
                const int x = xb0 + k * PXF_THREADS + threadIdx.x;
                if (x < T) { et[k] = __ldg(edgeRG + offT + x); ft[k] = __ldg(edgeB + offT + x); }
            }
        }
        if (offB >= 0) {
#pragma unroll
            for (int k = 0; k < PXF_IT; ++k) {
                const int x = xb0 + k * PXF_THREADS + threadIdx.x;
                if (x < T) { eb[k] = __ldg(edgeRG + offB + x); fb[k] = __ldg(edgeB + offB + x); }
            }
        }
#pragma unroll
        for (int k = 0; k < PXF_IT; ++k) { rg[k] += et[k] + eb[k]; b[k] += ft[k] + fb[k]; }
#pragma unroll
        for (int k = 0; k < PXF_IT; ++k) {
            const int xb = xb0 + k * PXF_THREADS;
            if (xb >= T) break;
            const int x = xb + threadIdx.x;
            if (x < T) {
                const uint32_t v[3] = {(uint32_t)rg[k], (uint32_t)(rg[k] >> 32), b[k]};
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) {
                    const float f = EXACT24 ? divM((float)v[ch]) : (float)((double)v[ch] / (double)M);
                    s_tile[threadIdx.x * 3 + ch] = POST ? contrast_srgb_one(f, lo, inv, scale) : f;
                    mn = fminf(mn, f);
                    mx = fmaxf(mx, f);
                }
            }
            // a warp's 32 pixels are 96 consecutive floats of the output row: the transposition stays inside the warp
            __syncwarp();
            const int wb = (threadIdx.x & ~31) * 3, lane = threadIdx.x & 31;
            const int ne = (min(PXF_THREADS, T - xb) - (int)(threadIdx.x & ~31)) * 3;     // elements this warp owns
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                const int e = lane + q * 32;
                if (e < ne) orow[(size_t)xb * 3 + wb + e] = s_tile[wb + e];
            }
            __syncwarp();
        }
        // every thread zeroes the words it read itself (no other thread reads them), and only now that its loads have
        // long been answered: a store that meets the pending fill of its own line in L2 is slow (zeroing right behind
        // each load made the pass 7x slower)
        if (zero_acc) {
#pragma unroll
            for (int k = 0; k < PXF_IT; ++k) {
                const int x = xb0 + k * PXF_THREADS + threadIdx.x;
                if (x < T) { accRG[arow + x] = 0ull; accB[arow + x] = 0u; }
            }
        }
    }
    }
    if (POST != 2 && mm != nullptr) block_minmax_commit(mn, mx, mm);
}

constexpr int PXF_ROWS_PER_CTA = 4;
// a few rows per CTA once there are enough rows to fill the GPU several times over (a single slice stays one row per CTA)
static dim3 finalize_grid(int rows) {
    const int per = rows >= 16 * PXF_ROWS_PER_CTA * sm_count() ? PXF_ROWS_PER_CTA : 1;
    return dim3((rows + per - 1) / per, 1);
}

struct PxPlan {
    bool ok;
    PxArgs a;
    size_t smem;
    int threads, ppt;
    bool clamp, exact24;
    size_t accRG_bytes, accB_bytes, edgeRG_bytes, edgeB_bytes;
    size_t scratch_bytes() const { return accRG_bytes + accB_bytes + edgeRG_bytes + edgeB_bytes; }
};

static int px_force_nxc() {          // tuning knob for A/B measurements: PCB_PX_NXC=1|2|3...
    const char *e = getenv("PCB_PX_NXC");
    return e ? atoi(e) : 0;
}

// Host-side planning (no device access).  Prefers the coarsest x chunking that still lets two CTAs share an
// SM (one stages while the other computes); falls back to one CTA per SM for long rows / large shifts.
static PxPlan plan_refocus_px(int M, int S, int T, int N, int max_abs_a, int max_active_per_row,
                              int n_active_total, const uint8_t *view_zeroed_host) {
    PxPlan pl;
    memset(&pl, 0, sizeof(pl));
    pl.ok = false;
    for (int j = 0; j < M; ++j) {        // every view-row must keep ONE contiguous run of views
        int runs = 0;
        for (int i = 0; i < M; ++i)
            if (!view_zeroed_host[j * M + i] && (i == 0 || view_zeroed_host[j * M + i - 1])) ++runs;
        if (runs > 1) return pl;
    }
    if (M > 16 || max_active_per_row > PX_MAXV || max_active_per_row < 1) return pl;
    const int c0 = M / 2;
    const long long shift = (long long)max_abs_a * c0;
    const int pad = (int)(shift > T ? T : shift);
    const size_t per_sm = 227 * 1024, per_cta_max = 227 * 1024 - 1024;
    int best_nxc = 0;
    for (int want_ctas = 2; want_ctas >= 1 && !best_nxc; --want_ctas) {
        for (int nxc = 1; nxc <= 32; ++nxc) {
            const int chunk = (T + nxc - 1) / nxc;
            if (chunk > 1024) continue;
            int threads = chunk <= 512 ? ((chunk + 31) & ~31) : ((((chunk + 1) / 2) + 31) & ~31);
            const int ppt = chunk <= 512 ? 1 : 2;
            const size_t smem = (size_t)max_active_per_row * (threads * ppt + 2 * pad) * 8;
            if (smem > per_cta_max) continue;
            if ((smem + 2048) * want_ctas > per_sm) continue;
            // chunking multiplies the staged (and loaded) pixels by (chunk + 2*pad)/chunk: stop splitting once the
            // rows fit; a finer split only pays through the second resident CTA
            best_nxc = nxc;
            break;
        }
    }
    if (px_force_nxc() > 0) best_nxc = px_force_nxc();
    if (!best_nxc) return pl;
    const int nxc = best_nxc;
    const int chunk = (T + nxc - 1) / nxc;
    if (chunk > 1024) return pl;
    pl.ppt = chunk <= 512 ? 1 : 2;
    pl.threads = chunk <= 512 ? ((chunk + 31) & ~31) : ((((chunk + 1) / 2) + 31) & ~31);
    pl.a.M = M; pl.a.S = S; pl.a.T = T; pl.a.N = N;
    pl.a.Tp = (T + 3) & ~3;
    pl.a.pad = pad; pl.a.chunk = chunk; pl.a.nxc = nxc;
    pl.a.rowPx = pl.threads * pl.ppt + 2 * pad;
    pl.a.nsl = N < RR_MAX_SLICES ? N : RR_MAX_SLICES;
    pl.smem = (size_t)max_active_per_row * pl.a.rowPx * 8;
    if (pl.smem > per_cta_max) return pl;
    pl.clamp = shift > T;
    pl.exact24 = (long long)n_active_total * 65535LL < (1LL << 24);
    if ((long long)N * S * pl.a.Tp >= (1LL << 31)) return pl;          // 32-bit accumulator offsets
    if ((long long)N * M * 2 * pl.a.Tp >= (1LL << 31)) return pl;      // ... and edge-table offsets
    pl.accRG_bytes = (size_t)N * S * pl.a.Tp * 8;
    pl.accB_bytes = (size_t)N * S * pl.a.Tp * 4;
    pl.edgeRG_bytes = (size_t)N * M * 2 * pl.a.Tp * 8;
    pl.edgeB_bytes = (size_t)N * M * 2 * pl.a.Tp * 4;
    pl.ok = true;
    return pl;
}

template <int PPT>
static int launch_px_variant(const PxArgs &a, const PxCtl &ctl, size_t smem, int threads, bool clamp,
                             cudaStream_t st) {
    dim3 grid(a.M, a.S, a.nxc), block(threads);
    if (clamp) {
        PCB_CUDA(cudaFuncSetAttribute(refocus_px_kernel<PPT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smem));
        refocus_px_kernel<PPT, true><<<grid, block, smem, st>>>(a, ctl);
    } else {
        PCB_CUDA(cudaFuncSetAttribute(refocus_px_kernel<PPT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smem));
        refocus_px_kernel<PPT, false><<<grid, block, smem, st>>>(a, ctl);
    }
    PCB_LAUNCH_OK();
    return PCB_OK;
}

static int launch_px_threads(const PxPlan &pl, const PxArgs &a, const PxCtl &ctl, cudaStream_t st) {
    return pl.ppt == 1 ? launch_px_variant<1>(a, ctl, pl.smem, pl.threads, pl.clamp, st)
                       : launch_px_variant<2>(a, ctl, pl.smem, pl.threads, pl.clamp, st);
}

static int launch_refocus_px(const PxPlan &pl, const uint16_t *vp, const int32_t *a_list_host,
                             const uint8_t *view_zeroed_host, void *scratch, float *out, pcb_minmax_t *mm,
                             cudaStream_t st) {
    PxArgs a = pl.a;
    char *sp = reinterpret_cast<char *>(scratch);
    a.vp = vp;
    a.accRG = reinterpret_cast<unsigned long long *>(sp);
    a.accB = reinterpret_cast<uint32_t *>(sp + pl.accRG_bytes);
    a.edgeRG = reinterpret_cast<unsigned long long *>(sp + pl.accRG_bytes + pl.accB_bytes);
    a.edgeB = reinterpret_cast<uint32_t *>(sp + pl.accRG_bytes + pl.accB_bytes + pl.edgeRG_bytes);
    PCB_REQUIRE(a.S <= 65535, "S <= 65535");
    PCB_CUDA(cudaMemsetAsync(sp, 0, pl.accRG_bytes + pl.accB_bytes, st));
    PxCtl ctl;
    memset(&ctl, 0, sizeof(ctl));
    for (int j = 0; j < a.M; ++j)
        for (int i = a.M - 1; i >= 0; --i)
            if (!view_zeroed_host[j * a.M + i]) {
                ctl.first[j] = (uint8_t)i;
                ctl.nact[j]++;
                ctl.row_active |= 1u << j;
            }
    const int N = a.N;
    for (int n0 = 0; n0 < N; n0 += RR_MAX_SLICES) {
        const int nloc = N - n0 < RR_MAX_SLICES ? N - n0 : RR_MAX_SLICES;
        for (int k = 0; k < RR_MAX_SLICES; ++k) ctl.a[k] = k < nloc ? a_list_host[n0 + k] : 0;
        a.n0 = n0;
        a.nsl = nloc;
        int rc = launch_px_threads(pl, a, ctl, st);
        if (rc) return rc;
        rc = launch_edge_cumsum(a.edgeRG, a.edgeB, ctl, a.M, a.Tp, n0, nloc, st);
        if (rc) return rc;
        const dim3 fgrid = finalize_grid(a.S * nloc);
        if (pl.exact24)
            refocus_px_finalize_kernel<true><<<fgrid, PXF_THREADS, 0, st>>>(a.accRG, a.accB, a.edgeRG, a.edgeB, ctl, out,
                                                                          a.M, a.S, a.T, a.Tp, n0, nloc, mm);
        else
            refocus_px_finalize_kernel<false><<<fgrid, PXF_THREADS, 0, st>>>(a.accRG, a.accB, a.edgeRG, a.edgeB, ctl,
                                                                           out, a.M, a.S, a.T, a.Tp, n0, nloc, mm);
        PCB_LAUNCH_OK();
    }
    return PCB_OK;
}

}  // namespace pcb
