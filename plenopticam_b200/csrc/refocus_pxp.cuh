// (c) Integer shift-and-sum, row-stationary, PERSISTENT form of refocus_px.cuh for uint16 RGB light fields.
//
// refocus_px_kernel is bound by shared-memory wavefronts (one LDS.64 per pixel and view), but its CTAs spend
// a third of their life staging rows from global memory, two co-resident CTAs tend to run in phase, and x
// chunking re-stages the shift margin of every chunk.  Here one CTA per SM walks the (r, j) tiles itself:
//
//   * the raw rows of tile n+1 are fetched by the TMA unit (cp.async.bulk, one bulk copy per view row,
//     completion on an mbarrier) into a raw buffer WHILE tile n is being summed out of the packed buffer, so no
//     thread ever waits on global memory;
//   * the whole image row is one chunk; each view keeps only the margin ITS shifts can reach
//     (|c0 - i| * max|a| pixels per side instead of (M/2) * max|a| for everyone), which is what lets the packed
//     rows (105 KB at Illum shape) and the raw rows of the next tile (57 KB) share one SM;
//   * raw -> packed (R | G << 21 | B << 42, edge-replicated) is a short shared-to-shared pass between tiles.
//
// Tiles are dealt round-robin in r-major order, so at any time the CTAs of the grid reduce into the same
// +-(M/2)|a| window of accumulator rows as before (L2-resident REDs).  Accumulator / edge-table layout and
// the finalize pass are those of refocus_px.cuh.
#pragma once
#include "common.cuh"
#include "refocus_px.cuh"

namespace pcb {

struct PxpCtl {                      // per-launch control block, passed by value (kernel-parameter bank)
    int32_t a[RR_MAX_SLICES];
    int32_t cb[16 * PX_MAXV];        // byte offset in the packed buffer of pixel 0 of view ii of view-row j
    uint8_t first[16];               // unmasked views of view-row j: columns first[j] .. first[j]+nact[j]-1
    uint8_t nact[16];
    uint8_t rows[16];                // the view-rows with at least one unmasked view
    int32_t nrows;
    int32_t maxa;                    // max |a| of this launch (per-view margins)
    int32_t gsz;                     // view rows per sweep (see pxp_tile)
};

struct PxpArgs {
    const uint16_t *vp;
    unsigned long long *accRG;
    uint32_t *accB;
    unsigned long long *edgeRG;
    uint32_t *edgeB;
    int M, S, T, N;
    int Tp;          // accumulator row pitch in pixels
    int W;           // threads * PPT: pixels of a packed row excluding margins
    int rawStride;   // bytes between raw rows in the raw buffer (multiple of 16)
    int rawOffset;   // byte offset of the raw buffer in dynamic shared memory (multiple of 16)
    int ntiles;      // nrows * S
    int nsl, n0;
    size_t vpBytes;  // size of the light field in bytes: no access outside [vp, vp + vpBytes)
    int Ma;          // 0: `vp` is the 4-D array vp[j,i,s,t,c];  > 0: `vp` is the ALIGNED micro-image array
                     //    (S*Ma, T*Ma, 3) of pitch Ma and view (j,i) is its sub-grid [ck+j :: Ma, ck+i :: Ma] - the
                     //    viewpoint gather (lfp_extractor/lfp_rearranger.py:101) and the central crop
                     //    (lfp_cropper.py:58-62) happen while the tile is staged
    int ck;          // crop margin (Ma - M) / 2
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ void red_add_u64(unsigned long long *p, unsigned long long v) {
    asm volatile("red.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void red_add_u32(uint32_t *p, uint32_t v) {
    asm volatile("red.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Tile order.  The view rows are walked in sweeps of `gsz` consecutive rows, each sweep r-major over the whole image:
// accumulator row y of slice a receives the tiles r = y + a (c0 - j), so it stays live while r crosses
// |a| * (span of c0 - j inside the sweep) image rows.  With all M view rows in one sweep the live accumulator rows of a
// 64-slice stack (109 MB) overflow L2 and cycle through DRAM; sweeps of 5 view rows keep them at about 31 MB.
__device__ __forceinline__ void pxp_tile(const PxpArgs &p, const PxpCtl &ctl, int t, int &j, int &r) {
    const int per_sweep = ctl.gsz * p.S;
    const int g = t / per_sweep;
    const int rem = t - g * per_sweep;
    const int gn = min(ctl.gsz, ctl.nrows - g * ctl.gsz);          // rows of this sweep (the last one may be short)
    r = rem / gn;
    j = ctl.rows[g * ctl.gsz + (rem - r * gn)];
}

// Elected thread: one bulk copy per active view row of tile (j, r).  cp.async.bulk needs 16-byte aligned addresses and
// sizes, so a row travels as the 16-byte aligned superset of its bytes (<= 15 bytes either side, which belong to the
// neighbouring rows of the light field) - CLAMPED to the 16-byte lines that lie entirely inside the caller's buffer
// [vp, vp + vpBytes): the few bytes of the first / last row that share a line with memory outside the buffer are
// fetched with plain 2-byte loads instead, so the kernel never touches a byte the caller does not own.
__device__ __forceinline__ void pxp_issue_tile(const PxpArgs &p, const PxpCtl &ctl, unsigned char *raw, uint64_t *bar,
                                               int j, int r) {
    // 4-D layout: one copy per active view row.  Aligned layout: ONE copy, the aligned row r*Ma + ck + j, which holds
    // row r of all M views of view-row j interleaved at pixel pitch Ma.
    const int nact = p.Ma ? 1 : (int)ctl.nact[j];
    const size_t rowB = p.Ma ? (size_t)p.T * p.Ma * 6 : (size_t)p.T * 6;
    const size_t g0 = p.Ma ? (size_t)p.vp + ((size_t)r * p.Ma + p.ck + j) * rowB
                           : (size_t)p.vp + (((size_t)j * p.M + ctl.first[j]) * p.S + r) * rowB;
    const size_t gstep = (size_t)p.S * rowB;
    const size_t lo16 = ((size_t)p.vp + 15) & ~(size_t)15;                   // first / one-past-last whole line inside
    const size_t hi16 = ((size_t)p.vp + p.vpBytes) & ~(size_t)15;
    uint32_t total = 0;
    for (int ii = 0; ii < nact; ++ii) {
        const size_t g = g0 + ii * gstep;
        const size_t b = max(g & ~(size_t)15, lo16), e = min((g + rowB + 15) & ~(size_t)15, hi16);
        if (e > b) total += (uint32_t)(e - b);
    }
    for (int ii = 0; ii < nact; ++ii) {                  // partial lines at the edges of the buffer (at most two rows)
        const size_t g = g0 + ii * gstep;
        const size_t ga = g & ~(size_t)15;
        const size_t b = max(ga, lo16), e = min((g + rowB + 15) & ~(size_t)15, hi16);
        if (b <= g && e >= g + rowB) continue;                               // the usual case: all of it by TMA
        unsigned char *slot = raw + (size_t)ii * p.rawStride;               // slot + (address - ga) holds that byte
        const size_t head_end = e > b ? min(b, g + rowB) : g + rowB;         // [g, head_end) and [tail_begin, g + rowB)
        const size_t tail_begin = e > b ? max(e, g) : g + rowB;
        for (size_t a = g; a < head_end; a += 2)
            *reinterpret_cast<uint16_t *>(slot + (a - ga)) = *reinterpret_cast<const uint16_t *>(a);
        for (size_t a = tail_begin; a < g + rowB; a += 2)
            *reinterpret_cast<uint16_t *>(slot + (a - ga)) = *reinterpret_cast<const uint16_t *>(a);
    }
    mbar_expect_tx(bar, total);
    for (int ii = 0; ii < nact; ++ii) {
        const size_t g = g0 + ii * gstep;
        const size_t b = max(g & ~(size_t)15, lo16), e = min((g + rowB + 15) & ~(size_t)15, hi16);
        if (e > b)
            tma_bulk_g2s(raw + (size_t)ii * p.rawStride + (b - (g & ~(size_t)15)), reinterpret_cast<const void *>(b),
                         (uint32_t)(e - b), bar);
    }
}

// raw rows (uint16 triplets at any 2-byte alignment) -> packed, edge-replicated 64-bit pixels
__device__ __forceinline__ void pxp_convert_tile(const PxpArgs &p, const PxpCtl &ctl, const unsigned char *raw,
                                                 unsigned char *packed, int j, int r, int tid, int THREADS) {
    const int nact = ctl.nact[j];
    const int c0 = p.M / 2;
    const size_t rowB = (size_t)p.T * 6;
    const size_t arow = p.Ma ? (size_t)p.vp + ((size_t)r * p.Ma + p.ck + j) * ((size_t)p.T * p.Ma * 6) : 0;
    const int pitch = p.Ma ? p.Ma * 6 : 6;                                   // bytes between pixels of one view
    for (int ii = 0; ii < nact; ++ii) {
        const int col = ctl.first[j] + ii;
        const size_t g = (size_t)p.vp + (((size_t)j * p.M + col) * p.S + r) * rowB;
        const unsigned char *src = p.Ma ? raw + (arow & 15) + (size_t)(p.ck + col) * 6
                                        : raw + (size_t)ii * p.rawStride + (g & 15);
        const int pad = min(abs(c0 - col) * ctl.maxa, p.T);
        unsigned long long *dst = reinterpret_cast<unsigned long long *>(packed + ctl.cb[j * PX_MAXV + ii]) - pad;
        const int npx = p.W + 2 * pad;
        for (int idx = tid; idx < npx; idx += THREADS) {
            const int x = min(max(idx - pad, 0), p.T - 1);
            const uint16_t *px = reinterpret_cast<const uint16_t *>(src + (size_t)x * pitch);
            dst[idx] = pack_px(px[0], px[1], px[2]);
        }
    }
}

// SYMC: the run of views is centred on c0 (true for every row of a circular aperture): the centre view is never
// shifted, so its pixel is read once per tile instead of once per slice.
template <int NACT, int PPT, bool CLAMP, int BATCH, bool SYMC>
__device__ __forceinline__ void pxp_slices(const PxpArgs &p, const PxpCtl &ctl, const unsigned char *packed,
                                           int THREADS, int j, int r, int tid, unsigned long long *rg_t,
                                           uint32_t *b_t) {
    constexpr int IC = SYMC ? (NACT - 1) / 2 : -1;
    const int c0 = p.M / 2;
    const bool is_top = (r == 0), is_bot = (r == p.S - 1);
    bool valid[PPT];
#pragma unroll
    for (int k = 0; k < PPT; ++k) valid[k] = tid + k * THREADS < p.T;
    const int dj = c0 - j;
    const int d0 = c0 - (int)ctl.first[j];
    int q[NACT];                           // byte offset of pixel `tid` of every staged view
#pragma unroll
    for (int ii = 0; ii < NACT; ++ii) q[ii] = ctl.cb[j * PX_MAXV + ii] + tid * 8;
    unsigned long long vc[PPT];
#pragma unroll
    for (int k = 0; k < PPT; ++k)
        vc[k] = SYMC ? *reinterpret_cast<const unsigned long long *>(packed + (q[SYMC ? IC : 0] + k * THREADS * 8)) : 0ull;

    for (int al = 0; al < p.nsl; ++al) {
        const int a = ctl.a[al];
        unsigned long long tot[PPT];
#pragma unroll
        for (int k = 0; k < PPT; ++k) tot[k] = vc[k];
#pragma unroll
        for (int b0 = 0; b0 < NACT; b0 += BATCH) {       // BATCH independent LDS.64 in flight, then their adds
            constexpr int NB = BATCH;
            unsigned long long v[NB][PPT];
#pragma unroll
            for (int u = 0; u < NB; ++u) {
                const int ii = b0 + u;
                if (ii < NACT && ii != IC) {
                    int s = a * (d0 - ii);
                    if (CLAMP) s = max(-p.T, min(p.T, s));
#pragma unroll
                    for (int k = 0; k < PPT; ++k)
                        v[u][k] = *reinterpret_cast<const unsigned long long *>(packed + (q[ii] + s * 8 + k * THREADS * 8));
                }
            }
#pragma unroll
            for (int k = 0; k < PPT; ++k) {
                unsigned long long s0 = 0ull, s1 = 0ull;
#pragma unroll
                for (int u = 0; u < NB; ++u) {
                    if (b0 + u < NACT && b0 + u != IC) {
                        if (u & 1) s1 += v[u][k]; else s0 += v[u][k];
                    }
                }
                tot[k] += s0 + s1;
            }
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
                red_add_u64(rg_t + ro + k * THREADS, rg);
                red_add_u32(b_t + ro + k * THREADS, B);
            }
            if (is_top || is_bot) {
                const size_t e = (size_t)(n * p.M + j) * 2 * p.Tp + tid + k * THREADS;
                if (is_top) { p.edgeRG[e] = rg; p.edgeB[e] = B; }
                if (is_bot) { p.edgeRG[e + p.Tp] = rg; p.edgeB[e + p.Tp] = B; }
            }
        }
    }
}

template <int MAXT, int PPT, bool CLAMP>
__global__ void __launch_bounds__(MAXT, 1)
refocus_pxp_kernel(PxpArgs p, PxpCtl ctl) {
    extern __shared__ __align__(128) unsigned char dsm[];
    __shared__ __align__(8) uint64_t s_bar;
    unsigned char *packed = dsm;
    unsigned char *raw = dsm + p.rawOffset;
    const int THREADS = blockDim.x;
    const int tid = threadIdx.x;

    if (tid == 0) {
        mbar_init(&s_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    int t = blockIdx.x;
    if (t >= p.ntiles) return;
    unsigned long long *rg_t = p.accRG + tid;      // kept in vector registers: re-deriving them from the parameter
    uint32_t *b_t = p.accB + tid;                  // bank inside the slice loop costs uniform-register spills
    asm volatile("" : "+l"(rg_t), "+l"(b_t));
    uint32_t parity = 0;
    int j, r;
    pxp_tile(p, ctl, t, j, r);
    if (tid == 0) pxp_issue_tile(p, ctl, raw, &s_bar, j, r);

    for (;;) {
        pxp_tile(p, ctl, t, j, r);
        // raw rows of tile t have landed -> packed buffer
        mbar_wait(&s_bar, parity);
        parity ^= 1u;
        pxp_convert_tile(p, ctl, raw, packed, j, r, tid, THREADS);
        __syncthreads();                       // packed buffer complete; raw buffer free again
        const int tn = t + gridDim.x;
        if (tid == 0 && tn < p.ntiles) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic reads of `raw` before async writes
            int jn, rn;
            pxp_tile(p, ctl, tn, jn, rn);
            pxp_issue_tile(p, ctl, raw, &s_bar, jn, rn);
        }
        constexpr int BATCH = MAXT <= 640 ? 16 : 8 / PPT;
        const bool symc = 2 * (p.M / 2 - (int)ctl.first[j]) == (int)ctl.nact[j] - 1;
#define PCB_PXP_NACT(K)                                                                              \
    case K:                                                                                          \
        if ((K & 1) && symc) pxp_slices<K, PPT, CLAMP, BATCH, (K & 1) != 0>(p, ctl, packed, THREADS, j, r, tid, rg_t, b_t); \
        else pxp_slices<K, PPT, CLAMP, BATCH, false>(p, ctl, packed, THREADS, j, r, tid, rg_t, b_t);   \
        break;
        switch (ctl.nact[j]) {
            PCB_PXP_NACT(1) PCB_PXP_NACT(2) PCB_PXP_NACT(3) PCB_PXP_NACT(4) PCB_PXP_NACT(5) PCB_PXP_NACT(6)
            PCB_PXP_NACT(7) PCB_PXP_NACT(8) PCB_PXP_NACT(9) PCB_PXP_NACT(10) PCB_PXP_NACT(11) PCB_PXP_NACT(12)
            PCB_PXP_NACT(13) PCB_PXP_NACT(14) PCB_PXP_NACT(15) PCB_PXP_NACT(16)
        }
#undef PCB_PXP_NACT
        if (tn >= p.ntiles) break;
        t = tn;
        __syncthreads();                       // every warp is done reading the packed buffer
    }
}

struct PxpPlan {
    bool ok;
    PxpArgs a;
    PxpCtl ctl;
    size_t smem;
    int threads, ppt, maxt;
    bool clamp, exact24;
};

static int pxp_sm_count() { return sm_count(); }

// Host-side planning (no device access).  `px` supplies the accumulator / edge-table geometry.
static PxpPlan plan_refocus_pxp(const PxPlan &px, int M, int S, int T, int N, int max_abs_a,
                                const uint8_t *view_zeroed_host, const void *vp, int Ma = 0) {
    PxpPlan pl;
    memset(&pl, 0, sizeof(pl));
    pl.ok = false;
    if (!px.ok || T > 2048 || ((size_t)vp & 1)) return pl;
    const int c0 = M / 2;
    pl.ppt = T <= 1024 ? 1 : 2;
    pl.threads = ((((T + pl.ppt - 1) / pl.ppt) + 31) & ~31);
    pl.maxt = (pl.threads <= 640 && pl.ppt == 1) ? 640 : 1024;      // the 640-thread build only exists for PPT = 1
    const int W = pl.threads * pl.ppt;
    PxpCtl &c = pl.ctl;
    int max_nact = 0;
    size_t max_packed = 0;
    for (int j = 0; j < M; ++j) {
        for (int i = M - 1; i >= 0; --i)
            if (!view_zeroed_host[j * M + i]) { c.first[j] = (uint8_t)i; c.nact[j]++; }
        if (!c.nact[j]) continue;
        c.rows[c.nrows++] = (uint8_t)j;
        if (c.nact[j] > max_nact) max_nact = c.nact[j];
        size_t off = 0;                                   // pixels
        for (int ii = 0; ii < c.nact[j]; ++ii) {
            const int d = c0 - (c.first[j] + ii);
            long long padl = (long long)(d < 0 ? -d : d) * max_abs_a;
            const int pad = (int)(padl > T ? T : padl);
            c.cb[j * PX_MAXV + ii] = (int32_t)((off + pad) * 8);
            off += (size_t)W + 2 * pad;
        }
        if (off * 8 > max_packed) max_packed = off * 8;
    }
    if (c.nrows == 0) return pl;
    c.maxa = max_abs_a;
    {   // sweeps of view rows: only worth it when the live accumulator rows would not fit L2 in one sweep
        const char *e = getenv("PCB_PXP_SWEEP_ROWS");
        int gsz = e ? atoi(e) : 0;
        if (gsz <= 0) {
            // live rows ~ sum over slices of (M - 1) |a|; the plan sees max |a| only: assume a symmetric range
            const double live_mb = (double)(M - 1) * max_abs_a * 0.5 * (N < RR_MAX_SLICES ? N : RR_MAX_SLICES) * px.a.Tp * 12.0 / 1e6;
            gsz = live_mb > 48.0 ? 5 : c.nrows;
        }
        if (gsz > c.nrows) gsz = c.nrows;
        c.gsz = gsz;
    }
    const int rawStride = Ma ? ((T * Ma * 6 + 30 + 15) & ~15) : ((T * 6 + 30 + 15) & ~15);
    const size_t rawOffset = (max_packed + 127) & ~(size_t)127;
    pl.smem = rawOffset + (size_t)(Ma ? 1 : max_nact) * rawStride;
    if (Ma && (size_t)T * Ma * 6 + 32 >= (1u << 20)) return pl;             // one bulk copy per row: mbarrier tx-count
    if (pl.smem > 227 * 1024 - 64) return pl;
    pl.a.M = M; pl.a.S = S; pl.a.T = T; pl.a.N = N;
    pl.a.Tp = px.a.Tp;
    pl.a.W = W;
    pl.a.rawStride = rawStride;
    pl.a.rawOffset = (int)rawOffset;
    pl.a.ntiles = c.nrows * S;
    pl.a.vpBytes = Ma ? (size_t)S * Ma * T * Ma * 6 : (size_t)M * M * S * T * 6;
    pl.a.Ma = Ma;
    pl.a.ck = Ma ? (Ma - M) / 2 : 0;
    pl.clamp = px.clamp;
    pl.exact24 = px.exact24;
    pl.ok = true;
    return pl;
}

template <int MAXT, int PPT>
static int launch_pxp_variant(const PxpPlan &pl, const PxpArgs &a, const PxpCtl &ctl, cudaStream_t st) {
    const int grid = a.ntiles < pxp_sm_count() ? a.ntiles : pxp_sm_count();
    if (pl.clamp) {
        PCB_CUDA(cudaFuncSetAttribute(refocus_pxp_kernel<MAXT, PPT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)pl.smem));
        refocus_pxp_kernel<MAXT, PPT, true><<<grid, pl.threads, pl.smem, st>>>(a, ctl);
    } else {
        PCB_CUDA(cudaFuncSetAttribute(refocus_pxp_kernel<MAXT, PPT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)pl.smem));
        refocus_pxp_kernel<MAXT, PPT, false><<<grid, pl.threads, pl.smem, st>>>(a, ctl);
    }
    PCB_LAUNCH_OK();
    return PCB_OK;
}

// Colour management fused into the finalize pass (pcb_refocus_int_srgb).
struct PxpFusedPost {
    double q_lo, q_hi;       // percentiles of slice 0: luma / all channels (lfp_contrast.py:252-255)
    float *ref_slice;        // (S, T, 3) float32 scratch: the linear slice 0
    double *lohi;            // [2] device
    void *pscratch;          // pcb_percentile_scratch_bytes()
};

template <int POST>
static int launch_pxp_finalize(const PxpPlan &pl, const PxpArgs &a, const PxCtl &fctl, float *out, int n0, int nloc,
                               pcb_minmax_t *mm, const double *lohi, int leave_zero, cudaStream_t st) {
    dim3 fgrid = finalize_grid(a.S * nloc);
    if (POST == 2) fgrid = dim3(min(a.S * nloc, 8 * sm_count()), 1);     // usually returns at once: keep the launch small
    if (pl.exact24)
        refocus_px_finalize_kernel<true, POST><<<fgrid, PXF_THREADS, 0, st>>>(a.accRG, a.accB, a.edgeRG, a.edgeB, fctl, out,
                                                                            a.M, a.S, a.T, a.Tp, n0, nloc, mm, lohi,
                                                                            leave_zero);
    else
        refocus_px_finalize_kernel<false, POST><<<fgrid, PXF_THREADS, 0, st>>>(a.accRG, a.accB, a.edgeRG, a.edgeB, fctl,
                                                                             out, a.M, a.S, a.T, a.Tp, n0, nloc, mm, lohi,
                                                                             leave_zero);
    PCB_LAUNCH_OK();
    return PCB_OK;
}

static int launch_refocus_pxp(const PxPlan &px, const PxpPlan &pl, const uint16_t *vp, const int32_t *a_list_host,
                              const uint8_t *view_zeroed_host, void *scratch, float *out, pcb_minmax_t *mm,
                              cudaStream_t st, const PxpFusedPost *post = nullptr, int scratch_flags = 0) {
    PxpArgs a = pl.a;
    char *sp = reinterpret_cast<char *>(scratch);
    a.vp = vp;
    a.accRG = reinterpret_cast<unsigned long long *>(sp);
    a.accB = reinterpret_cast<uint32_t *>(sp + px.accRG_bytes);
    a.edgeRG = reinterpret_cast<unsigned long long *>(sp + px.accRG_bytes + px.accB_bytes);
    a.edgeB = reinterpret_cast<uint32_t *>(sp + px.accRG_bytes + px.accB_bytes + px.edgeRG_bytes);
    PCB_REQUIRE(a.S <= 65535, "S <= 65535");
    // PCB_REFOCUS_SCRATCH_ZEROED: the caller vouches that the accumulators hold zeros (a previous call with
    // PCB_REFOCUS_LEAVE_ZEROED on the same scratch and stream order); PCB_REFOCUS_LEAVE_ZEROED: the pass that reads an
    // accumulator word last stores zero behind it (refocus_px_finalize_kernel)
    if (!(scratch_flags & PCB_REFOCUS_SCRATCH_ZEROED))
        PCB_CUDA(cudaMemsetAsync(sp, 0, px.accRG_bytes + px.accB_bytes, st));
    const int leave = (scratch_flags & PCB_REFOCUS_LEAVE_ZEROED) ? 1 : 0;
    PxpCtl ctl = pl.ctl;
    PxCtl fctl;                                            // finalize shares refocus_px.cuh's pass
    memset(&fctl, 0, sizeof(fctl));
    for (int j = 0; j < a.M; ++j) {
        fctl.first[j] = ctl.first[j];
        fctl.nact[j] = ctl.nact[j];
        if (ctl.nact[j]) fctl.row_active |= 1u << j;
    }
    const int N = a.N;
    // pass: 0 = accumulate + (linear finalize | nothing), 1 = fused finalize, 2 = its fix-up
    for (int pass = 0; pass < (post ? 3 : 1); ++pass) {
        if (pass == 1) {                                   // bounds from the linear slice 0
            for (int k = 0; k < RR_MAX_SLICES; ++k) fctl.a[k] = k == 0 ? a_list_host[0] : 0;
            int rc = launch_pxp_finalize<0>(pl, a, fctl, post->ref_slice, 0, 1, nullptr, nullptr, 0, st);
            if (rc) return rc;
            rc = run_contrast_bounds(post->ref_slice, (int64_t)a.S * a.T, post->q_lo, post->q_hi, post->lohi,
                                     post->pscratch, st);
            if (rc) return rc;
        }
        for (int n0 = 0; n0 < N; n0 += RR_MAX_SLICES) {
            const int nloc = N - n0 < RR_MAX_SLICES ? N - n0 : RR_MAX_SLICES;
            for (int k = 0; k < RR_MAX_SLICES; ++k) fctl.a[k] = ctl.a[k] = k < nloc ? a_list_host[n0 + k] : 0;
            a.n0 = n0;
            a.nsl = nloc;
            int rc = PCB_OK;
            if (pass == 0) {
                if (g_kev[0]) cudaEventRecord(g_kev[0], st);
                if (pl.maxt == 640)
                    rc = launch_pxp_variant<640, 1>(pl, a, ctl, st);
                else
                    rc = pl.ppt == 1 ? launch_pxp_variant<1024, 1>(pl, a, ctl, st) : launch_pxp_variant<1024, 2>(pl, a, ctl, st);
                if (g_kev[1]) cudaEventRecord(g_kev[1], st);
                if (rc == PCB_OK) rc = launch_edge_cumsum(a.edgeRG, a.edgeB, fctl, a.M, a.Tp, n0, nloc, st);
                if (rc == PCB_OK && !post) rc = launch_pxp_finalize<0>(pl, a, fctl, out, n0, nloc, mm, nullptr, leave, st);
            } else if (pass == 1) {
                rc = launch_pxp_finalize<1>(pl, a, fctl, out, n0, nloc, mm, post->lohi, leave, st);
            } else {
                rc = launch_pxp_finalize<2>(pl, a, fctl, out, n0, nloc, mm, post->lohi, leave, st);
            }
            if (rc) return rc;
        }
    }
    return PCB_OK;
}

}  // namespace pcb
