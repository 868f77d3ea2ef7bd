// (c) Integer shift-and-sum, row-stationary formulation for uint16 light fields.
//
// The direct form (refocus.cuh) re-reads every view row once per slice from L2: N * B_vp of L2->SM traffic.
// The only reuse the algorithm offers is that ONE row r of view (j,i) feeds EVERY slice a (at output row
// y = r - a(c0-j), shifted by a(c0-i) pixels).  So here one CTA owns (view-row j, image row r): it stages
// the <= M active rows vp[j, :, r] once in shared memory (edge-replicated on both sides so that the x clamp
// disappears, and in two copies offset by one element so that every shifted read is a 32-bit aligned LDS of
// two uint16), computes  G_a[x] = sum_i vp[j,i,r, x + a(c0-i)]  for all slices from shared memory, and folds
// G_a into the stack accumulator row (a, r - a(c0-j)) with integer REDs (exact, order independent).
// Rows whose source is clamped in y (y + a(c0-j) outside [0,S)) are served from a small edge table written
// by the r = 0 and r = S-1 CTAs and added by the finalize pass, which also applies the 1/M scaling, converts
// to float32 and reduces the stack min/max.
//
// Packed accumulation of a word w = lo | hi << 16:   sum += w  (ALU pipe);  hs = mad.hi(w, 65536, hs)  (FMA
// pipe); lo_total = sum - (hs << 16)  (mod 2^32, exact because lo_total < 2^32).  One op per element.
// Both 32-bit totals of a word go out as ONE 64-bit RED: the low lane never carries (total < 2^32).
//
// CTA order is r-major (blockIdx.x = j): at any time the accumulator rows being reduced into lie within
// +-(M/2)|a| rows of the current r, a window that fits the 126 MB L2, so the REDs resolve in L2 and DRAM sees
// the light field once plus the accumulator once.
#pragma once
#include "common.cuh"

namespace pcb {

constexpr int RR_MAX_THREADS = 512;
constexpr int RR_TBL_STRIDE = 16;    // table entries per slice (>= max active views per view-row, multiple of 4)
constexpr int RR_MAX_SLICES = 64;    // slices per launch (refocus parameters travel by value)
constexpr int RR_STAGE_UNROLL = 8;   // independent global loads in flight per thread while staging

struct RowsCtl {                     // per-launch control block, passed by value (kernel-parameter bank)
    int32_t a[RR_MAX_SLICES];
    uint32_t zeroed_bits[8];         // bit (j*M+i) set -> view masked by the circular aperture (M <= 16)
    uint32_t row_active;             // bit j set -> view-row j has at least one unmasked view
};

// The light field is read exactly once per launch: keep it from displacing the accumulator rows in L2.
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint16_t ld_stream_u16(const uint16_t *p, uint64_t pol) {
    uint16_t v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u16 %0, [%1], %2;" : "=h"(v) : "l"(p), "l"(pol));
    return v;
}

__device__ __forceinline__ bool ctl_zeroed(const RowsCtl &c, int idx) {
    return (c.zeroed_bits[idx >> 5] >> (idx & 31)) & 1u;
}

struct RowsArgs {
    const uint16_t *vp;
    uint32_t *acc;      // [N][S][pitch]
    uint32_t *edge;     // [N][M][2][pitch]
    int M, S, T, C, N;
    int E;              // T*C elements per row
    int pitch;          // accumulator row pitch (elements, multiple of 4)
    int padE;           // replicated elements on each side (C * max shift in px, capped at E), even
    int chunkE;         // output elements per CTA (even)
    int nxc;            // x chunks
    int rowWords;       // 32-bit words per staged row copy (>= padE + WPT*RR_THREADS: unguarded strided reads)
    int nsl;            // slices per launch (<= RR_MAX_SLICES), all handled by every CTA
    int n0;             // first slice of this launch
};

__device__ __forceinline__ int clamp_elem(int e, int E, int C) {
    if (e < 0) {
        const int ch = e % C;
        return ch < 0 ? ch + C : ch;
    }
    if (e >= E) return E - C + (e % C);
    return e;
}

template <int RR_THREADS, int WPT>
__global__ void __launch_bounds__(RR_THREADS, 1)
refocus_rows_kernel(RowsArgs p, RowsCtl ctl) {
    extern __shared__ __align__(16) uint32_t smem[];
    __shared__ int s_act[RR_TBL_STRIDE];
    __shared__ int s_nact;
    __shared__ long long s_rowoff[RR_MAX_SLICES];   // accumulator row (element offset) each slice reduces into, or -1

    const int j = blockIdx.x;           // fastest: the M view-rows of one image row r run together
    const int r = blockIdx.y;
    const int xc = blockIdx.z;
    const int nloc = p.nsl;
    const int c0 = p.M / 2;
    const int e0 = xc * p.chunkE;
    const int tid = threadIdx.x;

    if (tid == 0) {
        int n = 0;
        for (int i = 0; i < p.M; ++i)
            if (!ctl_zeroed(ctl, j * p.M + i)) s_act[n++] = i;
        s_nact = n;
    }
    __syncthreads();
    const int nact = s_nact;
    if (nact == 0) return;
    const int ngrp = (nact + 3) >> 2;

    // smem layout: [nact][2][rowWords] staged rows | zero row [rowWords] | table [nsl][RR_TBL_STRIDE]
    uint32_t *zero_row = smem + (size_t)nact * 2 * p.rowWords;
    int *tbl = reinterpret_cast<int *>(zero_row + p.rowWords);

    // ---- stage rows: copy A at element idx, copy B shifted by one element.  All (row, idx) pairs are
    // independent: walk them as one flat range with RR_STAGE_UNROLL loads in flight per thread.
    const int rowElems = p.rowWords * 2;
    const int needElems = p.chunkE + 2 * p.padE + 2;
    const size_t view_stride = (size_t)p.S * p.E;
    const uint16_t *row0 = p.vp + ((size_t)j * p.M * p.S + r) * p.E;
    uint16_t *s16 = reinterpret_cast<uint16_t *>(smem);
    const uint64_t pol = l2_evict_first_policy();
    for (int ii = 0; ii < nact; ii += 2) {          // two rows per pass: 2*RR_STAGE_UNROLL loads in flight
        const bool two = ii + 1 < nact;
        const uint16_t *srcA = row0 + (size_t)s_act[ii] * view_stride;
        const uint16_t *srcB = row0 + (size_t)s_act[two ? ii + 1 : ii] * view_stride;
        uint16_t *dA = s16 + (size_t)ii * 2 * rowElems;
        uint16_t *dB = dA + 2 * rowElems;
        for (int base = tid; base < needElems; base += RR_THREADS * RR_STAGE_UNROLL) {
            uint16_t va[RR_STAGE_UNROLL], vb[RR_STAGE_UNROLL];
#pragma unroll
            for (int u = 0; u < RR_STAGE_UNROLL; ++u) {
                const int idx = base + u * RR_THREADS;
                if (idx < needElems) {
                    const int e = clamp_elem(e0 - p.padE + idx, p.E, p.C);
                    va[u] = ld_stream_u16(srcA + e, pol);
                    vb[u] = ld_stream_u16(srcB + e, pol);
                }
            }
#pragma unroll
            for (int u = 0; u < RR_STAGE_UNROLL; ++u) {
                const int idx = base + u * RR_THREADS;
                if (idx < needElems) {
                    dA[idx] = va[u];                                  // copy A
                    if (idx > 0) dA[rowElems + idx - 1] = va[u];      // copy B: shifted by one element
                    if (two) {
                        dB[idx] = vb[u];
                        if (idx > 0) dB[rowElems + idx - 1] = vb[u];
                    }
                }
            }
        }
    }
    for (int idx = tid; idx < p.rowWords; idx += RR_THREADS) zero_row[idx] = 0u;
    // ---- offset table: BYTE offset of the staged word feeding output word 0, per (slice, active view)
    for (int t = tid; t < nloc * RR_TBL_STRIDE; t += RR_THREADS) {
        const int al = t / RR_TBL_STRIDE, ii = t - al * RR_TBL_STRIDE;
        int woff = (int)(zero_row - smem);
        if (ii < nact) {
            int s = ctl.a[al] * (c0 - s_act[ii]);
            s = max(-p.T, min(p.T, s));
            const int idx0 = p.C * s + p.padE;          // >= 0 by construction of padE
            const int copy = idx0 & 1;
            woff = (ii * 2 + copy) * p.rowWords + ((idx0 - copy) >> 1);
        }
        tbl[t] = woff * 4;
    }
    for (int al = tid; al < nloc; al += RR_THREADS) {
        const int y = r - ctl.a[al] * (c0 - j);
        s_rowoff[al] = (y >= 0 && y < p.S) ? ((long long)(p.n0 + al) * p.S + y) * p.pitch + e0 : -1LL;
    }
    __syncthreads();

    const int chunkWords = min(p.chunkE, p.E - e0 + 1) >> 1;     // words that hold at least one valid element
    uint32_t *acc_thread = p.acc + 2 * tid;
    const char *sbase = reinterpret_cast<const char *>(smem) + tid * 4;
    bool valid[WPT], full[WPT];
#pragma unroll
    for (int k = 0; k < WPT; ++k) {
        const int w = tid + k * RR_THREADS;
        valid[k] = w < chunkWords;
        full[k] = valid[k] && (e0 + 2 * w + 1 < p.E);
    }
    const bool is_top = (r == 0), is_bot = (r == p.S - 1);

    for (int al = 0; al < nloc; ++al) {
        uint32_t sum[WPT], hs[WPT], sum2[WPT], hs2[WPT];     // two chains per word: more ILP on the ALU pipe
#pragma unroll
        for (int k = 0; k < WPT; ++k) { sum[k] = 0u; hs[k] = 0u; sum2[k] = 0u; hs2[k] = 0u; }
        const int4 *trow = reinterpret_cast<const int4 *>(tbl + al * RR_TBL_STRIDE);
        for (int g = 0; g < ngrp; ++g) {
            const int4 o = trow[g];
            uint32_t v[4][WPT];
#pragma unroll
            for (int k = 0; k < WPT; ++k) {      // word k of a thread sits k*RR_THREADS*4 bytes further: immediate
                v[0][k] = *reinterpret_cast<const uint32_t *>(sbase + o.x + k * RR_THREADS * 4);
                v[1][k] = *reinterpret_cast<const uint32_t *>(sbase + o.y + k * RR_THREADS * 4);
                v[2][k] = *reinterpret_cast<const uint32_t *>(sbase + o.z + k * RR_THREADS * 4);
                v[3][k] = *reinterpret_cast<const uint32_t *>(sbase + o.w + k * RR_THREADS * 4);
            }
#pragma unroll
            for (int k = 0; k < WPT; ++k) {
                sum[k] += v[0][k] + v[1][k];
                sum2[k] += v[2][k] + v[3][k];
                hs[k] = __umulhi(v[0][k], 65536u) + hs[k];
                hs2[k] = __umulhi(v[1][k], 65536u) + hs2[k];
                hs[k] = __umulhi(v[2][k], 65536u) + hs[k];
                hs2[k] = __umulhi(v[3][k], 65536u) + hs2[k];
            }
        }
#pragma unroll
        for (int k = 0; k < WPT; ++k) { sum[k] += sum2[k]; hs[k] += hs2[k]; }
        const int n = p.n0 + al;
        const long long ro = s_rowoff[al];
        if (ro >= 0) {
            uint32_t *arow = acc_thread + ro;
#pragma unroll
            for (int k = 0; k < WPT; ++k) {
                const uint32_t hi = hs[k];
                const uint32_t lo = sum[k] - (hi << 16);
                uint32_t *dst = arow + k * RR_THREADS * 2;
                if (full[k])
                    atomicAdd(reinterpret_cast<unsigned long long *>(dst), ((unsigned long long)hi << 32) | lo);
                else if (valid[k])
                    atomicAdd(dst, lo);
            }
        }
        if (is_top || is_bot) {
#pragma unroll
            for (int k = 0; k < WPT; ++k) {
                if (!valid[k]) continue;
                const uint32_t hi = hs[k];
                const uint32_t lo = sum[k] - (hi << 16);
                const int e = e0 + 2 * (tid + k * RR_THREADS);
                if (is_top) {
                    uint32_t *d = p.edge + ((size_t)(n * p.M + j) * 2 + 0) * p.pitch + e;
                    d[0] = lo;
                    if (full[k]) d[1] = hi;
                }
                if (is_bot) {
                    uint32_t *d = p.edge + ((size_t)(n * p.M + j) * 2 + 1) * p.pitch + e;
                    d[0] = lo;
                    if (full[k]) d[1] = hi;
                }
            }
        }
    }
}

// out[n,y,e] = (acc[n,y,e] + sum_{j: source row clamped} edge[n,j,side,e]) / M ;  fused min/max.
// EXACT24: every total is < 2^24, so float(v) is exact and one IEEE float32 division gives the correctly
// rounded quotient; otherwise the division is done in float64.
template <bool EXACT24>
__global__ void __launch_bounds__(256)
refocus_rows_finalize_kernel(const uint32_t *__restrict__ acc, const uint32_t *__restrict__ edge, RowsCtl ctl,
                             float *__restrict__ out, int M, int S, int E, int pitch, int n0, pcb_minmax_t *mm) {
    __shared__ int s_rows[16];     // edge-table row index (n*M + j)*2 + side of every clamped contribution
    __shared__ int s_cnt;
    const int y = blockIdx.x, n = n0 + blockIdx.y;
    if (threadIdx.x < 32) {                    // M <= 16: one lane per view-row, compacted with a ballot
        const int j = threadIdx.x;
        const int ys = y + ctl.a[blockIdx.y] * (M / 2 - j);
        const bool clamped = j < M && ((ctl.row_active >> j) & 1u) && (ys < 0 || ys > S - 1);
        const unsigned m = __ballot_sync(0xffffffffu, clamped);
        if (clamped) s_rows[__popc(m & ((1u << j) - 1u))] = (n * M + j) * 2 + (ys < 0 ? 0 : 1);
        if (j == 0) s_cnt = __popc(m);
    }
    __syncthreads();
    const int cnt = s_cnt;
    const uint32_t *arow = acc + ((size_t)n * S + y) * pitch;
    float *orow = out + ((size_t)n * S + y) * E;
    const float fM = (float)M;
    float mn = INFINITY, mx = -INFINITY;
    for (int e = threadIdx.x; e < E; e += blockDim.x) {                // consecutive lanes, consecutive elements
        uint32_t v = arow[e];
        for (int c = 0; c < cnt; ++c) v += __ldg(edge + (size_t)s_rows[c] * pitch + e);
        const float f = EXACT24 ? __fdiv_rn((float)v, fM) : (float)((double)v / (double)M);
        orow[e] = f;
        mn = fminf(mn, f);
        mx = fmaxf(mx, f);
    }
    if (mm != nullptr) block_minmax_commit(mn, mx, mm);
}

struct RowsPlan {
    bool ok;
    RowsArgs a;
    size_t smem;
    int wpt, threads;
    bool exact24;
    size_t acc_bytes, edge_bytes;
};

static int rr_force_threads() {      // tuning knob for A/B measurements: PCB_RR_THREADS=256
    const char *e = getenv("PCB_RR_THREADS");
    return e ? atoi(e) : 0;
}

// Host-side planning (no device access): picks the x chunking so that the staged rows fit in shared memory.
static RowsPlan plan_refocus_rows(int M, int S, int T, int C, int N, int max_abs_a, int max_active_per_row,
                                  int n_active_total) {
    RowsPlan pl;
    memset(&pl, 0, sizeof(pl));
    pl.ok = false;
    if (M > 16 || max_active_per_row > RR_TBL_STRIDE || max_active_per_row < 1) return pl;
    const int E = T * C;
    const int c0 = M / 2;
    long long shift = (long long)max_abs_a * c0;
    if (shift > T) shift = T;
    int padE = (int)shift * C;
    padE += padE & 1;
    const int nsl = N < RR_MAX_SLICES ? N : RR_MAX_SLICES;
    const size_t budget = 224 * 1024;
    const size_t tbl_bytes = (size_t)nsl * RR_TBL_STRIDE * 4;
    for (int nxc = 1; nxc <= 64; ++nxc) {
        int chunkE = (E + nxc - 1) / nxc;
        chunkE += chunkE & 1;
        const int words = chunkE / 2;
        // thread shape: 512 threads x (1|2) words, or 256 threads x 4 words for long rows
        int threads = 512;
        int wpt = (words + threads - 1) / threads;
        if (wpt < 1) wpt = 1;
        if (wpt > 2 || rr_force_threads() == 256) {
            threads = 256;
            wpt = (words + threads - 1) / threads;
            if (wpt < 1) wpt = 1;
            if (wpt == 3) wpt = 4;
        }
        if (wpt > 4) continue;
        int rowWords = (chunkE + 2 * padE + 2 + 1) / 2;
        const int reach = padE + wpt * threads + 1;             // furthest word an unguarded strided read touches
        if (rowWords < reach) rowWords = reach;
        rowWords = (rowWords + 3) & ~3;                         // 16-byte rows
        const size_t smem = ((size_t)2 * max_active_per_row + 1) * rowWords * 4 + tbl_bytes;
        if (smem <= budget) {
            pl.ok = true;
            pl.smem = smem;
            pl.wpt = wpt;
            pl.threads = threads;
            pl.exact24 = (long long)n_active_total * 65535LL < (1LL << 24);
            pl.a.M = M; pl.a.S = S; pl.a.T = T; pl.a.C = C; pl.a.N = N; pl.a.E = E;
            pl.a.pitch = (E + 3) & ~3;
            pl.a.padE = padE; pl.a.chunkE = chunkE; pl.a.nxc = nxc; pl.a.rowWords = rowWords; pl.a.nsl = nsl;
            pl.acc_bytes = (size_t)N * S * pl.a.pitch * 4;
            pl.edge_bytes = (size_t)N * M * 2 * pl.a.pitch * 4;
            return pl;
        }
    }
    return pl;
}

template <int THREADS, int WPT>
static int launch_rows_wpt(const RowsArgs &a, const RowsCtl &ctl, size_t smem, cudaStream_t st) {
    PCB_CUDA(cudaFuncSetAttribute(refocus_rows_kernel<THREADS, WPT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem));
    dim3 grid(a.M, a.S, a.nxc), block(THREADS);
    refocus_rows_kernel<THREADS, WPT><<<grid, block, smem, st>>>(a, ctl);
    PCB_LAUNCH_OK();
    return PCB_OK;
}

static int launch_refocus_rows(const RowsPlan &pl, const uint16_t *vp, const int32_t *a_list_host,
                               const uint8_t *view_zeroed_host, void *scratch, float *out, pcb_minmax_t *mm,
                               cudaStream_t st) {
    RowsArgs a = pl.a;
    a.vp = vp;
    a.acc = reinterpret_cast<uint32_t *>(scratch);
    a.edge = reinterpret_cast<uint32_t *>(reinterpret_cast<char *>(scratch) + pl.acc_bytes);
    PCB_REQUIRE(a.S <= 65535, "S <= 65535");
    PCB_CUDA(cudaMemsetAsync(a.acc, 0, pl.acc_bytes, st));
    RowsCtl ctl;
    memset(&ctl, 0, sizeof(ctl));
    for (int k = 0; k < a.M * a.M; ++k) {
        if (view_zeroed_host[k]) ctl.zeroed_bits[k >> 5] |= 1u << (k & 31);
        else ctl.row_active |= 1u << (k / a.M);
    }
    const int N = a.N;
    for (int n0 = 0; n0 < N; n0 += RR_MAX_SLICES) {
        const int nloc = N - n0 < RR_MAX_SLICES ? N - n0 : RR_MAX_SLICES;
        for (int k = 0; k < RR_MAX_SLICES; ++k) ctl.a[k] = k < nloc ? a_list_host[n0 + k] : 0;
        a.n0 = n0;
        a.nsl = nloc;
        int rc;
        if (pl.threads == 512)
            rc = pl.wpt == 1 ? launch_rows_wpt<512, 1>(a, ctl, pl.smem, st) : launch_rows_wpt<512, 2>(a, ctl, pl.smem, st);
        else
            rc = pl.wpt == 1   ? launch_rows_wpt<256, 1>(a, ctl, pl.smem, st)
                 : pl.wpt == 2 ? launch_rows_wpt<256, 2>(a, ctl, pl.smem, st)
                               : launch_rows_wpt<256, 4>(a, ctl, pl.smem, st);
        if (rc) return rc;
        dim3 fgrid(a.S, nloc);
        if (pl.exact24)
            refocus_rows_finalize_kernel<true><<<fgrid, 256, 0, st>>>(a.acc, a.edge, ctl, out, a.M, a.S, a.E, a.pitch, n0, mm);
        else
            refocus_rows_finalize_kernel<false><<<fgrid, 256, 0, st>>>(a.acc, a.edge, ctl, out, a.M, a.S, a.E, a.pitch, n0, mm);
        PCB_LAUNCH_OK();
    }
    return PCB_OK;
}

}  // namespace pcb
