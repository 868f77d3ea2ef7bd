// (c) Integer shift-and-sum, row-stationary formulation for uint16 light fields.
//
// The direct form (refocus.cuh) re-reads every view row once per slice from L2: N * B_vp of L2->SM traffic.
// The only reuse the algorithm offers is that ONE row r of view (j,i) feeds EVERY slice a (at output row
// y = r - a(c0-j), shifted by a(c0-i) pixels).  So here one CTA owns (view-row j, image row r): it stages
// the <= M active rows vp[j, :, r] once in shared memory (edge-replicated on both sides so that the x clamp
// disappears, and in two copies offset by one element so that every shifted read is a 32-bit aligned LDS of
// two uint16), computes  G_a[x] = sum_i vp[j,i,r, x + a(c0-i)]  for all slices from shared memory, and folds
// G_a into the stack accumulator row (a, r - a(c0-j)) with 32-bit integer REDs (exact, order independent).
// Rows whose source is clamped in y (y + a(c0-j) outside [0,S)) are served from a small edge table written
// by the r = 0 and r = S-1 CTAs and added by the finalize pass, which also applies the 1/M scaling, converts
// to float32 and reduces the stack min/max.
//
// Packed accumulation of a word w = lo | hi << 16:   sum += w  (ALU pipe);  hs = mad.hi(w, 65536, hs)  (FMA
// pipe); lo_total = sum - (hs << 16)  (mod 2^32, exact because lo_total < 2^32).  One op per element.
#pragma once
#include "common.cuh"

namespace pcb {

constexpr int RR_THREADS = 512;
constexpr int RR_TBL_STRIDE = 16;    // table entries per slice (>= max active views per view-row, multiple of 4)

constexpr int RR_MAX_SLICES = 64;    // slices per launch (refocus parameters travel by value)

struct RowsCtl {                     // per-launch control block, passed by value (kernel-parameter bank)
    int32_t a[RR_MAX_SLICES];
    uint32_t zeroed_bits[8];         // bit (j*M+i) set -> view masked by the circular aperture (M <= 16)
};

__device__ __forceinline__ bool ctl_zeroed(const RowsCtl &c, int idx) { return (c.zeroed_bits[idx >> 5] >> (idx & 31)) & 1u; }

struct RowsArgs {
    const uint16_t *vp;
    uint32_t *acc;      // [N][S][pitch]
    uint32_t *edge;     // [N][M][2][pitch]
    int M, S, T, C, N;
    int E;              // T*C elements per row
    int pitch;          // accumulator row pitch (elements, multiple of 4)
    int padE;           // replicated elements on each side (C * max shift in px, capped at E)
    int chunkE;         // output elements per CTA (even)
    int nxc;            // x chunks
    int rowWords;       // 32-bit words per staged row copy
    int nsl;            // slices per launch (<= RR_MAX_SLICES), all handled by every CTA
    int n0;             // first slice of this launch
};

template <int WPT>
__global__ void __launch_bounds__(RR_THREADS, 1)
refocus_rows_kernel(RowsArgs p, RowsCtl ctl) {
    extern __shared__ __align__(16) uint32_t smem[];
    __shared__ int s_act[RR_TBL_STRIDE];
    __shared__ int s_nact;

    const int r = blockIdx.x;
    const int j = blockIdx.y;
    const int xc = blockIdx.z;
    const int n0 = p.n0;
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

    // ---- stage rows: copy A at element idx, copy B shifted by one element
    const int rowElems = p.rowWords * 2;
    const int needElems = p.chunkE + 2 * p.padE + 2;
    for (int ii = 0; ii < nact; ++ii) {
        const int i = s_act[ii];
        const uint16_t *src = p.vp + ((size_t)(j * p.M + i) * p.S + r) * p.E;
        uint16_t *A = reinterpret_cast<uint16_t *>(smem + (size_t)(ii * 2) * p.rowWords);
        uint16_t *B = A + rowElems;
        for (int idx = tid; idx < needElems; idx += RR_THREADS) {
            int e = e0 - p.padE + idx;
            if (e < 0) {
                int ch = e % p.C;
                e = ch < 0 ? ch + p.C : ch;
            } else if (e >= p.E) {
                e = p.E - p.C + (e % p.C);
            }
            const uint16_t v = __ldg(src + e);
            A[idx] = v;
            if (idx > 0) B[idx - 1] = v;
        }
    }
    for (int idx = tid; idx < p.rowWords; idx += RR_THREADS) zero_row[idx] = 0u;
    // ---- offset table: word index of the staged data feeding output word 0, per (slice, active view)
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
        tbl[t] = woff;
    }
    __syncthreads();

    const int chunkWords = min(p.chunkE, p.E - e0 + 1) >> 1;     // words that hold at least one valid element
    const int dy = c0 - j;
    int wk[WPT];                                                  // clamped word index (loads stay inside the row)
#pragma unroll
    for (int k = 0; k < WPT; ++k) wk[k] = min(tid + k * RR_THREADS, max(chunkWords - 1, 0));
    for (int al = 0; al < nloc; ++al) {
        uint32_t sum[WPT], hs[WPT];
#pragma unroll
        for (int k = 0; k < WPT; ++k) { sum[k] = 0u; hs[k] = 0u; }
        const int4 *trow = reinterpret_cast<const int4 *>(tbl + al * RR_TBL_STRIDE);
        for (int g = 0; g < ngrp; ++g) {
            const int4 o = trow[g];
            uint32_t v[4][WPT];
#pragma unroll
            for (int k = 0; k < WPT; ++k) {
                v[0][k] = smem[o.x + wk[k]];
                v[1][k] = smem[o.y + wk[k]];
                v[2][k] = smem[o.z + wk[k]];
                v[3][k] = smem[o.w + wk[k]];
            }
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int k = 0; k < WPT; ++k) {
                    sum[k] += v[q][k];
                    hs[k] = __umulhi(v[q][k], 65536u) + hs[k];
                }
        }
        const int n = n0 + al;
        const int a = ctl.a[al];
        const int y = r - a * dy;
        uint32_t *arow = (y >= 0 && y < p.S) ? p.acc + ((size_t)n * p.S + y) * p.pitch : nullptr;
        uint32_t *etop = (r == 0) ? p.edge + ((size_t)(n * p.M + j) * 2 + 0) * p.pitch : nullptr;
        uint32_t *ebot = (r == p.S - 1) ? p.edge + ((size_t)(n * p.M + j) * 2 + 1) * p.pitch : nullptr;
#pragma unroll
        for (int k = 0; k < WPT; ++k) {
            const int w = tid + k * RR_THREADS;
            if (w < chunkWords) {
                const uint32_t hi = hs[k];
                const uint32_t lo = sum[k] - (hi << 16);
                const int e = e0 + 2 * w;
                const bool has_hi = (e + 1 < p.E) && (2 * w + 1 < p.chunkE);
                if (arow) {
                    atomicAdd(arow + e, lo);
                    if (has_hi) atomicAdd(arow + e + 1, hi);
                }
                if (etop) { etop[e] = lo; if (has_hi) etop[e + 1] = hi; }
                if (ebot) { ebot[e] = lo; if (has_hi) ebot[e + 1] = hi; }
            }
        }
    }
}

// out[n,y,e] = (acc[n,y,e] + sum_{j: source row clamped} edge[n,j,side,e]) / M ;  fused min/max
__global__ void __launch_bounds__(256)
refocus_rows_finalize_kernel(const uint32_t *__restrict__ acc, const uint32_t *__restrict__ edge, RowsCtl ctl,
                             float *__restrict__ out, int M, int S, int E, int pitch, int n0, pcb_minmax_t *mm) {
    __shared__ int s_side[16];     // per j: -1 none, 0 top, 1 bottom
    const int y = blockIdx.x, n = n0 + blockIdx.y;
    const int a = ctl.a[blockIdx.y];
    const int c0 = M / 2;
    for (int j = threadIdx.x; j < M; j += blockDim.x) {
        bool any = false;
        for (int i = 0; i < M; ++i) any |= !ctl_zeroed(ctl, j * M + i);
        const int ys = y + a * (c0 - j);
        s_side[j] = !any ? -1 : (ys < 0 ? 0 : (ys > S - 1 ? 1 : -1));
    }
    __syncthreads();
    const uint32_t *arow = acc + ((size_t)n * S + y) * pitch;
    float *orow = out + ((size_t)n * S + y) * E;
    float mn = INFINITY, mx = -INFINITY;
    for (int e = threadIdx.x; e < E; e += blockDim.x) {
        uint32_t v = arow[e];
        for (int j = 0; j < M; ++j) {
            const int sd = s_side[j];
            if (sd >= 0) v += __ldg(edge + ((size_t)(n * M + j) * 2 + sd) * pitch + e);
        }
        const float f = (float)((double)v / (double)M);
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
    int wpt;
    size_t acc_bytes, edge_bytes;
};

// Host-side planning (no device access): picks the x chunking so that the staged rows fit in shared memory.
static RowsPlan plan_refocus_rows(int M, int S, int T, int C, int N, int max_abs_a, int max_active_per_row) {
    RowsPlan pl;
    memset(&pl, 0, sizeof(pl));
    pl.ok = false;
    if (M > 16 || max_active_per_row > RR_TBL_STRIDE || max_active_per_row < 1) return pl;
    const int E = T * C;
    const int c0 = M / 2;
    long long shift = (long long)max_abs_a * c0;
    if (shift > T) shift = T;
    const int padE = (int)shift * C;
    const int nsl = N < RR_MAX_SLICES ? N : RR_MAX_SLICES;
    const size_t budget = 220 * 1024;
    const size_t tbl_bytes = (size_t)nsl * RR_TBL_STRIDE * 4;
    for (int nxc = 1; nxc <= 64; ++nxc) {
        int chunkE = (E + nxc - 1) / nxc;
        chunkE += chunkE & 1;
        const int rowWords = ((chunkE + 2 * padE + 2 + 1) / 2 + 3) & ~3;     // multiple of 4 words (16 B rows)
        const size_t smem = ((size_t)2 * max_active_per_row + 1) * rowWords * 4 + tbl_bytes;
        const int words = chunkE / 2;
        const int wpt = (words + RR_THREADS - 1) / RR_THREADS;
        if (smem <= budget && wpt <= 4) {
            pl.ok = true;
            pl.smem = smem;
            pl.wpt = wpt < 1 ? 1 : wpt;
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

static int launch_refocus_rows(const RowsPlan &pl, const uint16_t *vp, const int32_t *a_list_host,
                               const uint8_t *view_zeroed_host, void *scratch, float *out, pcb_minmax_t *mm,
                               cudaStream_t st) {
    RowsArgs a = pl.a;
    a.vp = vp;
    a.acc = reinterpret_cast<uint32_t *>(scratch);
    a.edge = reinterpret_cast<uint32_t *>(reinterpret_cast<char *>(scratch) + pl.acc_bytes);
    PCB_CUDA(cudaMemsetAsync(a.acc, 0, pl.acc_bytes, st));
    RowsCtl ctl;
    memset(&ctl, 0, sizeof(ctl));
    for (int k = 0; k < a.M * a.M; ++k)
        if (view_zeroed_host[k]) ctl.zeroed_bits[k >> 5] |= 1u << (k & 31);
    switch (pl.wpt) {
    case 1: PCB_CUDA(cudaFuncSetAttribute(refocus_rows_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem)); break;
    case 2: PCB_CUDA(cudaFuncSetAttribute(refocus_rows_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem)); break;
    case 3: PCB_CUDA(cudaFuncSetAttribute(refocus_rows_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem)); break;
    default: PCB_CUDA(cudaFuncSetAttribute(refocus_rows_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem)); break;
    }
    const int N = a.N;
    for (int n0 = 0; n0 < N; n0 += RR_MAX_SLICES) {
        const int nloc = N - n0 < RR_MAX_SLICES ? N - n0 : RR_MAX_SLICES;
        for (int k = 0; k < RR_MAX_SLICES; ++k) ctl.a[k] = k < nloc ? a_list_host[n0 + k] : 0;
        a.n0 = n0;
        a.nsl = nloc;
        dim3 grid(a.S, a.M, a.nxc), block(RR_THREADS);
        switch (pl.wpt) {
        case 1: refocus_rows_kernel<1><<<grid, block, pl.smem, st>>>(a, ctl); break;
        case 2: refocus_rows_kernel<2><<<grid, block, pl.smem, st>>>(a, ctl); break;
        case 3: refocus_rows_kernel<3><<<grid, block, pl.smem, st>>>(a, ctl); break;
        default: refocus_rows_kernel<4><<<grid, block, pl.smem, st>>>(a, ctl); break;
        }
        PCB_LAUNCH_OK();
        dim3 fgrid(a.S, nloc);
        refocus_rows_finalize_kernel<<<fgrid, 256, 0, st>>>(a.acc, a.edge, ctl, out, a.M, a.S, a.E, a.pitch, n0, mm);
        PCB_LAUNCH_OK();
    }
    return PCB_OK;
}

}  // namespace pcb
