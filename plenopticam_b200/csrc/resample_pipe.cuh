// (a) Micro-image resampling, PERSISTENT warp-specialised form of resample_col.cuh for uint16 RGB exposures.
//
// resample_col_kernel is one tile per CTA: tables -> raw rows -> barrier -> columns, and what bounds it is that chain
// (ncu: barrier 27-31 %, long scoreboard 15-38 % of the stall samples at 65 % issue utilisation), not bytes or flops.
// Here a CTA walks many tiles and a dedicated PRODUCER warp runs one tile ahead of the CONSUMER warps:
//
//   producer : tile tables (hex-column entries + lens window, one global-memory latency), the tile's bounding box, then
//              ONE 2-D tensor copy of the box (cp.async.bulk.tensor.2d; or one cp.async.bulk per raw row when the base or
//              the row pitch is not 16-byte aligned) - TMA unit, no registers, completion on an mbarrier - into a uint16
//              ring of two slots; it also snapshots the running global bounds for the min/max pass' skip test;
//   consumers: wait for the slot, convert it to the float32 tile (shared -> shared, the same magic-number conversion as
//              before), and run the column loop of resample_col.cuh (rc_columns, unchanged arithmetic: identical bits).
//
// So the global-memory latency of tile t+1 is spent while tile t is interpolated, a dropped tile of the min/max pass costs
// its scan only (no CTA launch, no table / raw latency on the critical path), and there is no per-tile CTA prologue.
// Tiles whose aligned bulk copies would leave the caller's buffer, whose box does not fit the tile, or rows that are not
// 4-byte phase-aligned take the direct global path inside rc_columns, exactly like the one-tile kernel.
//
// Three forms of the slot (RpkRaw<>):
//   uint16_t       the slot is converted into a float tile (two slots + tile, 108 KB) - the passes that interpolate every
//                  tile (quantise, float output);
//   RpkU16InPlace  no float tile: the column loop reads the uint16 rows in place (conversion on load) and the ring has four
//                  slots - the min/max pass, which interpolates a fraction of its tiles (needs W % 8 == 0);
//   float          float32 exposures (after de-vignetting / rotation): the slot IS the tile, read in place (W % 4 == 0).
#pragma once
#include <cuda.h>
#include "common.cuh"
#include "resample_col.cuh"
#include "refocus_pxp.cuh"        // mbarrier / bulk-copy helpers

namespace pcb {

#ifndef PCB_RPK_KT
#define PCB_RPK_KT 16
#define PCB_RPK_TS 704
#define PCB_RPK_CONS 384
#endif
constexpr int RPK_KT = PCB_RPK_KT;                   // output lens columns per tile
constexpr int RPK_TS = PCB_RPK_TS;                   // float tile row stride (a multiple of 32)
constexpr int RPK_CONS = PCB_RPK_CONS;               // consumer threads
constexpr int RPK_THREADS = RPK_CONS + 32;           // + one producer warp
constexpr int RPK_RSTR = ((RPK_TS * 2 + 32 + 15) / 16) * 16;      // bytes per raw row in a ring slot
constexpr int RPK_SLOT = (RC_ROWS * RPK_RSTR + 127) / 128 * 128;      // bytes per ring slot (tensor copies land 128-byte aligned)

// Per raw element type.  uint16: the slot is converted into a float tile before the column loop.  float32 (what
// LfpAligner hands on after de-vignetting or rotation): the slot IS the tile - no float tile, no conversion pass; the
// column loop reads the staged rows in place, which needs every row of the box at the same 16-byte phase (row pitch a
// multiple of 16 bytes, i.e. W % 4 == 0).
struct RpkU16InPlace {};          // tag: uint16 exposure whose rows share one 16-byte phase (W % 8 == 0), read in place
template <typename T> struct RpkRaw;
template <> struct RpkRaw<uint16_t> {
    typedef uint16_t elem;
    static constexpr int ES = 2, RSTR = RPK_RSTR, SLOT = RPK_SLOT, NSLOT = 2;
    static constexpr bool INPLACE = false;
};
template <> struct RpkRaw<RpkU16InPlace> {
    typedef uint16_t elem;
    static constexpr int ES = 2, RSTR = RPK_RSTR, SLOT = RPK_SLOT, NSLOT = 4;
    static constexpr bool INPLACE = true;
};
template <> struct RpkRaw<float> {
    typedef float elem;
    static constexpr int ES = 4, RSTR = ((RPK_TS * 4 + 32 + 15) / 16) * 16, SLOT = (RC_ROWS * RSTR + 127) / 128 * 128, NSLOT = 2;
    static constexpr bool INPLACE = true;
};

struct RpkDesc {                  // what the producer hands over per tile
    int ly, k0, kn, xa, nsrc;     // ly < 0: no more tiles
    int ymin, xmin, R, Wt;
    int use_tile;                 // 1: raw rows are in the ring slot; 0: rc_columns reads global memory directly
    int ninv;
    int rowoff[RC_ROWS];          // byte offset of sample (xmin * 3) of row rr inside its slot row
    float g_mn, g_mx;             // running global bounds when the tile was prepared (min/max pass)
    pcb_lens_t lens[RC_MAX_SRC];
};

__device__ __forceinline__ void rpk_consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(RPK_CONS) : "memory"); }
__device__ __forceinline__ void tma_tensor2d_g2s(void *dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(dst)), "l"(reinterpret_cast<unsigned long long>(map)), "r"(c0), "r"(c1), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- the tables of one tile, per lane of ONE warp (used by the producer warp and by the tile-bounds kernel) ----------
struct RpkPending { pcb_hexcol_t h; pcb_lens_t Ls; int xa_est; };      // loads in flight: first / last hex column, lens window

__device__ __forceinline__ RpkPending rpk_request(const ResampleArgs &p, int t, int ntx, int lane) {
    RpkPending q;
    q.h.x0 = q.h.x1 = 0; q.h.w = 0.f; q.h.reserved = 0;
    const int ly = t / ntx, k0 = (t - ly * ntx) * RPK_KT;
    const int kn = min(RPK_KT, p.Xs - k0);
    const pcb_hexcol_t *hc = p.hexcol ? p.hexcol + ((ly + p.hex_odd) & 1) * p.Xs : nullptr;
    q.xa_est = k0;
    if (hc) {
        q.xa_est = max((int)floorf((float)k0 * (float)p.LX / (float)max(p.Xs - 1, 1)) - 1, 0);
        if (lane < 2) q.h = hc[lane == 0 ? k0 : k0 + kn - 1];
    }
    q.Ls = p.lens[ly * p.LX + min(q.xa_est + lane, p.LX - 1)];
    return q;
}

struct RpkBox {
    int ly, k0, kn, xa, nsrc;
    int ymin, xmin, R, Wt;
    bool use_tile, mine;          // mine: this lane holds source lens xa + lane in L
    unsigned inv;                 // ballot of the invalid source lenses
    size_t a0;                    // lane < R: 16-byte aligned start of the bulk copy of box row `lane`
    uint32_t nbytes;              // ... and its length (the aligned superset of the row, + one sample of slack)
    pcb_lens_t L;
};

// (as in resample_col_kernel) source lens range, bounding box of the raw samples, and whether the box can travel by bulk
// copies: 16-byte aligned supersets of every row inside the caller's buffer, 4-byte phase-aligned rows
template <typename TRAW = uint16_t>
__device__ __forceinline__ RpkBox rpk_box(const ResampleArgs &p, int M, int t, int ntx, int lane, const RpkPending &cur,
                                          size_t raw_lo, size_t raw_hi) {
    constexpr int ES = RpkRaw<TRAW>::ES;
    RpkBox b;
    b.ly = t / ntx; b.k0 = (t - b.ly * ntx) * RPK_KT;
    b.kn = min(RPK_KT, p.Xs - b.k0);
    int xa = b.k0, xb = b.k0 + b.kn - 1;
    if (p.hexcol) {
        xa = __shfl_sync(0xffffffffu, cur.h.x0, 0);
        const int bx0 = __shfl_sync(0xffffffffu, cur.h.x0, 1), bx1 = __shfl_sync(0xffffffffu, cur.h.x1, 1);
        const float bw = __shfl_sync(0xffffffffu, cur.h.w, 1);
        xb = bw == 0.f ? bx0 : bx1;
    }
    b.xa = xa;
    b.nsrc = xb - xa + 1;
    const int from = xa + lane - cur.xa_est;
    b.L.oy = __shfl_sync(0xffffffffu, cur.Ls.oy, from & 31);
    b.L.ox = __shfl_sync(0xffffffffu, cur.Ls.ox, from & 31);
    b.L.wy = __shfl_sync(0xffffffffu, cur.Ls.wy, from & 31);
    b.L.wx = __shfl_sync(0xffffffffu, cur.Ls.wx, from & 31);
    b.mine = lane < b.nsrc && b.nsrc <= RC_MAX_SRC;
    if (b.mine && (from < 0 || from > 31 || cur.xa_est + from > p.LX - 1)) b.L = p.lens[b.ly * p.LX + xa + lane];
    const bool valid = b.mine && b.L.oy != PCB_LENS_INVALID;
    int b0 = valid ? b.L.oy : INT32_MAX, b1 = valid ? b.L.oy : INT32_MIN;
    int b2 = valid ? b.L.ox : INT32_MAX, b3 = valid ? b.L.ox : INT32_MIN;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        b0 = min(b0, __shfl_xor_sync(0xffffffffu, b0, o)); b1 = max(b1, __shfl_xor_sync(0xffffffffu, b1, o));
        b2 = min(b2, __shfl_xor_sync(0xffffffffu, b2, o)); b3 = max(b3, __shfl_xor_sync(0xffffffffu, b3, o));
    }
    b.inv = __ballot_sync(0xffffffffu, b.mine && !valid);
    b.ymin = b0; b.xmin = b2;
    b.R = b1 - b.ymin + M + 1; b.Wt = (b3 - b.xmin + M + 1) * 3;
    const bool any_valid = b1 != INT32_MIN;
    bool use_tile = b.nsrc <= RC_MAX_SRC && any_valid && b.R <= RC_ROWS && b.Wt + (ES == 2 ? 2 : 4) <= RPK_TS &&
                    (raw_lo & 3) == 0 && ((p.W * 3 * ES) & (RpkRaw<TRAW>::INPLACE ? 15 : 3)) == 0;
    b.a0 = 0; b.nbytes = 0;
    if (use_tile && lane < b.R) {
        const size_t A = raw_lo + (((size_t)(b.ymin + lane)) * p.W + b.xmin) * (3 * ES);
        b.a0 = A & ~(size_t)15;
        b.nbytes = (uint32_t)(((A + (size_t)b.Wt * ES + ES + 15) & ~(size_t)15) - b.a0);
    }
    const bool row_ok = !(use_tile && lane < b.R) ||
                        (b.a0 >= raw_lo && b.a0 + b.nbytes <= raw_hi && b.nbytes <= (uint32_t)RpkRaw<TRAW>::RSTR);
    b.use_tile = use_tile && __all_sync(0xffffffffu, row_ok);
    return b;
}

template <int MODE, int MT, typename TRAW = uint16_t>
__global__ void __launch_bounds__(RPK_THREADS, 2)
resample_pipe_kernel(ResampleArgs p, float *__restrict__ out_f32, uint16_t *__restrict__ out_u16, pcb_minmax_t *mm_rw,
                     const pcb_minmax_t *mm_ro, int ntx, int ntiles, const __grid_constant__ CUtensorMap tmap, int use2d) {
    using RR = RpkRaw<TRAW>;
    typedef typename RR::elem TEL;
    constexpr int ES = RR::ES;
    extern __shared__ __align__(128) unsigned char dsm[];
    float *tile = reinterpret_cast<float *>(dsm);                                  // [RC_ROWS][RPK_TS] (uint16 only)
    unsigned char *ring = dsm + (RR::INPLACE ? 0 : (size_t)RC_ROWS * RPK_TS * sizeof(float));    // 2 slots of RR::SLOT bytes
    constexpr int NSLOT = RR::NSLOT;
    __shared__ RpkDesc s_desc[NSLOT];
    __shared__ __align__(8) uint64_t s_full[NSLOT], s_empty[NSLOT];
    __shared__ uint32_t s_smm[2];          // min / max of the staged raw samples (uint16 values, or ordered keys of floats)
    __shared__ uint32_t s_tmm[2];          // ordered keys of the min / max of the tile being interpolated
    const int M = MT ? MT : p.M;
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const size_t raw_lo = reinterpret_cast<size_t>(p.raw), raw_hi = raw_lo + (size_t)p.H * p.W * (3 * ES);
    const bool may_skip = MODE == RS_MINMAX && mm_rw != nullptr && p.skip;

    if (tid == 0) {
        s_tmm[0] = 0xffffffffu; s_tmm[1] = 0u;
        for (int k = 0; k < NSLOT; ++k) { mbar_init(&s_full[k], 1); mbar_init(&s_empty[k], RPK_CONS / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == RPK_CONS / 32) {
        // ================================================= producer =================================================
        int it = 0;                        // descriptors handed over so far
        // hand one tile (or the end marker, b == nullptr) to the consumers through the next ring slot
        auto hand_over = [&](const RpkBox *b, float g_mn, float g_mx) {
            const int slot = it % NSLOT;
            const bool use_tile = b != nullptr && b->use_tile;
            uint32_t total = use_tile && lane < b->R ? b->nbytes : 0u;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
            if (use_tile && use2d) total = (uint32_t)(RC_ROWS * RR::RSTR);         // the box of the tensor map, whole
            // the slot (ring rows AND descriptor) is free once the consumers are through with the tile before last
            if (it >= NSLOT) mbar_wait(&s_empty[slot], ((it / NSLOT) - 1) & 1);
            RpkDesc &d = s_desc[slot];
            if (b != nullptr) {
                if (b->mine) d.lens[lane] = b->L;
                if (use_tile && lane < b->R) {
                    const size_t A = raw_lo + (((size_t)(b->ymin + lane)) * p.W + b->xmin) * (3 * ES);
                    d.rowoff[lane] = (int)(A - b->a0);
                }
                if (lane == 0) {
                    d.ly = b->ly; d.k0 = b->k0; d.kn = b->kn; d.xa = b->xa; d.nsrc = b->nsrc;
                    d.ymin = b->ymin; d.xmin = b->xmin; d.R = b->R; d.Wt = b->Wt;
                    d.use_tile = use_tile ? 1 : 0;
                    d.ninv = __popc(b->inv);
                    d.g_mn = g_mn; d.g_mx = g_mx;
                }
            } else if (lane == 0) {
                d.ly = -1;
            }
            __syncwarp();
            if (lane == 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic reads of the slot before async writes
                mbar_expect_tx(&s_full[slot], total);                           // (arrive: the descriptor stores precede it)
            }
            __syncwarp();
            if (use_tile && use2d) {
                // ONE tensor copy of the whole box (19 rows x 1440 bytes, zero-filled outside the image): 19 row copies per
                // tile cost the pipeline 40 % more time than one copy of the same bytes
                if (lane == 0)
                    tma_tensor2d_g2s(ring + (size_t)slot * RR::SLOT, &tmap, (b->xmin * (3 * ES) & ~15) >> 3, b->ymin, &s_full[slot]);
            } else
            if (use_tile && lane < b->R)
                tma_bulk_g2s(ring + (size_t)slot * RR::SLOT + (size_t)lane * RR::RSTR, reinterpret_cast<const void *>(b->a0),
                             b->nbytes, &s_full[slot]);
            ++it;
        };
        const float nanf_ = __int_as_float(0x7fffffff);       // bounds "unknown": the tile's own bounds are always committed
        auto bounds = [&](uint32_t &kmn, uint32_t &kmx) {
            kmn = kmx = 0u;
            if (may_skip && lane == 0) {
                kmn = *reinterpret_cast<volatile uint32_t *>(&mm_rw->kmin);
                kmx = *reinterpret_cast<volatile uint32_t *>(&mm_rw->kmax);
            }
        };
        {
            // ---- static round-robin; the table loads of tile t + stride are issued before tile t is handed over
            // (register-carried), so their latency runs under the wait for a free slot
            RpkPending nxt = rpk_request(p, blockIdx.x, ntx, lane);
            uint32_t nkmn, nkmx;
            bounds(nkmn, nkmx);
            for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
                const RpkPending cur = nxt;
                const uint32_t kmn = nkmn, kmx = nkmx;
                if (t + (int)gridDim.x < ntiles) { nxt = rpk_request(p, t + gridDim.x, ntx, lane); bounds(nkmn, nkmx); }
                float g_mn = nanf_, g_mx = nanf_;
                if (may_skip) {
                    g_mn = key_to_float(__shfl_sync(0xffffffffu, kmn, 0));
                    g_mx = key_to_float(__shfl_sync(0xffffffffu, kmx, 0));
                }
                const RpkBox b = rpk_box<TRAW>(p, M, t, ntx, lane, cur, raw_lo, raw_hi);
                hand_over(&b, g_mn, g_mx);
            }
        }
        hand_over(nullptr, 0.f, 0.f);
        return;
    }

    // ===================================================== consumers =====================================================
    ResampleOut<MODE> out;
    out.f32 = out_f32; out.u16 = out_u16;
    out.qz.init(0.f, 0.f);
    if (MODE == RS_U16) out.qz.init(key_to_float(mm_ro->kmin), key_to_float(mm_ro->kmax));
    for (int it = 0;; ++it) {
        const int slot = it % NSLOT;
        mbar_wait(&s_full[slot], (it / NSLOT) & 1);
        const RpkDesc &d = s_desc[slot];
        if (d.ly < 0) break;
        const bool use_tile = d.use_tile != 0;
        const int R = d.R, Wt = d.Wt;
        int shift = 0;
        const unsigned char *srow0 = ring + (size_t)slot * RR::SLOT;
        int nw = 0;
        if (use_tile) {
            if (RR::INPLACE) {
                shift = d.rowoff[0] / ES;                                       // samples from the start of the slot row
                nw = (d.rowoff[0] + Wt * ES + 15) >> 4;                         // 16-byte vectors per row (scan)
            } else {
                shift = (d.rowoff[0] >> 1) & 1;                                 // same phase for every row (W * 3 even)
                nw = (Wt + shift + 1) >> 1;         // words per row from the word that holds the first sample
            }
        }
        bool skip = false;
        if (may_skip) {
            // min / max of the raw samples of the tile (a word / vector may hold samples before or after the box:
            // neighbours of the row, a superset is fine); most tiles end here
            if (tid == 0) { s_smm[0] = 0xffffffffu; s_smm[1] = 0u; }
            uint32_t lo, hi;
            if (RR::INPLACE && ES == 2) {
                uint32_t pmn = 0xffffffffu, pmx = 0u;
                if (use_tile) {
                    for (int rr = warp; rr < R; rr += RPK_CONS / 32) {
                        const uint4 *vsrc = reinterpret_cast<const uint4 *>(srow0 + (size_t)rr * RR::RSTR);
                        for (int m = lane; m < nw; m += 32) {
                            const uint4 v = vsrc[m];
                            pmn = __vminu2(__vminu2(pmn, v.x), __vminu2(__vminu2(v.y, v.z), v.w));
                            pmx = __vmaxu2(__vmaxu2(pmx, v.x), __vmaxu2(__vmaxu2(v.y, v.z), v.w));
                        }
                    }
                }
                lo = min(pmn & 0xffffu, pmn >> 16); hi = max(pmx & 0xffffu, pmx >> 16);
            } else if (RR::INPLACE) {
                float fmn = INFINITY, fmx = -INFINITY;
                if (use_tile) {
                    for (int rr = warp; rr < R; rr += RPK_CONS / 32) {
                        const float4 *vsrc = reinterpret_cast<const float4 *>(srow0 + (size_t)rr * RR::RSTR);
                        for (int m = lane; m < nw; m += 32) {
                            const float4 v = vsrc[m];
                            fmn = fminf(fmn, fminf(fminf(v.x, v.y), fminf(v.z, v.w)));
                            fmx = fmaxf(fmx, fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)));
                        }
                    }
                }
                lo = float_to_key(fmn); hi = float_to_key(fmx);                 // order-preserving keys
            } else {
                uint32_t pmn = 0xffffffffu, pmx = 0u;
                if (use_tile) {
                    for (int rr = warp; rr < R; rr += RPK_CONS / 32) {
                        const uint32_t *wsrc = reinterpret_cast<const uint32_t *>(srow0 + (size_t)rr * RR::RSTR + (d.rowoff[rr] & ~3));
#pragma unroll 4
                        for (int m = lane; m < nw; m += 32) {
                            const uint32_t v = wsrc[m];
                            pmn = __vminu2(pmn, v);
                            pmx = __vmaxu2(pmx, v);
                        }
                    }
                }
                lo = min(pmn & 0xffffu, pmn >> 16); hi = max(pmx & 0xffffu, pmx >> 16);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
                hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
            }
            rpk_consumer_sync();             // s_smm reset visible
            if (lane == 0 && use_tile) {
                atomicMin(&s_smm[0], lo);
                atomicMax(&s_smm[1], hi);
            }
            rpk_consumer_sync();             // s_smm complete
            if (use_tile && d.ninv == 0) {
                const float tmn = ES == 4 ? key_to_float(s_smm[0]) : (float)s_smm[0];
                const float tmx = ES == 4 ? key_to_float(s_smm[1]) : (float)s_smm[1];
                // the margins scale with the magnitude, whatever the sign (float exposures may hold negative samples)
                skip = tmn - fabsf(tmn) * RC_SKIP_EPS > d.g_mn && tmx + fabsf(tmx) * RC_SKIP_EPS < d.g_mx;     // uniform over the CTA
            }
        }
        if (!RR::INPLACE && use_tile && !skip) {
            // ring slot (uint16, any 2-byte phase) -> float tile: whole 32-bit words
            for (int rr = warp; rr < R; rr += RPK_CONS / 32) {
                const uint32_t *wsrc = reinterpret_cast<const uint32_t *>(srow0 + (size_t)rr * RR::RSTR + (d.rowoff[rr] & ~3));
                float2 *dst = reinterpret_cast<float2 *>(tile + rr * RPK_TS);
#pragma unroll 4
                for (int m = lane; m < nw; m += 32) {
                    const uint32_t v = wsrc[m];
                    dst[m] = make_float2(__uint_as_float(__byte_perm(v, 0x4B000000u, 0x7610)) - 8388608.0f,
                                         __uint_as_float(__byte_perm(v, 0x4B000000u, 0x7632)) - 8388608.0f);
                }
            }
        }
        if (!RR::INPLACE && !skip) rpk_consumer_sync();      // tile complete (uniform: every consumer took the same decision)
        if (!skip) {
            const pcb_hexcol_t *hc = p.hexcol ? p.hexcol + ((d.ly + p.hex_odd) & 1) * p.Xs : nullptr;
            float mn = INFINITY, mx = -INFINITY;
            if (RR::INPLACE)
                rc_columns<MODE, MT, RR::RSTR / ES, TEL, TEL>(p, out, reinterpret_cast<const TEL *>(srow0), shift, d.lens, d.xa,
                                                              d.nsrc <= RC_MAX_SRC, use_tile, d.ymin, d.xmin, d.ly, d.k0, d.kn,
                                                              hc, tid, RPK_CONS, mn, mx);
            else
                rc_columns<MODE, MT, RPK_TS, TEL>(p, out, tile, shift, d.lens, d.xa, d.nsrc <= RC_MAX_SRC, use_tile, d.ymin,
                                                  d.xmin, d.ly, d.k0, d.kn, hc, tid, RPK_CONS, mn, mx);
            if (MODE != RS_U16 && mm_rw != nullptr) {
                // fold the tile's bounds in shared memory first: ONE pair of global atomics per tile (and only when the tile
                // moves the bounds the producer saw) - every warp hitting the same two global words serialises in L2
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
                    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
                }
                if (lane == 0) {
                    if (mn != INFINITY) atomicMin(&s_tmm[0], float_to_key(mn));
                    if (mx != -INFINITY) atomicMax(&s_tmm[1], float_to_key(mx));
                }
            }
        }
        rpk_consumer_sync();                 // everybody is done with the tile and the descriptor
        if (MODE != RS_U16 && mm_rw != nullptr && tid == 0 && !skip) {
            const uint32_t kmn = s_tmm[0], kmx = s_tmm[1];
            s_tmm[0] = 0xffffffffu; s_tmm[1] = 0u;              // for the next tile (nobody touches it before the next fold)
            // d.g_* are NaN until the first commits have landed: the comparisons are then false -> commit
            if (kmn != 0xffffffffu && !(key_to_float(kmn) >= d.g_mn)) atomicMin(&mm_rw->kmin, kmn);
            if (kmx != 0u && !(key_to_float(kmx) <= d.g_mx)) atomicMax(&mm_rw->kmax, kmx);
        }
        if (lane == 0) mbar_arrive(&s_empty[slot]);
    }
}

static bool resample_pipe_applies(const ResampleArgs &a, int raw_dtype) {
    const char *e = getenv("PCB_RESAMPLE_KERNEL");
    if (e && (!strcmp(e, "tile") || !strcmp(e, "col"))) return false;
    if (raw_dtype == PCB_F32)           // the slot is the tile: every row of a box at the same 16-byte phase
        return a.C == 3 && a.M + 2 <= RC_ROWS && a.M >= 1 && (a.W & 3) == 0 && (reinterpret_cast<size_t>(a.raw) & 3) == 0;
    return raw_dtype == PCB_U16 && a.C == 3 && a.M + 2 <= RC_ROWS && a.M >= 1;
}

// The raw exposure as a 2-D tensor of 8-byte elements (rows of W*6 bytes), box = one ring slot.  Needs a 16-byte aligned
// base and row pitch; cuTensorMapEncodeTiled comes from the driver at run time (no link dependency on libcuda).
static bool rpk_make_tensor_map(const ResampleArgs &a, CUtensorMap *map) {
    const char *e = getenv("PCB_RESAMPLE_ROWCOPIES");
    if (e && atoi(e)) return false;
    const size_t row_bytes = (size_t)a.W * 6;
    if ((reinterpret_cast<size_t>(a.raw) & 15) != 0 || (row_bytes & 15) != 0) return false;
    typedef CUresult (*encode_t)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                 const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static encode_t encode = []() -> encode_t {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            return nullptr;
        return reinterpret_cast<encode_t>(fn);
    }();
    if (!encode) return false;
    const cuuint64_t gdim[2] = {(cuuint64_t)(row_bytes / 8), (cuuint64_t)a.H};
    const cuuint64_t gstr[1] = {(cuuint64_t)row_bytes};
    const cuuint32_t box[2] = {(cuuint32_t)(RPK_RSTR / 8), (cuuint32_t)RC_ROWS};
    const cuuint32_t estr[2] = {1, 1};
    return encode(map, CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, const_cast<void *>(a.raw), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int MODE, typename TRAW = uint16_t>
static int launch_resample_pipe(const ResampleArgs &a_in, float *out_f32, uint16_t *out_u16, pcb_minmax_t *mm_rw,
                                const pcb_minmax_t *mm_ro, cudaStream_t st) {
    using RR = RpkRaw<TRAW>;
    ResampleArgs a = a_in;
    a.skip = (getenv("PCB_RESAMPLE_NOSKIP") && atoi(getenv("PCB_RESAMPLE_NOSKIP"))) ? 0 : 1;
    const size_t smem = (RR::INPLACE ? 0 : (size_t)RC_ROWS * RPK_TS * sizeof(float)) + (size_t)RR::NSLOT * RR::SLOT;
    const int ntx = (a.Xs + RPK_KT - 1) / RPK_KT;
    const int ntiles = ntx * a.LY;
    const int grid = ntiles < 2 * sm_count() ? ntiles : 2 * sm_count();
    alignas(64) CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    const int use2d = (RR::ES == 2 && rpk_make_tensor_map(a, &tmap)) ? 1 : 0;       // float rows exceed one tensor box
    if (a.M == 15) {
        PCB_CUDA(cudaFuncSetAttribute(resample_pipe_kernel<MODE, 15, TRAW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        resample_pipe_kernel<MODE, 15, TRAW><<<grid, RPK_THREADS, smem, st>>>(a, out_f32, out_u16, mm_rw, mm_ro, ntx, ntiles, tmap, use2d);
    } else {
        PCB_CUDA(cudaFuncSetAttribute(resample_pipe_kernel<MODE, 0, TRAW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        resample_pipe_kernel<MODE, 0, TRAW><<<grid, RPK_THREADS, smem, st>>>(a, out_f32, out_u16, mm_rw, mm_ro, ntx, ntiles, tmap, use2d);
    }
    PCB_LAUNCH_OK();
    return PCB_OK;
}

}  // namespace pcb
