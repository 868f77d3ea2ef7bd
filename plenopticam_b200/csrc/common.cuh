// Shared helpers for libpcb200.so (sm_100a).  No torch types anywhere in this library.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>

#include "../../include/plenopticam_b200.h"

namespace pcb {

extern thread_local char g_err[512];
// pcb_debug_kernel_events: optional pair of events recorded right before / after the dominant kernel of the next calls
extern thread_local cudaEvent_t g_kev[2];

inline int fail(int code, const char *fmt, const char *a = "", long long b = 0) {
    snprintf(g_err, sizeof(g_err), fmt, a, b);
    return code;
}

#define PCB_REQUIRE(cond, what)                                                                   \
    do {                                                                                          \
        if (!(cond)) return pcb::fail(PCB_ERR_INVALID, "%s: requirement failed: " what, __func__); \
    } while (0)

#define PCB_LAUNCH_OK()                                                                           \
    do {                                                                                          \
        cudaError_t e__ = cudaGetLastError();                                                     \
        if (e__ != cudaSuccess) {                                                                 \
            snprintf(pcb::g_err, sizeof(pcb::g_err), "%s: CUDA error: %s", __func__,              \
                     cudaGetErrorString(e__));                                                    \
            return PCB_ERR_CUDA;                                                                  \
        }                                                                                         \
    } while (0)

#define PCB_CUDA(call)                                                                            \
    do {                                                                                          \
        cudaError_t e__ = (call);                                                                 \
        if (e__ != cudaSuccess) {                                                                 \
            snprintf(pcb::g_err, sizeof(pcb::g_err), "%s: %s failed: %s", __func__, #call,        \
                     cudaGetErrorString(e__));                                                    \
            return PCB_ERR_CUDA;                                                                  \
        }                                                                                         \
    } while (0)

// SM count of the CURRENT device (persistent / strip-mined grids); cached per device ordinal, so a process that
// drives several GPUs sizes every grid for the device it launches on
static inline int sm_count() {
    static int cache[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev < 0 || dev >= 64) dev = 0;
    int n = cache[dev];
    if (n == 0) {
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cache[dev] = n;
    }
    return n;
}

// ---- order-preserving float <-> uint32 keys (for atomic min/max and radix select) -----------------
__host__ __device__ __forceinline__ uint32_t float_to_key(float f) {
#ifdef __CUDA_ARCH__
    uint32_t u = __float_as_uint(f);
#else
    uint32_t u;
    memcpy(&u, &f, 4);
#endif
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float key_to_float(uint32_t k) {
    uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    float f;
    memcpy(&f, &u, 4);
    return f;
#endif
}

// ---- typed read-only loads returning float -------------------------------------------------------
template <typename T> __device__ __forceinline__ float ld_as_float(const T *p);
template <> __device__ __forceinline__ float ld_as_float<uint16_t>(const uint16_t *p) { return (float)__ldg(p); }
template <> __device__ __forceinline__ float ld_as_float<uint8_t>(const uint8_t *p) { return (float)__ldg(p); }
template <> __device__ __forceinline__ float ld_as_float<float>(const float *p) { return __ldg(p); }

// ---- block-wide min/max fold into a pcb_minmax_t ---------------------------------------------------
__device__ __forceinline__ void block_minmax_commit(float mn, float mx, pcb_minmax_t *mm) {
    __shared__ float s_mn[32], s_mx[32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { s_mn[warp] = mn; s_mx[warp] = mx; }
    __syncthreads();
    if (warp == 0) {
        const int nw = (blockDim.x + 31) >> 5;
        mn = lane < nw ? s_mn[lane] : INFINITY;
        mx = lane < nw ? s_mx[lane] : -INFINITY;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        }
        if (lane == 0) {
            atomicMin(&mm->kmin, float_to_key(mn));
            atomicMax(&mm->kmax, float_to_key(mx));
        }
    }
}

// Same, committed to the running min / max of every destination of a fan-out (atomics on peer-mapped memory travel
// over NVLink like any other access).
__device__ __forceinline__ void block_minmax_commit_fanout(float mn, float mx, const pcb_fanout_t &dst) {
    __shared__ float s_mn[32], s_mx[32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { s_mn[warp] = mn; s_mx[warp] = mx; }
    __syncthreads();
    if (warp == 0) {
        const int nw = (blockDim.x + 31) >> 5;
        mn = lane < nw ? s_mn[lane] : INFINITY;
        mx = lane < nw ? s_mx[lane] : -INFINITY;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        }
        if (lane < dst.n && lane < PCB_MAX_PEERS) {          // one lane per destination
            pcb_minmax_t *mm = nullptr;
#pragma unroll
            for (int d = 0; d < PCB_MAX_PEERS; ++d)
                if (d == lane) mm = dst.mm[d];
            if (mm != nullptr) {
                atomicMin(&mm->kmin, float_to_key(mn));
                atomicMax(&mm->kmax, float_to_key(mx));
            }
        }
    }
}

// Normalizer.norm_fun + uint16_norm (misc/normalizer.py:43-46,69-78) in float32.
// The divisor (max - min) is the same for every sample, so the IEEE division is done as Markstein's
// reciprocal-multiply with one FMA correction: with inv = RN(1/d),  q = RN(num*inv),  r = num - d*q (exact in an
// FMA),  q' = RN(q + r*inv)  is the correctly rounded num/d (4 instructions instead of the ~12 of a generic
// __fdiv_rn with its slow-path check).
// Correctly rounded a / d for a divisor that is the same for every sample (Markstein, see above); checked exhaustively
// against __fdiv_rn for a < 2^24, d = 1..16 (tools/checks/div_by_const.cu).
struct DivByConst {
    float d, inv;
    __device__ __forceinline__ void init(float d_) { d = d_; inv = __frcp_rn(d_); }
    __device__ __forceinline__ float operator()(float a) const {
        const float q = __fmul_rn(a, inv);
        return __fmaf_rn(__fmaf_rn(-d, q, a), inv, q);
    }
};

struct QuantU16 {
    float mn, d, inv;
    bool scale;
    __device__ __forceinline__ void init(float mn_, float mx_) {
        mn = mn_;
        scale = (mx_ != mn_) && (mx_ != 0.f);
        d = __fsub_rn(mx_, mn_);
        inv = scale ? __frcp_rn(d) : 1.f;
    }
    // clip((x - min) / (max - min), 0, 1), the division correctly rounded
    __device__ __forceinline__ float unit(float x) const {
        float n = x;
        if (scale) {
            const float num = __fsub_rn(x, mn);
            const float q = __fmul_rn(num, inv);
            const float r = __fmaf_rn(-d, q, num);
            n = __fmaf_rn(r, inv, q);
        }
        return fminf(fmaxf(n, 0.f), 1.f);
    }
    __device__ __forceinline__ uint16_t operator()(float x) const {
        return (uint16_t)__float2uint_rn(__fmul_rn(unit(x), 65535.0f));   // round-half-even like np.round
    }
};

}  // namespace pcb
