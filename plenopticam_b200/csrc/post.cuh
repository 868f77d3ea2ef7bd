// Reductions and element-wise epilogues of the hot path:
//   min/max fold                       (misc/normalizer.py:40-41)
//   Normalizer.uint16_norm             (misc/normalizer.py:43-46,69-78)
//   numpy-style percentiles            (lfp_extractor/lfp_contrast.py:252-255) by exact 4x8-bit radix select
//   auto_hist_align + sRGB             (lfp_contrast.py:261, normalizer.py:53-78, gamma_converter.py:41-42)
#pragma once
#include "common.cuh"
#include <cooperative_groups.h>

namespace pcb {

__global__ void minmax_reset_kernel(pcb_minmax_t *mm) {
    mm->kmin = 0xffffffffu;
    mm->kmax = 0u;
}

__global__ void __launch_bounds__(256)
minmax_f32_kernel(const float *__restrict__ data, int64_t n, pcb_minmax_t *mm) {
    float mn = INFINITY, mx = -INFINITY;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float v = __ldg(data + i);
        mn = fminf(mn, v);
        mx = fmaxf(mx, v);
    }
    block_minmax_commit(mn, mx, mm);
}

__global__ void __launch_bounds__(256)
quantize_u16_kernel(const float *__restrict__ in, int64_t n, const pcb_minmax_t *mm, uint16_t *__restrict__ out) {
    QuantU16 qz;
    qz.init(key_to_float(mm->kmin), key_to_float(mm->kmax));
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = qz(__ldg(in + i));
}

// Normalizer.uint8_norm (misc/normalizer.py:48-51): same normalisation, 255 levels
__global__ void __launch_bounds__(256)
quantize_u8_kernel(const float *__restrict__ in, int64_t n, const pcb_minmax_t *mm, uint8_t *__restrict__ out) {
    QuantU16 qz;
    qz.init(key_to_float(mm->kmin), key_to_float(mm->kmax));
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = (uint8_t)__float2uint_rn(__fmul_rn(qz.unit(__ldg(in + i)), 255.0f));
}

// ---- contrast + sRGB ------------------------------------------------------------------------------
// Normalizer(img, min=lo, max=hi).type_norm(): with NumPy >= 2 the float32 data meets float64 percentile
// scalars, so (x-lo)/(hi-lo) is evaluated in float64, clipped, then cast to float32 (SURVEY §8c iv).
// GammaConverter.srgb_conv then runs in float32.
// `inv` = 1/(hi-lo) in float64: x*inv - lo*inv differs from the reference's float64 division by a few ulp(double)
// of the [0, 1] result, invisible after the float32 cast.  y^(5/12) = exp2(5/12 * log2 y) on the SFU (ex2/lg2.approx, ~2 ulp of
// float32 on [0.003, 1]): |error| < 3e-7 on the [0,1] result, far inside the 1e-4 parity tolerance.
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ float lg2_approx(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ float contrast_srgb_one(float x, double lo, double inv, bool scale) {
    // the clip runs after the (monotone) cast to float32: same result, two float64 instructions fewer — the float64
    // pipe is what the fused refocus epilogue waits for
    // x*inv - lo*inv in one FMA (lo*inv is loop-invariant): within 2 ulp(double) of the reference's float64 quotient
    const float y = fminf(fmaxf(scale ? (float)fma((double)x, inv, -(lo * inv)) : x, 0.f), 1.f);
    // both branches are a handful of instructions: evaluate both and select (no divergent branch).  For y in
    // [0.0031308, 1] the argument is a normal number and the exponent lies in [-3.5, 0]: lg2.approx / ex2.approx need
    // none of log2f's / exp2f's range handling.
    const float pw = __fmaf_rn(1.055f, ex2_approx(__fmul_rn((float)(5.0 / 12.0), lg2_approx(fmaxf(y, 0.0031308f)))), -0.055f);
    const float ln = __fmul_rn(y, 12.92f);
    return y >= 0.0031308f ? pw : ln;
}

// four consecutive samples as float4 (uint16 / uint8 inputs are integer-valued, exact in float32 like the
// reference's np.asarray(data, dtype='float32'), misc/normalizer.py:36)
template <typename T> struct Vec4In;
template <> struct Vec4In<float> {
    static constexpr int ALIGN = 16;
    __device__ static __forceinline__ float4 ld(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }
};
template <> struct Vec4In<uint16_t> {
    static constexpr int ALIGN = 8;
    __device__ static __forceinline__ float4 ld(const uint16_t *p) {
        const uint2 w = __ldg(reinterpret_cast<const uint2 *>(p));
        return make_float4((float)(w.x & 0xffffu), (float)(w.x >> 16), (float)(w.y & 0xffffu), (float)(w.y >> 16));
    }
};
template <> struct Vec4In<uint8_t> {
    static constexpr int ALIGN = 4;
    __device__ static __forceinline__ float4 ld(const uint8_t *p) {
        const uint32_t w = __ldg(reinterpret_cast<const uint32_t *>(p));
        return make_float4((float)(w & 0xffu), (float)((w >> 8) & 0xffu), (float)((w >> 16) & 0xffu), (float)(w >> 24));
    }
};

template <typename T>
__global__ void __launch_bounds__(256)
contrast_srgb_kernel(const T *__restrict__ in, int64_t n, const double *__restrict__ lohi,
                     const pcb_minmax_t *__restrict__ mm, float *__restrict__ out) {
    double lo = lohi[0], hi = lohi[1];
    if (mm != nullptr) {   // misc/normalizer.py:40-41: a falsy bound falls back to the data's own min / max
        if (lo == 0.0) lo = (double)key_to_float(mm->kmin);
        if (hi == 0.0) hi = (double)key_to_float(mm->kmax);
    }
    const bool scale = (hi != lo) && (hi != 0.0);
    const double inv = scale ? 1.0 / (hi - lo) : 1.0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t n4 = n >> 2;
    float4 *out4 = reinterpret_cast<float4 *>(out);
    const bool vec_ok = (reinterpret_cast<uintptr_t>(in) & (Vec4In<T>::ALIGN - 1)) == 0 &&
                        (reinterpret_cast<uintptr_t>(out) & 15) == 0;
    if (vec_ok) {
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
            float4 v = Vec4In<T>::ld(in + 4 * i);
            v.x = contrast_srgb_one(v.x, lo, inv, scale);
            v.y = contrast_srgb_one(v.y, lo, inv, scale);
            v.z = contrast_srgb_one(v.z, lo, inv, scale);
            v.w = contrast_srgb_one(v.w, lo, inv, scale);
            out4[i] = v;
        }
        for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
            out[i] = contrast_srgb_one(ld_as_float(in + i), lo, inv, scale);
    } else {
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
            out[i] = contrast_srgb_one(ld_as_float(in + i), lo, inv, scale);
    }
}

// ---- exact percentiles by radix select -------------------------------------------------------------
// np.percentile(x, q) (default 'linear'): vidx = q/100*(n-1); lo = floor(vidx); t = vidx-lo;
// result = lerp(sorted[lo], sorted[lo+1], t).  We find the needed order statistics exactly with 8-bit
// histogram passes over order-preserving keys (4 passes for float32 data, 8 for float64); up to PS_MAXR ranks
// are resolved together.
constexpr int PS_MAXQ = 4;
constexpr int PS_MAXR = 2 * PS_MAXQ;

struct SelectState {
    uint32_t hist[PS_MAXR][256];
    unsigned long long prefix[PS_MAXR];   // key bits resolved so far (high bits)
    unsigned long long rank[PS_MAXR];     // remaining rank inside the current prefix bucket
    int nr;
    unsigned int done;                    // histogram CTAs finished in the current pass (last one runs the scan)
    double q[PS_MAXQ];
};

template <typename T> struct SelKey;
template <> struct SelKey<float> {
    static constexpr int PASSES = 4;
    __device__ static __forceinline__ unsigned long long key(float v) { return float_to_key(v); }
    __device__ static __forceinline__ double value(unsigned long long k) { return (double)key_to_float((uint32_t)k); }
};
template <> struct SelKey<double> {
    static constexpr int PASSES = 8;
    __device__ static __forceinline__ unsigned long long key(double v) {
        const unsigned long long u = (unsigned long long)__double_as_longlong(v);
        return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
    }
    __device__ static __forceinline__ double value(unsigned long long k) {
        const unsigned long long u = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
        return __longlong_as_double((long long)u);
    }
};

// rgb2gry (color_space_converter, third party, unpinned): BT.709 luma
__device__ __forceinline__ float select_value(const float *data, int64_t i, int luma) {
    if (!luma) return __ldg(data + i);
    const float r = __ldg(data + 3 * i), g = __ldg(data + 3 * i + 1), b = __ldg(data + 3 * i + 2);
    return (float)(0.2126 * (double)r + 0.7152 * (double)g + 0.0722 * (double)b);
}
__device__ __forceinline__ double select_value(const double *data, int64_t i, int luma) {
    if (!luma) return __ldg(data + i);
    const double r = __ldg(data + 3 * i), g = __ldg(data + 3 * i + 1), b = __ldg(data + 3 * i + 2);
    return __dadd_rn(__dadd_rn(__dmul_rn(r, 0.2126), __dmul_rn(g, 0.7152)), __dmul_rn(b, 0.0722));
}

__global__ void select_init_kernel(SelectState *st, int64_t n, int nq, double q0, double q1, double q2, double q3) {
    const double q[PS_MAXQ] = {q0, q1, q2, q3};
    if (threadIdx.x == 0) {
        st->nr = 2 * nq;
        st->done = 0u;
        for (int k = 0; k < nq; ++k) {
            st->q[k] = q[k];
            const double vidx = q[k] / 100.0 * (double)(n - 1);
            long long lo = (long long)floor(vidx);
            lo = lo < 0 ? 0 : (lo > n - 1 ? n - 1 : lo);
            const long long hi = lo + 1 > n - 1 ? n - 1 : lo + 1;
            st->rank[2 * k] = (unsigned long long)lo;
            st->rank[2 * k + 1] = (unsigned long long)hi;
            st->prefix[2 * k] = st->prefix[2 * k + 1] = 0;
        }
    }
    for (int i = threadIdx.x; i < PS_MAXR * 256; i += blockDim.x) (&st->hist[0][0])[i] = 0;
}

// One warp per rank: lane l owns bins 8l..8l+7; exclusive warp scan of the lane sums, then each lane looks for the
// bucket holding the rank among its own eight bins.  Runs in the LAST histogram CTA of a pass (PS_MAXR warps).
__device__ __forceinline__ void select_scan(SelectState *st, int pass) {
    const int r = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nr = st->nr;
    if (r < nr) {
        const volatile uint32_t *h = pass == 0 ? st->hist[0] : st->hist[r];
        uint32_t c[8];
#pragma unroll
        for (int b = 0; b < 8; ++b) c[b] = h[8 * lane + b];
        unsigned long long sum = 0;
#pragma unroll
        for (int b = 0; b < 8; ++b) sum += c[b];
        unsigned long long incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        unsigned long long cum = incl - sum;                    // counts below this lane's bins
        const unsigned long long k = st->rank[r];
        const bool mine = k >= cum && k < incl;
        const unsigned found = __ballot_sync(0xffffffffu, mine);
        if (mine) {
            int b = 0;
#pragma unroll
            for (; b < 7; ++b) {
                if (k < cum + c[b]) break;
                cum += c[b];
            }
            st->rank[r] = k - cum;
            st->prefix[r] = (st->prefix[r] << 8) | (unsigned long long)(8 * lane + b);
        } else if (found == 0u && lane == 31) {                  // rank beyond the histogram (cannot happen): last bin
            st->rank[r] = 0;
            st->prefix[r] = (st->prefix[r] << 8) | 255ull;
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < PS_MAXR * 256; i += blockDim.x) (&st->hist[0][0])[i] = 0;
    if (threadIdx.x == 0) st->done = 0u;
}

// pass 0 looks at the most significant byte; all ranks share the first-level histogram
template <typename T>
__global__ void __launch_bounds__(256)
select_hist_kernel(const T *__restrict__ data, int64_t n, int luma, int pass, SelectState *st) {
    __shared__ uint32_t sh[PS_MAXR][256];
    __shared__ unsigned long long s_prefix[PS_MAXR];
    const int nr = st->nr;
    for (int i = threadIdx.x; i < PS_MAXR * 256; i += blockDim.x) (&sh[0][0])[i] = 0;
    if (threadIdx.x < PS_MAXR) s_prefix[threadIdx.x] = st->prefix[threadIdx.x];
    __syncthreads();
    const int shift = (SelKey<T>::PASSES - 1 - pass) * 8;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const unsigned long long key = SelKey<T>::key(select_value(data, i, luma));
        const uint32_t bin = (uint32_t)(key >> shift) & 0xffu;
        if (pass == 0) {
            atomicAdd(&sh[0][bin], 1u);
        } else {
            const unsigned long long hi = key >> (shift + 8);
            for (int r = 0; r < nr; ++r)
                if (hi == s_prefix[r]) atomicAdd(&sh[r][bin], 1u);
        }
    }
    __syncthreads();
    const int rows = pass == 0 ? 1 : nr;
    for (int i = threadIdx.x; i < rows * 256; i += blockDim.x) {
        const uint32_t v = (&sh[0][0])[i];
        if (v) atomicAdd(&(&st->hist[0][0])[i], v);
    }
    // the last CTA to get here resolves one more key byte of every rank (no separate scan launch)
    __shared__ bool s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(&st->done, 1u) == gridDim.x - 1;
    __syncthreads();
    if (s_last) {
        __threadfence();
        select_scan(st, pass);
    }
}

template <typename T>
__global__ void select_finish_kernel(const SelectState *st, int64_t n, double *result) {
    const int k = threadIdx.x;
    if (2 * k < st->nr) {
        const double a = SelKey<T>::value(st->prefix[2 * k]);
        const double b = SelKey<T>::value(st->prefix[2 * k + 1]);
        const double vidx = st->q[k] / 100.0 * (double)(n - 1);
        const double t = vidx - floor(vidx);
        // numpy _lerp: a + (b-a)*t, and b - (b-a)*(1-t) where t >= 0.5
        // (explicit _rn intrinsics: no FMA contraction, so the result is bit-identical to numpy's)
        const double d = __dsub_rn(b, a);
        double v = __dadd_rn(a, __dmul_rn(d, t));
        if (t >= 0.5) v = __dsub_rn(b, __dmul_rn(d, __dsub_rn(1.0, t)));
        result[k] = v;
    }
}

// ---- both contrast bounds of LfpContrast.auto_hist_align in ONE cooperative launch ------------------------
// lo = percentile(rgb2gry(ref), q_lo), hi = percentile(ref, q_hi) (lfp_extractor/lfp_contrast.py:252-255) on the
// same reference image: ranks 0,1 select over the luma of the npx pixels, ranks 2,3 over the 3*npx channel values;
// the four 8-bit passes, their scans and the final interpolation are separated by grid-wide barriers instead of
// kernel boundaries (13 launches -> 1).
__global__ void __launch_bounds__(256)
contrast_bounds_kernel(const float *__restrict__ ref, int64_t npx, double q_lo, double q_hi, SelectState *st,
                       double *__restrict__ lohi) {
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    __shared__ uint32_t sh[4][256];
    __shared__ unsigned long long s_prefix[4];
    const int64_t nval[2] = {npx, 3 * npx};
    const double q[2] = {q_lo, q_hi};
    if (blockIdx.x == 0) {
        if (threadIdx.x < 2) {
            const int f = threadIdx.x;
            const double vidx = q[f] / 100.0 * (double)(nval[f] - 1);
            long long lo = (long long)floor(vidx);
            lo = lo < 0 ? 0 : (lo > nval[f] - 1 ? nval[f] - 1 : lo);
            const long long hi = lo + 1 > nval[f] - 1 ? nval[f] - 1 : lo + 1;
            st->rank[2 * f] = (unsigned long long)lo;
            st->rank[2 * f + 1] = (unsigned long long)hi;
            st->prefix[2 * f] = st->prefix[2 * f + 1] = 0;
        }
        for (int i = threadIdx.x; i < PS_MAXR * 256; i += blockDim.x) (&st->hist[0][0])[i] = 0;
    }
    __threadfence();
    grid.sync();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int pass = 0; pass < 4; ++pass) {
        for (int i = threadIdx.x; i < 4 * 256; i += blockDim.x) (&sh[0][0])[i] = 0;
        if (threadIdx.x < 4) s_prefix[threadIdx.x] = st->prefix[threadIdx.x];
        __syncthreads();
        const int shift = (3 - pass) * 8;
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < npx; i += stride) {
            const float r = __ldg(ref + 3 * i), g = __ldg(ref + 3 * i + 1), b = __ldg(ref + 3 * i + 2);
            const float y = (float)(0.2126 * (double)r + 0.7152 * (double)g + 0.0722 * (double)b);
            const uint32_t keys[4] = {float_to_key(y), float_to_key(r), float_to_key(g), float_to_key(b)};
#pragma unroll
            for (int v = 0; v < 4; ++v) {
                const int f = v == 0 ? 0 : 1;                           // value function: luma | channel values
                const uint32_t bin = (keys[v] >> shift) & 0xffu;
                if (pass == 0) {
                    atomicAdd(&sh[2 * f][bin], 1u);                     // both ranks of a function share pass 0
                } else {
                    const unsigned long long hi = (unsigned long long)keys[v] >> (shift + 8);
                    if (hi == s_prefix[2 * f]) atomicAdd(&sh[2 * f][bin], 1u);
                    if (hi == s_prefix[2 * f + 1]) atomicAdd(&sh[2 * f + 1][bin], 1u);
                }
            }
        }
        __syncthreads();
        for (int i = threadIdx.x; i < 4 * 256; i += blockDim.x) {
            const uint32_t v = (&sh[0][0])[i];
            if (v) atomicAdd(&(&st->hist[0][0])[i], v);
        }
        __threadfence();
        grid.sync();
        if (blockIdx.x == 0) {
            const int r = threadIdx.x >> 5, lane = threadIdx.x & 31;
            if (r < 4) {
                const volatile uint32_t *h = pass == 0 ? st->hist[r & ~1] : st->hist[r];
                uint32_t c[8];
#pragma unroll
                for (int b = 0; b < 8; ++b) c[b] = h[8 * lane + b];
                unsigned long long sum = 0;
#pragma unroll
                for (int b = 0; b < 8; ++b) sum += c[b];
                unsigned long long incl = sum;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const unsigned long long v = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += v;
                }
                unsigned long long cum = incl - sum;
                const unsigned long long k = st->rank[r];
                if (k >= cum && k < incl) {
                    int b = 0;
#pragma unroll
                    for (; b < 7; ++b) {
                        if (k < cum + c[b]) break;
                        cum += c[b];
                    }
                    st->rank[r] = k - cum;
                    st->prefix[r] = (st->prefix[r] << 8) | (unsigned long long)(8 * lane + b);
                }
            }
            __syncthreads();
            for (int i = threadIdx.x; i < PS_MAXR * 256; i += blockDim.x) (&st->hist[0][0])[i] = 0;
            __threadfence();
        }
        grid.sync();
    }
    if (blockIdx.x == 0 && threadIdx.x < 2) {
        const int f = threadIdx.x;
        const double a = (double)key_to_float((uint32_t)st->prefix[2 * f]);
        const double b = (double)key_to_float((uint32_t)st->prefix[2 * f + 1]);
        const double vidx = q[f] / 100.0 * (double)(nval[f] - 1);
        const double t = vidx - floor(vidx);
        const double d = __dsub_rn(b, a);
        double v = __dadd_rn(a, __dmul_rn(d, t));
        if (t >= 0.5) v = __dsub_rn(b, __dmul_rn(d, __dsub_rn(1.0, t)));
        lohi[f] = v;
    }
}

static int run_contrast_bounds(const float *ref, int64_t npx, double q_lo, double q_hi, double *lohi, void *scratch,
                               cudaStream_t st) {
    SelectState *ss = reinterpret_cast<SelectState *>(scratch);
    int per_sm = 0;
    PCB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, contrast_bounds_kernel, 256, 0));
    int blocks = (int)((npx + 256 * 2 - 1) / (256 * 2));
    const int cap = sm_count() * (per_sm < 1 ? 1 : (per_sm > 4 ? 4 : per_sm));
    blocks = blocks < 1 ? 1 : (blocks > cap ? cap : blocks);
    void *args[] = {(void *)&ref, (void *)&npx, (void *)&q_lo, (void *)&q_hi, (void *)&ss, (void *)&lohi};
    PCB_CUDA(cudaLaunchCooperativeKernel((const void *)contrast_bounds_kernel, dim3(blocks), dim3(256), args, 0, st));
    return PCB_OK;
}

template <typename T>
static int run_percentile(const T *data, int64_t n, int luma, const double *q_host, int nq, double *result,
                          void *scratch, cudaStream_t st) {
    SelectState *ss = reinterpret_cast<SelectState *>(scratch);
    double q[PS_MAXQ] = {0, 0, 0, 0};
    for (int k = 0; k < nq; ++k) q[k] = q_host[k];
    select_init_kernel<<<1, 256, 0, st>>>(ss, n, nq, q[0], q[1], q[2], q[3]);
    PCB_LAUNCH_OK();
    int blocks = (int)((n + 256 * 8 - 1) / (256 * 8));
    blocks = blocks < 1 ? 1 : (blocks > 148 * 8 ? 148 * 8 : blocks);
    for (int pass = 0; pass < SelKey<T>::PASSES; ++pass)
        select_hist_kernel<T><<<blocks, 256, 0, st>>>(data, n, luma, pass, ss);
    select_finish_kernel<T><<<1, 32, 0, st>>>(ss, n, result);
    PCB_LAUNCH_OK();
    return PCB_OK;
}

}  // namespace pcb
