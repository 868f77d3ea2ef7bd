// (b) Viewpoint gather: aligned micro-image array <-> 4-D sub-aperture array, with the optional central
// crop of each micro image folded in.  Pure copies, bit-exact, dtype preserved.
// Replaces lfp_extractor/lfp_rearranger.py:79-137 and lfp_extractor/lfp_cropper.py:58-62.
//
//   vp[j, i, s, t, c] = aligned[s*M_in + k + j, t*M_in + k + i, c],   k = (M_in - M_out)/2
//
// One CTA owns one aligned row (s, j) (a tile of it when the row exceeds the smem budget): the row is
// read once, fully coalesced, into shared memory and leaves as M_out contiguous runs, one per view i.
#pragma once
#include "common.cuh"

namespace pcb {

template <typename T, int CT>   // CT = compile-time channel count (0 = runtime)
__global__ void __launch_bounds__(256)
viewpoints_kernel(const T *__restrict__ aligned, T *__restrict__ vp, int cols_elems, int M_in, int M_out, int kc,
                  int S, int Tn, int Crt, int tile_t, size_t total_elems) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *row = reinterpret_cast<T *>(smem_raw);
    const int C = CT ? CT : Crt;
    const int s = blockIdx.x, j = blockIdx.y;
    const int t0 = blockIdx.z * tile_t;
    const int nt = min(tile_t, Tn - t0);
    if (nt <= 0) return;

    const T *src = aligned + (size_t)(s * M_in + kc + j) * cols_elems + (size_t)t0 * M_in * C;
    const int n_in = nt * M_in * C;
    // The row is copied with 16-byte asynchronous global -> shared copies (no register staging, the whole tile in
    // flight at once).  Rows start at arbitrary element alignment (row pitch 56 250 B at Illum size): the shared-memory
    // copy starts at the same offset inside a 16-byte line as the source, so the 16-byte interior lines up on both
    // sides; the few head / tail elements are copied by plain loads.
    {
        const int mis = (int)(reinterpret_cast<uintptr_t>(src) & 15);
        row = reinterpret_cast<T *>(smem_raw + mis);
        constexpr int EPL = 16 / (int)sizeof(T);                           // elements per 16-byte line
        const int head = min(n_in, mis ? (16 - mis) / (int)sizeof(T) : 0);
        const int nlines = (n_in - head) / EPL;
        const int tail0 = head + nlines * EPL;
        const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(row + head);
        const T *gbase = src + head;
        for (int k = threadIdx.x; k < nlines; k += 256)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sbase + 16u * k), "l"(gbase + (size_t)k * EPL)
                         : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
        if (threadIdx.x < head) row[threadIdx.x] = __ldg(src + threadIdx.x);
        if (threadIdx.x < n_in - tail0) row[tail0 + threadIdx.x] = __ldg(src + tail0 + threadIdx.x);
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();

    // every thread keeps its (t, c) decomposition and walks the M_out views: no division in the store loop
    const int n_out = nt * C;
    const int mc = M_in * C;
    for (int idx = threadIdx.x; idx < n_out; idx += 256) {
        const int t = idx / C;
        const int c = idx - t * C;
        const T *rp = row + t * mc + kc * C + c;
        T *dst = vp + (((size_t)(j * M_out) * S + s) * Tn + t0) * C + idx;
        const size_t view_stride = (size_t)S * Tn * C;
#pragma unroll 5
        for (int i = 0; i < M_out; ++i) dst[i * view_stride] = rp[i * C];
    }
}

template <typename T, int CT>
__global__ void __launch_bounds__(256)
viewpoints_inverse_kernel(const T *__restrict__ vp, T *__restrict__ aligned, int M, int S, int Tn, int Crt, int tile_t) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *row = reinterpret_cast<T *>(smem_raw);
    const int C = CT ? CT : Crt;
    const int s = blockIdx.x, j = blockIdx.y;
    const int t0 = blockIdx.z * tile_t;
    const int nt = min(tile_t, Tn - t0);
    if (nt <= 0) return;

    const int n_view = nt * C;
    for (int i = 0; i < M; ++i) {
        const T *src = vp + (((size_t)(j * M + i) * S + s) * Tn + t0) * C;
        for (int idx = threadIdx.x; idx < n_view; idx += blockDim.x) {
            const int t = idx / C;
            const int c = idx - t * C;
            row[(t * M + i) * C + c] = __ldg(src + idx);
        }
    }
    __syncthreads();
    T *dst = aligned + (size_t)(s * M + j) * ((size_t)Tn * M * C) + (size_t)t0 * M * C;
    const int n_out = nt * M * C;
    for (int idx = threadIdx.x; idx < n_out; idx += blockDim.x) dst[idx] = row[idx];
}

// micro images of (My_in x Mx_in) pixels -> central (M_out x M_out), margins ky / kx (lfp_cropper.py:34,58-62; the
// globally rectified hexagonal light field has a non-square pitch, e.g. 13 x 15)
template <typename T>
__global__ void __launch_bounds__(256)
crop_kernel(const T *__restrict__ aligned, T *__restrict__ cropped, int cols_elems_in, int My_in, int Mx_in, int M_out,
            int ky, int kx, int LX, int C) {
    // one CTA per output row
    const int orow = blockIdx.x;
    const int ly = orow / M_out, v = orow - ly * M_out;
    const T *src = aligned + (size_t)(ly * My_in + ky + v) * cols_elems_in;
    T *dst = cropped + (size_t)orow * ((size_t)LX * M_out * C);
    const int n = LX * M_out * C;
    const int mc = M_out * C;
    for (int idx = threadIdx.x; idx < n; idx += blockDim.x) {
        const int lx = idx / mc;
        const int rem = idx - lx * mc;   // u*C + c
        dst[idx] = __ldg(src + (lx * Mx_in + kx) * C + rem);
    }
}

static inline int vp_tile(int Tn, int M, int C, int esz, int *tile_t, int *ntile, size_t *smem) {
    const size_t budget = 48 * 1024 - 16;         // default dynamic shared-memory limit minus the alignment slack
    const size_t row_bytes = (size_t)Tn * M * C * esz;
    int nt = (int)((row_bytes + budget - 1) / budget);
    if (nt < 1) nt = 1;
    int tt = (Tn + nt - 1) / nt;
    if ((size_t)tt * M * C * esz > budget) return 1;   // a single micro-image row exceeds the budget
    *tile_t = tt;
    *ntile = (Tn + tt - 1) / tt;
    *smem = (size_t)tt * M * C * esz + 16;     // + one line: the staged row keeps the source's offset inside a 16-byte line
    return 0;
}

template <typename T>
static int launch_viewpoints(const void *aligned, void *vp, int rows, int cols, int C, int M_in, int M_out,
                             cudaStream_t st) {
    const int S = rows / M_in, Tn = cols / M_in, kc = (M_in - M_out) / 2;
    int tile_t, ntile;
    size_t smem;
    if (vp_tile(Tn, M_in, C, sizeof(T), &tile_t, &ntile, &smem))
        return fail(PCB_ERR_UNSUPPORTED, "%s: micro image row too large for shared memory", "pcb_viewpoints");
    dim3 grid(S, M_out, ntile), block(256);
    if (C == 3)
        viewpoints_kernel<T, 3><<<grid, block, smem, st>>>((const T *)aligned, (T *)vp, cols * C, M_in, M_out, kc, S, Tn, C, tile_t,
                                                           (size_t)rows * cols * C);
    else if (C == 1)
        viewpoints_kernel<T, 1><<<grid, block, smem, st>>>((const T *)aligned, (T *)vp, cols * C, M_in, M_out, kc, S, Tn, C, tile_t,
                                                           (size_t)rows * cols * C);
    else
        viewpoints_kernel<T, 0><<<grid, block, smem, st>>>((const T *)aligned, (T *)vp, cols * C, M_in, M_out, kc, S, Tn, C, tile_t,
                                                           (size_t)rows * cols * C);
    PCB_LAUNCH_OK();
    return PCB_OK;
}

template <typename T>
static int launch_viewpoints_inverse(const void *vp, void *aligned, int M, int S, int Tn, int C, cudaStream_t st) {
    int tile_t, ntile;
    size_t smem;
    if (vp_tile(Tn, M, C, sizeof(T), &tile_t, &ntile, &smem))
        return fail(PCB_ERR_UNSUPPORTED, "%s: micro image row too large for shared memory", "pcb_viewpoints_inverse");
    dim3 grid(S, M, ntile), block(256);
    if (C == 3)
        viewpoints_inverse_kernel<T, 3><<<grid, block, smem, st>>>((const T *)vp, (T *)aligned, M, S, Tn, C, tile_t);
    else if (C == 1)
        viewpoints_inverse_kernel<T, 1><<<grid, block, smem, st>>>((const T *)vp, (T *)aligned, M, S, Tn, C, tile_t);
    else
        viewpoints_inverse_kernel<T, 0><<<grid, block, smem, st>>>((const T *)vp, (T *)aligned, M, S, Tn, C, tile_t);
    PCB_LAUNCH_OK();
    return PCB_OK;
}

template <typename T>
static int launch_crop(const void *aligned, void *cropped, int rows, int cols, int C, int My_in, int Mx_in, int M_out,
                       cudaStream_t st) {
    const int LY = rows / My_in, LX = cols / Mx_in;
    crop_kernel<T><<<LY * M_out, 256, 0, st>>>((const T *)aligned, (T *)cropped, cols * C, My_in, Mx_in, M_out,
                                              (My_in - M_out) / 2, (Mx_in - M_out) / 2, LX, C);
    PCB_LAUNCH_OK();
    return PCB_OK;
}

}  // namespace pcb
