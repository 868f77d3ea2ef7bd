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

// Inverse: the M view runs of one aligned row (s, j) are requested together with 16-byte asynchronous copies (each run
// keeps its source's offset inside a 16-byte line, see viewpoints_kernel), then interleaved out of shared memory into
// one contiguous row.  run_pitch: bytes reserved per run (>= run bytes + 32, multiple of 16).
template <typename T, int CT>
__global__ void __launch_bounds__(256)
viewpoints_inverse_kernel(const T *__restrict__ vp, T *__restrict__ aligned, int M, int S, int Tn, int Crt, int tile_t,
                          int run_pitch) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int s_mis[64];                              // per view: byte offset of its run inside its slot
    const int C = CT ? CT : Crt;
    const int s = blockIdx.x, j = blockIdx.y;
    const int t0 = blockIdx.z * tile_t;
    const int nt = min(tile_t, Tn - t0);
    if (nt <= 0) return;

    const int n_view = nt * C;
    constexpr int EPL = 16 / (int)sizeof(T);
    for (int i = 0; i < M; ++i) {
        const T *src = vp + (((size_t)(j * M + i) * S + s) * Tn + t0) * C;
        const int mis = (int)(reinterpret_cast<uintptr_t>(src) & 15);
        T *run = reinterpret_cast<T *>(smem_raw + (size_t)i * run_pitch + mis);
        const int head = min(n_view, mis ? (16 - mis) / (int)sizeof(T) : 0);
        const int nlines = (n_view - head) / EPL;
        const int tail0 = head + nlines * EPL;
        const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(run + head);
        const T *gbase = src + head;
        for (int k = threadIdx.x; k < nlines; k += 256)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sbase + 16u * k), "l"(gbase + (size_t)k * EPL)
                         : "memory");
        if (threadIdx.x < head) run[threadIdx.x] = __ldg(src + threadIdx.x);
        if (threadIdx.x < n_view - tail0) run[tail0 + threadIdx.x] = __ldg(src + tail0 + threadIdx.x);
        if (threadIdx.x == 0) s_mis[i] = mis;
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    T *dst = aligned + (size_t)(s * M + j) * ((size_t)Tn * M * C) + (size_t)t0 * M * C;
    const int n_out = nt * M * C;
    const int mc = M * C;
    if (mc <= 256) {
        // a thread keeps its (view, channel) and walks the lens columns: no division in the copy loop; the
        // (256 / mc) * mc active threads of a pass write one contiguous stretch
        const int tq_n = 256 / mc;
        const int tq = threadIdx.x / mc, rem = threadIdx.x - tq * mc;
        if (tq < tq_n) {
            const int i = rem / C, c = rem - i * C;
            const T *run = reinterpret_cast<const T *>(smem_raw + (size_t)i * run_pitch + s_mis[i]) + c;
            for (int t = tq; t < nt; t += tq_n) dst[t * mc + rem] = run[t * C];
        }
    } else {
        for (int idx = threadIdx.x; idx < n_out; idx += 256) {
            const int t = idx / mc;
            const int rem = idx - t * mc;
            const int i = rem / C, c = rem - i * C;
            dst[idx] = reinterpret_cast<const T *>(smem_raw + (size_t)i * run_pitch + s_mis[i])[t * C + c];
        }
    }
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
    if (mc <= (int)blockDim.x) {
        // a thread keeps its position (u, c) inside the micro image and walks the lens columns: no division per element
        const int lq_n = blockDim.x / mc;
        const int lq = threadIdx.x / mc, rem = threadIdx.x - lq * mc;
        if (lq < lq_n) {
            const T *sp = src + kx * C + rem;
            constexpr int U = 4;
            int lx = lq;
            for (; lx + (U - 1) * lq_n < LX; lx += U * lq_n) {            // U independent loads in flight
                T v[U];
#pragma unroll
                for (int u = 0; u < U; ++u) v[u] = __ldg(sp + (size_t)(lx + u * lq_n) * Mx_in * C);
#pragma unroll
                for (int u = 0; u < U; ++u) dst[(size_t)(lx + u * lq_n) * mc + rem] = v[u];
            }
            for (; lx < LX; lx += lq_n) dst[(size_t)lx * mc + rem] = __ldg(sp + (size_t)lx * Mx_in * C);
        }
    } else {
        for (int idx = threadIdx.x; idx < n; idx += blockDim.x) {
            const int lx = idx / mc;
            const int rem = idx - lx * mc;   // u*C + c
            dst[idx] = __ldg(src + (lx * Mx_in + kx) * C + rem);
        }
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
    if (M > 64) return fail(PCB_ERR_UNSUPPORTED, "%s: more than 64 x 64 views", "pcb_viewpoints_inverse");
    // tiles of tile_t lens columns: M runs of tile_t * C elements, each in a slot with 32 bytes of alignment slack
    const size_t budget = 48 * 1024;
    int ntile = 1, tile_t = Tn;
    size_t run_pitch = 0;
    for (;; ++ntile) {
        tile_t = (Tn + ntile - 1) / ntile;
        run_pitch = (((size_t)tile_t * C * sizeof(T) + 32) + 15) & ~(size_t)15;
        if ((size_t)M * run_pitch <= budget) break;
        if (tile_t == 1) return fail(PCB_ERR_UNSUPPORTED, "%s: micro image row too large for shared memory", "pcb_viewpoints_inverse");
    }
    ntile = (Tn + tile_t - 1) / tile_t;
    const size_t smem = (size_t)M * run_pitch;
    dim3 grid(S, M, ntile), block(256);
    if (C == 3)
        viewpoints_inverse_kernel<T, 3><<<grid, block, smem, st>>>((const T *)vp, (T *)aligned, M, S, Tn, C, tile_t, (int)run_pitch);
    else if (C == 1)
        viewpoints_inverse_kernel<T, 1><<<grid, block, smem, st>>>((const T *)vp, (T *)aligned, M, S, Tn, C, tile_t, (int)run_pitch);
    else
        viewpoints_inverse_kernel<T, 0><<<grid, block, smem, st>>>((const T *)vp, (T *)aligned, M, S, Tn, C, tile_t, (int)run_pitch);
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
