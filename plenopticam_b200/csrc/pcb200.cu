// libpcb200.so — C-ABI entry points (include/plenopticam_b200.h) over the sm_100a kernels.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC
#include "common.cuh"
#include "post.cuh"
#include "devignette.cuh"
#include "scheimpflug.cuh"
#include "cfa.cuh"
#include "unpack.cuh"
#include "refocus.cuh"
#include "refocus_rows.cuh"
#include "refocus_px.cuh"
#include "refocus_pxp.cuh"
#include "refine.cuh"
#include "resample.cuh"
#include "resample_col.cuh"
#include "resample_pipe.cuh"
#include "global_resample.cuh"
#include "rotate.cuh"
#include "viewpoints.cuh"

namespace pcb {
thread_local char g_err[512] = "";
thread_local cudaEvent_t g_kev[2] = {nullptr, nullptr};
}

using namespace pcb;

static inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }

static inline int grid_for(int64_t n, int per_block) {
    int64_t b = (n + per_block - 1) / per_block;
    const int64_t cap = 148 * 16;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

template <int MODE>
static int dispatch_resample(const ResampleArgs &a, int raw_dtype, float *out_f32, uint16_t *out_u16,
                             pcb_minmax_t *mm_rw, const pcb_minmax_t *mm_ro, cudaStream_t st) {
    if (resample_pipe_applies(a, raw_dtype)) {
        if (raw_dtype == PCB_F32) return launch_resample_pipe<MODE, float>(a, out_f32, out_u16, mm_rw, mm_ro, st);
        // the min/max pass interpolates a fraction of its tiles only: it reads the uint16 rows in place (conversion in the
        // column loop) and spends the shared memory of the float tile on a ring of four slots (0.171 -> 0.160 ms); the passes
        // that interpolate every tile keep the converted tile (in place they are slower: 0.302 -> 0.332 ms)
        const char *e = getenv("PCB_RESAMPLE_INPLACE");
        const bool inplace = e ? atoi(e) != 0 : MODE == RS_MINMAX;
        if ((a.W & 7) == 0 && inplace)
            return launch_resample_pipe<MODE, RpkU16InPlace>(a, out_f32, out_u16, mm_rw, mm_ro, st);
        return launch_resample_pipe<MODE, uint16_t>(a, out_f32, out_u16, mm_rw, mm_ro, st);
    }
    if (resample_col_applies(a, raw_dtype)) return launch_resample_col<MODE>(a, out_f32, out_u16, mm_rw, mm_ro, st);
    return launch_resample<MODE>(a, raw_dtype, out_f32, out_u16, mm_rw, mm_ro, st);
}

template <typename TIn>
static int devig_divide_out(const void *lfp, int64_t npx, int C, const double *wn, void *out, int out_dtype,
                            cudaStream_t st) {
    const int g = devig_grid(npx * C);
    if (out_dtype == PCB_F64)
        devig_divide_kernel<TIn, double><<<g, 256, 0, st>>>((const TIn *)lfp, npx, C, wn, (double *)out);
    else if (out_dtype == PCB_F32)
        devig_divide_kernel<TIn, float><<<g, 256, 0, st>>>((const TIn *)lfp, npx, C, wn, (float *)out);
    else
        return fail(PCB_ERR_INVALID, "%s: output dtype %lld (float32 / float64)", "pcb_devig_divide", out_dtype);
    PCB_LAUNCH_OK();
    return PCB_OK;
}

extern "C" {

int pcb_version(void) { return PCB_ABI_VERSION; }

const char *pcb_last_error(void) { return g_err; }

int pcb_debug_kernel_events(void *ev_begin, void *ev_end) {
    g_kev[0] = reinterpret_cast<cudaEvent_t>(ev_begin);
    g_kev[1] = reinterpret_cast<cudaEvent_t>(ev_end);
    return PCB_OK;
}

float pcb_key_to_float(uint32_t key) { return key_to_float(key); }

int pcb_minmax_reset(pcb_minmax_t *mm, void *stream) {
    PCB_REQUIRE(mm != nullptr, "mm != NULL");
    minmax_reset_kernel<<<1, 1, 0, as_stream(stream)>>>(mm);
    PCB_LAUNCH_OK();
    return PCB_OK;
}

int pcb_minmax_f32(const float *data, int64_t n, pcb_minmax_t *mm, void *stream) {
    PCB_REQUIRE(data != nullptr && mm != nullptr && n > 0, "data, mm, n > 0");
    minmax_f32_kernel<<<grid_for(n, 256 * 8), 256, 0, as_stream(stream)>>>(data, n, mm);
    PCB_LAUNCH_OK();
    return PCB_OK;
}

static int make_resample_args(ResampleArgs &a, const void *raw, int H, int W, int C, const pcb_lens_t *lens, int LY,
                              int LX, int M, int pattern, const pcb_hexcol_t *hexcol, int Xs, int hex_odd, int flip) {
    PCB_REQUIRE(raw != nullptr && lens != nullptr, "raw, lens");
    PCB_REQUIRE(H > 0 && W > 0 && C > 0 && LY > 0 && LX > 0 && M > 0 && Xs > 0, "positive sizes");
    PCB_REQUIRE(LY <= 65535, "LY <= 65535");
    PCB_REQUIRE(pattern == PCB_REC || pattern == PCB_HEX, "pattern");
    PCB_REQUIRE(pattern == PCB_HEX ? hexcol != nullptr : Xs == LX, "hexcol for hex / Xs == LX for rec");
    PCB_REQUIRE((int64_t)LY * M * (int64_t)Xs * M * C < ((int64_t)1 << 40), "aligned size");
    PCB_REQUIRE((int64_t)Xs * M * C < ((int64_t)1 << 31), "aligned row length");
    a.raw = raw; a.H = H; a.W = W; a.C = C;
    a.lens = lens; a.LY = LY; a.LX = LX; a.M = M;
    a.hexcol = pattern == PCB_HEX ? hexcol : nullptr;
    a.Xs = Xs; a.hex_odd = hex_odd ? 1 : 0; a.flip = flip ? 1 : 0;
    a.ncols = Xs * M * C;
    a.KT = 0; a.cap_floats = 0; a.skip = 0;
    return PCB_OK;
}

int pcb_resample_minmax(const void *raw, int raw_dtype, int H, int W, int C, const pcb_lens_t *lens, int LY, int LX,
                        int M, int pattern, const pcb_hexcol_t *hexcol, int Xs, int hex_odd, int flip,
                        pcb_minmax_t *mm, void *stream) {
    ResampleArgs a;
    int rc = make_resample_args(a, raw, H, W, C, lens, LY, LX, M, pattern, hexcol, Xs, hex_odd, flip);
    if (rc) return rc;
    PCB_REQUIRE(mm != nullptr, "mm");
    return dispatch_resample<RS_MINMAX>(a, raw_dtype, nullptr, nullptr, mm, nullptr, as_stream(stream));
}

int pcb_resample_f32(const void *raw, int raw_dtype, int H, int W, int C, const pcb_lens_t *lens, int LY, int LX,
                     int M, int pattern, const pcb_hexcol_t *hexcol, int Xs, int hex_odd, int flip, float *aligned,
                     pcb_minmax_t *mm, void *stream) {
    ResampleArgs a;
    int rc = make_resample_args(a, raw, H, W, C, lens, LY, LX, M, pattern, hexcol, Xs, hex_odd, flip);
    if (rc) return rc;
    PCB_REQUIRE(aligned != nullptr, "aligned");
    return dispatch_resample<RS_F32>(a, raw_dtype, aligned, nullptr, mm, nullptr, as_stream(stream));
}

int pcb_resample_u16(const void *raw, int raw_dtype, int H, int W, int C, const pcb_lens_t *lens, int LY, int LX,
                     int M, int pattern, const pcb_hexcol_t *hexcol, int Xs, int hex_odd, int flip,
                     const pcb_minmax_t *mm, uint16_t *aligned, void *stream) {
    ResampleArgs a;
    int rc = make_resample_args(a, raw, H, W, C, lens, LY, LX, M, pattern, hexcol, Xs, hex_odd, flip);
    if (rc) return rc;
    PCB_REQUIRE(aligned != nullptr && mm != nullptr, "aligned, mm");
    return dispatch_resample<RS_U16>(a, raw_dtype, nullptr, aligned, nullptr, mm, as_stream(stream));
}

int pcb_quantize_u16(const float *in, int64_t n, const pcb_minmax_t *mm, uint16_t *out, void *stream) {
    PCB_REQUIRE(in != nullptr && out != nullptr && mm != nullptr && n > 0, "in, out, mm, n > 0");
    quantize_u16_kernel<<<grid_for(n, 256 * 8), 256, 0, as_stream(stream)>>>(in, n, mm, out);
    PCB_LAUNCH_OK();
    return PCB_OK;
}

int pcb_quantize_u8(const float *in, int64_t n, const pcb_minmax_t *mm, uint8_t *out, void *stream) {
    PCB_REQUIRE(in != nullptr && out != nullptr && mm != nullptr && n > 0, "in, out, mm, n > 0");
    quantize_u8_kernel<<<grid_for(n, 256 * 8), 256, 0, as_stream(stream)>>>(in, n, mm, out);
    PCB_LAUNCH_OK();
    return PCB_OK;
}

int pcb_spline_prefilter(const void *img, int dtype, int H, int W, int C, float *tmp, float *coef, void *stream) {
    PCB_REQUIRE(img != nullptr && tmp != nullptr && coef != nullptr, "img, tmp, coef");
    PCB_REQUIRE(H > 0 && W > 0 && C > 0 && H <= 65535 * PF_ROWS, "sizes");
    switch (dtype) {
    case PCB_U16: return launch_prefilter<uint16_t>(img, H, W, C, tmp, coef, as_stream(stream));
    case PCB_F32: return launch_prefilter<float>(img, H, W, C, tmp, coef, as_stream(stream));
    case PCB_U8:  return launch_prefilter<uint8_t>(img, H, W, C, tmp, coef, as_stream(stream));
    }
    return fail(PCB_ERR_INVALID, "%s: unsupported dtype %lld", __func__, dtype);
}

int pcb_rotate_sample(const float *coef, int H, int W, int C, double rad, float *out, void *stream) {
    PCB_REQUIRE(coef != nullptr && out != nullptr && H > 0 && W > 0 && C > 0 && H <= 65535, "coef, out, sizes");
    dim3 grid((W + 255) / 256, H), block(256);
    rotate_sample_kernel<<<grid, block, 0, as_stream(stream)>>>(coef, out, H, W, C, cos(rad), sin(rad),
                                                                (H - 1) / 2.0, (W - 1) / 2.0);
    PCB_LAUNCH_OK();
    return PCB_OK;
}

#define PCB_DISPATCH_COPY_DTYPE(dtype, FN, ...)                                           \
    switch (dtype) {                                                                      \
    case PCB_U16: return FN<uint16_t>(__VA_ARGS__);                                       \
    case PCB_F32: return FN<float>(__VA_ARGS__);                                          \
    case PCB_U8:  return FN<uint8_t>(__VA_ARGS__);                                        \
    default: return fail(PCB_ERR_INVALID, "%s: unsupported dtype %lld", __func__, dtype); \
    }

int pcb_viewpoints(const void *aligned, int dtype, int rows, int cols, int C, int M_in, int M_out, void *vp,
                   void *stream) {
    PCB_REQUIRE(aligned != nullptr && vp != nullptr, "aligned, vp");
    PCB_REQUIRE(rows > 0 && cols > 0 && C > 0 && M_in > 0 && M_out > 0 && M_out <= M_in, "sizes, M_out <= M_in");
    PCB_REQUIRE(rows / M_in > 0 && cols / M_in > 0, "at least one micro image");
    PCB_REQUIRE((M_in - M_out) % 2 == 0, "M_in - M_out even (lfp_cropper.py:34,60-62)");
    PCB_DISPATCH_COPY_DTYPE(dtype, launch_viewpoints, aligned, vp, rows, cols, C, M_in, M_out, as_stream(stream));
}

int pcb_viewpoints_inverse(const void *vp, int dtype, int M, int S, int T, int C, void *aligned, void *stream) {
    PCB_REQUIRE(aligned != nullptr && vp != nullptr, "aligned, vp");
    PCB_REQUIRE(M > 0 && S > 0 && T > 0 && C > 0 && M <= 65535, "sizes");
    PCB_DISPATCH_COPY_DTYPE(dtype, launch_viewpoints_inverse, vp, aligned, M, S, T, C, as_stream(stream));
}

int pcb_crop_micro_images(const void *aligned, int dtype, int rows, int cols, int C, int M_in, int M_out,
                          void *cropped, void *stream) {
    PCB_REQUIRE(aligned != nullptr && cropped != nullptr, "aligned, cropped");
    PCB_REQUIRE(rows > 0 && cols > 0 && C > 0 && M_in > 0 && M_out > 0 && M_out <= M_in, "sizes, M_out <= M_in");
    PCB_REQUIRE((M_in - M_out) % 2 == 0, "M_in - M_out even (lfp_cropper.py:34,60-62)");
    PCB_DISPATCH_COPY_DTYPE(dtype, launch_crop, aligned, cropped, rows, cols, C, M_in, M_in, M_out, as_stream(stream));
}

int pcb_crop_micro_images_yx(const void *aligned, int dtype, int rows, int cols, int C, int My_in, int Mx_in, int M_out,
                             void *cropped, void *stream) {
    PCB_REQUIRE(aligned != nullptr && cropped != nullptr, "aligned, cropped");
    PCB_REQUIRE(rows > 0 && cols > 0 && C > 0 && My_in > 0 && Mx_in > 0 && M_out > 0 && M_out <= My_in && M_out <= Mx_in,
                "sizes, M_out <= My_in, Mx_in");
    PCB_REQUIRE((My_in - M_out) % 2 == 0 && (Mx_in - M_out) % 2 == 0, "even margins (lfp_cropper.py:34,60-62)");
    PCB_REQUIRE(rows / My_in > 0 && cols / Mx_in > 0, "at least one micro image");
    PCB_DISPATCH_COPY_DTYPE(dtype, launch_crop, aligned, cropped, rows, cols, C, My_in, Mx_in, M_out, as_stream(stream));
}

// Planning shared by pcb_refocus_int_scratch_bytes and pcb_refocus_int (host-only, deterministic).
struct RefocusPlan {
    PxpPlan pxp;     // ... persistent TMA-fed form of it when the whole row and the next tile's raw rows fit one SM
    PxPlan px;       // uint16 RGB, M <= 16: packed-pixel row-stationary kernel
    RowsPlan rows;   // uint16, any C, M <= 16: element-pair row-stationary kernel
};

static int refocus_force_kernel() {   // A/B knob: PCB_REFOCUS_KERNEL=rows|direct (default: best available)
    const char *e = getenv("PCB_REFOCUS_KERNEL");
    if (!e) return 0;
    return !strcmp(e, "rows") ? 1 : (!strcmp(e, "direct") ? 2 : (!strcmp(e, "px") ? 3 : 0));
}

static RefocusPlan refocus_plan(int dtype, int M, int S, int T, int C, const int32_t *a_list_host, int N,
                                const uint8_t *view_zeroed_host, const void *vp = nullptr, int Ma = 0) {
    RefocusPlan none;
    memset(&none, 0, sizeof(none));
    if (dtype != PCB_U16 || M > 16 || refocus_force_kernel() == 2) return none;
    int max_abs_a = 0;
    for (int n = 0; n < N; ++n) {
        const long long v = a_list_host[n] < 0 ? -(long long)a_list_host[n] : a_list_host[n];
        if (v > 1000000) return none;
        if (v > max_abs_a) max_abs_a = (int)v;
    }
    int max_act = 0, tot_act = 0;
    for (int j = 0; j < M; ++j) {
        int n = 0;
        for (int i = 0; i < M; ++i) n += view_zeroed_host[j * M + i] ? 0 : 1;
        if (n > max_act) max_act = n;
        tot_act += n;
    }
    if (max_act == 0) return none;
    RefocusPlan pl = none;
    if (C == 3 && refocus_force_kernel() != 1) pl.px = plan_refocus_px(M, S, T, N, max_abs_a, max_act, tot_act, view_zeroed_host);
    if (pl.px.ok && vp != nullptr && refocus_force_kernel() != 3)
        pl.pxp = plan_refocus_pxp(pl.px, M, S, T, N, max_abs_a, view_zeroed_host, vp, Ma);
    if (!pl.px.ok) pl.rows = plan_refocus_rows(M, S, T, C, N, max_abs_a, max_act, tot_act);
    return pl;
}

static inline size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

int64_t pcb_refocus_int_scratch_bytes(int dtype, int M, int S, int T, int C, const int32_t *a_list_host, int N,
                                      const uint8_t *view_zeroed_host) {
    if (a_list_host == nullptr || view_zeroed_host == nullptr || M <= 0 || S <= 0 || T <= 0 || C <= 0 || N <= 0)
        return -1;
    const RefocusPlan pl = refocus_plan(dtype, M, S, T, C, a_list_host, N, view_zeroed_host);
    if (pl.px.ok) return (int64_t)pl.px.scratch_bytes();
    if (pl.rows.ok) return (int64_t)(pl.rows.acc_bytes + pl.rows.edge_bytes);
    return (int64_t)(align256((size_t)N * 4) + align256((size_t)M * M));
}

int pcb_refocus_int(const void *vp, int dtype, int M, int S, int T, int C, const int32_t *a_list_host, int N,
                    const uint8_t *view_zeroed_host, void *scratch, int64_t scratch_bytes, float *out,
                    pcb_minmax_t *mm, void *stream) {
    PCB_REQUIRE(vp != nullptr && a_list_host != nullptr && view_zeroed_host != nullptr && out != nullptr &&
                scratch != nullptr, "pointers");
    PCB_REQUIRE(M > 0 && S > 0 && T > 0 && C > 0 && N > 0 && N <= 65535, "sizes");
    PCB_REQUIRE(M % 2 == 1, "odd M (even M changes the slice shape, lfp_shiftandsum.py:123-124)");
    PCB_REQUIRE(scratch_bytes >= pcb_refocus_int_scratch_bytes(dtype, M, S, T, C, a_list_host, N, view_zeroed_host),
                "scratch_bytes >= pcb_refocus_int_scratch_bytes(...)");
    cudaStream_t st = as_stream(stream);
    const RefocusPlan pl = refocus_plan(dtype, M, S, T, C, a_list_host, N, view_zeroed_host, vp);
    if (pl.pxp.ok)
        return launch_refocus_pxp(pl.px, pl.pxp, (const uint16_t *)vp, a_list_host, view_zeroed_host, scratch, out, mm, st);
    if (pl.px.ok)
        return launch_refocus_px(pl.px, (const uint16_t *)vp, a_list_host, view_zeroed_host, scratch, out, mm, st);
    if (pl.rows.ok)
        return launch_refocus_rows(pl.rows, (const uint16_t *)vp, a_list_host, view_zeroed_host, scratch, out, mm, st);
    // direct form: control arrays staged at the head of the scratch buffer
    int32_t *a_dev = reinterpret_cast<int32_t *>(scratch);
    uint8_t *z_dev = reinterpret_cast<uint8_t *>(scratch) + align256((size_t)N * 4);
    PCB_CUDA(cudaMemcpyAsync(a_dev, a_list_host, (size_t)N * 4, cudaMemcpyHostToDevice, st));
    PCB_CUDA(cudaMemcpyAsync(z_dev, view_zeroed_host, (size_t)M * M, cudaMemcpyHostToDevice, st));
    switch (dtype) {
    case PCB_U16: return launch_refocus_int<uint16_t>(vp, M, S, T, C, a_dev, N, z_dev, out, mm, st);
    case PCB_F32: return launch_refocus_int<float>(vp, M, S, T, C, a_dev, N, z_dev, out, mm, st);
    case PCB_U8:  return launch_refocus_int<uint8_t>(vp, M, S, T, C, a_dev, N, z_dev, out, mm, st);
    }
    return fail(PCB_ERR_INVALID, "%s: unsupported dtype %lld", __func__, dtype);
}

// Shift-and-sum with the refocuser's colour management fused into the epilogue: the linear stack never reaches HBM.
int pcb_refocus_int_srgb(const void *vp, int dtype, int M, int S, int T, int C, const int32_t *a_list_host, int N,
                         const uint8_t *view_zeroed_host, void *scratch, int64_t scratch_bytes, double q_lo, double q_hi,
                         float *ref_slice, double *lohi, void *pscratch, float *out, pcb_minmax_t *mm, void *stream) {
    PCB_REQUIRE(vp != nullptr && a_list_host != nullptr && view_zeroed_host != nullptr && out != nullptr &&
                scratch != nullptr && ref_slice != nullptr && lohi != nullptr && pscratch != nullptr && mm != nullptr,
                "pointers");
    PCB_REQUIRE(M > 0 && S > 0 && T > 0 && C > 0 && N > 0 && N <= 65535, "sizes");
    PCB_REQUIRE(M % 2 == 1, "odd M (even M changes the slice shape, lfp_shiftandsum.py:123-124)");
    PCB_REQUIRE(scratch_bytes >= pcb_refocus_int_scratch_bytes(dtype, M, S, T, C, a_list_host, N, view_zeroed_host),
                "scratch_bytes >= pcb_refocus_int_scratch_bytes(...)");
    const RefocusPlan pl = refocus_plan(dtype, M, S, T, C, a_list_host, N, view_zeroed_host, vp);
    if (!pl.pxp.ok)
        return fail(PCB_ERR_UNSUPPORTED, "%s: the fused epilogue exists for the packed-pixel kernel only (uint16 RGB, "
                    "M <= 16, row <= 2048 px); use pcb_refocus_int + pcb_contrast_bounds + pcb_contrast_srgb", __func__);
    PxpFusedPost post;
    post.q_lo = q_lo; post.q_hi = q_hi; post.ref_slice = ref_slice; post.lohi = lohi; post.pscratch = pscratch;
    return launch_refocus_pxp(pl.px, pl.pxp, (const uint16_t *)vp, a_list_host, view_zeroed_host, scratch, out, mm,
                              as_stream(stream), &post);
}

// The same two entry points reading the ALIGNED micro-image array instead of the 4-D light field: the viewpoint gather
// (and LfpCropper's central crop when M < Ma) is folded into the staging of the refocus tiles, so a pipeline that only
// wants the stack never materialises vp[j,i,s,t,c].  Offered by the persistent packed-pixel kernel only.
static int refocus_aligned_common(const void *aligned, int dtype, int Ma, int M, int S, int T, int C,
                                  const int32_t *a_list_host, int N, const uint8_t *view_zeroed_host, void *scratch,
                                  int64_t scratch_bytes, float *out, pcb_minmax_t *mm, const PxpFusedPost *post,
                                  int scratch_flags, cudaStream_t st, const char *who) {
    PCB_REQUIRE(aligned != nullptr && a_list_host != nullptr && view_zeroed_host != nullptr && out != nullptr &&
                scratch != nullptr, "pointers");
    PCB_REQUIRE(M > 0 && S > 0 && T > 0 && C > 0 && N > 0 && N <= 65535, "sizes");
    PCB_REQUIRE(M % 2 == 1, "odd M (even M changes the slice shape, lfp_shiftandsum.py:123-124)");
    PCB_REQUIRE(Ma >= M && (Ma - M) % 2 == 0, "aligned pitch >= M with an even crop margin (lfp_cropper.py:34)");
    PCB_REQUIRE((scratch_flags & ~(PCB_REFOCUS_SCRATCH_ZEROED | PCB_REFOCUS_LEAVE_ZEROED)) == 0, "scratch_flags");
    PCB_REQUIRE(scratch_bytes >= pcb_refocus_int_scratch_bytes(dtype, M, S, T, C, a_list_host, N, view_zeroed_host),
                "scratch_bytes >= pcb_refocus_int_scratch_bytes(...)");
    const RefocusPlan pl = refocus_plan(dtype, M, S, T, C, a_list_host, N, view_zeroed_host, aligned, Ma);
    if (!pl.pxp.ok)
        return fail(PCB_ERR_UNSUPPORTED, "%s: reading the aligned image directly exists for the packed-pixel kernel only "
                    "(uint16 RGB, M <= 16, row <= 2048 px, tiles that fit one SM); run pcb_viewpoints first", who);
    return launch_refocus_pxp(pl.px, pl.pxp, (const uint16_t *)aligned, a_list_host, view_zeroed_host, scratch, out, mm,
                              st, post, scratch_flags);
}

int pcb_refocus_int_aligned(const void *aligned, int dtype, int Ma, int M, int S, int T, int C,
                            const int32_t *a_list_host, int N, const uint8_t *view_zeroed_host, void *scratch,
                            int64_t scratch_bytes, float *out, pcb_minmax_t *mm, int scratch_flags, void *stream) {
    return refocus_aligned_common(aligned, dtype, Ma, M, S, T, C, a_list_host, N, view_zeroed_host, scratch,
                                  scratch_bytes, out, mm, nullptr, scratch_flags, as_stream(stream), __func__);
}

int pcb_refocus_int_srgb_aligned(const void *aligned, int dtype, int Ma, int M, int S, int T, int C,
                                 const int32_t *a_list_host, int N, const uint8_t *view_zeroed_host, void *scratch,
                                 int64_t scratch_bytes, double q_lo, double q_hi, float *ref_slice, double *lohi,
                                 void *pscratch, float *out, pcb_minmax_t *mm, int scratch_flags, void *stream) {
    PCB_REQUIRE(ref_slice != nullptr && lohi != nullptr && pscratch != nullptr && mm != nullptr, "pointers");
    PxpFusedPost post;
    post.q_lo = q_lo; post.q_hi = q_hi; post.ref_slice = ref_slice; post.lohi = lohi; post.pscratch = pscratch;
    return refocus_aligned_common(aligned, dtype, Ma, M, S, T, C, a_list_host, N, view_zeroed_host, scratch,
                                  scratch_bytes, out, mm, &post, scratch_flags, as_stream(stream), __func__);
}

// The direct (slice-stationary) form only; kept addressable for A/B measurements and as the general path.
int pcb_refocus_int_direct(const void *vp, int dtype, int M, int S, int T, int C, const int32_t *a_list_dev, int N,
                           const uint8_t *view_zeroed_dev, float *out, pcb_minmax_t *mm, void *stream) {
    PCB_REQUIRE(vp != nullptr && a_list_dev != nullptr && view_zeroed_dev != nullptr && out != nullptr, "pointers");
    PCB_REQUIRE(M > 0 && S > 0 && T > 0 && C > 0 && N > 0 && N <= 65535 && M % 2 == 1, "sizes, odd M");
    cudaStream_t st = as_stream(stream);
    switch (dtype) {
    case PCB_U16: return launch_refocus_int<uint16_t>(vp, M, S, T, C, a_list_dev, N, view_zeroed_dev, out, mm, st);
    case PCB_F32: return launch_refocus_int<float>(vp, M, S, T, C, a_list_dev, N, view_zeroed_dev, out, mm, st);
    case PCB_U8:  return launch_refocus_int<uint8_t>(vp, M, S, T, C, a_list_dev, N, view_zeroed_dev, out, mm, st);
    }
    return fail(PCB_ERR_INVALID, "%s: unsupported dtype %lld", __func__, dtype);
}

// LfpShiftAndSum.refo_from_scratch (lfp_shiftandsum.py:141-221): zero outside the light field, every view.
int pcb_refocus_int_zero(const void *vp, int dtype, int M, int S, int T, int C, const int32_t *a_list_dev, int N,
                         const uint8_t *view_zeroed_dev, float *out, pcb_minmax_t *mm, void *stream) {
    PCB_REQUIRE(vp != nullptr && a_list_dev != nullptr && view_zeroed_dev != nullptr && out != nullptr, "pointers");
    PCB_REQUIRE(M > 0 && S > 0 && T > 0 && C > 0 && N > 0 && N <= 65535 && M % 2 == 1, "sizes, odd M");
    cudaStream_t st = as_stream(stream);
    switch (dtype) {
    case PCB_U16: return launch_refocus_int<uint16_t>(vp, M, S, T, C, a_list_dev, N, view_zeroed_dev, out, mm, st, true);
    case PCB_F32: return launch_refocus_int<float>(vp, M, S, T, C, a_list_dev, N, view_zeroed_dev, out, mm, st, true);
    case PCB_U8:  return launch_refocus_int<uint8_t>(vp, M, S, T, C, a_list_dev, N, view_zeroed_dev, out, mm, st, true);
    }
    return fail(PCB_ERR_INVALID, "%s: unsupported dtype %lld", __func__, dtype);
}

int pcb_refi_prefilter(const void *vp, int dtype, int M, int S, int T, int C, const float *py_vals,
                       const int32_t *py_first, int Kpy, const float *px_vals, const int32_t *px_first, int Kpx,
                       const uint8_t *view_zeroed, float *tmp, float *coef, void *stream) {
    PCB_REQUIRE(vp && py_vals && py_first && px_vals && px_first && view_zeroed && tmp && coef, "pointers");
    PCB_REQUIRE(M > 0 && M * M <= 65535 && S >= 4 && T >= 4 && S <= 65535 && Kpy > 0 && Kpx > 0, "sizes");
    PCB_REQUIRE(C == 3, "3 colour channels (lfp_shiftandsum.py:97)");
    cudaStream_t st = as_stream(stream);
    switch (dtype) {
    case PCB_U16: return launch_refi_prefilter<uint16_t>(vp, M, S, T, py_vals, py_first, Kpy, px_vals, px_first, Kpx, view_zeroed, tmp, coef, st);
    case PCB_F32: return launch_refi_prefilter<float>(vp, M, S, T, py_vals, py_first, Kpy, px_vals, px_first, Kpx, view_zeroed, tmp, coef, st);
    case PCB_U8:  return launch_refi_prefilter<uint8_t>(vp, M, S, T, py_vals, py_first, Kpy, px_vals, px_first, Kpx, view_zeroed, tmp, coef, st);
    }
    return fail(PCB_ERR_INVALID, "%s: unsupported dtype %lld", __func__, dtype);
}

// Same with the interior of both prefilter operators described as one symmetric filter (host arrays of 29 taps; rows /
// columns lo..hi equal it exactly, sample offset c): those rows run from the kernel-parameter bank with a register window.
int pcb_refi_prefilter_toeplitz(const void *vp, int dtype, int M, int S, int T, int C, const float *py_vals,
                                const int32_t *py_first, int Kpy, const float *px_vals, const int32_t *px_first, int Kpx,
                                const uint8_t *view_zeroed, float *tmp, float *coef, const float *hx_host, int ntx, int cx,
                                int x_lo, int x_hi, const float *hy_host, int nty, int cyy, int y_lo, int y_hi,
                                const pcb_refi_rows_t *rows_host, void *stream) {
    PCB_REQUIRE(vp && py_vals && py_first && px_vals && px_first && view_zeroed && tmp && coef, "pointers");
    PCB_REQUIRE(M > 0 && M * M <= 65535 && S >= 4 && T >= 4 && S <= 65535 && Kpy > 0 && Kpx > 0, "sizes");
    PCB_REQUIRE(C == 3, "3 colour channels (lfp_shiftandsum.py:97)");
    RpTaps tx, ty;
    memset(&tx, 0, sizeof(tx)); memset(&ty, 0, sizeof(ty));
    tx.lo = 1; tx.hi = 0; ty.lo = 1; ty.hi = 0;
    if (hx_host && ntx == RPT_K && cx >= 0 && x_lo - cx >= 0 && x_hi - cx + RPT_K - 1 < T && x_hi >= x_lo) {
        memcpy(tx.h, hx_host, sizeof(tx.h)); tx.c = cx; tx.lo = x_lo; tx.hi = x_hi;
    }
    if (hy_host && nty == RPT_K && cyy >= 0 && y_lo - cyy >= 0 && y_hi - cyy + RPT_K - 1 < S && y_hi >= y_lo) {
        memcpy(ty.h, hy_host, sizeof(ty.h)); ty.c = cyy; ty.lo = y_lo; ty.hi = y_hi;
    }
    cudaStream_t st = as_stream(stream);
    switch (dtype) {
    case PCB_U16: return launch_refi_prefilter<uint16_t>(vp, M, S, T, py_vals, py_first, Kpy, px_vals, px_first, Kpx, view_zeroed, tmp, coef, st, &tx, &ty, rows_host);
    case PCB_F32: return launch_refi_prefilter<float>(vp, M, S, T, py_vals, py_first, Kpy, px_vals, px_first, Kpx, view_zeroed, tmp, coef, st, &tx, &ty, rows_host);
    case PCB_U8:  return launch_refi_prefilter<uint8_t>(vp, M, S, T, py_vals, py_first, Kpy, px_vals, px_first, Kpx, view_zeroed, tmp, coef, st, &tx, &ty, rows_host);
    }
    return fail(PCB_ERR_INVALID, "%s: unsupported dtype %lld", __func__, dtype);
}

static int check_refocus_refi(const float *coef, int M, int S, int T, int C, int N, const float *dy_vals,
                              const int32_t *dy_first, int Kgy, const float *dx_vals, const int32_t *dx_first, int Kx,
                              const int32_t *iy, const int32_t *ix, const int32_t *xlo_host, const int32_t *xhi_host,
                              const uint8_t *view_zeroed_host, const uint8_t *view_zeroed, float *tmp) {
    PCB_REQUIRE(coef && dy_vals && dy_first && dx_vals && dx_first && iy && ix && xlo_host && xhi_host &&
                view_zeroed_host && view_zeroed && tmp, "pointers");
    PCB_REQUIRE(M > 0 && M <= RF_MAXM && M % 2 == 1 && S >= 4 && T >= 4 && N > 0 && N <= 65535 && Kgy > 0 && Kx > 0,
                "sizes, odd M <= 32");
    PCB_REQUIRE(C == 3, "3 colour channels (lfp_shiftandsum.py:97)");
    PCB_REQUIRE((S + RH_ROWS - 1) / RH_ROWS <= 65535, "S");
    return PCB_OK;
}

int pcb_refocus_refi(const float *coef, int M, int S, int T, int C, int N, const float *dy_vals,
                     const int32_t *dy_first, int Kgy, const float *dx_vals, const int32_t *dx_first, int Kx,
                     const int32_t *iy, const int32_t *ix, const int32_t *xlo_host, const int32_t *xhi_host,
                     const uint8_t *view_zeroed_host, const uint8_t *view_zeroed, float *tmp, float *out,
                     pcb_minmax_t *mm, void *stream) {
    int rc = check_refocus_refi(coef, M, S, T, C, N, dy_vals, dy_first, Kgy, dx_vals, dx_first, Kx, iy, ix, xlo_host,
                                xhi_host, view_zeroed_host, view_zeroed, tmp);
    if (rc) return rc;
    PCB_REQUIRE(out != nullptr, "out");
    pcb_fanout_t dst;
    memset(&dst, 0, sizeof(dst));
    dst.out[0] = out;
    dst.mm[0] = mm;
    dst.n = 1;
    return launch_refine(coef, M, S, T, N, dy_vals, dy_first, Kgy, dx_vals, dx_first, Kx, iy, ix, xlo_host, xhi_host,
                         view_zeroed_host, tmp, dst, as_stream(stream));
}

int pcb_refocus_refi_fanout(const float *coef, int M, int S, int T, int C, int N, const float *dy_vals,
                            const int32_t *dy_first, int Kgy, const float *dx_vals, const int32_t *dx_first, int Kx,
                            const int32_t *iy, const int32_t *ix, const int32_t *xlo_host, const int32_t *xhi_host,
                            const uint8_t *view_zeroed_host, const uint8_t *view_zeroed, float *tmp,
                            const pcb_fanout_t *dst_host, const pcb_refi_quads_t *quads_host,
                            const pcb_refi_rows_t *rows_host, void *stream) {
    int rc = check_refocus_refi(coef, M, S, T, C, N, dy_vals, dy_first, Kgy, dx_vals, dx_first, Kx, iy, ix, xlo_host,
                                xhi_host, view_zeroed_host, view_zeroed, tmp);
    if (rc) return rc;
    PCB_REQUIRE(dst_host != nullptr && dst_host->n >= 1 && dst_host->n <= PCB_MAX_PEERS, "1 <= destinations <= 16");
    pcb_fanout_t dst = *dst_host;
    for (int d = 0; d < PCB_MAX_PEERS; ++d) {
        if (d < dst.n) {
            PCB_REQUIRE(dst.out[d] != nullptr, "destination pointer");
            PCB_REQUIRE((dst.mm[d] != nullptr) == (dst.mm[0] != nullptr), "min/max targets: all or none");
        } else {
            dst.out[d] = nullptr;
            dst.mm[d] = nullptr;
        }
    }
    return launch_refine(coef, M, S, T, N, dy_vals, dy_first, Kgy, dx_vals, dx_first, Kx, iy, ix, xlo_host, xhi_host,
                         view_zeroed_host, tmp, dst, as_stream(stream), quads_host, rows_host);
}

int pcb_refi_upscaled_slice(const float *coef, int M, int S, int T, int C, int a, const int32_t *fx, const float *wx,
                            const int32_t *fy, const float *wy, const uint8_t *view_zeroed_host, float *tmp, float *out,
                            void *stream) {
    PCB_REQUIRE(coef && fx && wx && fy && wy && view_zeroed_host && tmp && out, "pointers");
    PCB_REQUIRE(M > 0 && M <= RF_MAXM && M % 2 == 1 && S >= 4 && T >= 4 && S <= 65535, "sizes, odd M <= 32");
    PCB_REQUIRE((int64_t)S * M <= 65535, "S * M <= 65535");
    PCB_REQUIRE(C == 3, "3 colour channels (lfp_shiftandsum.py:97)");
    return launch_refi_upscaled(coef, M, S, T, a, fx, wx, fy, wy, view_zeroed_host, tmp, out, as_stream(stream));
}

// ---- peer-mapped buffers (slice-sharded stacks) --------------------------------------------------------------------
int pcb_peer_alloc(int64_t bytes, void **dptr_host) {
    PCB_REQUIRE(bytes > 0 && dptr_host != nullptr, "bytes > 0, dptr");
    PCB_CUDA(cudaMalloc(dptr_host, (size_t)bytes));
    return PCB_OK;
}

int pcb_peer_free(void *dptr) {
    PCB_REQUIRE(dptr != nullptr, "dptr");
    PCB_CUDA(cudaFree(dptr));
    return PCB_OK;
}

int pcb_peer_export(void *dptr, unsigned char handle_host[64]) {
    PCB_REQUIRE(dptr != nullptr && handle_host != nullptr, "dptr, handle");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    PCB_CUDA(cudaIpcGetMemHandle(&h, dptr));
    memcpy(handle_host, &h, 64);
    return PCB_OK;
}

int pcb_peer_open(const unsigned char handle_host[64], void **dptr_host) {
    PCB_REQUIRE(handle_host != nullptr && dptr_host != nullptr, "handle, dptr");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle_host, 64);
    PCB_CUDA(cudaIpcOpenMemHandle(dptr_host, h, cudaIpcMemLazyEnablePeerAccess));
    return PCB_OK;
}

int pcb_peer_close(void *dptr) {
    PCB_REQUIRE(dptr != nullptr, "dptr");
    PCB_CUDA(cudaIpcCloseMemHandle(dptr));
    return PCB_OK;
}

int64_t pcb_percentile_scratch_bytes(void) { return (int64_t)sizeof(SelectState); }

int pcb_percentile_f32(const float *data, int64_t n, int luma, const double *q_host, int nq, double *result,
                       void *scratch, void *stream) {
    PCB_REQUIRE(data != nullptr && q_host != nullptr && result != nullptr && scratch != nullptr, "pointers");
    PCB_REQUIRE(n > 0 && nq > 0 && nq <= PS_MAXQ, "n > 0, 1 <= nq <= 4");
    for (int k = 0; k < nq; ++k) PCB_REQUIRE(q_host[k] >= 0.0 && q_host[k] <= 100.0, "0 <= q <= 100");
    return run_percentile<float>(data, n, luma, q_host, nq, result, scratch, as_stream(stream));
}

int pcb_contrast_bounds(const float *ref, int64_t npx, double q_lo, double q_hi, double *lohi, void *scratch,
                        void *stream) {
    PCB_REQUIRE(ref != nullptr && lohi != nullptr && scratch != nullptr && npx > 0, "ref, lohi, scratch, npx > 0");
    PCB_REQUIRE(q_lo >= 0.0 && q_lo <= 100.0 && q_hi >= 0.0 && q_hi <= 100.0, "0 <= q <= 100");
    return run_contrast_bounds(ref, npx, q_lo, q_hi, lohi, scratch, as_stream(stream));
}

int pcb_percentile_f64(const double *data, int64_t n, int luma, const double *q_host, int nq, double *result,
                       void *scratch, void *stream) {
    PCB_REQUIRE(data != nullptr && q_host != nullptr && result != nullptr && scratch != nullptr, "pointers");
    PCB_REQUIRE(n > 0 && nq > 0 && nq <= PS_MAXQ, "n > 0, 1 <= nq <= 4");
    for (int k = 0; k < nq; ++k) PCB_REQUIRE(q_host[k] >= 0.0 && q_host[k] <= 100.0, "0 <= q <= 100");
    return run_percentile<double>(data, n, luma, q_host, nq, result, scratch, as_stream(stream));
}

int pcb_devig_white(const void *wht, int dtype, int64_t npx, int C, double *wgray, void *stream) {
    PCB_REQUIRE(wht != nullptr && wgray != nullptr && npx > 0 && C >= 1, "wht, wgray, npx > 0, C >= 1");
    cudaStream_t st = as_stream(stream);
    const int g = devig_grid(npx);
    switch (dtype) {
    case PCB_U16: devig_white_kernel<uint16_t><<<g, 256, 0, st>>>((const uint16_t *)wht, npx, C, wgray); break;
    case PCB_F32: devig_white_kernel<float><<<g, 256, 0, st>>>((const float *)wht, npx, C, wgray); break;
    case PCB_U8:  devig_white_kernel<uint8_t><<<g, 256, 0, st>>>((const uint8_t *)wht, npx, C, wgray); break;
    case PCB_F64: devig_white_kernel<double><<<g, 256, 0, st>>>((const double *)wht, npx, C, wgray); break;
    default: return fail(PCB_ERR_INVALID, "%s: unsupported dtype %lld", __func__, dtype);
    }
    PCB_LAUNCH_OK();
    return PCB_OK;
}

int pcb_devig_normalize(double *wgray, int64_t npx, const double *divisor, void *stream) {
    PCB_REQUIRE(wgray != nullptr && divisor != nullptr && npx > 0, "wgray, divisor, npx > 0");
    devig_normalize_kernel<<<devig_grid(npx), 256, 0, as_stream(stream)>>>(wgray, npx, divisor);
    PCB_LAUNCH_OK();
    return PCB_OK;
}

int pcb_devig_noise_level(const double *wn, int H, int W, const double *gauss_host, int glen, double *tmp,
                          double *stats, void *stream) {
    PCB_REQUIRE(wn != nullptr && gauss_host != nullptr && tmp != nullptr && stats != nullptr, "pointers");
    PCB_REQUIRE(H > 0 && W > 0 && H <= 65535 && glen >= 1 && glen <= DV_MAXG && (glen & 1), "sizes, odd glen <= 64");
    cudaStream_t st = as_stream(stream);
    DevigGauss gk;
    memset(&gk, 0, sizeof(gk));
    gk.len = glen;
    for (int k = 0; k < glen; ++k) gk.g[k] = gauss_host[k];
    PCB_CUDA(cudaMemsetAsync(stats, 0, 3 * sizeof(double), st));
    dim3 grid((W + 255) / 256, H);
    devig_blur_rows_kernel<<<grid, 256, 0, st>>>(wn, H, W, gk, tmp);
    devig_blur_cols_stats_kernel<<<grid, 256, 0, st>>>(wn, tmp, H, W, gk, stats);
    devig_stats_finish_kernel<<<1, 1, 0, st>>>(stats, (double)H * (double)W);
    PCB_LAUNCH_OK();
    return PCB_OK;
}

int pcb_devig_divide(const void *lfp, int dtype, int64_t npx, int C, const double *wn, void *out, int out_dtype,
                     void *stream) {
    PCB_REQUIRE(lfp != nullptr && wn != nullptr && out != nullptr && npx > 0 && C >= 1, "pointers, npx > 0, C >= 1");
    cudaStream_t st = as_stream(stream);
    switch (dtype) {
    case PCB_U16: return devig_divide_out<uint16_t>(lfp, npx, C, wn, out, out_dtype, st);
    case PCB_F32: return devig_divide_out<float>(lfp, npx, C, wn, out, out_dtype, st);
    case PCB_U8:  return devig_divide_out<uint8_t>(lfp, npx, C, wn, out, out_dtype, st);
    case PCB_F64: return devig_divide_out<double>(lfp, npx, C, wn, out, out_dtype, st);
    }
    return fail(PCB_ERR_INVALID, "%s: unsupported dtype %lld", __func__, dtype);
}

int pcb_image_minmax(const void *img, int dtype, int64_t n, pcb_minmax_t *mm, void *stream) {
    PCB_REQUIRE(img != nullptr && mm != nullptr && n > 0, "img, mm, n > 0");
    cudaStream_t st = as_stream(stream);
    const int g = grid_for(n, 256 * 8);
    switch (dtype) {
    case PCB_U16: image_minmax_kernel<uint16_t><<<g, 256, 0, st>>>((const uint16_t *)img, n, mm); break;
    case PCB_F32: image_minmax_kernel<float><<<g, 256, 0, st>>>((const float *)img, n, mm); break;
    case PCB_U8:  image_minmax_kernel<uint8_t><<<g, 256, 0, st>>>((const uint8_t *)img, n, mm); break;
    default: return fail(PCB_ERR_INVALID, "%s: unsupported dtype %lld", __func__, dtype);
    }
    PCB_LAUNCH_OK();
    return PCB_OK;
}

int pcb_warp_projective(const void *img, int dtype, int H, int W, int C, const double *inv_matrix_host, int rows,
                        int cols, const pcb_minmax_t *mm, float *out, void *stream) {
    PCB_REQUIRE(img != nullptr && inv_matrix_host != nullptr && mm != nullptr && out != nullptr, "pointers");
    PCB_REQUIRE(H > 0 && W > 0 && C > 0 && rows > 0 && rows <= 65535 && cols > 0, "sizes");
    cudaStream_t st = as_stream(stream);
    WarpMat wm;
    for (int k = 0; k < 9; ++k) wm.m[k] = inv_matrix_host[k];
    dim3 grid((cols + 255) / 256, rows);
    switch (dtype) {
    case PCB_U16: warp_projective_kernel<uint16_t><<<grid, 256, 0, st>>>((const uint16_t *)img, H, W, C, wm, rows, cols, mm, out); break;
    case PCB_F32: warp_projective_kernel<float><<<grid, 256, 0, st>>>((const float *)img, H, W, C, wm, rows, cols, mm, out); break;
    case PCB_U8:  warp_projective_kernel<uint8_t><<<grid, 256, 0, st>>>((const uint8_t *)img, H, W, C, wm, rows, cols, mm, out); break;
    default: return fail(PCB_ERR_INVALID, "%s: unsupported dtype %lld", __func__, dtype);
    }
    PCB_LAUNCH_OK();
    return PCB_OK;
}

int pcb_hex_align(const float *prj, int prows, int pcols, int C, int strips, int py, int px, int hex_odd, int nk,
                  const int32_t *i0, const int32_t *i1, const float *w, const float *wx_vals, const int32_t *wx_first,
                  int K, int x_len, int crop_lo, int out_cols, const float *wy, int y_len, float *stretched, float *tmpx,
                  float *out, void *stream) {
    PCB_REQUIRE(prj && i0 && i1 && w && wx_vals && wx_first && wy && stretched && tmpx && out, "pointers");
    PCB_REQUIRE(prows > 0 && pcols > 0 && C > 0 && strips > 0 && strips <= 65535 && py >= 4 && py <= HX_MAXPY &&
                px >= 2 && px % 2 == 0 && nk > 0 && K > 0 && x_len > 0 && crop_lo >= 0 && out_cols > 0 &&
                crop_lo + out_cols <= x_len && y_len > 0, "sizes (even px, 4 <= py <= 64)");
    HexArgs a;
    a.prj = prj; a.prows = prows; a.pcols = pcols; a.C = C; a.strips = strips; a.py = py; a.px = px;
    a.hex_odd = hex_odd ? 1 : 0; a.nk = nk; a.i0 = i0; a.i1 = i1; a.w = w; a.wx_vals = wx_vals; a.wx_first = wx_first;
    a.K = K; a.x_len = x_len; a.crop_lo = crop_lo; a.out_cols = out_cols; a.wy = wy; a.y_len = y_len;
    const size_t smem = (size_t)nk * px * C * sizeof(float);
    if (smem > 220 * 1024) return fail(PCB_ERR_UNSUPPORTED, "%s: a stretched lens row does not fit shared memory", __func__);
    cudaStream_t st = as_stream(stream);
    PCB_REQUIRE(py <= 65535, "py");
    hex_merge_stretch_kernel<<<dim3((nk * px * C + 255) / 256, py, strips), 256, 0, st>>>(a, stretched);
    PCB_LAUNCH_OK();
    PCB_CUDA(cudaFuncSetAttribute(hex_resize_x_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    hex_resize_x_kernel<<<dim3(py, strips), HX_THREADS, smem, st>>>(a, stretched, tmpx);
    PCB_LAUNCH_OK();
    hex_strip_y_kernel<<<dim3((out_cols * C + 255) / 256, strips), 256, 0, st>>>(a, tmpx, out);
    PCB_LAUNCH_OK();
    return PCB_OK;
}

int pcb_unpack_bayer(const uint8_t *buf, int64_t nbytes, int bits, uint16_t *out, void *stream) {
    PCB_REQUIRE(buf != nullptr && out != nullptr && nbytes > 0, "buf, out, nbytes > 0");
    PCB_REQUIRE(bits == 10 || bits == 12, "10- or 12-bit packing (lfp_decoder.py:280-312)");
    PCB_REQUIRE((reinterpret_cast<uintptr_t>(out) & 7) == 0, "out 8-byte aligned");
    cudaStream_t st = as_stream(stream);
    if (bits == 10) {
        const int64_t ng = nbytes / 5;
        PCB_REQUIRE(ng > 0, "at least one 5-byte group");
        unpack10_kernel<<<grid_for(ng, 256 * 4), 256, 0, st>>>(buf, ng, out);
    } else {
        const int64_t ng = nbytes / 3;
        PCB_REQUIRE(ng > 0, "at least one 3-byte group");
        unpack12_kernel<<<grid_for(ng, 256 * 4), 256, 0, st>>>(buf, ng, out);
    }
    PCB_LAUNCH_OK();
    return PCB_OK;
}

int pcb_cfa_safe_awb(const float *bay, const float *wht, int H, int W, const double *gains_host, int soft_clip,
                     double max_lum, double b, double log1pb, double *out, void *stream) {
    PCB_REQUIRE(bay && out && gains_host, "null pointer");
    PCB_REQUIRE(H > 0 && W > 0 && (H & 1) == 0 && (W & 1) == 0, "mosaic needs even, positive dimensions");
    PCB_REQUIRE((reinterpret_cast<uintptr_t>(bay) & 7) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 &&
                (wht == nullptr || (reinterpret_cast<uintptr_t>(wht) & 7) == 0), "buffers must be 8 / 16 byte aligned");
    PCB_REQUIRE(!soft_clip || max_lum > 0.0, "max_lum must be positive");
    CfaArgs a;
    a.bay = bay; a.wht = wht; a.H = H; a.W = W; a.out = out;
    a.fact = gains_host[0];
    for (int i = 0; i < 4; ++i) { a.g[i] = gains_host[i]; if (gains_host[i] > a.fact) a.fact = gains_host[i]; }
    a.soft_clip = soft_clip; a.max_lum = max_lum; a.b = b; a.log1pb = log1pb;
    dim3 grid((W / 2 + 255) / 256, H / 2);
    if (wht) cfa_safe_awb_kernel<true><<<grid, 256, 0, as_stream(stream)>>>(a);
    else cfa_safe_awb_kernel<false><<<grid, 256, 0, as_stream(stream)>>>(a);
    PCB_LAUNCH_OK();
    return PCB_OK;
}

int pcb_scheimpflug(const float *stack, int N, int S, int T, int C, const int32_t *a_y, const int32_t *a_x, int mode,
                    float *out, void *stream) {
    PCB_REQUIRE(stack != nullptr && a_y != nullptr && a_x != nullptr && out != nullptr, "pointers");
    PCB_REQUIRE(N > 0 && S > 0 && S <= 65535 && T > 0 && C > 0 && mode >= 0 && mode <= 2, "sizes, mode in 0..2");
    const int E = T * C;
    scheimpflug_kernel<<<dim3((E + 255) / 256, S), 256, 0, as_stream(stream)>>>(stack, N, S, E, C, a_y, a_x, mode, out);
    PCB_LAUNCH_OK();
    return PCB_OK;
}

int pcb_contrast_srgb_typed(const void *in, int dtype, int64_t n, const double *lohi, const pcb_minmax_t *mm, float *out,
                            void *stream) {
    PCB_REQUIRE(in != nullptr && out != nullptr && lohi != nullptr && n > 0, "in, out, lohi, n > 0");
    const int g = grid_for(n, 256 * 8);
    cudaStream_t st = as_stream(stream);
    switch (dtype) {
    case PCB_F32: contrast_srgb_kernel<float><<<g, 256, 0, st>>>((const float *)in, n, lohi, mm, out); break;
    case PCB_U16: contrast_srgb_kernel<uint16_t><<<g, 256, 0, st>>>((const uint16_t *)in, n, lohi, mm, out); break;
    case PCB_U8:  contrast_srgb_kernel<uint8_t><<<g, 256, 0, st>>>((const uint8_t *)in, n, lohi, mm, out); break;
    default: return fail(PCB_ERR_INVALID, "%s: unsupported dtype %lld", "pcb_contrast_srgb_typed", dtype);
    }
    PCB_LAUNCH_OK();
    return PCB_OK;
}

int pcb_contrast_srgb(const float *in, int64_t n, const double *lohi, const pcb_minmax_t *mm, float *out,
                      void *stream) {
    return pcb_contrast_srgb_typed(in, PCB_F32, n, lohi, mm, out, stream);
}

}  // extern "C"
