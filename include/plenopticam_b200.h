/*
 * plenopticam_b200 — C-ABI of the B200-native light-field hot path (libpcb200.so).
 *
 * Drop-in boundary for hahnec/plenopticam's data-parallel path.  The reference has no FFI of its own
 * (it is pure Python); each entry point below replaces the numpy/scipy body of one reference method and
 * is what a ctypes binding inside that method would call.  Citations are
 * file:line in /root/reference/plenopticam/.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller unless the name ends in `_host`;
 *   - no allocation, no synchronisation inside: work is enqueued on `stream` (a cudaStream_t passed as
 *     void*; NULL = legacy default stream) and the call returns immediately;
 *   - images are dense, row-major, channel-interleaved (H, W, C) exactly like the reference's ndarrays;
 *     the 4-D light field is (M, M, S, T, C) = vp_img_arr[j, i, s, t, c];
 *   - return value: PCB_OK or an error code; pcb_last_error() gives the text (thread-local).
 */
#ifndef PLENOPTICAM_B200_H
#define PLENOPTICAM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ABI revision: bumped whenever an entry point is added or its arguments change.  pcb_version() returns the value the
 * library was built with; a binding refuses a library of another revision (plenopticam_b200/_native.py). */
#define PCB_ABI_VERSION 202

#define PCB_OK               0
#define PCB_ERR_INVALID      1   /* bad argument (null pointer, non-positive size, unsupported dtype)  */
#define PCB_ERR_CUDA         2   /* a CUDA runtime call / launch failed                                  */
#define PCB_ERR_UNSUPPORTED  3   /* valid in the reference but not offered by this library              */

/* element types of image buffers */
#define PCB_U16 0
#define PCB_F32 1
#define PCB_U8  2
#define PCB_F64 3   /* de-vignetting only (the reference works on float64 copies, lfp_aligner/top_level.py:39-40) */

/* micro-lens arrangement: cfg.calibs['pat_type'] */
#define PCB_REC 0
#define PCB_HEX 1

/* One calibrated micro-lens, prepared on the host in float64 so that `rint` (round-half-even,
 * misc/type_checks.py:57-58) is bit-exact:  for centroid (my, mx), ry = rint(my), fy = my - ry,
 *   oy = ry - M/2 + floor(fy)   first raw row read for patch row v = 0
 *   wy = fy - floor(fy)         bilinear weight of row oy+v+1                     (same for x)
 * A lens whose (M+2)^2 window leaves the image (lfp_local_resampler.py:50,58-60 -> zero patch) has
 * oy = PCB_LENS_INVALID. */
#define PCB_LENS_INVALID INT32_MIN
typedef struct { int32_t oy, ox; float wy, wx; } pcb_lens_t;

/* One output lens column k of the hex->rectilinear stretch (lfp_local_resampler.py:129-136):
 * out = P[x0] + (P[x1] - P[x0]) * w.  Two tables of Xs entries: row parity (ly + hex_odd) % 2 = 0, 1.
 * For PCB_REC pass NULL (identity). */
typedef struct { int32_t x0, x1; float w; int32_t reserved; } pcb_hexcol_t;

/* Order-preserving uint32 keys of the running float min / max (keys[0] = min, keys[1] = max). */
typedef struct { uint32_t kmin, kmax; } pcb_minmax_t;

int         pcb_version(void);
const char *pcb_last_error(void);
float       pcb_key_to_float(uint32_t key);         /* host helper: decode a pcb_minmax_t key           */
/* Measurement hook (bench.py's roofline): two cudaEvent_t (as void*; NULL, NULL switches it off) that the calling thread's
 * next pcb_refocus_int* calls record on their stream right before / after the persistent shift-and-sum kernel, so that the
 * dominant kernel can be timed inside the stage without a profiler. */
int         pcb_debug_kernel_events(void *ev_begin, void *ev_end);

/* Reset a pcb_minmax_t to (+inf, -inf). */
int pcb_minmax_reset(pcb_minmax_t *mm, void *stream);
/* Fold n float32 values into *mm (atomic). */
int pcb_minmax_f32(const float *data, int64_t n, pcb_minmax_t *mm, void *stream);

/* ---------------------------------------------------------------------------------------------------
 * (a) Micro-image resampling — replaces LfpLocalResampler.resample_rec / resample_hex / _patch_align
 *     (lfp_aligner/lfp_local_resampler.py:44-148, method 'linear') and, for the _u16 form, the
 *     quantisation in LfpResampler._write_lfp_align (lfp_aligner/lfp_resampler.py:60,
 *     misc/normalizer.py:33-46,69-78).
 *   raw (H,W,C) of raw_dtype;  lens[LY*LX];  hexcol[2*Xs] or NULL;  Xs = LX for PCB_REC.
 *   Output aligned image: (LY*M, Xs*M, C).
 *   pcb_resample_minmax : no image output, folds min/max of the float32 aligned values into *mm
 *   pcb_resample_f32    : float32 aligned image (and min/max if mm != NULL)
 *   pcb_resample_u16    : recomputes and writes uint16(round_half_even(clip((x-min)/(max-min))*65535))
 *                         using *mm read on the device (no host round trip between the two passes)
 * ------------------------------------------------------------------------------------------------ */
int pcb_resample_minmax(const void *raw, int raw_dtype, int H, int W, int C,
                        const pcb_lens_t *lens, int LY, int LX, int M,
                        int pattern, const pcb_hexcol_t *hexcol, int Xs, int hex_odd, int flip,
                        pcb_minmax_t *mm, void *stream);
int pcb_resample_f32(const void *raw, int raw_dtype, int H, int W, int C,
                     const pcb_lens_t *lens, int LY, int LX, int M,
                     int pattern, const pcb_hexcol_t *hexcol, int Xs, int hex_odd, int flip,
                     float *aligned, pcb_minmax_t *mm, void *stream);
int pcb_resample_u16(const void *raw, int raw_dtype, int H, int W, int C,
                     const pcb_lens_t *lens, int LY, int LX, int M,
                     int pattern, const pcb_hexcol_t *hexcol, int Xs, int hex_odd, int flip,
                     const pcb_minmax_t *mm, uint16_t *aligned, void *stream);
/* Normalizer.uint16_norm on an existing float32 buffer (misc/normalizer.py:43-46). */
int pcb_quantize_u16(const float *in, int64_t n, const pcb_minmax_t *mm, uint16_t *out, void *stream);
/* Normalizer.uint8_norm (misc/normalizer.py:48-51): what LfpExporter / save_img_file write to 8-bit files
 * (lfp_extractor/lfp_exporter.py:195, misc/file_rw.py:86). */
int pcb_quantize_u8(const float *in, int64_t n, const pcb_minmax_t *mm, uint8_t *out, void *stream);

/* ---------------------------------------------------------------------------------------------------
 * (a') Rotation — replaces LfpRotator._rotate_img (lfp_aligner/lfp_rotator.py:129-145):
 *     scipy.ndimage.rotate(img, deg, reshape=False, order=3, mode='constant', cval=0).
 *   pcb_spline_prefilter : cubic B-spline coefficients (mirror boundary), separable, float32;
 *                          `tmp` and `coef` are (H,W,C) float32 scratch/outputs.
 *   pcb_rotate_sample    : out[y,x] = sum 4x4 B-spline taps of coef at R(y,x), 0 outside [0,H-1]x[0,W-1]
 * ------------------------------------------------------------------------------------------------ */
int pcb_spline_prefilter(const void *img, int dtype, int H, int W, int C, float *tmp, float *coef, void *stream);
int pcb_rotate_sample(const float *coef, int H, int W, int C, double rad, float *out, void *stream);

/* ---------------------------------------------------------------------------------------------------
 * (b) Viewpoint gather — replaces LfpRearranger.compose_viewpoints / decompose_viewpoints
 *     (lfp_extractor/lfp_rearranger.py:79-137) with LfpCropper.crop_micro_image
 *     (lfp_extractor/lfp_cropper.py:58-62) folded into the index math.
 *   aligned (rows, cols, C) with micro images of M_in px; vp (M_out, M_out, rows/M_in, cols/M_in, C);
 *   crop margin k = (M_in - M_out)/2.  dtype preserved (bit-exact copy).
 * ------------------------------------------------------------------------------------------------ */
int pcb_viewpoints(const void *aligned, int dtype, int rows, int cols, int C, int M_in, int M_out,
                   void *vp, void *stream);
int pcb_viewpoints_inverse(const void *vp, int dtype, int M, int S, int T, int C, void *aligned, void *stream);
int pcb_crop_micro_images(const void *aligned, int dtype, int rows, int cols, int C, int M_in, int M_out,
                          void *cropped, void *stream);
/* non-square micro images (My_in x Mx_in), e.g. the 13 x 15 pitch of a globally rectified hexagonal light field */
int pcb_crop_micro_images_yx(const void *aligned, int dtype, int rows, int cols, int C, int My_in, int Mx_in,
                             int M_out, void *cropped, void *stream);

/* ---------------------------------------------------------------------------------------------------
 * (c) Shift-and-sum refocus — replaces LfpShiftAndSum.refo_from_vp without refinement
 *     (lfp_refocuser/lfp_shiftandsum.py:77-139) including circular_view_aperture
 *     (lfp_extractor/lfp_viewpoints.py:229-250) and the /M scaling (:89):
 *       out[n,y,x,c] = (1/M) * sum_{(j,i): !view_zeroed[j*M+i]} vp[j,i,clamp(y+a_n(c0-j)),clamp(x+a_n(c0-i)),c]
 *   uint16 input is accumulated exactly in integers; out is float32 (N,S,T,C).  M must be odd.
 *   `a_list_host` (N refocus parameters) and `view_zeroed_host` (M*M flags) are small HOST arrays.
 *   `scratch` is device memory of at least pcb_refocus_int_scratch_bytes(...) bytes (same arguments).
 *   uint16 light fields with M <= 16 take the row-stationary kernel (each view row staged once in shared
 *   memory and used for every slice); other inputs take the direct kernel.  Both are exact for integer
 *   data; pcb_refocus_int_direct exposes the direct kernel alone (device-resident control arrays).
 * ------------------------------------------------------------------------------------------------ */
int64_t pcb_refocus_int_scratch_bytes(int dtype, int M, int S, int T, int C,
                                      const int32_t *a_list_host, int N, const uint8_t *view_zeroed_host);
int pcb_refocus_int(const void *vp, int dtype, int M, int S, int T, int C,
                    const int32_t *a_list_host, int N, const uint8_t *view_zeroed_host,
                    void *scratch, int64_t scratch_bytes,
                    float *out, pcb_minmax_t *mm /* nullable: fused min/max of the stack */, void *stream);
/* pcb_refocus_int / pcb_refocus_int_srgb reading the ALIGNED micro-image array (S*Ma, T*Ma, 3) of LfpResampler instead
 * of the 4-D light field: view (j, i) is the sub-grid [ck + j :: Ma, ck + i :: Ma], ck = (Ma - M) / 2, i.e.
 * LfpCropper.crop_micro_image (lfp_extractor/lfp_cropper.py:58-62) and LfpRearranger.compose_viewpoints
 * (lfp_extractor/lfp_rearranger.py:93-101) are folded into the staging of the refocus tiles (one bulk copy per aligned
 * row).  Same results bit for bit.  PCB_ERR_UNSUPPORTED when the persistent packed-pixel kernel does not apply
 * (then: pcb_viewpoints + pcb_refocus_int).
 * scratch_flags (0 = the call zeroes the accumulators itself and leaves them dirty, like pcb_refocus_int):
 *   PCB_REFOCUS_LEAVE_ZEROED    the pass that reads an accumulator word last stores zero behind it, so the scratch comes
 *                               back with clean accumulators;
 *   PCB_REFOCUS_SCRATCH_ZEROED  the caller vouches that the accumulators ARE zero - the previous call on this scratch was
 *                               one of these two with PCB_REFOCUS_LEAVE_ZEROED, ordered before this one, and nothing wrote
 *                               to the scratch since - and the call skips its memset (209 MB per Illum exposure).
 * A sequence of exposures passes LEAVE_ZEROED on the first call and both flags from then on. */
#define PCB_REFOCUS_SCRATCH_ZEROED 1
#define PCB_REFOCUS_LEAVE_ZEROED 2
int pcb_refocus_int_aligned(const void *aligned, int dtype, int Ma, int M, int S, int T, int C,
                            const int32_t *a_list_host, int N, const uint8_t *view_zeroed_host,
                            void *scratch, int64_t scratch_bytes, float *out, pcb_minmax_t *mm, int scratch_flags,
                            void *stream);
int pcb_refocus_int_srgb_aligned(const void *aligned, int dtype, int Ma, int M, int S, int T, int C,
                                 const int32_t *a_list_host, int N, const uint8_t *view_zeroed_host,
                                 void *scratch, int64_t scratch_bytes, double q_lo, double q_hi,
                                 float *ref_slice, double *lohi, void *pscratch, float *out, pcb_minmax_t *mm,
                                 int scratch_flags, void *stream);
int pcb_refocus_int_direct(const void *vp, int dtype, int M, int S, int T, int C,
                           const int32_t *a_list_dev, int N, const uint8_t *view_zeroed_dev,
                           float *out, pcb_minmax_t *mm, void *stream);
/* LfpShiftAndSum.refo_from_scratch (lfp_refocuser/lfp_shiftandsum.py:141-221; the alternate entry taken when the class
 * is given the aligned image instead of the viewpoints and opt_refi is off): the same shifts a (c0 - j), a (c0 - i), but
 * samples outside the light field count as ZERO ("leave pixel empty", :180-186,:203-206) and, as that method never calls
 * circular_view_aperture, the caller passes an all-zero mask.  Direct kernel, device-resident control arrays. */
int pcb_refocus_int_zero(const void *vp, int dtype, int M, int S, int T, int C,
                         const int32_t *a_list_dev, int N, const uint8_t *view_zeroed_dev,
                         float *out, pcb_minmax_t *mm, void *stream);

/* pcb_refocus_int with LfpRefocuser.shift_and_sum's colour management (lfp_refocuser/top_level.py:68-70) fused into the
 * epilogue — the "fused normalisation and contrast clipping" form of the stack: out receives
 * srgb(clip((x - lo) / (hi - lo))) with lo = percentile(rgb2gry(slice 0), q_lo), hi = percentile(slice 0, q_hi), the
 * linear stack is never written (one pass over N slices less).  Same values as pcb_refocus_int + pcb_contrast_bounds +
 * pcb_contrast_srgb, bit for bit.  ref_slice: (S,T,C) float32 scratch (linear slice 0), lohi: double[2] (device, out),
 * pscratch: pcb_percentile_scratch_bytes(), mm: reset by the caller, receives min / max of the LINEAR stack (a bound
 * that is exactly 0 falls back to them, misc/normalizer.py:40-41, through a fix-up launch that otherwise returns at
 * once).  PCB_ERR_UNSUPPORTED when the packed-pixel kernel does not apply (not uint16 RGB, M > 16, rows > 2048 px). */
int pcb_refocus_int_srgb(const void *vp, int dtype, int M, int S, int T, int C, const int32_t *a_list_host, int N,
                         const uint8_t *view_zeroed_host, void *scratch, int64_t scratch_bytes, double q_lo,
                         double q_hi, float *ref_slice, double *lohi, void *pscratch, float *out, pcb_minmax_t *mm,
                         void *stream);

/* ---------------------------------------------------------------------------------------------------
 * (c') Sub-pixel "refocus refinement" — replaces refo_from_vp with opt_refi (lfp_refocuser/lfp_shiftandsum.py:93-135)
 *     and the two misc.img_resize calls per view / per slice (misc/data_proc.py:57-92).
 *     The up-sample (not-a-knot bicubic) / shift / sum / down-sample chain is linear and separable:
 *       out[n] = (1/M) * sum_{j,i active}  A_{a_n(c0-j)} . vp[j,i] . B_{a_n(c0-i)}^T ,     A_s = D_s . P
 *     P (samples -> cubic B-spline coefficients) is independent of the slice, D_s is a band of <= 5 taps.  All
 *     operators are prepared on the host in float64 (plenopticam_b200/lfp_refocuser/refinement.py).
 *   pcb_refi_prefilter : once per light field, coef[j,i] = Py . vp[j,i] . Px^T for every unmasked view
 *       py_vals[ceil(S/8)][Kpy][8], py_first[ceil(S/8)] (groups of 8 output rows share a window);
 *       px_vals[Kpx][T], px_first[T];  view_zeroed[M*M] (device)
 *       tmp, coef: float32 (M,M,S,T,3)
 *   pcb_refocus_refi   : a batch of N slices from the coefficients
 *       dy_vals[nsy][ceil(S/8)][Kgy][8], dy_first[nsy][ceil(S/8)]  (groups of 8 output rows share a window of Kgy
 *                                 sample rows that lies inside the image: dy_first + Kgy <= S)
 *       dx_vals[nsx][Kx][T], dx_first[nsx][T]
 *       iy[N][M], ix[N][M]: operator index for slice n and view-row j / view-column i (device)
 *       xlo_host[M], xhi_host[M]: per view column, smallest / largest coefficient offset relative to the output
 *                                 pixel that any slice of the batch reads (sizes the shared-memory windows)
 *       view_zeroed_host / view_zeroed: the M*M mask on the host and on the device
 *       tmp: float32 scratch (N,M,S,T,3); out: float32 (N,S,T,3); mm: nullable fused min/max of the stack.
 * ------------------------------------------------------------------------------------------------ */
/* The summed UP-SAMPLED image of ONE refined slice, i.e. `final_img` of lfp_shiftandsum.py:123-124 before the down-sampling
 * of :135 - what the reference histogram-aligns, sRGB-encodes and writes as upscale_<a/M>.png (:126-132).
 *   coef: B-spline coefficients from pcb_refi_prefilter;  a: shift in up-sampled pixels;
 *   fx[T*M], wx[T*M][4] / fy[S*M], wy[S*M][4]: first coefficient and the four cubic B-spline weights of every up-sampled
 *   sample (img_resize(view, M), misc/data_proc.py:57-92; device);  tmp: float32 (M, S, T*M, 3);  out: float32 (S*M, T*M, 3).
 * A slow path (732 MB per slice at Illum size) for callers that want the file; contrast / sRGB: pcb_contrast_bounds +
 * pcb_contrast_srgb on `out`. */
int pcb_refi_upscaled_slice(const float *coef, int M, int S, int T, int C, int a,
                            const int32_t *fx, const float *wx, const int32_t *fy, const float *wy,
                            const uint8_t *view_zeroed_host, float *tmp, float *out, void *stream);

/* Row restriction of the refinement passes (host struct, nullable everywhere: NULL = the whole image).  For a ROW-sharded
 * stack a rank produces the output rows [out_lo, out_hi) of every slice (out_lo a multiple of 8; out_hi too unless it is S);
 * for that it needs the rows [h_lo, h_hi) of the horizontal pass = of the B-spline coefficients, and the prefilter's x pass
 * for the rows [tmp_lo, tmp_hi) (plenopticam_b200/lfp_refocuser/refinement.py::RefinementPlan.rows derives all three). */
typedef struct { int32_t out_lo, out_hi, h_lo, h_hi, tmp_lo, tmp_hi; } pcb_refi_rows_t;
int pcb_refi_prefilter(const void *vp, int dtype, int M, int S, int T, int C,
                       const float *py_vals, const int32_t *py_first, int Kpy,
                       const float *px_vals, const int32_t *px_first, int Kpx,
                       const uint8_t *view_zeroed, float *tmp, float *coef, void *stream);
/* pcb_refi_prefilter with the interior of both operators given as ONE symmetric filter: away from the border the rows of
 * P are the cubic B-spline prefilter sqrt(3) (sqrt(3) - 2)^|k| (29 taps at 1e-8).  hx_host / hy_host: HOST arrays of
 * ntx = nty = 29 taps; rows x_lo..x_hi (y_lo..y_hi) of the tables equal them exactly with first = row - cx (cyy).  Those
 * rows run from the kernel-parameter bank with a register window (same summation order: bit-identical to the table
 * kernels), the border rows keep the tables.  A NULL / mismatching description falls back to the tables. */
int pcb_refi_prefilter_toeplitz(const void *vp, int dtype, int M, int S, int T, int C,
                                const float *py_vals, const int32_t *py_first, int Kpy,
                                const float *px_vals, const int32_t *px_first, int Kpx,
                                const uint8_t *view_zeroed, float *tmp, float *coef,
                                const float *hx_host, int ntx, int cx, int x_lo, int x_hi,
                                const float *hy_host, int nty, int cyy, int y_lo, int y_hi,
                                const pcb_refi_rows_t *rows_host /* nullable: see below */, void *stream);
int pcb_refocus_refi(const float *coef, int M, int S, int T, int C, int N,
                     const float *dy_vals, const int32_t *dy_first, int Kgy,
                     const float *dx_vals, const int32_t *dx_first, int Kx,
                     const int32_t *iy, const int32_t *ix,
                     const int32_t *xlo_host, const int32_t *xhi_host,
                     const uint8_t *view_zeroed_host, const uint8_t *view_zeroed,
                     float *tmp, float *out, pcb_minmax_t *mm, void *stream);

/* ---------------------------------------------------------------------------------------------------
 * Multi-GPU, slice-sharded stack (no counterpart in the single-process reference): every GPU
 * evaluates its own block of refocus parameters and the kernel that finishes a slice stores it, tile by tile, into the
 * stack buffer of EVERY GPU of the job through peer-mapped pointers (NVLink / NVSwitch stores) - compute and exchange
 * in one kernel, no separate all-gather.
 *   pcb_peer_alloc / pcb_peer_free : a buffer other processes can map (plain cudaMalloc on the current device: IPC
 *                        handles cover whole allocations, so a sub-allocated pointer cannot be exported).  The one
 *                        place this library allocates.
 *   pcb_peer_export    : 64-byte handle of such a buffer (cudaIpcGetMemHandle), to be passed to the other processes
 *   pcb_peer_open / _close : map / unmap a peer's buffer from its handle (enables peer access on first use)
 *   pcb_fanout_t       : where a finished block goes: n destinations (index 0..n-1; the local buffer is one of them),
 *                        out[d] = address of the FIRST slice of this call inside destination d, mm[d] = that
 *                        destination's running min / max (nullable: all or none)
 *   pcb_refocus_refi_fanout : pcb_refocus_refi with the destinations in *dst_host (HOST struct, copied at the call)
 * ------------------------------------------------------------------------------------------------ */
#define PCB_MAX_PEERS 16
typedef struct { float *out[PCB_MAX_PEERS]; pcb_minmax_t *mm[PCB_MAX_PEERS]; int32_t n; int32_t reserved; } pcb_fanout_t;
/* Horizontal operators of a batch of slices regrouped in QUADS of four consecutive slices (host struct; nullable in
 * pcb_refocus_refi_fanout): per quad g, view column i and output pixel x ONE sample window [fu, fu + W[i]) covering the
 * bands of the four slices, and their weights padded with zeros to that window,
 *     wq[woff[i] + (g * W[i] + t) * T + x] = float4 { slice 4g, 4g+1, 4g+2, 4g+3 } of tap t      (device, float4 units)
 *     fu[(g * M + i) * T + x]                                                                    (device)
 * lo[i] / hi[i]: smallest / largest sample offset relative to x that any window of view column i touches.
 * With it the horizontal pass loads every sample once per quad instead of once per slice (same sums, bit for bit). */
typedef struct { const float *wq; const int32_t *fu; int32_t G; int32_t reserved;
                 int32_t woff[32]; int32_t W[32]; int32_t lo[32]; int32_t hi[32]; } pcb_refi_quads_t;
int pcb_peer_alloc(int64_t bytes, void **dptr_host);
int pcb_peer_free(void *dptr);
int pcb_peer_export(void *dptr, unsigned char handle_host[64]);
int pcb_peer_open(const unsigned char handle_host[64], void **dptr_host);
int pcb_peer_close(void *dptr);
int pcb_refocus_refi_fanout(const float *coef, int M, int S, int T, int C, int N,
                            const float *dy_vals, const int32_t *dy_first, int Kgy,
                            const float *dx_vals, const int32_t *dx_first, int Kx,
                            const int32_t *iy, const int32_t *ix,
                            const int32_t *xlo_host, const int32_t *xhi_host,
                            const uint8_t *view_zeroed_host, const uint8_t *view_zeroed,
                            float *tmp, const pcb_fanout_t *dst_host, const pcb_refi_quads_t *quads_host /* nullable */,
                            const pcb_refi_rows_t *rows_host /* nullable */, void *stream);

/* ---------------------------------------------------------------------------------------------------
 * Post-processing of the stack — replaces LfpContrast.auto_hist_align + GammaConverter.srgb_conv as used
 * by LfpRefocuser.shift_and_sum (lfp_refocuser/top_level.py:68-70; lfp_extractor/lfp_contrast.py:249-263;
 * misc/normalizer.py:53-78; misc/gamma_converter.py:32-44).
 *   pcb_percentile_f32 : numpy-style linear-interpolated percentiles (exact order statistics by radix
 *                        select).  luma != 0 -> values are BT.709 luma of C=3 pixels (rgb2gry), n pixels.
 *                        `q_host` are percents in [0,100]; results (float64) go to `result` (device, nq).
 *                        `scratch` >= pcb_percentile_scratch_bytes() bytes.
 *   pcb_contrast_srgb  : y = clip((x-lo)/(hi-lo),0,1) evaluated in float64, cast to float32, then sRGB OETF.
 *                        lo/hi are read from the device (`lohi` = 2 doubles); a bound that is exactly 0 is
 *                        replaced by the data min / max in *mm (Normalizer's falsy fallback, normalizer.py:40-41).
 * ------------------------------------------------------------------------------------------------ */
int64_t pcb_percentile_scratch_bytes(void);
int pcb_percentile_f32(const float *data, int64_t n, int luma, const double *q_host, int nq,
                       double *result, void *scratch, void *stream);
int pcb_percentile_f64(const double *data, int64_t n, int luma, const double *q_host, int nq,
                       double *result, void *scratch, void *stream);   /* 8 x 8-bit passes over 64-bit keys */
/* Both bounds of LfpContrast.auto_hist_align (lfp_extractor/lfp_contrast.py:252-255) for a 3-channel reference image
 * in ONE cooperative launch: lohi[0] = percentile(rgb2gry(ref), q_lo), lohi[1] = percentile(ref, q_hi); same exact
 * order statistics as two pcb_percentile_f32 calls (13 launches).  scratch >= pcb_percentile_scratch_bytes(). */
int pcb_contrast_bounds(const float *ref, int64_t npx, double q_lo, double q_hi, double *lohi,
                        void *scratch, void *stream);
int pcb_contrast_srgb(const float *in, int64_t n, const double *lohi,
                      const pcb_minmax_t *mm /* nullable: fallback for a bound that is exactly 0 */,
                      float *out, void *stream);
/* Same with a uint16 / uint8 / float32 input (dtype PCB_U16 / PCB_U8 / PCB_F32): LfpContrast.main of the extractor
 * (lfp_extractor/lfp_contrast.py:43-67) runs auto_hist_align + srgb_conv over the integer-typed 4-D light field;
 * Normalizer reads it as float32 (misc/normalizer.py:36), which is exact for these types. */
int pcb_contrast_srgb_typed(const void *in, int dtype, int64_t n, const double *lohi, const pcb_minmax_t *mm,
                            float *out, void *stream);

/* ---------------------------------------------------------------------------------------------------
 * Global rectification ("next" row; the reference's default smp_meth) — device side of
 * LfpGlobalResampler (lfp_aligner/lfp_global_resampler.py:32-167).  Matrices and spline operators come from
 * plenopticam_b200/lfp_aligner/lfp_global_resampler.py (host, float64).
 *   pcb_image_minmax    : min / max of an image of any element type (clip range of the warp)
 *   pcb_warp_projective : skimage.transform.warp(img, ProjectiveTransform(tra_mat).inverse, output_shape) (:81-82):
 *                         (x,y,z) = inv_matrix @ (col,row,1), bilinear, outside = 0, clipped to the input range.
 *                         inv_matrix_host: 9 doubles, row major.  out: float32 (rows, cols, C).
 *                         skimage is not installed in the build container: parity of this call is UNPINNED.
 *   pcb_hex_align       : hexagonal_alignment + hex_stretch + per-strip misc.img_resize (:84-167) for even (py, px):
 *                         i0/i1/w[nk]: stretch lerp; wx_vals[K][x_len] / wx_first[x_len]: banded x resize of a strip of
 *                         nk*px samples; wy[y_len][py]: dense y resize; kept columns [crop_lo, crop_lo + out_cols).
 *                         stretched: float32 (strips*py, nk*px, C) scratch; tmpx: float32 (strips*py, out_cols, C)
 *                         scratch; out: float32 (strips*y_len, out_cols, C).
 * ------------------------------------------------------------------------------------------------ */
int pcb_image_minmax(const void *img, int dtype, int64_t n, pcb_minmax_t *mm, void *stream);
int pcb_warp_projective(const void *img, int dtype, int H, int W, int C, const double *inv_matrix_host,
                        int rows, int cols, const pcb_minmax_t *mm, float *out, void *stream);
int pcb_hex_align(const float *prj, int prows, int pcols, int C, int strips, int py, int px, int hex_odd, int nk,
                  const int32_t *i0, const int32_t *i1, const float *w,
                  const float *wx_vals, const int32_t *wx_first, int K, int x_len, int crop_lo, int out_cols,
                  const float *wy, int y_len, float *stretched, float *tmpx, float *out, void *stream);

/* ---------------------------------------------------------------------------------------------------
 * Lytro raw bit-unpacking ("next" row) — replaces LfpDecoder.comp_bayer
 * (lfp_reader/lfp_decoder.py:273-326).  buf: packed bytes on the device; bits = cfg.lfpimg['bit'] (10: 5 bytes ->
 * 4 pixels, 12: 3 bytes -> 2 pixels); out: floor(nbytes/5)*4 or floor(nbytes/3)*2 uint16 pixels in file order.
 * ------------------------------------------------------------------------------------------------ */
int pcb_unpack_bayer(const uint8_t *buf, int64_t nbytes, int bits, uint16_t *out, void *stream);

/* ---------------------------------------------------------------------------------------------------
 * Bayer white balance with highlight protection and soft clipping ("next" row f1) — replaces
 * CfaProcessor.safe_bayer_awb (lfp_aligner/cfa_processor.py:233-250) = _reshape_bayer (:96-114) +
 * _correct_bayer_highlights (:171-203) + apply_awb (:142-169) + clamp + soft_clipping (:252-255) on a 2-D mosaic.
 *   bay, wht : (H, W) float32 on the device (wht may be NULL: a white image of ones), H and W even
 *   gains_host[4] : HOST pointer, gains of the quad positions (0,0) (0,1) (1,0) (1,1) as CfaProcessor.set_gains
 *                   (:257-279) orders them for the Bayer pattern
 *   soft_clip : 1 when cfg.lfpimg has 'exp'; max_lum = 2^-exp; b = exp(7), log1pb = log(1 + exp(7)) as the caller's
 *               libm rounds them (so the constants equal numpy's)
 *   out : (H, W) float64.  Bit-identical to the reference up to the soft clipping (exp/log: <= 1 ulp apart).
 * ------------------------------------------------------------------------------------------------ */
int pcb_cfa_safe_awb(const float *bay, const float *wht, int H, int W, const double *gains_host, int soft_clip,
                     double max_lum, double b, double log1pb, double *out, void *stream);

/* ---------------------------------------------------------------------------------------------------
 * Scheimpflug composite ("next" row) — replaces LfpScheimpflug.scheimpflug_from_stack
 * (lfp_refocuser/lfp_scheimpflug.py:68-114): out[y,x,:] = stack[a_map[y,x]][y,x,:] with
 *   mode 0: a_map = a_y[y]   mode 1: a_map = a_x[x]   mode 2: a_map = (a_x[x] + a_y[y]) / 2  (integer mean, :94)
 * stack (N,S,T,C) float32, a_y[S], a_x[T] int32 (device), out (S,T,C) float32.  The uint16 image the reference
 * saves is pcb_minmax_f32 + pcb_quantize_u16 of `out` (misc/normalizer.py:43-46).
 * ------------------------------------------------------------------------------------------------ */
int pcb_scheimpflug(const float *stack, int N, int S, int T, int C, const int32_t *a_y, const int32_t *a_x,
                    int mode, float *out, void *stream);

/* ---------------------------------------------------------------------------------------------------
 * De-vignetting (first "next" row of SURVEY §8f; runs immediately before the resampler) — replaces
 * LfpDevignetter.__init__ / main / wht_img_divide / _estimate_noise_level
 * (lfp_aligner/lfp_devignetter.py:33-57,59-93,160-176).  float64 like the reference; divisions are IEEE, so the
 * de-vignetted image is bit-identical.
 *   pcb_devig_white       : wgray[npx] = rgb2gry(wht) (C == 3: ((r*0.2126 + g*0.7152) + b*0.0722)) or channel 0
 *   pcb_percentile_f64    : the 99.9th percentile of wgray (exact order statistics, see above)
 *   pcb_devig_normalize   : wgray /= *divisor                     (divisor on the device: no host round trip)
 *   pcb_devig_noise_level : std(wn - convolve2d(wn, g (x) g, 'same')) -> stats[2]  (stats: 3 doubles; tmp: H*W doubles)
 *                           gauss_host: normalised 1-D Gaussian of odd length glen <= 64 (misc/data_proc.py:31-42)
 *   pcb_devig_divide      : out[i,c] = wn[i] != 0 ? lfp[i,c] / wn[i] : 0;  out_dtype PCB_F64 (API parity) or
 *                           PCB_F32 (input of the resampler when the image stays on the device)
 * ------------------------------------------------------------------------------------------------ */
int pcb_devig_white(const void *wht, int dtype, int64_t npx, int C, double *wgray, void *stream);
int pcb_devig_normalize(double *wgray, int64_t npx, const double *divisor, void *stream);
int pcb_devig_noise_level(const double *wn, int H, int W, const double *gauss_host, int glen,
                          double *tmp, double *stats, void *stream);
int pcb_devig_divide(const void *lfp, int dtype, int64_t npx, int C, const double *wn,
                     void *out, int out_dtype, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* PLENOPTICAM_B200_H */
