"""Device-level operators: torch tensors in HBM in, torch tensors out, all work done by libpcb200.so.

torch is used for device memory, streams and pinned staging only.  Every function enqueues on the current
torch stream and does not synchronise.  There is no CPU fallback: without a CUDA device these raise.
"""
import ctypes as C

import os

import numpy as np
import torch

from . import _native as nat

_TORCH2PCB = {torch.uint16: nat.PCB_U16, torch.float32: nat.PCB_F32, torch.uint8: nat.PCB_U8}
_NP2TORCH = {np.dtype("uint16"): torch.uint16, np.dtype("float32"): torch.float32, np.dtype("uint8"): torch.uint8}


def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("plenopticam_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr())


def _pcb_dtype(t):
    try:
        return _TORCH2PCB[t.dtype]
    except KeyError:
        raise TypeError("unsupported device dtype %s (uint8, uint16, float32 are)" % t.dtype)


def _chk_dev(*tensors):
    for t in tensors:
        if not (t.is_cuda and t.is_contiguous()):
            raise ValueError("expected contiguous CUDA tensors")


# ------------------------------------------------------------------------------------------------------
# host <-> device staging
# ------------------------------------------------------------------------------------------------------
def device_dtype_for(arr):
    """Narrowest dtype the kernels read that holds `arr` exactly enough: uint8/uint16 stay, every float and
    wider integer goes to float32 unless it is integer-valued within uint16 (post-demosaic RGB is uint16
    stored as float64: lfp_aligner/top_level.py:39, cfa_processor.py:73)."""
    dt = np.asarray(arr).dtype
    if dt in (np.dtype("uint8"), np.dtype("uint16")):
        return dt
    return np.dtype("float32")


class HostArray(np.ndarray):
    """numpy view of a PINNED host buffer (a torch tensor it keeps alive in `_pcb_pin`).  `to_host` hands these out, so
    that the next stage's `to_device` can copy straight from the page-locked memory (no staging pass, full PCIe rate).
    The data is always read from host memory at call time - nothing is cached on the device - so modifying the array in
    place is safe; copies / arrays derived by arithmetic are ordinary pageable arrays and take the staged path."""
    _pcb_pin = None

    def __array_finalize__(self, obj):
        self._pcb_pin = getattr(obj, '_pcb_pin', None)


def _pinned_alias(arr):
    """torch view of `arr`'s bytes if they lie inside the pinned buffer `arr` descends from (and are contiguous)."""
    pin = getattr(arr, '_pcb_pin', None)
    if pin is None or not arr.flags.c_contiguous or arr.nbytes == 0:
        return None
    off = arr.ctypes.data - pin.data_ptr()
    if off < 0 or off + arr.nbytes > pin.numel() * pin.element_size():
        return None
    tdt = _NP2TORCH.get(arr.dtype)
    if tdt is None:
        return None
    return pin.view(torch.uint8).reshape(-1)[off:off + arr.nbytes].view(tdt).view(tuple(arr.shape))


def to_device(arr, dtype=None, exact_u16=False):
    """numpy -> contiguous CUDA tensor: straight from the pinned buffer for arrays `to_host` handed out, else through
    a pinned staging buffer."""
    dev = require_cuda()
    arr = np.asarray(arr) if not isinstance(arr, np.ndarray) else arr
    if (dtype is None or np.dtype(dtype) == arr.dtype) and arr.dtype in _NP2TORCH:
        alias = _pinned_alias(arr)
        if alias is not None:
            return alias.to(dev, non_blocking=True)
    arr = np.asarray(arr)
    if dtype is None:
        dtype = device_dtype_for(arr)
        if exact_u16 and dtype == np.dtype("float32") and arr.size:
            lo, hi = arr.min(), arr.max()
            if lo >= 0 and hi <= 65535 and np.array_equal(arr, np.rint(arr)):
                dtype = np.dtype("uint16")
    host = np.ascontiguousarray(arr, dtype=dtype)
    if host.dtype == np.dtype("uint16"):
        pinned = torch.empty(host.shape, dtype=torch.uint16, pin_memory=True)
        pinned.view(torch.int16).numpy()[...] = host.view(np.int16)
    else:
        pinned = torch.empty(host.shape, dtype=_NP2TORCH[host.dtype], pin_memory=True)
        pinned.numpy()[...] = host
    return pinned.to(dev, non_blocking=True)


def to_host(t):
    """CUDA tensor -> numpy (synchronises the current stream).  The result lives in pinned memory (`HostArray`): the
    copy runs at the full PCIe rate and a later `to_device` of the same array needs no staging.  torch's caching host
    allocator recycles the buffers once the arrays are dropped."""
    if not t.is_cuda:
        return t.numpy()
    t = t.contiguous()
    pinned = torch.empty(tuple(t.shape), dtype=t.dtype, pin_memory=True)
    pinned.copy_(t, non_blocking=True)
    torch.cuda.current_stream(t.device).synchronize()
    if t.dtype == torch.uint16:
        arr = pinned.view(torch.int16).numpy().view(np.uint16)
    else:
        arr = pinned.numpy()
    out = arr.view(HostArray)
    out._pcb_pin = pinned
    return out


def struct_to_device(arr):
    """Upload a numpy structured array (pcb_lens_t / pcb_hexcol_t tables) as raw bytes."""
    dev = require_cuda()
    raw = np.ascontiguousarray(arr).view(np.uint8).reshape(-1)
    return torch.from_numpy(raw.copy()).to(dev)


# ------------------------------------------------------------------------------------------------------
# min / max
# ------------------------------------------------------------------------------------------------------
def new_minmax():
    dev = require_cuda()
    mm = torch.empty(2, dtype=torch.int32, device=dev)
    nat.check(nat.lib().pcb_minmax_reset(_ptr(mm), _stream()))
    return mm


def minmax_values(mm):
    """(min, max) as Python floats (synchronises)."""
    keys = mm.cpu().numpy().view(np.uint32)
    L = nat.lib()
    return float(L.pcb_key_to_float(int(keys[0]))), float(L.pcb_key_to_float(int(keys[1])))


def minmax_f32(data, mm=None):
    _chk_dev(data)
    mm = new_minmax() if mm is None else mm
    nat.check(nat.lib().pcb_minmax_f32(_ptr(data), data.numel(), _ptr(mm), _stream()))
    return mm


# ------------------------------------------------------------------------------------------------------
# (a) resampling
# ------------------------------------------------------------------------------------------------------
class LensTables:
    """Device copies of the host-prepared tables (see microlens.build_tables)."""

    def __init__(self, tables):
        self.LY, self.LX, self.M = tables["LY"], tables["LX"], tables["M"]
        self.pattern = nat.PCB_HEX if tables["pat_type"] == "hex" else nat.PCB_REC
        self.Xs = tables["Xs"]
        self.hex_odd = int(bool(tables["hex_odd"]))
        self.lens = struct_to_device(tables["lens"])
        self.hexcol = struct_to_device(tables["hexcol"]) if tables["hexcol"] is not None else None

    def _args(self, raw, flip):
        H, W, Cn = raw.shape
        hexp = _ptr(self.hexcol) if self.hexcol is not None else C.c_void_p(0)
        return [_ptr(raw), _pcb_dtype(raw), H, W, Cn, _ptr(self.lens), self.LY, self.LX, self.M,
                self.pattern, hexp, self.Xs, self.hex_odd, int(bool(flip))]

    def out_shape(self, raw):
        return (self.LY * self.M, self.Xs * self.M, raw.shape[2])


def resample_f32(raw, tabs, flip=False, want_minmax=False):
    """float32 aligned image (LY*M, Xs*M, C) [+ running min/max keys]."""
    _chk_dev(raw)
    out = torch.empty(tabs.out_shape(raw), dtype=torch.float32, device=raw.device)
    mm = new_minmax() if want_minmax else None
    args = tabs._args(raw, flip) + [_ptr(out), _ptr(mm) if mm is not None else C.c_void_p(0), _stream()]
    nat.check(nat.lib().pcb_resample_f32(*args))
    return (out, mm) if want_minmax else out


def resample_u16(raw, tabs, flip=False, out=None, mm=None):
    """LfpResampler's product: the aligned image min-max normalised to uint16.  Two launches, no host sync:
    a min/max pass and a recompute-and-quantise pass."""
    _chk_dev(raw)
    if out is None:
        out = torch.empty(tabs.out_shape(raw), dtype=torch.uint16, device=raw.device)
    if mm is None:
        mm = torch.empty(2, dtype=torch.int32, device=raw.device)
    L = nat.lib()
    nat.check(L.pcb_minmax_reset(_ptr(mm), _stream()))
    nat.check(L.pcb_resample_minmax(*(tabs._args(raw, flip) + [_ptr(mm), _stream()])))
    nat.check(L.pcb_resample_u16(*(tabs._args(raw, flip) + [_ptr(mm), _ptr(out), _stream()])))
    return out, mm


def quantize_u16(data, mm, out=None):
    _chk_dev(data)
    if out is None:
        out = torch.empty(data.shape, dtype=torch.uint16, device=data.device)
    nat.check(nat.lib().pcb_quantize_u16(_ptr(data), data.numel(), _ptr(mm), _ptr(out), _stream()))
    return out


def quantize_u8(data, mm, out=None):
    _chk_dev(data)
    if out is None:
        out = torch.empty(data.shape, dtype=torch.uint8, device=data.device)
    nat.check(nat.lib().pcb_quantize_u8(_ptr(data), data.numel(), _ptr(mm), _ptr(out), _stream()))
    return out


def uint16_norm(data, out=None):
    """misc.Normalizer(data).uint16_norm() of a float32 CUDA tensor (global min / max; misc/normalizer.py:43-46):
    what LfpExporter writes for viewpoints and refocus slices (lfp_extractor/lfp_exporter.py:84,134)."""
    return quantize_u16(data, minmax_f32(data), out=out)


def uint8_norm(data, out=None):
    """misc.Normalizer(data).uint8_norm() (misc/normalizer.py:48-51; lfp_exporter.py:195)."""
    return quantize_u8(data, minmax_f32(data), out=out)


# ------------------------------------------------------------------------------------------------------
# (a') rotation
# ------------------------------------------------------------------------------------------------------
def rotate(img, rad):
    """scipy.ndimage.rotate(img, deg, reshape=False, order=3) -> float32 (H,W,C)."""
    _chk_dev(img)
    H, W, Cn = img.shape
    tmp = torch.empty((H, W, Cn), dtype=torch.float32, device=img.device)
    coef = torch.empty_like(tmp)
    L = nat.lib()
    nat.check(L.pcb_spline_prefilter(_ptr(img), _pcb_dtype(img), H, W, Cn, _ptr(tmp), _ptr(coef), _stream()))
    nat.check(L.pcb_rotate_sample(_ptr(coef), H, W, Cn, float(rad), _ptr(tmp), _stream()))
    return tmp


# ------------------------------------------------------------------------------------------------------
# (b) viewpoints
# ------------------------------------------------------------------------------------------------------
def viewpoints(aligned, M_in, M_out=None, out=None):
    _chk_dev(aligned)
    M_out = M_in if M_out is None else M_out
    rows, cols, Cn = aligned.shape
    S, T = rows // M_in, cols // M_in
    if out is None:
        out = torch.empty((M_out, M_out, S, T, Cn), dtype=aligned.dtype, device=aligned.device)
    nat.check(nat.lib().pcb_viewpoints(_ptr(aligned), _pcb_dtype(aligned), rows, cols, Cn, M_in, M_out,
                                       _ptr(out), _stream()))
    return out


def viewpoints_inverse(vp):
    _chk_dev(vp)
    M, M2, S, T, Cn = vp.shape
    if M != M2:
        raise ValueError("square angular resolution expected")
    out = torch.empty((S * M, T * M, Cn), dtype=vp.dtype, device=vp.device)
    nat.check(nat.lib().pcb_viewpoints_inverse(_ptr(vp), _pcb_dtype(vp), M, S, T, Cn, _ptr(out), _stream()))
    return out


def crop_micro_images(aligned, M_in, M_out):
    """Central M_out x M_out crop of every micro image; M_in is the aligned pitch, a scalar or (My, Mx)."""
    _chk_dev(aligned)
    rows, cols, Cn = aligned.shape
    My, Mx = (int(M_in), int(M_in)) if np.ndim(M_in) == 0 else (int(M_in[0]), int(M_in[1]))
    LY, LX = rows // My, cols // Mx
    out = torch.empty((LY * M_out, LX * M_out, Cn), dtype=aligned.dtype, device=aligned.device)
    nat.check(nat.lib().pcb_crop_micro_images_yx(_ptr(aligned), _pcb_dtype(aligned), rows, cols, Cn, My, Mx, M_out,
                                                 _ptr(out), _stream()))
    return out


# ------------------------------------------------------------------------------------------------------
# (c) refocus
# ------------------------------------------------------------------------------------------------------
def aperture_mask(M):
    """View (j,i) is zeroed iff int(np.round(sqrt((i-c)^2+(j-c)^2))) > c
    (lfp_extractor/lfp_viewpoints.py:229-250; np.round is round-half-even)."""
    c = M // 2
    jj, ii = np.meshgrid(np.arange(M) - c, np.arange(M) - c, indexing="ij")
    return (np.rint(np.sqrt(ii.astype(np.float64) ** 2 + jj.astype(np.float64) ** 2)).astype(np.int64) > c)


class RefocusPlan(object):
    """Host control arrays + device scratch for pcb_refocus_int; reusable across exposures of one camera.

    aligned_pitch: None - `launch*` take the 4-D light field vp[j,i,s,t,c];  Ma (>= M) - they take LfpResampler's aligned
    micro-image array (S*Ma, T*Ma, 3) and the viewpoint gather / central crop is folded into the kernel's staging
    (pcb_refocus_int_aligned); NotImplementedError when the library has no such form for the configuration."""

    def __init__(self, vp_shape, vp_dtype, a_list, view_zeroed, device, aligned_pitch=None):
        M, M2, S, T, Cn = vp_shape
        self.Ma = None if aligned_pitch is None else int(aligned_pitch)
        if M != M2:
            raise ValueError("square angular resolution expected")
        self.shape = (int(M), int(S), int(T), int(Cn))
        self.dtype = _TORCH2PCB[vp_dtype]
        self.a_host = np.ascontiguousarray(np.asarray(list(a_list), dtype=np.int32))
        self.N = int(self.a_host.size)
        if view_zeroed is None:
            view_zeroed = aperture_mask(M)
        self.z_host = np.ascontiguousarray(np.asarray(view_zeroed, dtype=np.uint8).reshape(-1))
        if self.z_host.size != M * M:
            raise ValueError("view mask must have M*M entries")
        self.a_ptr = self.a_host.ctypes.data_as(C.c_void_p)
        self.z_ptr = self.z_host.ctypes.data_as(C.c_void_p)
        nbytes = int(nat.lib().pcb_refocus_int_scratch_bytes(self.dtype, M, S, T, Cn, self.a_ptr, self.N, self.z_ptr))
        if nbytes < 0:
            raise ValueError("invalid refocus configuration")
        self.scratch = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=device)
        self.fused = None            # None: not tried yet; True / False: pcb_refocus_int_srgb applies / does not
        self.ref_slice = None

    # The accumulators in `scratch` must be zero when the kernel starts.  The aligned entry points can hand them back
    # zeroed (the finalize pass stores zero behind every word it reads) and skip their memset the next time; the state
    # travels with the scratch TENSOR, because plans of one geometry share it (LightFieldPipeline).
    SCRATCH_ZEROED, LEAVE_ZEROED = 1, 2

    def _take_scratch(self, keep_clean):
        """Flags for the next launch; the scratch counts as dirty until `_scratch_left_clean` says otherwise (an
        exception in between leaves it dirty: the next launch zeroes it itself)."""
        was_clean = bool(getattr(self.scratch, "_pcb_acc_zero", False))
        self.scratch._pcb_acc_zero = False
        if not keep_clean or os.environ.get("PCB_REFOCUS_MEMSET") == "1":      # A/B knob: always memset
            return 0
        return self.LEAVE_ZEROED | (self.SCRATCH_ZEROED if was_clean else 0)

    def _scratch_left_clean(self):
        self.scratch._pcb_acc_zero = True

    def launch(self, vp, out, mm=None):
        M, S, T, Cn = self.shape
        mmp = _ptr(mm) if mm is not None else C.c_void_p(0)
        if self.Ma is not None:
            flags = self._take_scratch(True)
            nat.check(nat.lib().pcb_refocus_int_aligned(_ptr(vp), self.dtype, self.Ma, M, S, T, Cn, self.a_ptr, self.N,
                                                        self.z_ptr, _ptr(self.scratch), self.scratch.numel(), _ptr(out),
                                                        mmp, flags, _stream()))
            self._scratch_left_clean()
            return out
        self._take_scratch(False)
        nat.check(nat.lib().pcb_refocus_int(_ptr(vp), self.dtype, M, S, T, Cn, self.a_ptr, self.N, self.z_ptr,
                                            _ptr(self.scratch), self.scratch.numel(), _ptr(out), mmp, _stream()))
        return out


def _refocus_plan_launch_srgb(self, vp, out, mm, lohi, pscratch, q_lo=0.005, q_hi=99.9):
    """Stack with LfpRefocuser's colour management fused into the epilogue (pcb_refocus_int_srgb).  Returns False when
    the library has no fused form for this configuration (the caller then runs the separate kernels)."""
    if self.fused is False:
        return False
    M, S, T, Cn = self.shape
    if self.ref_slice is None:
        self.ref_slice = torch.empty((S, T, Cn), dtype=torch.float32, device=self.scratch.device)
    try:
        if self.Ma is not None:
            flags = self._take_scratch(True)
            nat.check(nat.lib().pcb_refocus_int_srgb_aligned(_ptr(vp), self.dtype, self.Ma, M, S, T, Cn, self.a_ptr,
                                                             self.N, self.z_ptr, _ptr(self.scratch), self.scratch.numel(),
                                                             q_lo, q_hi, _ptr(self.ref_slice), _ptr(lohi), _ptr(pscratch),
                                                             _ptr(out), _ptr(mm), flags, _stream()))
            self._scratch_left_clean()
        else:
            self._take_scratch(False)
            nat.check(nat.lib().pcb_refocus_int_srgb(_ptr(vp), self.dtype, M, S, T, Cn, self.a_ptr, self.N, self.z_ptr,
                                                     _ptr(self.scratch), self.scratch.numel(), q_lo, q_hi,
                                                     _ptr(self.ref_slice), _ptr(lohi), _ptr(pscratch), _ptr(out),
                                                     _ptr(mm), _stream()))
    except NotImplementedError:
        self.fused = False
        return False
    self.fused = True
    return True


RefocusPlan.launch_srgb = _refocus_plan_launch_srgb


def refocus_int(vp, a_list, view_zeroed=None, out=None, mm=None, plan=None):
    """(N,S,T,C) float32 stack of integer shift-and-sum slices (divided by M); `mm` (from new_minmax())
    optionally receives the stack's min/max from the same launch."""
    _chk_dev(vp)
    if plan is None:
        plan = RefocusPlan(tuple(vp.shape), vp.dtype, a_list, view_zeroed, vp.device)
    M, S, T, Cn = plan.shape
    if out is None:
        out = torch.empty((plan.N, S, T, Cn), dtype=torch.float32, device=vp.device)
    return plan.launch(vp, out, mm)


def refocus_int_from_aligned(aligned, M_aligned, a_list, M=None, view_zeroed=None, out=None, mm=None, plan=None):
    """refocus_int(viewpoints(aligned, M_aligned, M), a_list) without materialising the 4-D light field: the gather and
    the central crop happen inside the refocus kernel's staging.  Raises NotImplementedError when the library has no
    such form for the configuration (then gather first)."""
    _chk_dev(aligned)
    Ma = int(M_aligned)
    M = Ma if M is None else int(M)
    rows, cols, Cn = (int(v) for v in aligned.shape)
    S, T = rows // Ma, cols // Ma
    if plan is None:
        plan = RefocusPlan((M, M, S, T, Cn), aligned.dtype, a_list, view_zeroed, aligned.device, aligned_pitch=Ma)
    if out is None:
        out = torch.empty((plan.N, S, T, Cn), dtype=torch.float32, device=aligned.device)
    return plan.launch(aligned, out, mm)


def refocus_int_direct(vp, a_list, view_zeroed=None, out=None, mm=None):
    """Same result through the direct (slice-stationary) kernel only; used for A/B checks."""
    _chk_dev(vp)
    M, M2, S, T, Cn = vp.shape
    a_host = np.ascontiguousarray(np.asarray(list(a_list), dtype=np.int32))
    N = int(a_host.size)
    if view_zeroed is None:
        view_zeroed = aperture_mask(M)
    a_dev = torch.from_numpy(a_host).to(vp.device)
    z_dev = torch.from_numpy(np.ascontiguousarray(view_zeroed, dtype=np.uint8).reshape(-1)).to(vp.device)
    if out is None:
        out = torch.empty((N, S, T, Cn), dtype=torch.float32, device=vp.device)
    nat.check(nat.lib().pcb_refocus_int_direct(_ptr(vp), _pcb_dtype(vp), M, S, T, Cn, _ptr(a_dev), N, _ptr(z_dev),
                                               _ptr(out), _ptr(mm) if mm is not None else C.c_void_p(0), _stream()))
    return out


def refocus_from_scratch(aligned, M, a_list, out=None):
    """LfpShiftAndSum.refo_from_scratch (lfp_refocuser/lfp_shiftandsum.py:141-221) for an aligned image (H*M, J*M, P):
    shift-and-sum over ALL views with zeros outside the light field; float32 (N, H, J, P)."""
    _chk_dev(aligned)
    vp = viewpoints(aligned, int(M))
    Mi, _, S, T, Cn = (int(v) for v in vp.shape)
    a_host = np.ascontiguousarray(np.asarray(list(a_list), dtype=np.int32))
    a_dev = torch.from_numpy(a_host).to(vp.device)
    z_dev = torch.zeros(Mi * Mi, dtype=torch.uint8, device=vp.device)
    if out is None:
        out = torch.empty((int(a_host.size), S, T, Cn), dtype=torch.float32, device=vp.device)
    nat.check(nat.lib().pcb_refocus_int_zero(_ptr(vp), _pcb_dtype(vp), Mi, S, T, Cn, _ptr(a_dev), int(a_host.size),
                                             _ptr(z_dev), _ptr(out), C.c_void_p(0), _stream()))
    return out


# ------------------------------------------------------------------------------------------------------
# post-processing
# ------------------------------------------------------------------------------------------------------
def percentile(data, q, luma=False):
    """np.percentile(data, q) (or of the BT.709 luma of (...,3) pixels) -> float64 tensor of len(q)."""
    _chk_dev(data)
    if data.dtype != torch.float32:
        raise TypeError("float32 expected")
    L = nat.lib()
    q = [float(v) for v in np.atleast_1d(q)]
    n = data.numel() // 3 if luma else data.numel()
    if luma and data.shape[-1] != 3:
        raise ValueError("luma percentiles need 3 channels")
    scratch = torch.empty(int(L.pcb_percentile_scratch_bytes()), dtype=torch.uint8, device=data.device)
    res = torch.empty(len(q), dtype=torch.float64, device=data.device)
    qarr = (C.c_double * len(q))(*q)
    nat.check(L.pcb_percentile_f32(_ptr(data), n, int(bool(luma)), qarr, len(q), _ptr(res), _ptr(scratch), _stream()))
    return res


def contrast_bounds(ref, q_lo=0.005, q_hi=99.9, lohi=None, scratch=None):
    """[percentile(rgb2gry(ref), q_lo), percentile(ref, q_hi)] of a (..., 3) float32 reference image as a float64 CUDA
    tensor of 2 (lfp_extractor/lfp_contrast.py:252-255), one cooperative launch."""
    _chk_dev(ref)
    if ref.dtype != torch.float32 or ref.shape[-1] != 3:
        raise TypeError("float32 (..., 3) reference image expected")
    L = nat.lib()
    if lohi is None:
        lohi = torch.empty(2, dtype=torch.float64, device=ref.device)
    if scratch is None:
        scratch = torch.empty(int(L.pcb_percentile_scratch_bytes()), dtype=torch.uint8, device=ref.device)
    nat.check(L.pcb_contrast_bounds(_ptr(ref), ref.numel() // 3, float(q_lo), float(q_hi), _ptr(lohi), _ptr(scratch),
                                    _stream()))
    return lohi


def contrast_srgb(stack, lohi, mm=None, out=None):
    """clip((x-lo)/(hi-lo)) then sRGB; `lohi` is a float64 CUDA tensor [lo, hi]; a bound that is exactly 0
    is replaced on the device by the data min/max in `mm` (misc/normalizer.py:40-41)."""
    _chk_dev(stack, lohi)
    if out is None:
        out = torch.empty(tuple(stack.shape), dtype=torch.float32, device=stack.device)
    nat.check(nat.lib().pcb_contrast_srgb_typed(_ptr(stack), _pcb_dtype(stack), stack.numel(), _ptr(lohi),
                                                _ptr(mm) if mm is not None else C.c_void_p(0), _ptr(out), _stream()))
    return out


def image_minmax(data, mm=None):
    """Running min / max keys of a uint8 / uint16 / float32 CUDA tensor."""
    _chk_dev(data)
    mm = new_minmax() if mm is None else mm
    nat.check(nat.lib().pcb_image_minmax(_ptr(data), _pcb_dtype(data), data.numel(), _ptr(mm), _stream()))
    return mm


def viewpoints_contrast(vp, ref_view=None, out=None):
    """LfpContrast.main with its options off (lfp_extractor/lfp_contrast.py:62-65): histogram alignment of the whole
    light field against the central view (lo = P0.005 of its luma, hi = P99.9) and sRGB; float32 like the reference."""
    _chk_dev(vp)
    if ref_view is None:
        c = int(vp.shape[0]) // 2
        ref_view = vp[c, c]
    ref32 = ref_view.to(torch.float32).contiguous()
    lohi = contrast_bounds(ref32)
    return contrast_srgb(vp, lohi, mm=image_minmax(vp), out=out)


def refocus_int_srgb(vp, a_list, view_zeroed=None, out=None, plan=None):
    """LfpShiftAndSum + the colour management of LfpRefocuser.shift_and_sum in one go: float32 sRGB stack.  Uses the
    fused epilogue when the library offers it for this light field, else the separate kernels (same values)."""
    _chk_dev(vp)
    if plan is None:
        plan = RefocusPlan(tuple(vp.shape), vp.dtype, a_list, view_zeroed, vp.device)
    M, S, T, Cn = plan.shape
    if out is None:
        out = torch.empty((plan.N, S, T, Cn), dtype=torch.float32, device=vp.device)
    mm = new_minmax()
    lohi = torch.empty(2, dtype=torch.float64, device=vp.device)
    pscratch = torch.empty(int(nat.lib().pcb_percentile_scratch_bytes()), dtype=torch.uint8, device=vp.device)
    if plan.launch_srgb(vp, out, mm, lohi, pscratch):
        return out
    stack = plan.launch(vp, torch.empty_like(out), mm)
    return refocuser_postprocess(stack, mm=mm, out=out)


def refocuser_postprocess(stack, mm=None, out=None):
    """LfpRefocuser.shift_and_sum's colour management (lfp_refocuser/top_level.py:68-70), on the device and
    without a host round trip: lo = P0.005(luma(stack[0])), hi = P99.9(stack[0]), normalise, sRGB.
    `mm`: min/max keys of the stack (computed here if not supplied by the refocus launch)."""
    lohi = contrast_bounds(stack[0])
    if mm is None:
        mm = minmax_f32(stack)
    return contrast_srgb(stack, lohi, mm=mm, out=out)


# ------------------------------------------------------------------------------------------------------
# de-vignetting (lfp_aligner/lfp_devignetter.py)
# ------------------------------------------------------------------------------------------------------
_TORCH2PCB_F64 = dict(_TORCH2PCB)
_TORCH2PCB_F64[torch.float64] = nat.PCB_F64


def to_device_any(arr):
    """numpy -> CUDA tensor keeping uint8 / uint16 / float32 / float64 as they are (other dtypes -> float64)."""
    dev = require_cuda()
    arr = np.asarray(arr)
    if arr.dtype in (np.dtype("uint8"), np.dtype("uint16"), np.dtype("float32")):
        return to_device(arr, dtype=arr.dtype)
    host = np.ascontiguousarray(arr, dtype=np.float64)
    pinned = torch.empty(host.shape, dtype=torch.float64, pin_memory=True)
    pinned.numpy()[...] = host
    return pinned.to(dev, non_blocking=True)


def percentile_f64(data, q):
    """np.percentile of a float64 CUDA tensor (exact order statistics, 8 radix passes) -> float64 tensor of len(q)."""
    _chk_dev(data)
    if data.dtype != torch.float64:
        raise TypeError("float64 expected")
    L = nat.lib()
    q = [float(v) for v in np.atleast_1d(q)]
    scratch = torch.empty(int(L.pcb_percentile_scratch_bytes()), dtype=torch.uint8, device=data.device)
    res = torch.empty(len(q), dtype=torch.float64, device=data.device)
    qarr = (C.c_double * len(q))(*q)
    nat.check(L.pcb_percentile_f64(_ptr(data), data.numel(), 0, qarr, len(q), _ptr(res), _ptr(scratch), _stream()))
    return res


def devignette_divide(lfp, wn, out=None, out_dtype=torch.float32):
    """lfp / wn where wn != 0, else 0 (lfp_devignetter.py:83-93) with an already normalised gray white image `wn`
    ((H, W) float64 CUDA tensor, e.g. the second result of `devignette`): the per-exposure part of de-vignetting."""
    _chk_dev(lfp, wn)
    H, W = int(lfp.shape[0]), int(lfp.shape[1])
    Cl = int(lfp.shape[2]) if lfp.dim() == 3 else 1
    if wn.dtype != torch.float64 or tuple(wn.shape) != (H, W):
        raise TypeError("normalised white image must be a float64 (H, W) tensor")
    if out is None:
        out = torch.empty(tuple(lfp.shape), dtype=out_dtype, device=lfp.device)
    nat.check(nat.lib().pcb_devig_divide(_ptr(lfp), _TORCH2PCB_F64[lfp.dtype], H * W, Cl, _ptr(wn), _ptr(out),
                                         _TORCH2PCB_F64[out.dtype], _stream()))
    return out


def devignette(lfp, wht, gauss_1d, out_dtype=torch.float64):
    """White-image division on the device.

    lfp: (H,W[,C]) CUDA tensor (uint8/uint16/float32/float64); wht: (H,W[,Cw]) CUDA tensor, Cw in (1, 3);
    gauss_1d: normalised odd-length 1-D Gaussian (numpy float64) of the noise-level estimate.
    Returns (out (same shape as lfp, out_dtype), wn (H,W) float64 normalised gray white image, stats) where
    stats[2] is the noise level (float64 CUDA tensor of 3; no synchronisation here)."""
    _chk_dev(lfp, wht)
    L, st = nat.lib(), _stream()
    H, W = int(lfp.shape[0]), int(lfp.shape[1])
    Cl = int(lfp.shape[2]) if lfp.dim() == 3 else 1
    Cw = int(wht.shape[2]) if wht.dim() == 3 else 1
    if tuple(wht.shape[:2]) != (H, W) or Cw not in (1, 3):
        raise ValueError("white image must be (H, W), (H, W, 1) or (H, W, 3) with the light-field image's H, W")
    npx = H * W
    dev = lfp.device
    wn = torch.empty((H, W), dtype=torch.float64, device=dev)
    nat.check(L.pcb_devig_white(_ptr(wht), _TORCH2PCB_F64[wht.dtype], npx, Cw, _ptr(wn), st))
    p = percentile_f64(wn, [99.9])
    nat.check(L.pcb_devig_normalize(_ptr(wn), npx, _ptr(p), st))
    g = np.ascontiguousarray(gauss_1d, dtype=np.float64)
    tmp = torch.empty((H, W), dtype=torch.float64, device=dev)
    stats = torch.empty(3, dtype=torch.float64, device=dev)
    nat.check(L.pcb_devig_noise_level(_ptr(wn), H, W, g.ctypes.data_as(C.POINTER(C.c_double)), int(g.size), _ptr(tmp),
                                      _ptr(stats), st))
    out = torch.empty(tuple(lfp.shape), dtype=out_dtype, device=dev)
    nat.check(L.pcb_devig_divide(_ptr(lfp), _TORCH2PCB_F64[lfp.dtype], npx, Cl, _ptr(wn), _ptr(out),
                                 _TORCH2PCB_F64[out_dtype], st))
    return out, wn, stats


# ------------------------------------------------------------------------------------------------------
# Scheimpflug composite (lfp_refocuser/lfp_scheimpflug.py:68-114)
# ------------------------------------------------------------------------------------------------------
PFLU_VALS = ('vertical', 'horizontal', 'skew up', 'skew down')


def scheimpflug_maps(n_slices, S, T):
    """a_y (S,), a_x (T,) of lfp_scheimpflug.py:81-82: integer-typed linspace without its end points."""
    a_x = np.linspace(0, n_slices, T + 2, dtype='int')[1:-1]
    a_y = np.linspace(0, n_slices, S + 2, dtype='int')[1:-1]
    return a_y.astype(np.int32), a_x.astype(np.int32)


def scheimpflug_mode(direction):
    """Which map the reference ends up using (:86-94).  Its test `== (PFLU_VALS[2] or PFLU_VALS[3])` only ever
    matches 'skew up'; 'skew down' therefore keeps the default vertical map."""
    if direction == PFLU_VALS[1]:
        return 1
    if direction == PFLU_VALS[2]:
        return 2
    return 0


def scheimpflug(stack, direction):
    """(S,T,C) float32 composite and its Normalizer.uint16_norm (uint16 CUDA tensors stay on the device)."""
    _chk_dev(stack)
    if stack.dtype != torch.float32 or stack.dim() != 4:
        raise TypeError("float32 (N,S,T,C) stack expected")
    N, S, T, Cn = (int(v) for v in stack.shape)
    a_y, a_x = scheimpflug_maps(N, S, T)
    ay_d, ax_d = torch.from_numpy(a_y).to(stack.device), torch.from_numpy(a_x).to(stack.device)
    out = torch.empty((S, T, Cn), dtype=torch.float32, device=stack.device)
    nat.check(nat.lib().pcb_scheimpflug(_ptr(stack), N, S, T, Cn, _ptr(ay_d), _ptr(ax_d), scheimpflug_mode(direction),
                                        _ptr(out), _stream()))
    mm = minmax_f32(out)
    return out, quantize_u16(out, mm)


# ------------------------------------------------------------------------------------------------------
# global rectification (lfp_aligner/lfp_global_resampler.py)
# ------------------------------------------------------------------------------------------------------
def warp_projective(img, tra_mat, oshape):
    """skimage.transform.warp(img, ProjectiveTransform(tra_mat).inverse, output_shape=oshape) -> float32 (rows, cols, C)."""
    _chk_dev(img)
    H, W, Cn = (int(v) for v in img.shape)
    rows, cols = int(oshape[0]), int(oshape[1])
    L, st = nat.lib(), _stream()
    mm = new_minmax()
    nat.check(L.pcb_image_minmax(_ptr(img), _pcb_dtype(img), img.numel(), _ptr(mm), st))
    inv = np.ascontiguousarray(np.linalg.inv(np.asarray(tra_mat, dtype=np.float64)))
    out = torch.empty((rows, cols, Cn), dtype=torch.float32, device=img.device)
    nat.check(L.pcb_warp_projective(_ptr(img), _pcb_dtype(img), H, W, Cn, inv.ctypes.data_as(C.POINTER(C.c_double)),
                                    rows, cols, _ptr(mm), _ptr(out), st))
    return out


def hex_align(prj, plan, LY):
    """hexagonal_alignment of the rectified image with the host plan of lfp_global_resampler.hex_plan."""
    _chk_dev(prj)
    dev = prj.device
    prows, pcols, Cn = (int(v) for v in prj.shape)
    cache = plan.setdefault("_device", {})               # operator tables stay on the device with the (cached) plan
    if str(dev) not in cache:
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        cache[str(dev)] = tuple(t(plan[k]) for k in ("i0", "i1", "w", "wx_vals", "wx_first", "wy"))
    i0, i1, w, wxv, wxf, wy = cache[str(dev)]
    strips = int(LY) - 1
    out_cols = plan["x_len"] + plan["crop_hi"] - plan["crop_lo"]
    stretched = torch.empty((strips * plan["py"], plan["m"], Cn), dtype=torch.float32, device=dev)
    tmpx = torch.empty((strips * plan["py"], out_cols, Cn), dtype=torch.float32, device=dev)
    out = torch.empty((strips * plan["y_len"], out_cols, Cn), dtype=torch.float32, device=dev)
    nat.check(nat.lib().pcb_hex_align(_ptr(prj), prows, pcols, Cn, strips, plan["py"], plan["px"], plan["hex_odd"],
                                      plan["lens_new_x"], _ptr(i0), _ptr(i1), _ptr(w), _ptr(wxv), _ptr(wxf),
                                      int(plan["wx_vals"].shape[0]), plan["x_len"], plan["crop_lo"], out_cols, _ptr(wy),
                                      plan["y_len"], _ptr(stretched), _ptr(tmpx), _ptr(out), _stream()))
    return out


# ------------------------------------------------------------------------------------------------------
# Bayer white balance (lfp_aligner/cfa_processor.py:233-255)
# ------------------------------------------------------------------------------------------------------
def cfa_safe_awb(bay, wht, gains, exp_bias=None, out=None):
    """CfaProcessor.safe_bayer_awb on a 2-D float32 mosaic (CUDA): float64 (H, W).  `gains`: the four quad-position
    gains as CfaProcessor.set_gains orders them; `wht` float32 (H, W) or None."""
    _chk_dev(bay) if wht is None else _chk_dev(bay, wht)
    if bay.dtype != torch.float32 or bay.dim() != 2 or (wht is not None and (wht.dtype != torch.float32 or
                                                                          tuple(wht.shape) != tuple(bay.shape))):
        raise TypeError("float32 (H, W) mosaic (and white image of the same shape) expected")
    H, W = (int(v) for v in bay.shape)
    g = np.ascontiguousarray(np.asarray(gains, dtype=np.float64).ravel())
    if g.size != 4:
        raise ValueError("four Bayer gains expected")
    if out is None:
        out = torch.empty((H, W), dtype=torch.float64, device=bay.device)
    soft = exp_bias is not None
    max_lum = float(2 ** (-exp_bias)) if soft else 1.0
    b = float(np.exp(7))
    nat.check(nat.lib().pcb_cfa_safe_awb(_ptr(bay), None if wht is None else _ptr(wht), H, W,
                                         g.ctypes.data_as(C.POINTER(C.c_double)), int(soft), max_lum, b,
                                         float(np.log(1 + b)), _ptr(out), _stream()))
    return out
