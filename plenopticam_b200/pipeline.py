"""Device-resident light-field pipeline: raw exposure -> aligned uint16 LF -> 4-D viewpoints -> refocus stack
-> contrast/sRGB, with every buffer pre-allocated in HBM and nothing but kernel launches in the steady state.

This is the batch / serving form of the three stage classes: the same kernels, same results, but the aligned
image, the viewpoint array and the stack never leave the GPU unless asked for.  `run()` is the host-buffer
entry point (pinned H2D of the raw exposure in, D2H of the finished stack out).
"""
import ctypes as C
import functools

import numpy as np
import torch

from . import _native as nat
from . import ops
from .lfp_aligner.lfp_microlenses import build_tables


def _on_own_device(method):
    """Run a launch entry point with the executor's GPU as the current CUDA device, whatever the caller's current
    device is: the kernels go to `ops._stream()` (the current stream of the CURRENT device) and the library sizes its
    grids for the current device, so launching cuda:1's buffers while cuda:0 is current would queue them on the wrong
    GPU's stream."""
    @functools.wraps(method)
    def wrapper(self, *args, **kwargs):
        dev = self.device if hasattr(self, "device") else self.pipe.device
        if torch.cuda.current_device() == dev.index:
            return method(self, *args, **kwargs)
        with torch.cuda.device(dev):
            return method(self, *args, **kwargs)
    return wrapper


class LightFieldPipeline(object):
    """One calibrated camera (centroid table + angular size) -> reusable executor.

    centroids : (N,4) table rows [y, x, ly, lx] (cfg.calibs['mic_list'])
    raw_shape : (H, W, C) of the exposures;  raw_dtype: numpy dtype of the device copy (uint16/uint8/float32)
    M         : micro-image size (cfg.params['ptc_leng'], already validated by LfpMicroLenses.safe_pitch_eval)
    ran_refo  : cfg.params['ran_refo'] -> slices a in range(*ran_refo)
    """

    def __init__(self, centroids, raw_shape, raw_dtype, M, pat_type, ran_refo=(0, 2), flip=False,
                 postprocess=True, device=None, fuse_post=True, wht_img=None, gather_fused=True):
        self.device = ops.require_cuda() if device is None else torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.M = int(M)
        self.flip = bool(flip)
        self.postprocess = bool(postprocess)
        self.fuse_post = bool(fuse_post)     # colour management in the refocus epilogue (pcb_refocus_int_srgb)
        self.raw_shape = tuple(int(v) for v in raw_shape)
        self.raw_dtype = np.dtype(raw_dtype)
        if self.M % 2 == 0:
            raise ValueError("odd micro-image size required for refocusing (lfp_shiftandsum.py:123-124)")
        if self.raw_shape[2] != 3:
            raise ValueError("the refocus accumulator is 3-channel (lfp_shiftandsum.py:97)")
        self.tables = build_tables(centroids, self.raw_shape, self.M, pat_type)
        with torch.cuda.device(self.device):
            self.tabs = ops.LensTables(self.tables)
            LY, Xs, Cn = self.tables["LY"], self.tables["Xs"], self.raw_shape[2]
            self.S, self.T = LY, Xs
            self.a_list = list(range(*ran_refo))
            self.N = len(self.a_list)
            dev = self.device
            tdt = ops._NP2TORCH[self.raw_dtype]
            self.raw = torch.empty(self.raw_shape, dtype=tdt, device=dev)
            self.aligned = torch.empty((LY * self.M, Xs * self.M, Cn), dtype=torch.uint16, device=dev)
            self.vp = torch.empty((self.M, self.M, LY, Xs, Cn), dtype=torch.uint16, device=dev)
            self.stack = torch.empty((self.N, LY, Xs, Cn), dtype=torch.float32, device=dev)
            self.post = torch.empty_like(self.stack) if self.postprocess else None
            self.mm_align = torch.empty(2, dtype=torch.int32, device=dev)
            self.mm_stack = torch.empty(2, dtype=torch.int32, device=dev)
            self.refocus_plan = ops.RefocusPlan(tuple(self.vp.shape), self.vp.dtype, self.a_list,
                                                ops.aperture_mask(self.M), dev)
            # the same stack straight from the aligned image: the viewpoint gather rides in the refocus kernel's TMA
            # staging and the 4-D array is materialised only for callers that ask for it (`viewpoints()`)
            self.refocus_plan_al = ops.RefocusPlan(tuple(self.vp.shape), self.vp.dtype, self.a_list,
                                                   ops.aperture_mask(self.M), dev, aligned_pitch=self.M)
            self.refocus_plan_al.scratch = self.refocus_plan.scratch          # same geometry: share the accumulators
            self.gather_fused = gather_fused
            self.lohi = torch.empty(2, dtype=torch.float64, device=dev)
            self.scratch = torch.empty(int(nat.lib().pcb_percentile_scratch_bytes()), dtype=torch.uint8, device=dev)
            self._q_lo = (C.c_double * 1)(0.005)
            self._q_hi = (C.c_double * 1)(99.9)
            # de-vignetting (LfpAligner with opt_vign, lfp_aligner/top_level.py:51-57): the normalised gray white image
            # depends on the calibration only and is prepared once; every exposure is divided by it (float64 division,
            # float32 image) and then takes the float resampler
            self.wn = None
            if wht_img is not None:
                from .lfp_aligner.lfp_devignetter import gauss_kernel_1d
                wht = wht_img if hasattr(wht_img, 'is_cuda') else ops.to_device_any(np.asarray(wht_img))
                probe = torch.zeros((self.raw_shape[0], self.raw_shape[1], 1), dtype=torch.float32, device=dev)
                _, self.wn, stats = ops.devignette(probe, wht, gauss_kernel_1d(int(self.M)), out_dtype=torch.float32)
                self.noise_lev = float(stats[2].item())
                self.devig = torch.empty(self.raw_shape, dtype=torch.float32, device=dev)
                del probe
            # pinned staging for the host-buffer entry point
            self._pin_in = None
            self._pin_out = None
        self.n_active_views = int((~ops.aperture_mask(self.M)).sum())

    # -- stage launches (enqueue only) ---------------------------------------------------------------------
    @_on_own_device
    def align(self, raw=None):
        raw = self.raw if raw is None else raw
        if self.wn is not None:
            raw = ops.devignette_divide(raw, self.wn, out=self.devig)
        ops.resample_u16(raw, self.tabs, flip=self.flip, out=self.aligned, mm=self.mm_align)
        return self.aligned

    @_on_own_device
    def viewpoints(self):
        ops.viewpoints(self.aligned, self.M, out=self.vp)
        return self.vp

    @_on_own_device
    def refocus(self):
        L, p, st = nat.lib(), ops._ptr, ops._stream()
        nat.check(L.pcb_minmax_reset(p(self.mm_stack), st))
        self.refocus_plan.launch(self.vp, self.stack, self.mm_stack)
        return self.stack

    @_on_own_device
    def post_process(self, out=None):
        out = self.post if out is None else out
        L, p, st = nat.lib(), ops._ptr, ops._stream()
        ops.contrast_bounds(self.stack[0], 0.005, 99.9, lohi=self.lohi, scratch=self.scratch)
        nat.check(L.pcb_contrast_srgb(p(self.stack), self.stack.numel(), p(self.lohi), p(self.mm_stack),
                                      p(out), st))
        return out

    @_on_own_device
    def refocus_post(self, out=None):
        """Stack + colour management with the fused epilogue (the linear stack is not materialised); False when the
        library has no fused form for this light field."""
        out = self.post if out is None else out
        nat.check(nat.lib().pcb_minmax_reset(ops._ptr(self.mm_stack), ops._stream()))
        return out if self.refocus_plan.launch_srgb(self.vp, out, self.mm_stack, self.lohi, self.scratch) else False

    @_on_own_device
    def _refocus_from_aligned(self, out=None):
        """Stack (+ colour management) read straight from `self.aligned`; False when the library cannot."""
        pl = self.refocus_plan_al
        nat.check(nat.lib().pcb_minmax_reset(ops._ptr(self.mm_stack), ops._stream()))
        try:
            if self.postprocess and self.fuse_post:
                out = self.post if out is None else out
                return out if pl.launch_srgb(self.aligned, out, self.mm_stack, self.lohi, self.scratch) else False
            pl.launch(self.aligned, self.stack, self.mm_stack)
        except NotImplementedError:
            return False
        if self.postprocess:
            return self.post_process(out)
        if out is not None:
            out.copy_(self.stack)
            return out
        return self.stack

    @_on_own_device
    def run_device(self, raw=None, out=None):
        """All stages on the current stream; returns the final device tensor (sRGB stack, or the linear
        stack if postprocess=False)."""
        self.align(raw)
        if self.gather_fused:
            res = self._refocus_from_aligned(out)
            if res is not False:
                return res
            self.gather_fused = False             # no such kernel form for this camera: gather explicitly from now on
        self.viewpoints()
        if self.postprocess and self.fuse_post:
            res = self.refocus_post(out)
            if res is not False:
                return res
        self.refocus()
        if self.postprocess:
            return self.post_process(out)
        if out is not None:
            out.copy_(self.stack)
            return out
        return self.stack

    @_on_own_device
    def capture(self, raw, out=None):
        """The launches of `run_device(raw, out)` as a CUDA graph bound to these two tensors (see CapturedExposure): for a
        stream of exposures of one camera that arrive in the same device buffer(s)."""
        return CapturedExposure(self, raw, out)

    # -- host-buffer entry point -----------------------------------------------------------------------------
    @_on_own_device
    def _staging(self):
        if self._pin_in is None:
            tdt = ops._NP2TORCH[self.raw_dtype]
            self._pin_in = torch.empty(self.raw_shape, dtype=tdt, pin_memory=True)
            self._pin_out = torch.empty(tuple(self.stack.shape), dtype=torch.float32, pin_memory=True)
        return self._pin_in, self._pin_out

    def pinned_input(self):
        """Pinned host buffer the caller may fill directly (numpy view) to skip one host copy."""
        pin_in, _ = self._staging()
        return pin_in.view(torch.int16).numpy().view(np.uint16) if pin_in.dtype == torch.uint16 else pin_in.numpy()

    @_on_own_device
    def run(self, raw_host=None):
        """raw exposure (numpy, host) -> finished stack (numpy view of a pinned buffer, valid until next run)."""
        pin_in, pin_out = self._staging()
        if raw_host is not None:
            self.pinned_input()[...] = raw_host
        self.raw.copy_(pin_in, non_blocking=True)
        out = self.run_device()
        pin_out.copy_(out, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return pin_out.numpy()

    def exposure_stream(self, depth=2, out_dtype=np.float32):
        """Pipelined host-buffer entry point for a sequence of exposures (see ExposureStream)."""
        return ExposureStream(self, depth, out_dtype)

    # -- byte accounting used by bench.py / DESIGN.md --------------------------------------------------------
    def algorithmic_bytes(self):
        H, W, Cn = self.raw_shape
        b_raw = H * W * Cn * self.raw_dtype.itemsize
        b_al = self.aligned.numel() * 2
        b_sl = self.S * self.T * 3 * 4
        b_vp_active = self.n_active_views * self.S * self.T * 3 * 2
        return {
            "resample_minmax": b_raw,
            "resample_u16": b_raw + b_al,
            "viewpoints": 2 * b_al,
            "refocus_int": b_vp_active + self.N * b_sl,
            # reading the aligned image moves whole aligned rows (masked views included); the ALGORITHMIC bytes stay those
            # of the active views
            "refocus_int_from_aligned": b_vp_active + self.N * b_sl,
            "post": 2 * self.N * b_sl + 2 * 4 * b_sl,
        }


class CapturedExposure(object):
    """`LightFieldPipeline.run_device(raw, out)` captured once as a CUDA graph and replayed: the ten kernels of an exposure
    are then launched by one call, and the gaps between them shrink (1.606 -> 1.593 ms per Illum exposure).  The graph is
    bound to the `raw` tensor (copy or generate the next exposure into it) and to the result tensor `self.out`.

        step = pipe.capture(raw_dev)
        ...fill raw_dev...; stack = step.replay()      # enqueues on the current stream, returns `step.out`

    The refocus scratch comes back zeroed from every replay and the captured launches rely on that (no memset inside
    the graph); if another launch dirtied it in between, replay() zeroes it first."""

    def __init__(self, pipe, raw, out=None):
        self.pipe, self.raw = pipe, raw
        with torch.cuda.device(pipe.device):
            self.out = out if out is not None else torch.empty_like(pipe.post if pipe.postprocess else pipe.stack)
            cur = torch.cuda.current_stream(pipe.device)
            side = torch.cuda.Stream(device=pipe.device)
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                for _ in range(2):                        # settles the plans (fused forms, scratch state) before the capture
                    pipe.run_device(raw, out=self.out)
                side.synchronize()
                self.graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.graph, stream=side):
                    res = pipe.run_device(raw, out=self.out)
                assert res is self.out
            cur.wait_stream(side)
        self._scratch = pipe.refocus_plan.scratch

    def replay(self):
        with torch.cuda.device(self.pipe.device):
            if not getattr(self._scratch, "_pcb_acc_zero", False):
                self._scratch.zero_()
            self._scratch._pcb_acc_zero = False
            self.graph.replay()
            self._scratch._pcb_acc_zero = True
        return self.out


class ExposureStream(object):
    """Host-buffer entry point for a SEQUENCE of exposures of one camera (batch / serving form of `run`).

    Three CUDA streams work on consecutive exposures at the same time: exposure k+1 is copied host->device
    while exposure k runs through align/viewpoints/refocus/post and the finished stack of exposure k-1 is
    copied device->host, so the steady-state cost per exposure is the longest of the three instead of their
    sum.  Every exposure still crosses PCIe in both directions.

        es = pipe.exposure_stream()
        for raw_pinned in exposures:            # pinned torch tensors (or use es.pinned_input(k))
            k = es.submit(raw_pinned)
            if k >= 1: stack = es.result(k - 1)  # numpy view of a pinned buffer, valid until slot reuse
        stack = es.result(k)
    """

    def __init__(self, pipe, depth=2, out_dtype=np.float32):
        """out_dtype: float32 (the reference's `refo_stack`), or uint16 / uint8 = what LfpExporter would write for it
        (`Normalizer(stack).uint16_norm()` / `.uint8_norm()`, lfp_extractor/lfp_exporter.py:134,195), quantised on the
        device so that a half / a quarter of the bytes cross PCIe."""
        self.pipe = pipe
        self.depth = int(depth)
        self.out_dtype = np.dtype(out_dtype)
        if self.out_dtype not in (np.dtype(np.float32), np.dtype(np.uint16), np.dtype(np.uint8)):
            raise TypeError("out_dtype must be float32, uint16 or uint8")
        dev = pipe.device
        tdt = ops._NP2TORCH[pipe.raw_dtype]
        with torch.cuda.device(dev):
            self.raw_dev = [torch.empty(pipe.raw_shape, dtype=tdt, device=dev) for _ in range(self.depth)]
            self.out_dev = [torch.empty(tuple(pipe.stack.shape), dtype=torch.float32, device=dev)
                            for _ in range(self.depth)]
            self.pin_in = [None] * self.depth
            tout = {2: torch.uint16, 1: torch.uint8, 4: torch.float32}[self.out_dtype.itemsize]
            self.q_dev = [torch.empty(tuple(pipe.stack.shape), dtype=tout, device=dev) for _ in range(self.depth)] \
                if tout != torch.float32 else None
            self.pin_out = [torch.empty(tuple(pipe.stack.shape), dtype=tout, pin_memory=True)
                            for _ in range(self.depth)]
            self.s_in = torch.cuda.Stream(device=dev)
            self.s_out = torch.cuda.Stream(device=dev)
            self.ev_in = [torch.cuda.Event() for _ in range(self.depth)]      # raw slot filled
            self.ev_run = [torch.cuda.Event() for _ in range(self.depth)]     # slot consumed / result slot produced
            self.ev_out = [torch.cuda.Event() for _ in range(self.depth)]     # result slot copied to the host
        self.k = 0

    def pinned_input(self, k=None):
        """Pinned host buffer of slot k % depth that the caller may fill directly (numpy view)."""
        s = (self.k if k is None else k) % self.depth
        if self.pin_in[s] is None:
            self.pin_in[s] = torch.empty(self.pipe.raw_shape, dtype=ops._NP2TORCH[self.pipe.raw_dtype], pin_memory=True)
        t = self.pin_in[s]
        return t.view(torch.int16).numpy().view(np.uint16) if t.dtype == torch.uint16 else t.numpy()

    @_on_own_device
    def submit(self, raw_pinned=None):
        """Enqueue one exposure (a pinned host tensor, or the slot's own pinned_input buffer); returns its index."""
        k, s = self.k, self.k % self.depth
        src = self.pin_in[s] if raw_pinned is None else raw_pinned
        cur = torch.cuda.current_stream(self.pipe.device)
        with torch.cuda.stream(self.s_in):
            if k >= self.depth:
                self.s_in.wait_event(self.ev_run[s])          # the kernels of exposure k-depth have read this slot
            self.raw_dev[s].copy_(src, non_blocking=True)
            self.ev_in[s].record(self.s_in)
        cur.wait_event(self.ev_in[s])
        if k >= self.depth:
            cur.wait_event(self.ev_out[s])                    # result slot has left for the host
        self.pipe.run_device(self.raw_dev[s], out=self.out_dev[s])
        res = self.out_dev[s]
        if self.q_dev is not None:
            res = (ops.uint16_norm if self.out_dtype.itemsize == 2 else ops.uint8_norm)(res, out=self.q_dev[s])
        self.ev_run[s].record(cur)
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(self.ev_run[s])
            self.pin_out[s].copy_(res, non_blocking=True)
            self.ev_out[s].record(self.s_out)
        self.k += 1
        return k

    @_on_own_device
    def result(self, k):
        """Block until exposure k has arrived on the host; numpy view valid until slot k % depth is resubmitted."""
        if not (self.k - self.depth <= k < self.k):
            raise IndexError("result %d is not in flight" % k)
        s = k % self.depth
        self.ev_out[s].synchronize()
        t = self.pin_out[s]
        return t.view(torch.int16).numpy().view(np.uint16) if t.dtype == torch.uint16 else t.numpy()
