#!/usr/bin/env python
"""Benchmark of the light-field hot path on synthetic Lytro-Illum-shaped data (BASELINE.json configs 1-4).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One step = one exposure through the whole path: micro-image resampling (min/max pass + uint16 quantise pass), sub-aperture
gather, 64-slice integer shift-and-sum stack, percentile contrast + sRGB.  In the default (fused) form the gather rides in
the refocus kernel's TMA staging (it reads the aligned image directly) and the colour management in its epilogue; the
explicit forms (separate viewpoints kernel, separate contrast kernels) are timed beside it and produce the same bytes.
Metric: refocus slices/s of the whole job (N GPUs each process their own exposures: weak scaling, no data-path collective).

The JSON line also carries:
  stage_ms / stage_gbs   per-stage CUDA-event times inside the timed region and achieved algorithmic GB/s
  roofline               dominant stage against the measured HBM peak (MEASURED_PEAKS.json), DRAM traffic from ncu
  roofline_smem          the persistent shift-and-sum kernel against the shared-memory roofline that actually bounds it
  aligned_lfs_per_s      BASELINE configs[1]: align + explicit sub-aperture gather only (rotate_ms: LfpRotator on one exposure,
                         an optional stage of the aligner that is off by default and therefore outside the step)
  refined_stack          BASELINE configs[3] on one GPU: 60 sub-pixel-refined slices of one light field
  sharded_stack          (N > 1) the same 60 slices sharded over the ranks, exchange fused into the finishing kernel
                         (peer stores over NVLink) and, beside it, the NCCL all-gather form; equality with the unsharded stack
  batch256               BASELINE configs[4]: 256 exposures, 256/N per rank, through the host-buffer API
  e2e                    same metric through LightFieldPipeline.exposure_stream with HOST buffers (pinned H2D of every raw
                         exposure, D2H of every finished stack inside the timed region)
  e2e_classes            the drop-in route: LfpAligner -> LfpExtractor -> LfpRefocuser with host ndarrays in and out
  cpu_baseline           the UNMODIFIED reference (oracle/_ref through oracle/ref_shim.py) on a bounded sample, 1 core
  clocks                 nvidia-smi samples taken during the timed region

`--impl reference` times the reference's own classes on all host cores (oracle/ref_runner.py).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from plenopticam_b200.misc import synth  # noqa: E402  (seeded data generators; no reference / oracle code)

ILLUM = synth.ILLUM


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--slices", type=int, default=64)
    ap.add_argument("--small", action="store_true", help="reduced shape (smoke-testing the harness only)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the side records (explicit forms, configs[1]/[3]/[4], class route): main loop only")
    ap.add_argument("--no-fuse-post", action="store_true",
                    help="linear stack + separate contrast kernels instead of the fused refocus epilogue")
    ap.add_argument("--no-fuse-gather", action="store_true",
                    help="separate viewpoints kernel + refocus from the 4-D array instead of reading the aligned image")
    ap.add_argument("--no-config1", action="store_true", help="reference arm: skip the one-off configs[0] run")
    ap.add_argument("--no-graph", action="store_true",
                    help="launch the kernels of a step one by one on the stream instead of replaying a CUDA graph")
    return ap.parse_args()


def workload(args):
    return synth.illum_workload(args.slices, small=getattr(args, "small", False))


def make_centroids(w, seed=0):
    return synth.illum_centroids(w, seed=seed)


def synth_raw_device(torch, H, W, C, seed, device):
    """uint16 exposure generated on the device: low-frequency sinusoid mix + uniform noise."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    yy = torch.arange(H, device=device, dtype=torch.float32)[:, None, None]
    xx = torch.arange(W, device=device, dtype=torch.float32)[None, :, None]
    ph = torch.arange(C, device=device, dtype=torch.float32)[None, None, :] + 0.37 * seed
    img = 0.5 + 0.2 * torch.sin(yy / 37.0 + ph) * torch.cos(xx / 53.0 - ph) + 0.15 * torch.sin((xx + yy) / 11.0 + 2 * ph)
    img = img + 0.1 * torch.rand((H, W, C), generator=g, device=device)
    img = (img.clamp_(0, 1) * 65535.0).round_()
    return img.to(torch.int32).to(torch.uint16)


# ---------------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------------
class ClockSampler(object):
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------------------
# CPU arms.  The only places bench.py executes oracle/ code as a workload.
# ---------------------------------------------------------------------------------------------------------
def reference_sample(w, n_slices_full, procs, area_div=16, n_slices=4):
    """One bounded sample of the workload through the UNMODIFIED reference classes (oracle/ref_runner.py)."""
    from oracle import ref_runner
    return ref_runner.illum_sample(w["LY"], w["LX"], w["pitch"], w["M"], n_slices_full, area_div=area_div,
                                   n_slices=n_slices, procs=procs)


def port_sample(w, n_slices, area_div=4):
    """Fallback when the reference snapshot is absent: the numpy oracle port on one core (crop x slices)."""
    from oracle import oracle_np as onp
    f = float(np.sqrt(area_div))
    LY, LX = max(8, int(round(w["LY"] / f))), max(8, int(round(w["LX"] / f)))
    cent = synth.synth_centroids(LY, LX, w["pitch"], w["pat"], hex_odd=True, jitter=0.15, origin=w["origin"], seed=0)
    H, W = int(cent[:, 0].max() + w["origin"][0] + 1), int(cent[:, 1].max() + w["origin"][1] + 1)
    raw = synth.synth_raw(H, W, 3, seed=0)
    M = w["M"]
    t0 = time.perf_counter()
    q = onp.uint16_norm(onp.resample_hex(raw, cent, M))
    vp = onp.compose_viewpoints(q, M)
    t1 = time.perf_counter()
    vp64 = onp._apply_aperture(vp, M)
    vp64 /= M
    a_all = list(range(-(n_slices // 2), n_slices - n_slices // 2))
    stack = np.asarray([onp.refocus_int_slice(vp64, a, M) for a in a_all])
    onp.refocuser_postprocess(stack)
    t2 = time.perf_counter()
    lens_ratio = (w["LY"] * w["LX"]) / float(LY * LX)
    n_full = w["ran_refo"][1] - w["ran_refo"][0]
    t_full = (t1 - t0) * lens_ratio + (t2 - t1) * lens_ratio * n_full / float(n_slices)
    return t2 - t0, {"lens_grid": [LY, LX], "slices": n_slices, "full_exposure_s_one_core_extrapolated": t_full}


def run_reference_arm(args):
    """`--impl reference`: the reference's own classes (LfpResampler / LfpRearranger / LfpShiftAndSum + the refocuser's
    colour management, unmodified, from oracle/_ref) on ALL host cores - one independent exposure per core, because the
    reference is single-threaded Python.  A step is a bounded sample of the GPU arm's workload (1/16-area crop of the
    lens grid, 4 of the 64 slices); `ms_per_step` is the wall time actually spent per step, `value` the slices/s of the
    full workload that this throughput extrapolates to (linear in lenslets x slices; `extrapolation` has the factors)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = workload(args)
    make_centroids(w)
    cores = os.cpu_count() or 1
    n_total = w["ran_refo"][1] - w["ran_refo"][0]
    from oracle import ref_runner
    steps, warm = max(1, min(args.steps, 2)), min(args.warmup, 1)
    if ref_runner.available():
        kind = "reference"
        area_div, n_sl = (64, 2) if args.small else (16, 4)
        for _ in range(warm):
            reference_sample(w, n_total, cores, area_div, n_sl)
        walls, info = [], None
        for _ in range(steps):
            wall, info = reference_sample(w, n_total, cores, area_div, n_sl)
            walls.append(wall)
        wall = float(np.mean(walls))
        t_full = info["full_exposure_s_all_cores_extrapolated"]
        sample = ("UNMODIFIED reference classes (LfpResampler.main incl. uint16 normalisation + pickle, LfpRearranger.main, "
                  "LfpShiftAndSum.main, auto_hist_align + srgb_conv) from oracle/_ref: %d processes x one exposure of a "
                  "1/%d-area crop (%dx%d hex lenslets, M=%d) and %d of %d slices per step"
                  % (cores, area_div, info["lens_grid"][0], info["lens_grid"][1], w["M"], n_sl, n_total))
    else:
        kind = "port"
        wall, info = port_sample(w, min(8, n_total))
        t_full = info["full_exposure_s_one_core_extrapolated"]
        cores = 1
        sample = "oracle/_ref absent: numpy oracle port, 1 core, %dx%d lenslets, %d slices" % (
            info["lens_grid"][0], info["lens_grid"][1], info["slices"])
    val = n_total / t_full
    line = {
        "impl": "reference", "metric": "refocus_slices_per_s", "value": val, "unit": "slices/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": wall * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(w, args),
        "extrapolation": {"what": "value = slices of the full workload / (full-workload seconds per exposure on all cores); "
                                  "ms_per_step is the measured wall time of the bounded sample, not of a full exposure",
                          "full_exposure_s_all_cores": t_full, "detail": info},
        "cpu_baseline": {"value": val, "unit": "slices/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": "slices/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    if kind == "reference" and not args.no_config1 and not args.small:
        # BASELINE configs[0] in full, once, through LfpAligner -> LfpExtractor -> LfpRefocuser (single process)
        res, _ = ref_runner.config1_full()
        line["config1_reference"] = res
    print(json.dumps(line))


def workload_config(w, args):
    fuse_g = not getattr(args, "no_fuse_gather", False)
    return {"workload": "illum_align+viewpoints+refocus%d (raw %dx%dx%d u16, %dx%d hex lenslets, M=%d, a in [%d,%d))"
                        % (args.slices, w["W"], w["H"], w["C"], w["LY"], w["LX"], w["M"], w["ran_refo"][0], w["ran_refo"][1]),
            "exposures_per_gpu_per_step": 1, "l2_policy": "inputs larger than L2 (249 MB raw per exposure, ring of 3)",
            "parallelism": "exposures sharded over GPUs, no data-path collective",
            "gather": "sub-aperture gather folded into the refocus kernel's TMA staging (reads the aligned image; the 4-D "
                      "array is materialised on request: see unfused / aligned_lfs_per_s)" if fuse_g
                      else "separate viewpoints kernel",
            "post": "separate contrast kernels" if getattr(args, "no_fuse_post", False)
                    else "contrast + sRGB fused into the refocus epilogue",
            "launch": "kernels launched one by one on the stream" if getattr(args, "no_graph", False) or not fuse_g
                      or getattr(args, "no_fuse_post", False)
                      else "the launches of a step replayed as one CUDA graph per ring buffer "
                           "(LightFieldPipeline.capture); stage_ms from a second pass with stream launches and events"}


# ---------------------------------------------------------------------------------------------------------
# GPU arm helpers
# ---------------------------------------------------------------------------------------------------------
def _maxed(torch, dist, world, device, ms):
    if world > 1:
        t = torch.tensor([ms], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    return float(ms)


def classes_route(w, cent, raw_host, pat, M, ran_refo, wht=None):
    """The drop-in route a user of the reference takes: the three stage classes with host ndarrays in and out (incl. the
    lfp_img_align.pkl checkpoint the reference writes).  Returns stage seconds and the stack."""
    from plenopticam_b200.cfg import PlenopticamConfig
    from plenopticam_b200.lfp_aligner import LfpAligner
    from plenopticam_b200.lfp_extractor import LfpExtractor
    from plenopticam_b200.lfp_refocuser import LfpRefocuser
    from plenopticam_b200.misc import PlenopticamStatus
    t = {}
    with tempfile.TemporaryDirectory() as tmp:
        cfg, sta = PlenopticamConfig(), PlenopticamStatus()
        cfg.params[cfg.lfp_path] = os.path.join(tmp, "exposure.png")
        for k, v in ((cfg.smp_meth, 'local'), (cfg.ptc_leng, int(M)), (cfg.ran_refo, list(ran_refo)), (cfg.opt_prnt, False),
                     (cfg.opt_vign, wht is not None), (cfg.opt_rota, False), (cfg.opt_view, True), (cfg.opt_refo, True),
                     (cfg.opt_refi, False), (cfg.opt_colo, False), (cfg.opt_cont, False), (cfg.opt_lier, False),
                     (cfg.opt_arti, False), (cfg.opt_dpth, False), (cfg.opt_pflu, False)):
            cfg.params[k] = v
        cfg.calibs[cfg.mic_list] = cent
        cfg.calibs[cfg.pat_type] = pat
        cfg.calibs[cfg.ptc_mean] = [float(M), float(M)]
        t0 = time.perf_counter()
        al = LfpAligner(raw_host, cfg=cfg, sta=sta, wht_img=wht)
        al.main()
        aligned = al.lfp_img
        t["align"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        ex = LfpExtractor(aligned, cfg=cfg, sta=sta)
        ex.main()
        vp = ex.vp_img_linear
        t["extract"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        rf = LfpRefocuser(vp, cfg=cfg, sta=sta)
        rf.main()
        stack = rf.refo_stack
        t["refocus"] = time.perf_counter() - t0
    t["total"] = t["align"] + t["extract"] + t["refocus"]
    return t, aligned, stack


# ---------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------
def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
        return

    import ctypes as C
    import torch
    import torch.distributed as dist
    from plenopticam_b200 import _native, ops
    from plenopticam_b200 import parallel as par
    from plenopticam_b200.pipeline import LightFieldPipeline

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    L, p = _native.lib(), ops._ptr

    w = workload(args)
    cent = make_centroids(w)
    fuse_post = not args.no_fuse_post
    fuse_gather = not args.no_fuse_gather
    pipe = LightFieldPipeline(cent, (w["H"], w["W"], w["C"]), np.uint16, w["M"], w["pat"], ran_refo=w["ran_refo"],
                              fuse_post=fuse_post, gather_fused=fuse_gather)
    ring = [synth_raw_device(torch, w["H"], w["W"], w["C"], seed=1000 * rank + s, device=device) for s in range(3)]
    K, W_ = args.steps, max(args.warmup, 3)
    stream = torch.cuda.current_stream()
    n_slices = pipe.N

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def make_step(gather_fused, post_fused):
        """One exposure, stage by stage, with optional events between the stages."""
        names = ["resample_minmax", "resample_u16"] + ([] if gather_fused else ["viewpoints"]) + ["refocus_int"] + \
                ([] if post_fused else ["post"])

        def step(raw, ev=None):
            st = ops._stream()
            k = 0
            if ev: ev[k].record(stream); k += 1
            _native.check(L.pcb_minmax_reset(p(pipe.mm_align), st))
            _native.check(L.pcb_resample_minmax(*(pipe.tabs._args(raw, pipe.flip) + [p(pipe.mm_align), st])))
            if ev: ev[k].record(stream); k += 1
            _native.check(L.pcb_resample_u16(*(pipe.tabs._args(raw, pipe.flip) + [p(pipe.mm_align), p(pipe.aligned), st])))
            if ev: ev[k].record(stream); k += 1
            src, plan = (pipe.aligned, pipe.refocus_plan_al) if gather_fused else (None, pipe.refocus_plan)
            if not gather_fused:
                src = pipe.viewpoints()
                if ev: ev[k].record(stream); k += 1
            _native.check(L.pcb_minmax_reset(p(pipe.mm_stack), st))
            if post_fused:
                if not plan.launch_srgb(src, pipe.post, pipe.mm_stack, pipe.lohi, pipe.scratch):
                    raise SystemExit("bench.py: the fused refocus epilogue does not apply to this workload")
                if ev: ev[k].record(stream); k += 1
                return
            plan.launch(src, pipe.stack, pipe.mm_stack)
            if ev: ev[k].record(stream); k += 1
            pipe.post_process()
            if ev: ev[k].record(stream); k += 1
        return step, names

    def timed(step_fn, n_ev, steps, warm):
        for i in range(warm):
            step_fn(ring[i % 3])
        barrier()
        events = [[torch.cuda.Event(enable_timing=True) for _ in range(n_ev)] for _ in range(steps)]
        t_b, t_e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        t_b.record(stream)
        for i in range(steps):
            step_fn(ring[i % 3], events[i])
        t_e.record(stream)
        barrier()
        return _maxed(torch, dist, world, device, t_b.elapsed_time(t_e)), events

    step, stage_names = make_step(fuse_gather, fuse_post)
    try:
        step(ring[0])
    except NotImplementedError:
        fuse_gather = False
        step, stage_names = make_step(False, fuse_post)
    for i in range(W_):
        step(ring[i % 3])
    use_graph = fuse_gather and fuse_post and not args.no_graph
    sampler = ClockSampler(local_rank)
    graph_note = None
    if use_graph:
        # the same launches (LightFieldPipeline.run_device == `step`), captured once per ring buffer and replayed
        try:
            caps = [pipe.capture(r, out=pipe.post) for r in ring]
            for i in range(max(W_, 3)):
                caps[i % 3].replay()
            torch.cuda.synchronize()
        except Exception as exc:                          # a driver that refuses the capture: time the stream launches
            graph_note = "graph capture failed (%s: %s): kernels launched one by one" % (type(exc).__name__, str(exc)[:120])
            use_graph = False
            try:
                torch.cuda.synchronize()
            except Exception:
                pass
            for i in range(max(W_, 3)):
                step(ring[i % 3])
    if use_graph:
        if rank == 0:
            sampler.start()
        t_b, t_e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        t_b.record(stream)
        for i in range(K):
            caps[i % 3].replay()
        t_e.record(stream)
        barrier()
        elapsed_ms = _maxed(torch, dist, world, device, t_b.elapsed_time(t_e))
        clocks = sampler.stop() if rank == 0 else None
        _, events = timed(step, len(stage_names) + 1, min(K, 50), 2)          # per-stage times: stream launches + events
        K_ev = min(K, 50)
    else:
        if rank == 0:
            sampler.start()
        elapsed_ms, events = timed(step, len(stage_names) + 1, K, 0)
        clocks = sampler.stop() if rank == 0 else None
        K_ev = K
    value = world * K * n_slices / (elapsed_ms * 1e-3)
    stage_ms = {nm: float(np.mean([events[i][k].elapsed_time(events[i][k + 1]) for i in range(K_ev)]))
                for k, nm in enumerate(stage_names)}
    abytes = pipe.algorithmic_bytes()
    if fuse_gather:
        abytes["refocus_int"] = abytes["refocus_int_from_aligned"]
    stage_gbs = {nm: abytes[nm] / (stage_ms[nm] * 1e-3) / 1e9 for nm in stage_names}
    launches_per_step = 2 + 2 + (0 if fuse_gather else 1) + (6 if fuse_post else 3 + 2)      # incl. edge_cumsum_kernel

    extras = {}
    if not args.no_extras:
        Kx = max(5, min(K, 30))
        # ---- the dominant kernel alone (events recorded by the library around the persistent shift-and-sum kernel)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream); e1.record(stream)
        torch.cuda.synchronize()
        ks = []
        for i in range(Kx):
            _native.check(L.pcb_debug_kernel_events(C.c_void_p(e0.cuda_event), C.c_void_p(e1.cuda_event)))
            step(ring[i % 3])
            _native.check(L.pcb_debug_kernel_events(None, None))
            torch.cuda.synchronize()
            ks.append(e0.elapsed_time(e1))
        extras["pxp_kernel_ms"] = float(np.mean(ks))
        # ---- explicit forms (same bytes out): separate viewpoints kernel; separate contrast kernels
        for tag, (gf, pf) in (("explicit_gather", (False, fuse_post)), ("explicit_gather_and_post", (False, False))):
            s2, n2 = make_step(gf, pf)
            ms2, ev2 = timed(s2, len(n2) + 1, Kx, 2)
            extras[tag] = {"ms_per_step": ms2 / Kx,
                           "stage_ms": {nm: float(np.mean([ev2[i][k].elapsed_time(ev2[i][k + 1]) for i in range(Kx)]))
                                        for k, nm in enumerate(n2)}}
        # ---- BASELINE configs[1]: align + sub-aperture extraction only
        def align_vp(raw, ev=None):
            pipe.align(raw)
            pipe.viewpoints()
        ms_av, _ = timed(align_vp, 1, Kx, 2)
        extras["aligned_lfs_per_s"] = world * Kx / (ms_av * 1e-3)
        extras["align_viewpoints_ms"] = ms_av / Kx
        # ---- LfpRotator (opt_rota, off by default in the reference: cfg/constants.py): cubic B-spline rotation of one exposure
        def rot(raw, ev=None):
            ops.rotate(raw, 0.004)
        ms_rot, _ = timed(rot, 1, 5, 2)
        extras["rotate_ms"] = ms_rot / 5

    # ---- BASELINE configs[3]: 60 refined slices of ONE light field (same on every rank), unsharded and sharded
    refined = None
    sharded = None
    if not args.no_extras:
        from plenopticam_b200.lfp_refocuser.refinement import refocus_refi
        same = synth_raw_device(torch, w["H"], w["W"], w["C"], seed=4242, device=device)
        pipe.align(same)
        vp = pipe.viewpoints().clone()
        M = w["M"]
        a_refi = list(range(-2 * M, 2 * M))
        mask = ops.aperture_mask(M)

        def unsharded():
            return ops.refocuser_postprocess(refocus_refi(vp, a_refi, mask))
        ref_stack = unsharded()
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            unsharded()
            e1.record(stream)
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        single_ms = float(np.median(ts))
        refined = {"slices": len(a_refi), "ms_per_stack": single_ms, "slices_per_s": len(a_refi) / (single_ms * 1e-3),
                   "what": "prefilter + 60 slices (a in [-30, 30) up-sampled pixels) + contrast / sRGB on ONE GPU"}
        if world > 1:
            sharded = {"slices": len(a_refi), "single_gpu_ms": single_ms}
            peer = None
            for tag, transport, shard in (("p2p_rows", "p2p", "rows"), ("p2p", "p2p", "slices"), ("nccl", "nccl", "slices")):
                try:
                    if transport == "p2p" and peer is None:
                        peer = par.PeerStack((len(a_refi),) + tuple(vp.shape[2:]))
                    run = lambda: par.sharded_refocus_stack(vp, a_refi, refi=True, transport=transport, peer_stack=peer,
                                                            shard=shard)
                    full = run()
                    torch.cuda.synchronize()
                    ok = bool(torch.equal(full, ref_stack))
                    okt = torch.tensor([1 if ok else 0], device=device)
                    dist.all_reduce(okt, op=dist.ReduceOp.MIN)
                    ts = []
                    for _ in range(7):
                        barrier()
                        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        e0.record(stream)
                        run()
                        e1.record(stream)
                        torch.cuda.synchronize()
                        ts.append(_maxed(torch, dist, world, device, e0.elapsed_time(e1)))
                    ms = float(np.median(ts))
                    sharded[tag] = {"ms_per_stack": ms, "slices_per_s": len(a_refi) / (ms * 1e-3),
                                          "equals_unsharded_on_every_rank": bool(int(okt.item())),
                                          "speedup_vs_one_gpu": single_ms / ms}
                except Exception as exc:                                   # e.g. no peer access on this box
                    sharded[tag] = {"error": "%s: %s" % (type(exc).__name__, exc)}
            sharded["what"] = ("one light field, 60 refined slices, every rank ends with the whole sRGB stack.  p2p_rows: each "
                               "rank computes a band of image rows of ALL slices (prefilter and horizontal pass restricted to "
                               "the band + halo) and the kernel that finishes a tile stores it into every rank's stack through "
                               "peer-mapped pointers (NVLink), 4-byte all-reduces as barriers; p2p: the same exchange with a "
                               "contiguous block of SLICES per rank (every rank prefilters everything); nccl: slice blocks + "
                               "all_gather.  CUDA events, max over ranks, median of 7")
            if peer is not None:
                peer.close()
        del vp, ref_stack

    # ---- end to end through the host-buffer API: pinned H2D raw in, D2H finished stack out, every step
    e2e = None
    batch = None
    if not args.no_e2e:
        host_ring = []
        for r in ring[:2]:
            hb = torch.empty(r.shape, dtype=torch.uint16, pin_memory=True)
            hb.copy_(r)
            host_ring.append(hb)
        torch.cuda.synchronize()
        Ke = max(4, min(K, 20))
        DEPTH = 3
        es = pipe.exposure_stream(depth=DEPTH)
        s_first, s_last = es.s_in, es.s_out

        def e2e_run(n, e_b=None, e_e=None):
            """n exposures through the pipelined host-buffer API; every result is read on the host."""
            acc = 0.0
            if e_b is not None:
                e_b.record(s_first)
            k0 = es.k
            for i in range(n):
                k = es.submit(host_ring[i % 2])
                if k - k0 >= DEPTH - 1:
                    acc += float(es.result(k - (DEPTH - 1))[0, 0, 0, 0])
            for k in range(max(k0, es.k - (DEPTH - 1)), es.k):
                acc += float(es.result(k)[0, 0, 0, 0])
            if e_e is not None:
                e_e.record(s_last)
            return acc

        e2e_run(3)
        barrier()
        t0 = time.perf_counter()
        e_b, e_e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e2e_run(Ke, e_b, e_e)
        barrier()
        wall_ms = (time.perf_counter() - t0) * 1e3
        e_ms = _maxed(torch, dist, world, device, max(e_b.elapsed_time(e_e), 0.0))
        e2e = {"value": world * Ke * n_slices / (e_ms * 1e-3), "unit": "slices/s",
               "h2d_bytes_per_step": int(pipe.raw.numel() * 2), "d2h_bytes_per_step": int(pipe.stack.numel() * 4),
               "ms_per_step": e_ms / Ke, "wall_ms_per_step": wall_ms / Ke, "steps": Ke,
               "api": "LightFieldPipeline.exposure_stream: pinned raw -> H2D | align, (gather+)refocus, post | D2H sRGB "
                      "stack, three streams, three exposures in flight; every exposure crosses PCIe both ways inside the "
                      "timed region"}
        if not args.no_extras:
            # ---- BASELINE configs[4]: 256 exposures, 256/N per rank, same API
            n_mine = len(par.shard_exposures(256, rank, world)) if not args.small else 8
            barrier()
            e_b, e_e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            e2e_run(n_mine, e_b, e_e)
            barrier()
            wall = time.perf_counter() - t0
            b_ms = _maxed(torch, dist, world, device, max(e_b.elapsed_time(e_e), 0.0))
            n_all = 256 if not args.small else 8 * world
            batch = {"exposures": n_all, "per_rank": n_mine, "ms_total": b_ms, "wall_s": wall,
                     "lfs_per_s": n_all / (b_ms * 1e-3), "slices_per_s": n_all * n_slices / (b_ms * 1e-3),
                     "what": "configs[4]: every exposure H2D from pinned host memory, full pipeline, sRGB stack D2H"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- the drop-in class route with host ndarrays (rank 0, one GPU)
    e2e_classes = None
    if not args.no_e2e and not args.no_extras:
        raw_host = ops.to_host(ring[0])
        classes_route(w, cent, raw_host, w["pat"], w["M"], w["ran_refo"])              # warm-up (plans, pinned pools)
        t_cls, _, stack_cls = classes_route(w, cent, raw_host, w["pat"], w["M"], w["ran_refo"])
        e2e_classes = {"illum": {"stage_s": t_cls, "slices_per_s": n_slices / t_cls["total"],
                                 "what": "plenopticam_b200.LfpAligner -> LfpExtractor -> LfpRefocuser, host ndarrays in / out, "
                                         "lfp_img_align.pkl written, float32 sRGB views + stack returned like the reference"}}
        if not args.small:
            w1 = synth.config1_workload()
            c1, lfp1, wht1 = synth.config1_data(w1)
            classes_route(w1, c1, lfp1, 'rec', w1["M"], w1["ran_refo"], wht=wht1)
            t1, al1, st1 = classes_route(w1, c1, lfp1, 'rec', w1["M"], w1["ran_refo"], wht=wht1)
            e2e_classes["config1"] = {"stage_s": t1, "lens_grid": [w1["LY"], w1["LX"]], "M": w1["M"],
                                      "aligned_sum": int(al1.astype(np.int64).sum()),
                                      "stack_sum": float(st1.astype(np.float64).sum()),
                                      "what": "BASELINE configs[0] in full (rec 186x279 lenslets, M=9, white image, "
                                              "ran_refo [-1,1]) through the drop-in classes; compare with "
                                              "config1_reference of the reference arm"}

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        with open(peaks_path) as f:
            peak, peak_src = float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    dom = max(stage_names, key=lambda nm: stage_ms[nm])
    roofline = {"bound": "hbm", "kernel": dom, "achieved": stage_gbs[dom], "peak": peak, "unit": "GB/s",
                "frac": stage_gbs[dom] / peak, "traffic": None, "peak_source": peak_src,
                "peak_nominal": 8000.0, "frac_nominal": stage_gbs[dom] / 8000.0,
                "algorithmic_bytes_per_launch": abytes[dom], "ms_per_launch": stage_ms[dom],
                "whole_step": {"algorithmic_bytes": sum(abytes[nm] for nm in stage_names),
                               "achieved": sum(abytes[nm] for nm in stage_names) / (elapsed_ms / K * 1e-3) / 1e9,
                               "frac": sum(abytes[nm] for nm in stage_names) / (elapsed_ms / K * 1e-3) / 1e9 / peak}}
    tr_path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tr_path):
        with open(tr_path) as f:
            roofline["traffic"] = json.load(f).get(dom + ("_from_aligned" if fuse_gather and dom == "refocus_int" else ""))
    roofline_smem = None
    if dom == "refocus_int" and "pxp_kernel_ms" in extras:
        # what bounds the stage: one 8-byte shared-memory read per (pixel, view, slice); 128 B / clk / SM
        sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
        sm_n = torch.cuda.get_device_properties(device).multi_processor_count
        smem_bytes = float(n_slices) * pipe.S * pipe.T * pipe.n_active_views * 8
        smem_peak = 128.0 * sm_n * sm_mhz * 1e6 / 1e9
        ach = smem_bytes / (extras["pxp_kernel_ms"] * 1e-3) / 1e9
        roofline_smem = {"bound": "smem", "kernel": "refocus_pxp_kernel", "achieved": ach, "peak": smem_peak, "unit": "GB/s",
                         "frac": ach / smem_peak, "bytes_per_launch": smem_bytes, "ms_per_launch": extras["pxp_kernel_ms"],
                         "peak_source": "128 B/clk/SM x %d SMs x %.0f MHz (clock seen during the run)" % (sm_n, sm_mhz),
                         "what": "N x S x T x active views pixel-view additions, one LDS.64 each"}

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        from oracle import ref_runner
        if ref_runner.available():
            area_div, n_sl = (64, 2) if args.small else (16, 4)
            wall, info = reference_sample(w, n_slices, 1, area_div, n_sl)
            t_full = info["full_exposure_s_one_core_extrapolated"]
            cpu_baseline = {"value": n_slices / t_full, "unit": "slices/s", "cores": 1, "kind": "reference",
                            "sample": "UNMODIFIED reference classes from oracle/_ref, 1 process: 1/%d-area crop (%dx%d lenslets), "
                                      "%d of %d slices, %.1f s measured, extrapolated linearly in lenslets x slices"
                                      % (area_div, info["lens_grid"][0], info["lens_grid"][1], n_sl, n_slices, wall),
                            "detail": info}
        else:
            wall, info = port_sample(w, min(8, n_slices))
            t_full = info["full_exposure_s_one_core_extrapolated"]
            cpu_baseline = {"value": n_slices / t_full, "unit": "slices/s", "cores": 1, "kind": "port",
                            "sample": "oracle/_ref absent: numpy oracle port, 1 core, %dx%d lenslets, %d slices, extrapolated"
                                      % (info["lens_grid"][0], info["lens_grid"][1], info["slices"]), "detail": info}

    line = {
        "metric": "refocus_slices_per_s", "value": value, "unit": "slices/s", "n_gpus": world, "steps": K, "warmup": W_,
        "ms_per_step": elapsed_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u16 in / u32 accumulate / f32 out", "data": "synthetic",
        "config": dict(workload_config(w, args), **({"launch": graph_note} if graph_note else {})),
        "lfs_per_s": world * K / (elapsed_ms * 1e-3),
        "aligned_lfs_per_s": extras.get("aligned_lfs_per_s"),
        "stage_ms": stage_ms, "stage_gbs": stage_gbs,
        "roofline": roofline, "roofline_smem": roofline_smem, "cpu_baseline": cpu_baseline, "e2e": e2e,
        "e2e_classes": e2e_classes, "unfused": {k: extras[k] for k in ("explicit_gather", "explicit_gather_and_post")
                                                if k in extras},
        "align_viewpoints_ms": extras.get("align_viewpoints_ms"), "rotate_ms": extras.get("rotate_ms"),
        "refined_stack": refined, "sharded_stack": sharded, "batch256": batch,
        "gpu_launches": launches_per_step * K, "clocks": clocks,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
