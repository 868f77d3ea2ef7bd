#!/usr/bin/env python
"""Benchmark of the light-field hot path on synthetic Lytro-Illum-shaped data (BASELINE.json configs 1+2).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One step = one exposure through the whole path: micro-image resampling (min/max pass + uint16 quantise pass),
viewpoint gather, 64-slice integer shift-and-sum stack, percentile contrast + sRGB.  Metric: refocus slices/s
of the whole job (N GPUs each process their own exposures: weak scaling, no data-path collective).

The JSON line also carries: per-stage times and achieved GB/s, the roofline of the dominant kernel against
MEASURED_PEAKS.json, the end-to-end number through the host-buffer API (pinned H2D of the raw exposure and D2H
of the finished stack inside the timed region), the CPU baseline (numpy oracle port on a bounded sample) and
the clocks seen during the timed region.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

ILLUM = dict(H=5368, W=7728, C=3, LY=434, LX=541, M=15, pitch=14.25, origin=(9.5, 9.5), pat='hex')
SMALL = dict(H=0, W=0, C=3, LY=60, LX=75, M=15, pitch=14.25, origin=(9.5, 9.5), pat='hex')      # --small (tests)


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
    ap.add_argument("--no-fuse-post", action="store_true",
                    help="linear stack + separate contrast kernels instead of the fused refocus epilogue")
    return ap.parse_args()


def workload(args):
    cfgw = dict(SMALL if args.small else ILLUM)
    n = args.slices
    cfgw["ran_refo"] = (-(n // 2), n - n // 2)
    return cfgw


def make_centroids(w, seed=0):
    from oracle import oracle_np as onp          # synthetic-data generator only (not the measured path)
    cent = onp.synth_centroids(w["LY"], w["LX"], w["pitch"], w["pat"], hex_odd=True, jitter=0.15,
                               origin=w["origin"], seed=seed)
    if not w["H"]:
        w["H"] = int(cent[:, 0].max() + w["origin"][0] + 1)
        w["W"] = int(cent[:, 1].max() + w["origin"][1] + 1)
    return cent


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
# CPU arms (oracle port).  The only places bench.py executes oracle/ code as a workload.
# ---------------------------------------------------------------------------------------------------------
_G = {}


def _cpu_align_rows(job):
    from oracle import oracle_np as onp
    lo, hi = job
    cent = _G["cent"]
    sel = (cent[:, 2] >= lo) & (cent[:, 2] < hi)
    sub = cent[sel].copy()
    sub[:, 2] -= lo
    # row parity of the stretch follows the absolute lens row: flip hex_odd for chunks starting on an odd row
    return onp.resample_hex(_G["raw"], sub, _G["M"], hex_odd=_G["hex_odd"] if lo % 2 == 0 else not _G["hex_odd"])


def _cpu_slices(job):
    from oracle import oracle_np as onp
    return [onp.refocus_int_slice(_G["vp64"], a, _G["M"]) for a in job]


def cpu_pipeline_sample(w, n_slices, cores, area_div=4):
    """Oracle port on a bounded sample: a 1/area_div-area crop of the lens grid and `n_slices` slices, the
    independent parts (lens rows, slices) spread over `cores` forked processes.  Returns (seconds, dict)."""
    import multiprocessing as mp
    from oracle import oracle_np as onp
    f = float(np.sqrt(area_div))
    LY, LX = max(8, int(round(w["LY"] / f))), max(8, int(round(w["LX"] / f)))
    cent = onp.synth_centroids(LY, LX, w["pitch"], w["pat"], hex_odd=True, jitter=0.15, origin=w["origin"], seed=0)
    H, W = int(cent[:, 0].max() + w["origin"][0] + 1), int(cent[:, 1].max() + w["origin"][1] + 1)
    raw = onp.synth_raw(H, W, 3, seed=0)
    M = w["M"]
    _G.update(cent=cent, raw=raw, M=M, hex_odd=onp.get_hex_direction(cent))
    ctx = mp.get_context("fork")
    t0 = time.perf_counter()
    # align: lens rows are independent; min/max + quantise once at the end
    if cores > 1:
        step = -(-LY // cores)
        step += step % 2            # keep row parity per chunk
        jobs = [(lo, min(lo + step, LY)) for lo in range(0, LY, step)]
        with ctx.Pool(min(cores, len(jobs))) as pool:
            parts = pool.map(_cpu_align_rows, jobs)
        aligned = np.concatenate(parts, axis=0)
    else:
        aligned = onp.resample_hex(raw, cent, M)
    q = onp.uint16_norm(aligned)
    t1 = time.perf_counter()
    vp = onp.compose_viewpoints(q, M)
    t2 = time.perf_counter()
    vp64 = onp._apply_aperture(vp, M)
    vp64 /= M
    _G["vp64"] = vp64
    a_all = list(range(-(n_slices // 2), n_slices - n_slices // 2))
    if cores > 1:
        jobs = [a_all[i::cores] for i in range(cores) if a_all[i::cores]]
        with ctx.Pool(len(jobs)) as pool:
            res = pool.map(_cpu_slices, jobs)
        stack = np.zeros((n_slices,) + vp64.shape[2:])
        for i, r in enumerate(res):
            stack[i::cores] = np.asarray(r)
    else:
        stack = np.asarray(_cpu_slices(a_all))
    t3 = time.perf_counter()
    post = onp.refocuser_postprocess(stack)
    t4 = time.perf_counter()
    assert post.dtype == np.float32
    lens_ratio = (w["LY"] * w["LX"]) / float(LY * LX)
    slice_ratio = (w["ran_refo"][1] - w["ran_refo"][0]) / float(n_slices)
    t_full = ((t1 - t0) + (t2 - t1)) * lens_ratio + ((t3 - t2) + (t4 - t3)) * lens_ratio * slice_ratio
    info = {"lens_grid": [LY, LX], "slices": n_slices, "t_align_s": t1 - t0, "t_viewpoints_s": t2 - t1,
            "t_refocus_s": t3 - t2, "t_post_s": t4 - t3, "t_sample_s": t4 - t0, "t_full_extrapolated_s": t_full}
    return t_full, info


def run_reference_arm(args):
    """`--impl reference`: the reference's own algorithm (oracle port; the Python reference cannot travel to
    the GPU box) on the host cores, same workload/metric, each step a bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = workload(args)
    make_centroids(w)
    cores = os.cpu_count() or 1
    n_slices = min(16, args.slices)
    steps, warm = max(1, min(args.steps, 3)), min(args.warmup, 1)
    for _ in range(warm):
        cpu_pipeline_sample(w, n_slices, cores)
    ts, info = [], None
    for _ in range(steps):
        t, info = cpu_pipeline_sample(w, n_slices, cores)
        ts.append(t)
    t_full = float(np.mean(ts))
    n_total = w["ran_refo"][1] - w["ran_refo"][0]
    val = n_total / t_full
    sample = ("1/4-area crop of the Illum lens grid (%dx%d lenslets) and %d of %d slices per step, "
              "extrapolated linearly in lenslets x slices" % (info["lens_grid"][0], info["lens_grid"][1], n_slices, n_total))
    line = {
        "impl": "reference", "metric": "refocus_slices_per_s", "value": val, "unit": "slices/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": t_full * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(w, args),
        "cpu_baseline": {"value": val, "unit": "slices/s", "cores": cores, "kind": "port", "sample": sample,
                         "detail": info},
        "e2e": {"value": val, "unit": "slices/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def workload_config(w, args):
    return {"workload": "illum_align+viewpoints+refocus%d (raw %dx%dx%d u16, %dx%d hex lenslets, M=%d, a in [%d,%d))"
                        % (args.slices, w["W"], w["H"], w["C"], w["LY"], w["LX"], w["M"], w["ran_refo"][0], w["ran_refo"][1]),
            "exposures_per_gpu_per_step": 1, "l2_policy": "inputs larger than L2 (249 MB raw per exposure, ring of 3)",
            "parallelism": "exposures sharded over GPUs, no data-path collective",
            "post": "separate contrast kernels" if getattr(args, "no_fuse_post", False)
                    else "contrast + sRGB fused into the refocus epilogue"}


# ---------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------
def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    import torch.distributed as dist
    from plenopticam_b200 import _native
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
    _native.lib()

    w = workload(args)
    cent = make_centroids(w)
    pipe = LightFieldPipeline(cent, (w["H"], w["W"], w["C"]), np.uint16, w["M"], w["pat"], ran_refo=w["ran_refo"])
    ring = [synth_raw_device(torch, w["H"], w["W"], w["C"], seed=1000 * rank + s, device=device) for s in range(3)]
    K, W_ = args.steps, max(args.warmup, 3)
    stream = torch.cuda.current_stream()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # default: colour management fused into the refocus epilogue (pcb_refocus_int_srgb): four stages; --no-fuse-post
    # runs the linear stack + contrast kernels separately (five stages, one more pass over the stack)
    fused = not args.no_fuse_post
    stage_names = ["resample_minmax", "resample_u16", "viewpoints", "refocus_int"] + ([] if fused else ["post"])
    from plenopticam_b200 import ops
    L, p = _native.lib(), ops._ptr

    def step(raw, ev=None):
        st = ops._stream()
        if ev: ev[0].record(stream)
        _native.check(L.pcb_minmax_reset(p(pipe.mm_align), st))
        _native.check(L.pcb_resample_minmax(*(pipe.tabs._args(raw, pipe.flip) + [p(pipe.mm_align), st])))
        if ev: ev[1].record(stream)
        _native.check(L.pcb_resample_u16(*(pipe.tabs._args(raw, pipe.flip) + [p(pipe.mm_align), p(pipe.aligned), st])))
        if ev: ev[2].record(stream)
        pipe.viewpoints()
        if ev: ev[3].record(stream)
        if fused:
            if pipe.refocus_post() is False:
                raise SystemExit("bench.py: the fused refocus epilogue does not apply to this workload")
            if ev: ev[4].record(stream)
            return
        pipe.refocus()
        if ev: ev[4].record(stream)
        pipe.post_process()
        if ev: ev[5].record(stream)

    for i in range(W_):
        step(ring[i % 3])
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    events = [[torch.cuda.Event(enable_timing=True) for _ in range(6)] for _ in range(K)]
    t_begin, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_begin.record(stream)
    for i in range(K):
        step(ring[i % 3], events[i])
    t_end.record(stream)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    elapsed_ms = t_begin.elapsed_time(t_end)
    if world > 1:
        t = torch.tensor([elapsed_ms], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    n_slices = pipe.N
    value = world * K * n_slices / (elapsed_ms * 1e-3)

    stage_ms = {nm: float(np.mean([events[i][k].elapsed_time(events[i][k + 1]) for i in range(K)]))
                for k, nm in enumerate(stage_names)}
    abytes = pipe.algorithmic_bytes()
    stage_gbs = {nm: abytes[nm] / (stage_ms[nm] * 1e-3) / 1e9 for nm in stage_names}

    # ---- end to end through the host-buffer API: pinned H2D raw in, D2H finished stack out, every step
    e2e = None
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
        e_ms = max(e_b.elapsed_time(e_e), 0.0)
        if world > 1:
            t = torch.tensor([e_ms], device=device, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e_ms = float(t.item())
        e2e = {"value": world * Ke * n_slices / (e_ms * 1e-3), "unit": "slices/s",
               "h2d_bytes_per_step": int(pipe.raw.numel() * 2), "d2h_bytes_per_step": int(pipe.stack.numel() * 4),
               "ms_per_step": e_ms / Ke, "wall_ms_per_step": wall_ms / Ke, "steps": Ke,
               "api": "LightFieldPipeline.exposure_stream: pinned raw -> H2D | align, viewpoints, refocus, post | D2H sRGB "
                      "stack, three streams, three exposures in flight; every exposure crosses PCIe both ways inside the timed region"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        with open(peaks_path) as f:
            peak, peak_src = float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    dom = max(stage_names, key=lambda nm: stage_ms[nm])
    roofline = {"bound": "hbm", "kernel": dom, "achieved": stage_gbs[dom], "peak": peak, "unit": "GB/s",
                "frac": stage_gbs[dom] / peak, "traffic": None, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": abytes[dom], "ms_per_launch": stage_ms[dom]}
    if dom == "refocus_int":
        # the stage is not HBM-limited: one shared-memory read per (pixel, view, slice) is what bounds it
        roofline["on_chip_limiter"] = ("shared-memory wavefronts: LSU data pipe 80-84 % busy in refocus_pxp_kernel "
                                       "(ncu, profiles/r01c_prof_step.txt; DESIGN.md 4.3)")
    tr_path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tr_path):
        with open(tr_path) as f:
            roofline["traffic"] = json.load(f).get(dom)

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        t_full, info = cpu_pipeline_sample(w, min(8, n_slices), 1, area_div=2)
        cpu_baseline = {"value": n_slices / t_full, "unit": "slices/s", "cores": 1, "kind": "port",
                        "sample": "numpy oracle, 1 core: 1/2-area crop (%dx%d lenslets), %d of %d slices, extrapolated "
                                  "linearly in lenslets x slices" % (info["lens_grid"][0], info["lens_grid"][1],
                                                                     info["slices"], n_slices),
                        "detail": info}

    line = {
        "metric": "refocus_slices_per_s", "value": value, "unit": "slices/s", "n_gpus": world, "steps": K, "warmup": W_,
        "ms_per_step": elapsed_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u16 in / u32 accumulate / f32 out", "data": "synthetic",
        "config": workload_config(w, args),
        "aligned_lfs_per_s": world * K / (elapsed_ms * 1e-3),
        "stage_ms": stage_ms, "stage_gbs": stage_gbs,
        "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e,
        "gpu_launches": (10 if fused else 9) * K, "clocks": clocks,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
