"""Timed runs of the UNMODIFIED reference classes (through oracle/ref_shim.py) for bench.py's reference arm.

TEST / BENCHMARK INFRASTRUCTURE ONLY: nothing under plenopticam_b200/ imports this.  What is timed is the reference's
own code on the host CPU:

  * `LfpResampler.main`   lfp_aligner/lfp_resampler.py:17-47  (smp_meth='local': lfp_local_resampler.py:106-148 /
                          76-104, uint16 normalisation and pickle included, as the reference does them)
  * `LfpRearranger.main`  lfp_extractor/lfp_rearranger.py:70-107
  * `LfpShiftAndSum.main` lfp_refocuser/lfp_shiftandsum.py:48-139
  * the colour management of `LfpRefocuser.shift_and_sum` (lfp_refocuser/top_level.py:68-70):
    `LfpContrast.auto_hist_align` + `GammaConverter.srgb_conv`

The reference is single-threaded Python.  To use all host cores the runner starts one process per core, each running
the reference on its own exposure (independent exposures are how the reference would be scaled out), and reports
the aggregate.  The data generators are the same seeded ones the GPU arm uses (plenopticam_b200/misc/synth.py).
"""
import atexit
import multiprocessing as mp
import os
import shutil
import sys
import tempfile
import time
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from plenopticam_b200.misc import synth  # noqa: E402


def available():
    return ref_shim.reference_available()


def _private_reference():
    """Give THIS process its own byte-identical copy of the reference package before importing it.  The reference rewrites
    `plenopticam/cfg/cfg.json` inside its package directory whenever a config object is created or saved
    (cfg/cfg.py:62-120); N worker processes doing that to one file race (one of them reads an empty file)."""
    if ref_shim._LOADED is not None or "plenopticam" in sys.modules:
        return
    src = os.path.join(ref_shim.REFERENCE_ROOT, "plenopticam")
    if _RUN_DIR is not None:                               # pool worker: the parent removes the run directory
        dst = os.path.join(_RUN_DIR, "p%d" % os.getpid())
        os.makedirs(dst)
    else:
        dst = tempfile.mkdtemp(prefix="pcb_ref_%d_" % os.getpid())
        atexit.register(shutil.rmtree, dst, True)
    shutil.copytree(src, os.path.join(dst, "plenopticam"), ignore=shutil.ignore_patterns("__pycache__"))
    ref_shim.REFERENCE_ROOT = dst


_RUN_DIR = None


def _quiet_cfg(tmp, cent, pat, M, ran_refo):
    cfg, sta = ref_shim.make_cfg_sta(tmp, cent, pat, M, ran_refo=ran_refo)
    cfg.lfpimg = {}
    for k, v in ((cfg.opt_vign, False), (cfg.opt_rota, False), (cfg.opt_view, True), (cfg.opt_refo, True),
                 (cfg.opt_colo, False), (cfg.opt_cont, False), (cfg.opt_awb_, False), (cfg.opt_sat_, False),
                 (cfg.opt_lier, False), (cfg.opt_arti, False), (cfg.opt_dpth, False), (cfg.opt_pflu, False)):
        cfg.params[k] = v
    return cfg, sta


def hot_path_once(raw, cent, pat, M, ran_refo):
    """One exposure through the reference's own hot-path classes; returns (stage seconds, sRGB stack)."""
    _private_reference()
    ref_shim.load_reference()
    from plenopticam.lfp_aligner.lfp_resampler import LfpResampler
    from plenopticam.lfp_extractor.lfp_rearranger import LfpRearranger
    from plenopticam.lfp_extractor.lfp_contrast import LfpContrast
    from plenopticam.lfp_refocuser.lfp_shiftandsum import LfpShiftAndSum
    from plenopticam.misc import GammaConverter
    t = {}
    with tempfile.TemporaryDirectory() as tmp, warnings.catch_warnings():
        warnings.simplefilter("ignore")
        cfg, sta = _quiet_cfg(tmp, cent, pat, M, ran_refo)
        t0 = time.perf_counter()
        rs = LfpResampler(lfp_img=raw, cfg=cfg, sta=sta, method='linear')
        assert rs.main() is True
        aligned = rs.lfp_img_align
        t["align"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        ra = LfpRearranger(aligned, cfg=cfg, sta=sta)
        ra.main()
        vp = ra.vp_img_arr
        t["viewpoints"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        ss = LfpShiftAndSum(vp_img_arr=vp, cfg=cfg, sta=sta)
        assert ss.main() is True
        stack = ss.refo_stack
        t["refocus"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        # lfp_refocuser/top_level.py:68-70
        stack = LfpContrast().auto_hist_align(stack, ref_img=stack[0], opt=True)
        stack = GammaConverter().srgb_conv(stack)
        t["post"] = time.perf_counter() - t0
    return t, np.asarray(stack)


def _illum_job(job):
    seed, LY, LX, pitch, M, ran_refo = job
    cent = synth.synth_centroids(LY, LX, pitch, 'hex', hex_odd=True, jitter=0.15, origin=(9.5, 9.5), seed=0)
    H, W = int(cent[:, 0].max() + 10.5), int(cent[:, 1].max() + 10.5)
    raw = synth.synth_raw(H, W, 3, seed=seed)
    t, stack = hot_path_once(raw, cent, 'hex', M, ran_refo)
    t["checksum"] = float(np.asarray(stack, dtype=np.float64).sum())
    return t


def illum_sample(LY_full, LX_full, pitch, M, n_slices_full, area_div=16, n_slices=4, procs=None, seed0=0):
    """`procs` processes, each one exposure of a 1/area_div-area crop of the Illum lens grid and `n_slices` slices
    (centred range).  Returns wall seconds of the step and a dict with per-stage CPU seconds and the linear
    extrapolation to the full workload (lenslets x slices; optimistic for the reference, whose centroid lookup is
    quadratic in the number of lenslets: lfp_aligner/lfp_microlenses.py:113-119)."""
    procs = procs or os.cpu_count() or 1
    f = float(np.sqrt(area_div))
    LY, LX = max(8, int(round(LY_full / f))), max(8, int(round(LX_full / f)))
    ran = (-(n_slices // 2), n_slices - n_slices // 2)
    jobs = [(seed0 + k, LY, LX, pitch, M, ran) for k in range(procs)]
    global _RUN_DIR
    t0 = time.perf_counter()
    if procs > 1:
        _RUN_DIR = tempfile.mkdtemp(prefix="pcb_ref_run_")     # inherited by the forked workers
        try:
            with mp.get_context("fork").Pool(procs) as pool:
                res = pool.map(_illum_job, jobs)
        finally:
            shutil.rmtree(_RUN_DIR, True)
            _RUN_DIR = None
    else:
        res = [_illum_job(jobs[0])]
    wall = time.perf_counter() - t0
    mean = {k: float(np.mean([r[k] for r in res])) for k in ("align", "viewpoints", "refocus", "post")}
    lens_ratio = (LY_full * LX_full) / float(LY * LX)
    slice_ratio = n_slices_full / float(n_slices)
    # one full exposure on one core, linearly extrapolated from the per-process stage times
    t_full_1core = (mean["align"] + mean["viewpoints"]) * lens_ratio + (mean["refocus"] + mean["post"]) * lens_ratio * slice_ratio
    # the step processed `procs` sample exposures in `wall` seconds: aggregate throughput over all cores
    t_full_allcores = wall * (t_full_1core / sum(mean.values())) / procs
    info = {"lens_grid": [LY, LX], "slices": n_slices, "processes": procs, "stage_s_per_process": mean,
            "wall_s": wall, "lens_ratio": lens_ratio, "slice_ratio": slice_ratio,
            "full_exposure_s_one_core_extrapolated": t_full_1core,
            "full_exposure_s_all_cores_extrapolated": t_full_allcores}
    return wall, info


def config1_full(with_white=True):
    """BASELINE.json configs[0] in full through the reference's stage classes: rec grid 186 x 279 lenslets, pitch 11 px,
    M = 9, white image, LfpAligner (de-vignetting + local resampling) -> LfpExtractor -> LfpRefocuser, ran_refo [-1, 1].
    One process (the reference is single-threaded).  Returns stage seconds + output checksums."""
    _private_reference()
    ref_shim.load_reference()
    from plenopticam.lfp_aligner import LfpAligner
    from plenopticam.lfp_extractor import LfpExtractor
    from plenopticam.lfp_refocuser import LfpRefocuser
    w = synth.config1_workload()
    cent, lfp, wht = synth.config1_data(w)
    t = {}
    with tempfile.TemporaryDirectory() as tmp, warnings.catch_warnings():
        warnings.simplefilter("ignore")
        cfg, sta = _quiet_cfg(tmp, cent, 'rec', w["M"], w["ran_refo"])
        cfg.params[cfg.opt_vign] = bool(with_white)
        t0 = time.perf_counter()
        al = LfpAligner(lfp_img=lfp, cfg=cfg, sta=sta, wht_img=wht if with_white else None)
        assert al.main() is True
        aligned = np.asarray(al.lfp_img)
        t["align"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        ex = LfpExtractor(aligned, cfg=cfg, sta=sta)
        assert ex.main() is True
        vp = np.asarray(ex.vp_img_linear)
        t["extract"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        rf = LfpRefocuser(vp, cfg=cfg, sta=sta)
        assert rf.main() is True
        stack = np.asarray(rf.refo_stack)
        t["refocus"] = time.perf_counter() - t0
    t["total"] = t["align"] + t["extract"] + t["refocus"]
    return {"stage_s": t, "lens_grid": [w["LY"], w["LX"]], "M": w["M"], "ran_refo": list(w["ran_refo"]),
            "raw_shape": list(lfp.shape), "aligned_shape": list(aligned.shape), "stack_shape": list(stack.shape),
            "aligned_sum": int(aligned.astype(np.int64).sum()), "stack_sum": float(stack.astype(np.float64).sum())}, \
        (aligned, vp, stack)


if __name__ == "__main__":
    import json
    if len(sys.argv) > 1 and sys.argv[1] == "config1":
        print(json.dumps(config1_full()[0]))
    else:
        wall, info = illum_sample(434, 541, 14.25, 15, 64)
        print(json.dumps(info))
