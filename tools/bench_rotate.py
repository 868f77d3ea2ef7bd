#!/usr/bin/env python
"""Time LfpRotator's image rotation (scipy.ndimage.rotate, order 3, reshape=False) at Illum size on the device, per
kernel, and scipy itself on a crop on the host.   python tools/bench_rotate.py > gpurun_out/rotate.txt"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from scipy import ndimage
    import bench
    from plenopticam_b200 import ops, _native as nat
    H, W, C = 5368, 7728, 3
    raw = bench.synth_raw_device(torch, H, W, C, seed=3, device=torch.device("cuda", 0))
    rad = 0.0021
    tmp = torch.empty((H, W, C), dtype=torch.float32, device=raw.device)
    coef = torch.empty_like(tmp)
    L, p, st = nat.lib(), ops._ptr, ops._stream()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=raw.device)
    t_pf, t_rot = [], []
    for it in range(9):
        flush.zero_()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        nat.check(L.pcb_spline_prefilter(p(raw), ops._pcb_dtype(raw), H, W, C, p(tmp), p(coef), st))
        e[1].record()
        nat.check(L.pcb_rotate_sample(p(coef), H, W, C, rad, p(tmp), st))
        e[2].record()
        torch.cuda.synchronize()
        t_pf.append(e[0].elapsed_time(e[1])); t_rot.append(e[1].elapsed_time(e[2]))
    pf, rt = float(np.median(t_pf[2:])), float(np.median(t_rot[2:]))
    b_raw, b_f32 = H * W * C * 2, H * W * C * 4
    print("Illum exposure %d x %d x %d uint16, angle %.4f rad" % (H, W, C, rad))
    print("spline prefilter (rows + columns): %.3f ms  (algorithmic %d MB: uint16 in, float32 coefficients out -> %.0f GB/s)"
          % (pf, (b_raw + b_f32) >> 20, (b_raw + b_f32) / pf / 1e6))
    print("rotation gather (4 x 4 taps):      %.3f ms  (algorithmic %d MB: coefficients in, float32 out -> %.0f GB/s)"
          % (rt, (2 * b_f32) >> 20, 2 * b_f32 / rt / 1e6))
    print("total %.3f ms" % (pf + rt))
    crop = ops.to_host(raw[:1024, :1024].contiguous()).astype(np.float64)
    t0 = time.perf_counter()
    ref = ndimage.rotate(crop, np.degrees(rad), reshape=False, order=3)
    dt = time.perf_counter() - t0
    got = ops.to_host(ops.rotate(raw[:1024, :1024].contiguous(), rad))
    print("scipy.ndimage.rotate on a 1024 x 1024 crop: %.2f s -> %.0f s extrapolated to the exposure (1 core); "
          "max |device - scipy| / 65535 on the crop = %.2e" % (dt, dt * H * W / (1024 * 1024),
                                                                 float(np.abs(got - ref).max()) / 65535))


if __name__ == "__main__":
    main()
