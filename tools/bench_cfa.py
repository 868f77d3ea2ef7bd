#!/usr/bin/env python
"""Time CfaProcessor.safe_bayer_awb at the Illum sensor size (5368 x 7728 mosaic) on the device and, on a crop,
the numpy restatement on the host.   python tools/bench_cfa.py > gpurun_out/cfa.txt"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from plenopticam_b200 import ops
    from oracle import oracle_np as onp
    H, W = 5368, 7728
    g = torch.Generator(device='cuda').manual_seed(1)
    bay = torch.rand((H, W), generator=g, device='cuda', dtype=torch.float32) * 1.1
    wht = 0.5 + 0.6 * torch.rand((H, W), generator=g, device='cuda', dtype=torch.float32)
    gains = [1.6, 1.0, 1.3, 1.0]
    out = torch.empty((H, W), dtype=torch.float64, device='cuda')
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
    for name, w, e in (("white image + soft clipping", wht, -0.5), ("no white image, no soft clipping", None, None)):
        ts = []
        for it in range(13):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            ops.cfa_safe_awb(bay, w, gains, e, out=out)
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ms = float(np.median(ts[3:]))
        nbytes = H * W * (4 + (4 if w is not None else 0) + 8)
        print("%-36s %.3f ms  %.0f GB/s algorithmic (%d MB)" % (name, ms, nbytes / ms / 1e6, nbytes >> 20))
    hb, hw = bay[:1024, :1024].cpu().numpy(), wht[:1024, :1024].cpu().numpy()
    t0 = time.perf_counter()
    ref = onp.cfa_safe_bayer_awb(hb, hw, [1.3, 1.0, 1.6, 1.0], "GRBG", -0.5)
    dt = time.perf_counter() - t0
    dev = ops.cfa_safe_awb(bay[:1024, :1024].contiguous(), wht[:1024, :1024].contiguous(), gains, -0.5).cpu().numpy()
    print("numpy restatement, 1024 x 1024 crop: %.3f s -> %.1f s extrapolated to the full mosaic (1 core); "
          "max |device - numpy| on the crop = %.2e" % (dt, dt * H * W / (1024 * 1024), float(np.abs(dev - ref).max())))


if __name__ == "__main__":
    main()
