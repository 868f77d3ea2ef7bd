#!/usr/bin/env python
"""Time the de-vignetting row at Illum size on the device (per kernel, CUDA events) and the numpy oracle beside it.

    python tools/bench_devignette.py [--reps 10]
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=10)
    args = ap.parse_args()
    import torch
    from plenopticam_b200 import ops
    from plenopticam_b200.lfp_aligner.lfp_devignetter import gauss_kernel_1d
    from oracle import oracle_np as onp

    H, W, Cn = 5368, 7728, 3
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev)
    g.manual_seed(0)
    lfp = torch.randint(0, 65536, (H, W, Cn), generator=g, device=dev, dtype=torch.int32).to(torch.uint16)
    wht = torch.randint(20000, 65536, (H, W, Cn), generator=g, device=dev, dtype=torch.int32).to(torch.uint16)
    gk = gauss_kernel_1d(14)
    for out_dtype, name in ((torch.float32, "f32 out (device pipeline)"), (torch.float64, "f64 out (API parity)")):
        ts = []
        for _ in range(args.reps + 2):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out, wn, stats = ops.devignette(lfp, wht, gk, out_dtype=out_dtype)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = float(np.median(ts[2:]))
        npx = H * W
        # white read + gray write | 8 select passes read gray | normalize rw | blur rows r/w, cols 2r | divide
        alg = npx * (Cn * 2 + 8) + 8 * npx * 8 + 2 * npx * 8 + 3 * npx * 8 + npx * 8 + npx * Cn * (2 + out.element_size())
        print(json.dumps({"row": "devignette", "variant": name, "ms": ms, "algorithmic_GB": alg / 1e9,
                          "achieved_GBps": alg / ms / 1e6, "shape": [H, W, Cn]}))
    # numpy oracle on a 1/16-area crop, extrapolated
    h, w = H // 4, W // 4
    lh, wh = ops.to_host(lfp[:h, :w].contiguous()), ops.to_host(wht[:h, :w].contiguous())
    t0 = time.perf_counter()
    onp.devignette(lh, wh, 14.0)
    t = (time.perf_counter() - t0) * 16
    print(json.dumps({"row": "devignette", "variant": "numpy oracle, 1 core, 1/16-area crop x16", "ms": t * 1e3}))


if __name__ == "__main__":
    main()
