#!/usr/bin/env python
"""Resampling a float32 exposure (what LfpAligner hands to LfpResampler after de-vignetting or rotation) against the
uint16 exposure of the benchmark, Illum size.   python tools/bench_resample_f32.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import argparse
    import torch
    import bench
    from plenopticam_b200 import ops
    from plenopticam_b200.lfp_aligner import build_tables
    w = bench.workload(argparse.Namespace(small=False, slices=64))
    cent = bench.make_centroids(w)
    dev = torch.device("cuda", 0)
    tabs = ops.LensTables(build_tables(cent, (w["H"], w["W"], w["C"]), w["M"], w["pat"]))
    raw16 = bench.synth_raw_device(torch, w["H"], w["W"], w["C"], seed=1, device=dev)
    raw32 = raw16.to(torch.float32)
    outs = {}
    for name, raw in (("uint16", raw16), ("float32", raw32)):
        out = torch.empty(tabs.out_shape(raw), dtype=torch.uint16, device=dev)
        ts = []
        for it in range(8):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            ops.resample_u16(raw, tabs, out=out)
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        outs[name] = out
        print("%-8s exposure: min/max pass + uint16 pass %.3f ms" % (name, float(np.median(ts[2:]))))
    d = (outs["uint16"].view(torch.int16).to(torch.int32) & 0xffff) - (outs["float32"].view(torch.int16).to(torch.int32) & 0xffff)
    print("max |uint16-path - float32-path| = %d LSB, differing samples %.4f %%" % (int(d.abs().max()), float((d != 0).float().mean()) * 100))


if __name__ == "__main__":
    main()
