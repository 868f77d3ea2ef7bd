#!/usr/bin/env python
"""Illum-size timings of the small stages either side of the benchmarked path (CUDA events, median of 7, L2 flushed):
LfpCropper, LfpRearranger.decompose, LfpContrast of the extractor, LfpScheimpflug, exporter quantisation, raw unpack,
and the host-side centroid tables.   python tools/bench_misc.py > gpurun_out/misc.txt"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import argparse
    import torch
    import bench
    from plenopticam_b200 import ops
    from plenopticam_b200.lfp_aligner import build_tables
    from plenopticam_b200.lfp_reader.bayer import comp_bayer_device
    dev = torch.device("cuda", 0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    gen = torch.Generator(device=dev).manual_seed(3)

    def timed(fn, nbytes, name):
        ts = []
        for it in range(9):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ms = float(np.median(ts[2:]))
        print("%-58s %.3f ms  %5.0f GB/s algorithmic (%d MB)" % (name, ms, nbytes / ms / 1e6, nbytes >> 20))

    M, S, T = 15, 434, 625
    aligned = torch.randint(0, 65536, (S * M, T * M, 3), generator=gen, device=dev, dtype=torch.int32).to(torch.uint16)
    b_al = aligned.numel() * 2
    timed(lambda: ops.crop_micro_images(aligned, 15, 13), b_al + b_al * 169 // 225, "LfpCropper 15 -> 13 (crop_kernel)")
    vp = ops.viewpoints(aligned, M)
    timed(lambda: ops.viewpoints_inverse(vp), 2 * b_al, "LfpRearranger.decompose_viewpoints")
    timed(lambda: ops.viewpoints_contrast(vp), b_al + 2 * b_al, "LfpContrast.main of the extractor (uint16 LF -> sRGB float32)")
    del aligned
    stack = torch.rand((64, S, T, 3), generator=gen, device=dev, dtype=torch.float32)
    b_st = stack.numel() * 4
    timed(lambda: ops.scheimpflug(stack, 'skew up'), 2 * S * T * 3 * 4 + S * T * 3 * 2, "LfpScheimpflug (64 slices, skew up) + uint16_norm")
    timed(lambda: ops.uint16_norm(stack), 2 * b_st + b_st // 2, "exporter uint16_norm of the 64-slice stack (min/max + quantise)")
    timed(lambda: ops.uint8_norm(stack), 2 * b_st + b_st // 4, "exporter uint8_norm of the 64-slice stack")
    W, H = 7728, 5368
    packed = torch.randint(0, 256, (W * H * 10 // 8,), generator=gen, device=dev, dtype=torch.int32).to(torch.uint8)
    timed(lambda: comp_bayer_device(packed, (W, H), 10), packed.numel() + W * H * 2, "LfpDecoder.comp_bayer, 10-bit Illum mosaic")
    w = bench.workload(argparse.Namespace(small=False, slices=64))
    cent = bench.make_centroids(w)
    t0 = time.perf_counter()
    build_tables(cent, (w["H"], w["W"], w["C"]), w["M"], w["pat"])
    print("%-58s %.1f ms (host, numpy; 234 794 centroids -> lens / hex tables)" % ("LfpMicroLenses tables (build_tables)",
                                                                                   (time.perf_counter() - t0) * 1e3))


if __name__ == "__main__":
    main()
