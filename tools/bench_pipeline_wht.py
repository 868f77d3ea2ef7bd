#!/usr/bin/env python
"""Illum-size exposure through LightFieldPipeline with and without a white image (de-vignetting inside the executor)."""
import argparse, sys, numpy as np, torch
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from plenopticam_b200.pipeline import LightFieldPipeline
w = bench.workload(argparse.Namespace(small=False, slices=64))
cent = bench.make_centroids(w)
dev = torch.device("cuda", 0)
raw = bench.synth_raw_device(torch, w["H"], w["W"], w["C"], seed=1, device=dev)
g = torch.Generator(device=dev).manual_seed(2)
wht = (20000 + 30000 * torch.rand((w["H"], w["W"], 3), generator=g, device=dev)).to(torch.int32).to(torch.uint16)
for name, kw in (("no white image", {}), ("white image", {"wht_img": wht})):
    pipe = LightFieldPipeline(cent, (w["H"], w["W"], w["C"]), np.uint16, w["M"], w["pat"], ran_refo=w["ran_refo"], **kw)
    ts, ta = [], []
    for it in range(8):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record(); pipe.align(raw); e[1].record(); pipe.viewpoints(); pipe.refocus_post(); e[2].record()
        torch.cuda.synchronize()
        ta.append(e[0].elapsed_time(e[1])); ts.append(e[0].elapsed_time(e[2]))
    print("%-16s align %.3f ms, whole exposure %.3f ms" % (name, float(np.median(ta[2:])), float(np.median(ts[2:]))))
