#!/usr/bin/env python
"""Run the Illum-shaped pipeline a few times (for ncu captures of single kernels).

    python tools/profile_stage.py [--steps 2] [--slices 64] [--small]

Prints per-stage CUDA-event times.  Use with
    ncu --set full --clock-control none --import-source on -k regex:<kernel> -s <skip> -c 1 -o gpurun_out/prof ...
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--slices", type=int, default=64)
    ap.add_argument("--small", action="store_true")
    ap.add_argument("--no-fuse-post", action="store_true")
    ap.add_argument("--explicit-gather", action="store_true",
                    help="separate viewpoints kernel + refocus from the 4-D array (default: refocus reads the aligned image)")
    args = ap.parse_args()

    import torch
    import bench
    from plenopticam_b200.pipeline import LightFieldPipeline

    bargs = argparse.Namespace(small=args.small, slices=args.slices)
    w = bench.workload(bargs)
    cent = bench.make_centroids(w)
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    pipe = LightFieldPipeline(cent, (w["H"], w["W"], w["C"]), np.uint16, w["M"], w["pat"], ran_refo=w["ran_refo"])
    raw = bench.synth_raw_device(torch, w["H"], w["W"], w["C"], seed=1, device=dev)
    names = ["align", "viewpoints", "refocus", "post"]
    for it in range(args.steps):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        ev[0].record()
        pipe.align(raw)
        ev[1].record()
        if args.explicit_gather or args.no_fuse_post:
            pipe.viewpoints()
            ev[2].record()
            if args.no_fuse_post or pipe.refocus_post() is False:
                pipe.refocus()
                ev[3].record()
                pipe.post_process()
            else:
                ev[3].record()              # fused epilogue: "refocus" below is 0, "post" carries the whole stage
        else:
            ev[2].record()                  # no separate gather: the refocus kernel reads the aligned image
            ev[3].record()
            if pipe._refocus_from_aligned() is False:
                raise SystemExit("no aligned-direct refocus form for this workload")
        ev[4].record()
        torch.cuda.synchronize()
        print("step %d: " % it + ", ".join("%s %.3f ms" % (n, ev[k].elapsed_time(ev[k + 1])) for k, n in enumerate(names)))


if __name__ == "__main__":
    main()
