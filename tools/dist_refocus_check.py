#!/usr/bin/env python
"""Slice-sharded refocus stack on N GPUs (BASELINE config 4): correctness against the unsharded result on rank 0 and
device timing (max over ranks) of the integer and the refinement path.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/dist_refocus_check.py [--small]

Prints one JSON line per path on rank 0.
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--small", action="store_true")
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from plenopticam_b200 import ops, parallel as par
    from plenopticam_b200.lfp_refocuser.refinement import refocus_refi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist.init_process_group("nccl", device_id=dev)
    M, S, T = (9, 40, 52) if args.small else (15, 434, 625)
    g = torch.Generator(device=dev)
    g.manual_seed(0)                                           # every rank holds the same light field
    vp = torch.randint(0, 65536, (M, M, S, T, 3), generator=g, device=dev, dtype=torch.int32).to(torch.uint16)
    mask = ops.aperture_mask(M)
    jobs = [("int", list(range(-32, 32)), False), ("refi", list(range(-2 * M, 2 * M)), True)]
    for name, a_list, refi in jobs:
        full = par.sharded_refocus_stack(vp, a_list, postprocess=True, refi=refi)        # warm-up + result
        torch.cuda.synchronize()
        ok = None
        if rank == 0:
            if refi:
                ref = ops.refocuser_postprocess(refocus_refi(vp, a_list, mask))
            else:
                ref = ops.refocuser_postprocess(ops.refocus_int(vp, a_list))
            ok = bool(torch.equal(full, ref))
        ts = []
        for _ in range(args.reps):
            dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            par.sharded_refocus_stack(vp, a_list, postprocess=True, refi=refi)
            e1.record()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ts.append(float(t.item()))
        if rank == 0:
            ms = float(np.median(ts))
            print(json.dumps({"path": name, "n_gpus": world, "slices": len(a_list), "shape": [M, M, S, T, 3],
                              "equals_unsharded": ok, "ms_per_stack": ms, "slices_per_s": len(a_list) / (ms * 1e-3),
                              "timing": "CUDA events, max over ranks, median of %d; includes prefilter (refi), "
                                        "contrast/sRGB and the NCCL all-gather of the finished slices" % args.reps}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
