#!/usr/bin/env python
"""Slice-sharded refined stack on N GPUs (BASELINE configs[3]): where the time goes.  Per rank, CUDA events around the
prefilter, the entry barrier, the horizontal + vertical passes (the vertical pass stores into every rank's stack), the
exit barrier and the replicated colour management; max over ranks of each.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/dist_refocus_check.py [--small]
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
    ap.add_argument("--reps", type=int, default=7)
    ap.add_argument("--shard", default="rows", choices=["rows", "slices"])
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from plenopticam_b200 import ops, parallel as par
    from plenopticam_b200.lfp_refocuser.refinement import device_plan, prefilter, refocus_refi

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
    a_list = list(range(-2 * M, 2 * M))
    plan = device_plan(S, T, M, a_list, mask, dev)
    lo, hi = par.shard_bounds(len(a_list), rank, world)
    peer = par.PeerStack((len(a_list), S, T, 3))
    stream = torch.cuda.current_stream()
    names = ["prefilter", "barrier_in", "passes+peer_stores", "barrier_out", "finish"]
    acc = []
    for rep in range(args.reps + 2):
        dist.barrier()
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
        ev[0].record(stream)
        if args.shard == "rows":
            y_lo, y_hi = par.row_band(S, rank, world, plan.halo_rows())
            rows = plan.rows(0, len(a_list), y_lo, y_hi)
            coef = prefilter(vp, plan, rows=rows)
        else:
            coef = prefilter(vp, plan)
        ev[1].record(stream)
        peer.barrier()
        ev[2].record(stream)
        if args.shard == "rows":
            refocus_refi(vp, a_list, mask, plan=plan, fanout=peer.fanout(0), coef=coef, row_range=(y_lo, y_hi))
        else:
            refocus_refi(vp, a_list, mask, plan=plan, fanout=peer.fanout(lo), coef=coef, n_range=(lo, hi))
        ev[3].record(stream)
        peer.barrier()
        ev[4].record(stream)
        out = par.finish_stack(peer.local)
        ev[5].record(stream)
        torch.cuda.synchronize()
        t = torch.tensor([ev[k].elapsed_time(ev[k + 1]) for k in range(5)] + [ev[0].elapsed_time(ev[5])], device=dev,
                         dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rep >= 2:
            acc.append(t.cpu().numpy())
    if rank == 0:
        med = np.median(np.asarray(acc), axis=0)
        if args.shard == "rows":
            print("rank 0 rows:", {k: getattr(rows, k) for k in ("out_lo", "out_hi", "h_lo", "h_hi", "tmp_lo", "tmp_hi")})
        print(json.dumps({"n_gpus": world, "shard": args.shard, "slices": len(a_list), "ms_max_over_ranks": dict(zip(names + ["total"], med.tolist()))}))
    peer.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
