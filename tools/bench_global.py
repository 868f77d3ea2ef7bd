#!/usr/bin/env python
"""Time smp_meth='global' at Illum size: host planning (per camera) and the device passes (per exposure)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import bench
    import argparse
    from oracle import oracle_np as onp
    from plenopticam_b200 import ops
    from plenopticam_b200.lfp_aligner import lfp_global_resampler as glob
    w = bench.workload(argparse.Namespace(small=False, slices=64))
    cent = bench.make_centroids(w)
    dev = torch.device("cuda", 0)
    raw = bench.synth_raw_device(torch, w["H"], w["W"], w["C"], seed=1, device=dev)
    hex_odd = onp.get_hex_direction(cent)
    t0 = time.perf_counter()
    tra, oshape, pitch = glob.projective_plan(cent, (w["H"], w["W"], 3), 'hex', hex_odd)
    t1 = time.perf_counter()
    ts = []
    plan = None
    for it in range(6):
        torch.cuda.synchronize()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        e[0].record()
        prj = ops.warp_projective(raw, tra, oshape)
        e[1].record()
        if plan is None:
            th = time.perf_counter()
            plan = glob.hex_plan(pitch, w["LY"], w["LX"], int(prj.shape[1]), hex_odd)
            t_hexplan = time.perf_counter() - th
            e[1].record()
        al = ops.hex_align(prj, plan, w["LY"])
        e[2].record()
        q = ops.quantize_u16(al, ops.minmax_f32(al))
        e[3].record()
        torch.cuda.synchronize()
        ts.append([e[k].elapsed_time(e[k + 1]) for k in range(3)])
    ts = np.median(np.array(ts[1:]), axis=0)
    print(json.dumps({"row": "global_resample", "host_plan_s": {"grid_fit+matrices": t1 - t0, "hex spline operators": t_hexplan},
                      "device_ms": {"warp": float(ts[0]), "hex_align": float(ts[1]), "uint16_norm": float(ts[2])},
                      "pitch": [int(v) for v in pitch], "rectified": [int(v) for v in oshape], "aligned": list(q.shape)}))


if __name__ == "__main__":
    main()
