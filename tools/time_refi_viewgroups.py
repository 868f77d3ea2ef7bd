#!/usr/bin/env python
"""Feasibility of sharding the refined stack by VIEW ROWS: time prefilter + passes of one GPU when only the view rows of
one of eight balanced groups are active (the others masked), against the whole aperture."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from plenopticam_b200 import ops
from plenopticam_b200.lfp_refocuser import refinement as rf

M, S, T = 15, 434, 625
a_list = list(range(-2 * M, 2 * M))
g = torch.Generator(device="cuda"); g.manual_seed(0)
vp = torch.randint(0, 65536, (M, M, S, T, 3), generator=g, device="cuda", dtype=torch.int32).to(torch.uint16)
full = np.asarray(ops.aperture_mask(M)).reshape(M, M).astype(bool)        # True = zeroed
groups = [[7], [6], [8], [5, 2], [9, 12], [4, 1, 0], [10, 13, 14], [3, 11]]
print("views per group:", [int((~full[g_]).sum()) for g_ in groups])
for name, rows in [("all", list(range(M)))] + [("group %d %s" % (k, g_), g_) for k, g_ in enumerate(groups)]:
    mask = np.ones((M, M), dtype=bool)
    mask[rows] = full[rows]
    plan = rf.device_plan(S, T, M, a_list, mask, vp.device)
    ts = []
    for it in range(4):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        ev[0].record()
        coef = rf.prefilter(vp, plan)
        ev[1].record()
        out = rf.refocus_refi(vp, a_list, mask, plan=plan, coef=coef)
        ev[2].record(); torch.cuda.synchronize()
        ts.append((ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])))
    print("%-28s prefilter %.3f ms  passes %.3f ms" % (name, min(t[0] for t in ts), min(t[1] for t in ts)))
