#!/usr/bin/env python
"""Time the refinement path at Illum shape: operator construction (host, one-off) and the device passes."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from plenopticam_b200 import ops
from plenopticam_b200.lfp_refocuser import refinement as rf

M, S, T = 15, 434, 625
r0, r1 = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (-2, 2)
a_list = list(range(M * r0, M * r1))
t = time.time()
Ay, Bx = rf.build_operators(S, T, M, a_list)
print("operators: %d shifts, Ky=%d Kx=%d, %.1f s host" % (len(Ay.shifts), Ay.K, Bx.K, time.time() - t))
g = torch.Generator(device="cuda"); g.manual_seed(0)
vp = torch.randint(0, 65536, (M, M, S, T, 3), generator=g, device="cuda", dtype=torch.int32).to(torch.uint16)
mask = ops.aperture_mask(M)
for it in range(3):
    torch.cuda.synchronize(); t = time.time()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = rf.refocus_refi(vp, a_list, mask)
    e1.record(); torch.cuda.synchronize()
    print("run %d: %d slices, device %.2f ms (%.3f ms/slice), wall %.2f s" % (it, len(a_list), e0.elapsed_time(e1), e0.elapsed_time(e1) / len(a_list), time.time() - t))
