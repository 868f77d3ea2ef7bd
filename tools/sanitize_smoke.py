#!/usr/bin/env python
"""Workload for `compute-sanitizer --tool memcheck`: smoke() plus one small pass of every kernel family whose index
arithmetic depends on the data layout (full aperture mask -> the first / last view rows reach the ends of the light-field
buffer: the clamped TMA copies of refocus_pxp; rec and hex resampling; rotation; refinement; viewpoints both ways; crop;
contrast).  Sizes are small because memcheck slows kernels ~50x."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import __graft_entry__ as ge
    from plenopticam_b200 import ops
    from plenopticam_b200.lfp_aligner import build_tables
    from plenopticam_b200.lfp_refocuser.refinement import refocus_refi
    from plenopticam_b200.misc import synth
    ge.smoke()
    rng = np.random.default_rng(0)
    # a light field whose byte size is NOT a multiple of 16 and whose base address is only 2-byte aligned inside a larger
    # allocation, every view active: the edge rows take the plain-load path of pxp_issue_tile
    M, S, T = 5, 19, 37
    n = M * M * S * T * 3
    pool = torch.zeros(n + 64, dtype=torch.int16, device="cuda")
    for off in (0, 1, 3, 8):
        vp = pool[off:off + n].view(torch.uint16).view(M, M, S, T, 3)
        host = rng.integers(0, 65536, size=(M, M, S, T, 3), dtype=np.uint16)
        vp.view(torch.int16).copy_(torch.from_numpy(host.view(np.int16)).cuda())
        full = np.zeros((M, M), dtype=bool)
        a = ops.refocus_int(vp, range(-3, 4), view_zeroed=full)
        b = ops.refocus_int_direct(vp, range(-3, 4), view_zeroed=full)
        assert torch.equal(a, b), off
    for pat in ("rec", "hex"):
        cent = synth.synth_centroids(9, 11, 10.3, pat, jitter=0.2, origin=(7.5, 7.5), seed=3)
        raw = synth.synth_raw(int(cent[:, 0].max() + 9), int(cent[:, 1].max() + 9), 3, seed=4)
        tabs = ops.LensTables(build_tables(cent, raw.shape, 9, pat))
        al, _ = ops.resample_u16(ops.to_device(raw), tabs)
        vp = ops.viewpoints(al, 9)
        assert torch.equal(ops.viewpoints_inverse(vp).view(torch.int16), al.view(torch.int16))
        ops.crop_micro_images(al, 9, 7)
        ops.rotate(ops.to_device(raw), 0.01)
        ops.viewpoints_contrast(vp)
        refocus_refi(vp, range(-9, 9), ops.aperture_mask(9))
    torch.cuda.synchronize()
    print("sanitize workload ok")


if __name__ == "__main__":
    main()
