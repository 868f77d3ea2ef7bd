#!/usr/bin/env python
"""Randomised parity sweep of the CUDA path against the numpy oracle (GPU box; minutes).

    python tools/stress_parity.py [--cases 60] [--seed 0]

Resampling (rec / hex, flips, odd micro-image sizes, jittered and rotated centroids, border lenses, uint8 / uint16 /
float32 exposures) and the integer refocus stack (angular sizes 3..15, random parameter lists incl. non-consecutive
and huge ones, circular / random / contiguous view masks, narrow and wide rows) through every kernel form.
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=60)
    ap.add_argument("--seed", type=int, default=0)
    args = ap.parse_args()
    import torch
    from oracle import oracle_np as onp
    from plenopticam_b200 import ops
    from plenopticam_b200.lfp_aligner import build_tables

    rng = np.random.default_rng(args.seed)
    bad = 0
    # ---------------------------------------------------------------- resampling
    for case in range(args.cases):
        M = int(rng.choice([3, 5, 7, 9, 11, 13, 15, 17, 19]))
        pat = str(rng.choice(['rec', 'hex']))
        LY, LX = int(rng.integers(3, 14)), int(rng.integers(3, 40))
        pitch = M * float(rng.uniform(0.9, 1.3)) + 0.3
        jitter = float(rng.choice([0.0, 0.15, 0.45]))
        rot = float(rng.choice([0.0, 0.0, 0.08]))
        origin = (float(rng.uniform(M / 2 - 2, M / 2 + 3)), float(rng.uniform(M / 2 - 2, M / 2 + 3)))   # some border lenses
        cent = onp.synth_centroids(LY, LX, pitch, pat, hex_odd=bool(rng.integers(0, 2)), jitter=jitter, rot_deg=rot,
                                   origin=origin, seed=int(rng.integers(1 << 30)))
        H = int(cent[:, 0].max() + rng.integers(M // 2, M + 3))
        W = int(cent[:, 1].max() + rng.integers(M // 2, M + 3))
        dt = rng.choice([np.uint16, np.uint16, np.uint8, np.float32])
        raw = onp.synth_raw(H, W, 3, seed=case, dtype=np.uint16)
        raw = (raw >> 8).astype(np.uint8) if dt == np.uint8 else (raw / 65535.0).astype(np.float32) if dt == np.float32 else raw
        flip = bool(rng.integers(0, 2))
        try:
            tabs = ops.LensTables(build_tables(cent, raw.shape, M, pat))
            got, _ = ops.resample_f32(ops.to_device(raw), tabs, flip=flip, want_minmax=False), None
            got = ops.to_host(got[0] if isinstance(got, tuple) else got)
            ref = onp.resample(raw, cent, M, pat, flip=flip)
            scale = max(float(np.abs(ref).max()), 1e-12)
            err = float(np.abs(got - ref).max() / scale)
            q, _ = ops.resample_u16(ops.to_device(raw), tabs, flip=flip)
            dq = np.abs(ops.to_host(q).astype(np.int64) - onp.uint16_norm(ref).astype(np.int64))
            ok = err <= 1e-5 and dq.max() <= 1
        except Exception as e:                                   # noqa: BLE001
            ok, err, dq = False, repr(e), np.zeros(1)
        if not ok:
            bad += 1
            print("RESAMPLE FAIL", dict(case=case, M=M, pat=pat, LY=LY, LX=LX, dt=np.dtype(dt).name, flip=flip, err=err,
                                        lsb=int(dq.max())))
    print("resampling: %d cases, %d failures" % (args.cases, bad))
    # ---------------------------------------------------------------- refocus
    bad_r = 0
    for case in range(args.cases):
        M = int(rng.choice([3, 5, 7, 9, 11, 13, 15]))
        S, T = int(rng.integers(1, 50)), int(rng.choice([4, 17, 61, 130, 700, 1100, 2300]))
        if S * T * M * M > 6e6:
            S = max(1, int(6e6 / (T * M * M)))
        kind = int(rng.integers(0, 4))
        if kind == 0:
            lo = int(rng.integers(-40, 10)); a_list = list(range(lo, lo + int(rng.integers(1, 70))))
        elif kind == 1:
            a_list = [int(v) for v in rng.integers(-12, 13, size=int(rng.integers(1, 9)))]
        elif kind == 2:
            a_list = [int(v) for v in rng.choice([-500, -77, -1, 0, 3, 200], size=3)]
        else:
            a_list = list(range(-3, 4))
        mk = int(rng.integers(0, 3))
        if mk == 0:
            mask = ops.aperture_mask(M)
        elif mk == 1:
            mask = rng.random((M, M)) < 0.35
            mask[M // 2, M // 2] = False
        else:
            mask = np.ones((M, M), dtype=bool)
            for j in range(M):
                lo = int(rng.integers(0, M)); mask[j, lo:lo + int(rng.integers(0, M + 1))] = False
            mask[0, 0] = False
        vp = rng.integers(0, 65536, size=(M, M, S, T, 3), dtype=np.uint16)
        dev = ops.to_device(vp)
        c = M // 2
        ref = np.zeros((len(a_list), S, T, 3))
        for k, a in enumerate(a_list):
            for j in range(M):
                yi = np.clip(np.arange(S) + a * (c - j), 0, S - 1)
                for i in range(M):
                    if not mask[j, i]:
                        xi = np.clip(np.arange(T) + a * (c - i), 0, T - 1)
                        ref[k] += vp[j, i][yi][:, xi]
        ref = (ref / M).astype(np.float32)
        for kern in ("", "px", "rows", "direct"):
            if kern:
                os.environ["PCB_REFOCUS_KERNEL"] = kern
            else:
                os.environ.pop("PCB_REFOCUS_KERNEL", None)
            try:
                got = ops.to_host(ops.refocus_int(dev, a_list, view_zeroed=mask))
                ok = np.array_equal(got, ref)
            except Exception as e:                               # noqa: BLE001
                ok, got = False, repr(e)
            if not ok:
                bad_r += 1
                print("REFOCUS FAIL", dict(case=case, kern=kern, M=M, S=S, T=T, a=a_list[:6], n=len(a_list), mask=mk,
                                           info=got if isinstance(got, str) else float(np.abs(got - ref).max())))
        os.environ.pop("PCB_REFOCUS_KERNEL", None)
        # colour management fused into the epilogue == the separate kernels, bit for bit (whatever form applies)
        try:
            mm = ops.new_minmax()
            sep = ops.refocuser_postprocess(ops.refocus_int(dev, a_list, view_zeroed=mask, mm=mm), mm=mm)
            fus = ops.refocus_int_srgb(dev, a_list, view_zeroed=mask)
            ok = bool(torch.equal(sep, fus))
        except Exception as e:                                   # noqa: BLE001
            ok = False
            print(repr(e))
        if not ok:
            bad_r += 1
            print("FUSED POST FAIL", dict(case=case, M=M, S=S, T=T, a=a_list[:6], n=len(a_list), mask=mk))
    print("refocus: %d cases x 4 kernel forms + fused epilogue, %d failures" % (args.cases, bad_r))
    # ---------------------------------------------------------------- refinement, viewpoints, post-processing
    from plenopticam_b200.lfp_refocuser.refinement import refocus_refi
    bad_o = 0
    for case in range(max(4, args.cases // 6)):
        M = int(rng.choice([3, 5, 7, 9]))
        S, T = int(rng.integers(6, 70)), int(rng.integers(6, 200))
        r0 = int(rng.integers(-2, 1)); ran = (r0, r0 + int(rng.integers(1, 3)))
        dt = rng.choice([np.uint16, np.float32])
        vp = (rng.random((M, M, S, T, 3)) * (65535 if dt == np.uint16 else 1)).astype(dt)
        try:
            got = ops.to_host(refocus_refi(ops.to_device(vp), range(M * ran[0], M * ran[1]), ops.aperture_mask(M)))
            ref = onp.refocus_refi(vp, ran, M)
            err = float(np.abs(got - ref).max() / max(np.abs(ref).max(), 1e-12))
            ok = err <= 2e-5
        except Exception as e:                                   # noqa: BLE001
            ok, err = False, repr(e)
        if not ok:
            bad_o += 1
            print("REFI FAIL", dict(case=case, M=M, S=S, T=T, ran=ran, dt=np.dtype(dt).name, err=err))
    for case in range(args.cases):
        M_in = int(rng.choice([3, 5, 7, 9, 11, 15])); M_out = int(rng.choice([m for m in (3, 5, 7, 9, 11, 15) if m <= M_in]))
        S, T, Cn = int(rng.integers(1, 40)), int(rng.integers(1, 300)), int(rng.choice([1, 3, 4]))
        dt = rng.choice([np.uint8, np.uint16, np.float32])
        al = (rng.random((S * M_in, T * M_in, Cn)) * 255).astype(dt)
        try:
            vpd = ops.viewpoints(ops.to_device(al), M_in, M_out)
            ref = onp.compose_viewpoints(onp.crop_micro_images(al, M_in, M_out), M_out)
            ok = np.array_equal(ops.to_host(vpd), ref)
            back = ops.to_host(ops.viewpoints_inverse(vpd))
            ok = ok and np.array_equal(back, onp.crop_micro_images(al, M_in, M_out))
        except Exception as e:                                   # noqa: BLE001
            ok = False
            print(repr(e))
        if not ok:
            bad_o += 1
            print("VIEWPOINTS FAIL", dict(case=case, M_in=M_in, M_out=M_out, S=S, T=T, C=Cn, dt=np.dtype(dt).name))
    for case in range(max(4, args.cases // 4)):
        N, S, T = int(rng.integers(1, 9)), int(rng.integers(2, 60)), int(rng.integers(2, 200))
        stack = (rng.random((N, S, T, 3)) ** 2 * float(rng.choice([1.0, 300.0, 70000.0]))).astype(np.float64)
        if rng.integers(0, 3) == 0:
            stack[0, : S // 2] = 0.0                            # zero lower percentile -> Normalizer's min fallback
        try:
            got = ops.to_host(ops.refocuser_postprocess(ops.to_device(stack.astype(np.float32))))
            ref = onp.refocuser_postprocess(stack.astype(np.float32).astype(np.float64))
            err = float(np.abs(got.astype(np.float64) - ref).max())
            ok = err <= 1e-5
        except Exception as e:                                   # noqa: BLE001
            ok, err = False, repr(e)
        if not ok:
            bad_o += 1
            print("POST FAIL", dict(case=case, N=N, S=S, T=T, err=err))
    print("refinement / viewpoints / post: %d failures" % bad_o)
    torch.cuda.synchronize()
    sys.exit(1 if bad or bad_r or bad_o else 0)


if __name__ == "__main__":
    main()
