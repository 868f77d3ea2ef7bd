"""CPU-only checks: the C-ABI library loads and exports every symbol the header declares, and the host-side
logic (table preparation, pitch evaluation, rotation estimate, status/interrupt plumbing) agrees with the
oracle / golden vectors.  No kernel is launched here."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT, golden, golden_names, make_cfg
from oracle import oracle_np as onp


# ------------------------------------------------------------------------------------------------------
# C-ABI
# ------------------------------------------------------------------------------------------------------
def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "plenopticam_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(pcb_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from plenopticam_b200 import _native
    syms = _header_symbols()
    assert len(syms) >= 18
    assert set(syms) == set(_native.exported_symbols())          # binding and header agree
    handle = ctypes.CDLL(_native.build())                        # builds with nvcc if stale; loads without a GPU
    for s in syms:
        assert hasattr(handle, s), s
    L = _native.lib()
    assert L.pcb_version() >= 100
    assert L.pcb_percentile_scratch_bytes() > 0
    # host helper round trip of the order-preserving key encoding
    for f in (0.0, 1.5, -2.25, 65535.0, -1e-30):
        u = np.float32(f).view(np.uint32)
        key = (~u) & 0xFFFFFFFF if u & 0x80000000 else u | 0x80000000
        assert L.pcb_key_to_float(int(key)) == np.float32(f)


def _c_kind(decl):
    """'p' pointer, 'q' int64, 'd' double, 'f' float, 'u' uint32, 'i' int - of one C parameter / return declaration."""
    if "*" in decl or "[" in decl:                     # an array parameter is a pointer
        return "p"
    for key, kind in (("int64_t", "q"), ("double", "d"), ("float", "f"), ("uint32_t", "u"), ("int", "i")):
        if re.search(r"\b%s\b" % key, decl):
            return kind
    raise AssertionError("unclassified C declaration: %r" % decl)


def _ctypes_kind(t):
    if t in (ctypes.c_void_p, ctypes.c_char_p) or issubclass(t, (ctypes._Pointer, ctypes.Structure, ctypes.Array)):
        return "p"
    return {ctypes.c_int64: "q", ctypes.c_double: "d", ctypes.c_float: "f", ctypes.c_uint32: "u", ctypes.c_int: "i"}[t]


def test_binding_signatures_match_the_header_prototypes():
    """Every entry point: the ctypes argument list of the binding has the number and the kinds (pointer / int / int64 /
    double / float) of the parameters the header declares, and the same kind of result - a stale binding would pass
    wrong words to the library without any error."""
    from plenopticam_b200 import _native
    txt = open(os.path.join(ROOT, "include", "plenopticam_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    txt = re.sub(r"^\s*#.*$", "", txt, flags=re.M)
    protos = re.findall(r"([A-Za-z_][A-Za-z0-9_ ]*?[\s\*]+)(pcb_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", txt, flags=re.S)
    seen = {}
    for ret, name, args in protos:
        args = " ".join(args.split())
        params = [] if args in ("", "void") else [a.strip() for a in args.split(",")]
        seen[name] = (_c_kind(ret + " x"), [_c_kind(a) for a in params])
    assert set(seen) == set(_native._SIGNATURES)
    for name, (res, argtypes) in _native._SIGNATURES.items():
        want_res, want_args = seen[name]
        assert _ctypes_kind(res) == want_res, name
        assert [_ctypes_kind(t) for t in argtypes] == want_args, (name, want_args)


def test_argument_validation_without_gpu():
    """Bad arguments are rejected before any CUDA call is made."""
    from plenopticam_b200 import _native
    L = _native.lib()
    assert L.pcb_minmax_reset(None, None) == 1
    assert b"mm" in L.pcb_last_error()
    a = (ctypes.c_int32 * 2)(0, 1)
    z = (ctypes.c_uint8 * 9)()
    assert L.pcb_refocus_int_scratch_bytes(0, 3, 4, 5, 3, a, 2, z) > 0
    assert L.pcb_refocus_int_scratch_bytes(0, 3, 0, 5, 3, a, 2, z) == -1
    with pytest.raises(ValueError):
        _native.check(L.pcb_viewpoints(None, 0, 4, 4, 3, 2, 2, None, None))


def test_no_cpu_fallback():
    import torch
    from plenopticam_b200 import ops
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError):
        ops.to_device(np.zeros((4, 4, 3), dtype=np.uint16))


def test_product_never_imports_oracle():
    for dp, _, files in os.walk(os.path.join(ROOT, "plenopticam_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dp, f)).read()
                assert "oracle" not in src, os.path.join(dp, f)


# ------------------------------------------------------------------------------------------------------
# host table preparation vs oracle
# ------------------------------------------------------------------------------------------------------
def _eval_tables(raw, t, flip=False):
    """Evaluate the kernel's closed form from the prepared tables in numpy float64."""
    raw = np.asarray(raw, dtype=np.float64)
    LY, LX, M, Xs = t["LY"], t["LX"], t["M"], t["Xs"]
    C = raw.shape[2]
    P = np.zeros((LY, LX, M, M, C))
    lens = t["lens"].reshape(LY, LX)
    v = np.arange(M)
    for ly in range(LY):
        for lx in range(LX):
            L = lens[ly, lx]
            if L["oy"] == -2 ** 31:
                continue
            y0, x0 = L["oy"] + v[:, None], L["ox"] + v[None, :]
            wy, wx = float(L["wy"]), float(L["wx"])
            top = raw[y0, x0] + (raw[y0, x0 + 1] - raw[y0, x0]) * wx
            bot = raw[y0 + 1, x0] + (raw[y0 + 1, x0 + 1] - raw[y0 + 1, x0]) * wx
            P[ly, lx] = top + (bot - top) * wy
    if flip:
        P = P[:, :, ::-1, ::-1]
    out = np.zeros((LY * M, Xs * M, C))
    for ly in range(LY):
        if t["hexcol"] is None:
            st = P[ly]
        else:
            hc = t["hexcol"][((ly + int(t["hex_odd"])) & 1) * Xs:][:Xs]
            p0, p1 = P[ly, hc["x0"]], P[ly, hc["x1"]]
            st = p0 + (p1 - p0) * hc["w"].astype(np.float64)[:, None, None, None]
        out[ly * M:(ly + 1) * M] = st.transpose(1, 0, 2, 3).reshape(M, Xs * M, C)
    return out


@pytest.mark.parametrize("name", golden_names("resample_"))
def test_build_tables_reproduce_reference(name):
    from plenopticam_b200.lfp_aligner import build_tables
    g = golden(name)
    t = build_tables(g["centroids"], g["raw"].shape, int(g["M_used"]), str(g["pat_type"]))
    got = _eval_tables(g["raw"], t, flip=bool(g["flip"]))
    ref = g["aligned_f64"]
    assert got.shape == ref.shape
    scale = max(1.0, np.abs(ref).max())
    assert np.abs(got - ref).max() <= 2e-7 * scale          # weights are stored as float32 in the tables
    # centroid indexing is bit-exact: same rounding, same validity decisions
    tab = onp.dense_centroid_table(g["centroids"])
    M = int(g["M_used"])
    ry, rx = onp.rint_arr(tab[..., 0]), onp.rint_arr(tab[..., 1])
    oy = ry - M // 2 + np.floor(tab[..., 0] - ry).astype(np.int64)
    valid = t["lens"]["oy"].reshape(oy.shape) != -2 ** 31
    assert np.array_equal(t["lens"]["oy"].reshape(oy.shape)[valid], oy[valid])
    H, W = g["raw"].shape[:2]
    c = M // 2
    ref_valid = (ry - c - 1 >= 0) & (ry + c + 2 <= H) & (rx - c - 1 >= 0) & (rx + c + 2 <= W)
    assert np.array_equal(valid, ref_valid)


def test_rint_is_round_half_even():
    from plenopticam_b200.lfp_aligner import build_tables
    from plenopticam_b200.misc import rint
    assert [rint(v) for v in (0.5, 1.5, 2.5, -0.5, -1.5)] == [0, 2, 2, 0, -2]
    cent = np.array([[10.5, 11.5, 0, 0], [10.5, 30.5, 0, 1], [30.5, 11.5, 1, 0], [30.5, 30.5, 1, 1]])
    t = build_tables(cent, (60, 60, 3), 5, 'rec')
    # rint(10.5) = 10, rint(11.5) = 12, rint(30.5) = 30: half-way centroids round to even like Python's round
    assert t["lens"]["oy"].tolist() == [8, 8, 28, 28]       # ry - 2 + floor(+0.5) = ry - 2
    assert t["lens"]["ox"].tolist() == [9, 28, 9, 28]       # rx(11.5) = 12, f = -0.5 -> 12 - 2 - 1
    assert np.allclose(t["lens"]["wy"], 0.5) and np.allclose(t["lens"]["wx"], 0.5)


def test_hex_direction_and_pitch_match_oracle():
    from plenopticam_b200.lfp_aligner.lfp_microlenses import avg_pitch, hex_direction
    for odd in (True, False):
        cent = onp.synth_centroids(9, 11, 10.3, 'hex', hex_odd=odd, jitter=0.1, seed=3)
        assert hex_direction(cent) == onp.get_hex_direction(cent) == odd
        assert np.array_equal(avg_pitch(cent), onp.centroid_avg_pitch(cent))


def test_missing_lenslet_raises_index_error():
    from plenopticam_b200.lfp_aligner import build_tables
    cent = onp.synth_centroids(4, 4, 9.0, 'rec', jitter=0.0, seed=0)
    with pytest.raises(IndexError):
        build_tables(np.delete(cent, 5, axis=0), (60, 60, 3), 5, 'rec')


def test_safe_pitch_eval_side_effects():
    """lfp_microlenses.py:121-157: user pitch is clamped to the centroid pitch and written back to cfg."""
    from plenopticam_b200.lfp_aligner import LfpMicroLenses
    cent = onp.synth_centroids(6, 7, 9.3, 'rec', jitter=0.05, seed=1)
    for user, expect in ((7, 7), (11, 11), (12, 12), (21, 10), (1, 10)):
        cfg, sta = make_cfg(cent, 'rec', user)
        obj = LfpMicroLenses(cfg=cfg, sta=sta, lfp_img=np.zeros((80, 80, 3), dtype=np.uint16))
        assert obj._size_pitch == expect == cfg.params[cfg.ptc_leng]
        assert obj._size_pitch == onp.safe_pitch(min(onp.centroid_avg_pitch(cent)), user)
        assert obj._cent_pitch == expect // 2 and obj._DIMS == (80, 80, 3)


@pytest.mark.parametrize("name", ["rotator_hex", "rotator_rec_f32"])
def test_rotation_estimate_and_centroids(name):
    from plenopticam_b200.lfp_aligner.lfp_rotator import estimate_rad, rotate_centroids
    g = golden(name)
    rad = estimate_rad(g["centroids"])
    assert abs(rad - float(g["rad"])) < 1e-12
    assert np.abs(rotate_centroids(g["centroids"], rad, g["raw"].shape) - g["rot_centroids"]).max() < 1e-9


def test_aperture_mask_matches_oracle():
    from plenopticam_b200 import ops
    for M in (1, 3, 5, 7, 9, 11, 13, 15, 21):
        assert np.array_equal(ops.aperture_mask(M), onp.aperture_mask(M))


def test_status_interrupt_and_range_validation():
    from plenopticam_b200.lfp_refocuser import LfpShiftAndSum
    from plenopticam_b200.misc import PlenopticamStatus
    sta = PlenopticamStatus()
    seen = []
    sta.bind_to_stat(seen.append)
    sta.bind_to_interrupt(lambda: seen.append("INT"))
    sta.status_msg("hello")
    sta.interrupt = True
    assert sta.interrupt and "hello" in seen and "INT" in seen
    sta.status_msg("ignored")
    assert sta.stat_var == "Canceling"
    sta.interrupt = False
    sta.error = True
    assert sta.interrupt and sta.error
    # validate_range bumps an empty range like lfp_shiftandsum.py:306-315 and rejects bad lengths
    cfg, sta = make_cfg(M=5, ran_refo=[2, 2])
    LfpShiftAndSum(vp_img_arr=None, cfg=cfg, sta=sta)
    assert cfg.params[cfg.ran_refo] == [2, 3]
    cfg.params[cfg.ran_refo] = [0, 1, 2]
    with pytest.raises(ValueError):
        LfpShiftAndSum(vp_img_arr=None, cfg=cfg, sta=sta)


def test_unsupported_paths_fail_loudly():
    from plenopticam_b200.lfp_aligner import LfpResampler
    cent = onp.synth_centroids(4, 4, 9.0, 'rec', jitter=0.0, seed=0)
    cfg, sta = make_cfg(cent, 'rec', 7)
    with pytest.raises(NotImplementedError):
        LfpResampler(lfp_img=np.zeros((50, 50, 3)), cfg=cfg, sta=sta, method='cubic')
    # smp_meth='global' is implemented (tests/test_global_resampler.py); without a CUDA device every path raises
    import torch
    if not torch.cuda.is_available():
        cfg.params[cfg.smp_meth] = 'global'
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            LfpResampler(lfp_img=np.zeros((50, 50, 3)), cfg=cfg, sta=sta).main()
