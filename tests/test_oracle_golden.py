"""The numpy oracle against golden vectors produced by the UNMODIFIED reference (oracle/gen_golden.py).
CPU only.  Tolerances: float64 paths 1e-12 relative to the data scale; integer/uint16 paths bit-exact."""
import numpy as np
import pytest

from conftest import golden, golden_names
from oracle import oracle_np as onp


@pytest.mark.parametrize("name", golden_names("resample_"))
def test_resample_matches_reference(name):
    g = golden(name)
    M = int(g["M_used"])
    al = onp.resample(g["raw"], g["centroids"], M, str(g["pat_type"]), flip=bool(g["flip"]))
    scale = max(1.0, float(np.abs(g["aligned_f64"]).max()))
    assert al.shape == g["aligned_f64"].shape
    assert np.abs(al - g["aligned_f64"]).max() <= 1e-12 * scale
    assert np.array_equal(onp.uint16_norm(al), g["aligned_u16"])          # bit-exact quantisation


def test_border_lenslets_are_zero_patches():
    g = golden("resample_hex_border_u16")
    assert (g["aligned_f64"] == 0).mean() > 0.2                            # the case really exercises it


@pytest.mark.parametrize("name", ["rotator_hex", "rotator_rec_f32"])
def test_rotator_matches_reference(name):
    g = golden(name)
    rad = onp.estimate_rad(g["centroids"])
    assert abs(rad - float(g["rad"])) < 1e-15
    img = onp.rotate_img(g["raw"].astype(np.float64), rad)
    assert np.array_equal(img, g["rot_img"])
    scale = max(1.0, float(np.abs(g["rot_img"]).max()))
    assert np.abs(onp.rotate_img_closed_form(g["raw"], rad) - g["rot_img"]).max() <= 1e-12 * scale
    assert np.abs(onp.rotate_centroids(g["centroids"], rad, g["raw"].shape) - g["rot_centroids"]).max() < 1e-10


def test_rearranger_and_cropper_bit_exact():
    g = golden("rearranger_u16")
    vp = onp.compose_viewpoints(g["aligned"], int(g["M"]))
    assert vp.dtype == g["vp"].dtype and np.array_equal(vp, g["vp"])
    assert np.array_equal(onp.decompose_viewpoints(g["vp"]), g["back"])
    g = golden("cropper_u16")
    assert np.array_equal(onp.crop_micro_images(g["aligned"], int(g["M_in"]), int(g["M_out"])), g["cropped"])


@pytest.mark.parametrize("name", ["refocus_int_u16", "refocus_int_f32"])
def test_refocus_int_matches_reference_exactly(name):
    g = golden(name)
    st = onp.refocus_int(g["vp"], tuple(int(v) for v in g["ran_refo"]), int(g["M"]))
    assert st.shape == g["stack"].shape
    assert np.array_equal(st, g["stack"])                                  # same float64 sum order -> 0.0


def test_refocus_int_closed_form_equals_literal_pad_add():
    g = golden("refocus_int_u16")
    M = int(g["M"])
    vp64 = onp._apply_aperture(g["vp"], M) / M
    for a in (-3, -1, 0, 2):
        assert np.array_equal(onp.refocus_int_slice(vp64, a, M), onp.refocus_int_slice_padadd(vp64, a, M))


@pytest.mark.parametrize("name", ["refocus_refi_u16", "refocus_refi_m7_neg_u16"])
def test_refocus_refi_matches_reference(name):
    g = golden(name)
    ran, M = tuple(int(v) for v in g["ran_refo"]), int(g["M"])
    st = onp.refocus_refi(g["vp"], ran, M)
    assert st.shape == g["stack"].shape
    assert np.abs(st - g["stack"]).max() <= 1e-11 * np.abs(g["stack"]).max()
    # a subset of the range evaluates the same slices
    full = list(range(M * ran[0], M * ran[1]))
    sub = [full[0], full[len(full) // 2], full[-1]]
    st_sub = onp.refocus_refi(g["vp"], ran, M, a_list=sub)
    assert np.array_equal(st_sub, st[[full.index(a) for a in sub]])


def test_refocuser_postprocess_matches_reference():
    g = golden("refocuser_post_u16")
    st = onp.refocuser_postprocess(onp.refocus_int(g["vp"], tuple(int(v) for v in g["ran_refo"]), int(g["M"])))
    assert st.dtype == np.float32
    assert np.array_equal(st, g["stack"])


def test_aperture_mask_counts():
    assert int((~onp.aperture_mask(15)).sum()) == 177                     # 177 of 225 views kept
    assert int((~onp.aperture_mask(9)).sum()) == 69


@pytest.mark.parametrize("name", ["refocus_scratch_u16", "refocus_scratch_m7_u16", "refocus_scratch_f32"])
def test_refocus_from_scratch_matches_reference(name):
    """LfpShiftAndSum.refo_from_scratch of the unmodified reference (oracle/gen_golden_scratch.py)."""
    g = golden(name)
    st = onp.refocus_from_scratch(g["aligned"], tuple(int(v) for v in g["ran_refo"]), int(g["M"]))
    assert st.shape == g["stack"].shape
    assert np.abs(st - g["stack"]).max() <= 1e-12 * np.abs(g["stack"]).max()
