"""Seeded synthetic inputs of BASELINE.json's configurations: centroid tables, raw exposures and
white images.  Data generators only - shared by bench.py, the tests, the golden-vector scripts and the benchmark's
reference arm so that every party sees the same bytes.  No reference or checker code is involved."""
import numpy as np

ILLUM = dict(H=5368, W=7728, C=3, LY=434, LX=541, M=15, pitch=14.25, origin=(9.5, 9.5), pat='hex')


def synth_centroids(LY, LX, pitch, pat_type='hex', hex_odd=True, jitter=0.15, rot_deg=0.0, origin=None, seed=0):
    """Centroid table rows [y, x, ly, lx] for a rec / hex MLA with Gaussian jitter and optional rotation.
    Same row layout as `GridFitter.grid_gen` (lfp_calibrator/grid_fitter.py:215-256)."""
    rng = np.random.default_rng(seed)
    py = pitch * (np.sqrt(3) / 2 if pat_type == 'hex' else 1.0)
    oy, ox = origin if origin is not None else (pitch / 2 + 2.0, pitch / 2 + 2.0)
    ly, lx = np.meshgrid(np.arange(LY), np.arange(LX), indexing='ij')
    y = oy + ly * py
    x = ox + lx * pitch
    if pat_type == 'hex':
        x = x + (pitch / 2) * np.mod(ly + (0 if hex_odd else 1), 2)
    y = y + rng.normal(0, jitter, y.shape)
    x = x + rng.normal(0, jitter, x.shape)
    if rot_deg:
        t = np.deg2rad(rot_deg)
        cy, cx = y.mean(), x.mean()
        y, x = (cy + np.cos(t) * (y - cy) + np.sin(t) * (x - cx),
                cx - np.sin(t) * (y - cy) + np.cos(t) * (x - cx))
    return np.stack([y.ravel(), x.ravel(), ly.ravel().astype(float), lx.ravel().astype(float)], axis=1)


def synth_raw(H, W, C=3, seed=0, dtype=np.uint16):
    """Low-frequency sinusoid mix + uniform noise, uint16 range (post-demosaic RGB is uint16,
    lfp_aligner/cfa_processor.py:73)."""
    rng = np.random.default_rng(seed)
    yy = np.arange(H, dtype=np.float32)[:, None, None]
    xx = np.arange(W, dtype=np.float32)[None, :, None]
    ph = np.arange(C, dtype=np.float32)[None, None, :]
    img = 0.5 + 0.2 * np.sin(yy / 37.0 + ph) * np.cos(xx / 53.0 - ph) + 0.15 * np.sin((xx + yy) / 11.0 + 2 * ph)
    img = img + 0.1 * rng.random((H, W, C), dtype=np.float32)
    img = np.clip(img, 0, 1)
    if np.issubdtype(dtype, np.integer):
        return np.round(img * 65535.0).astype(dtype)
    return img.astype(dtype)


def white_image(H, W, cent, pitch, C, seed, dtype):
    """cos^4-like vignette per lenslet + noise, a few dead (zero) pixels."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float64)
    d2 = np.full((H, W), 1e9)
    for y, x in cent[:, :2]:
        y0, y1 = int(max(0, y - pitch)), int(min(H, y + pitch + 1))
        x0, x1 = int(max(0, x - pitch)), int(min(W, x + pitch + 1))
        sub = (yy[y0:y1, x0:x1] - y) ** 2 + (xx[y0:y1, x0:x1] - x) ** 2
        d2[y0:y1, x0:x1] = np.minimum(d2[y0:y1, x0:x1], sub)
    vig = np.cos(np.clip(np.sqrt(d2) / pitch, 0, 1) * 1.2) ** 4
    img = vig[..., None] * np.array([0.9, 1.0, 0.8])[None, None, :C] if C > 1 else vig
    img = img * (0.9 + 0.1 * rng.random(img.shape))
    img = np.round(img * 60000).astype(dtype) if np.issubdtype(dtype, np.integer) else img.astype(dtype)
    flat = img.reshape(-1, C) if C > 1 else img.reshape(-1, 1)
    dead = rng.choice(flat.shape[0], size=7, replace=False)
    flat[dead] = 0
    return img


def illum_workload(n_slices=64, small=False):
    """BASELINE.json configs[1]+[2]: Illum-shaped exposure, hex MLA, M = 15, `n_slices` integer refocus slices."""
    w = dict(ILLUM)
    if small:
        w.update(H=0, W=0, LY=60, LX=75)
    w["ran_refo"] = (-(n_slices // 2), n_slices - n_slices // 2)
    return w


def illum_centroids(w, seed=0):
    cent = synth_centroids(w["LY"], w["LX"], w["pitch"], w["pat"], hex_odd=True, jitter=0.15, origin=w["origin"],
                           seed=seed)
    if not w["H"]:
        w["H"] = int(cent[:, 0].max() + w["origin"][0] + 1)
        w["W"] = int(cent[:, 1].max() + w["origin"][1] + 1)
    return cent


def config1_workload():
    """BASELINE.json configs[0]: custom-SPC-like rec grid, 186 x 279 lenslets, 11 px pitch, M = 9,
    ran_refo = [-1, 1], raw + white image."""
    return dict(LY=186, LX=279, pitch=11.0, M=9, pat='rec', origin=(8.0, 8.0), ran_refo=(-1, 1), C=3)


def config1_data(w, seed=61):
    cent = synth_centroids(w["LY"], w["LX"], w["pitch"], 'rec', jitter=0.2, origin=w["origin"], seed=seed)
    H, W = int(cent[:, 0].max() + 9), int(cent[:, 1].max() + 9)
    lfp = synth_raw(H, W, 3, seed=seed + 1, dtype=np.uint16)
    wht = white_image(H, W, cent, w["pitch"], 3, seed + 2, np.uint16)
    return cent, lfp, wht
