"""Global light-field rectification: host-side planning of `LfpGlobalResampler`
(lfp_aligner/lfp_global_resampler.py:32-167), the reference's default `smp_meth`.

The reference fits a projective grid to the centroids (`GridFitter`, lfp_calibrator/grid_fitter.py), builds the
transfer matrix onto an ideal grid with an integer micro-image pitch, warps the whole exposure with
`skimage.transform.warp`, and for hexagonal MLAs merges neighbouring lens rows, stretches them by 2/sqrt(3) and
resizes every row strip to an odd pitch with a bicubic spline.  Everything per-camera (grid fit, matrices, spline
operators) is prepared here in float64; the per-pixel work runs on the device (csrc/global_resample.cuh).

scikit-image is not installed in the build container, so `transform.warp` itself could not be run for comparison:
the warp follows its published algorithm (order 1, mode 'constant', cval 0, clip to the input range) and its
parity is UNPINNED; everything else on this path is pinned against the unmodified reference
(tests/test_global_resampler.py).
"""
import numpy as np
import scipy.sparse.linalg as spla
from scipy.interpolate import BSpline
from scipy.optimize import leastsq


def grid_points(dims, pat_type='rec', hex_odd=0):
    """Unit lens grid centred on the origin, rows [y, x, ly, lx] (GridFitter.grid_gen, grid_fitter.py:215-256)."""
    ny, nx = int(dims[0]), int(dims[1])
    ly, lx = np.meshgrid(np.arange(ny, dtype=np.float64), np.arange(nx, dtype=np.float64), indexing='ij')
    y, x = ly.copy(), lx.copy()
    if pat_type == 'hex':
        y *= np.sqrt(3) / 2
        x[int(hex_odd)::2, :] += .5                      # every other lens ROW is shifted by half a pitch
    y -= y.max() / 2
    x -= x.max() / 2
    return np.stack([y.ravel(), x.ravel(), ly.ravel(), lx.ravel()], axis=1)


def project(p, grid, flip_xy=False):
    """(y, x) of `grid` mapped by the 3x3 projective matrix p (8 or 9 coefficients, last one forced to 1);
    flip_xy applies the matrix to (x, y) instead (GridFitter.apply_transform, grid_fitter.py:166-194).
    One matrix product over homogeneous points, like the reference: the Levenberg-Marquardt iteration below
    stops on a flat optimum, so its result is only reproducible if the residuals round the same way."""
    p = np.asarray(p, dtype=np.float64).ravel()
    m = np.append(p, 1.0).reshape(3, 3) if p.size == 8 else p.reshape(3, 3).copy()
    m[2, 2] = 1.0
    pts = np.concatenate((grid[:, :2].T, np.ones(len(grid))[np.newaxis, :]), axis=0)
    if flip_xy:
        pts[:2, :] = pts[:2, :][::-1]
    prj = np.dot(m, pts)
    if flip_xy:
        prj[:2, :] = prj[:2, :][::-1]
    yx = np.divide(prj[:2, :], prj[2, :], out=np.zeros(prj[:2, :].shape), where=prj[2, :] != 0)
    return yx[0], yx[1]


def fit_projective(coords):
    """3x3 matrix of GridFitter(coords, flip_xy=True).main().pmat: Levenberg-Marquardt (MINPACK) fit of a
    projected RECTANGULAR unit grid to the points' (y, x), residual = Euclidean distance per point
    (grid_fitter.py:88-131,139-164; the global resampler never passes pat_type, lfp_global_resampler.py:35,72)."""
    coords = np.asarray(coords, dtype=np.float64)
    dims = (int(coords[:, 2].max() + 1), int(coords[:, 3].max() + 1))
    if not (dims[0] > 3 and dims[1] > 3):
        return np.identity(3)                            # grid_fitter.py:62-66: too few points, no fit
    grid = grid_points(dims)
    # `project(p, grid, flip_xy=True)` with everything that does not depend on p hoisted out of the ~500 residual
    # evaluations of a fit (234 794 points at Illum size); the floating-point operations and their order are the same,
    # so the iteration — and the matrix — stay bit-identical to the straightforward form (fit_projective_plain)
    pts = np.concatenate((grid[:, :2].T, np.ones(len(grid))[np.newaxis, :]), axis=0)
    pts[:2, :] = pts[:2, :][::-1]
    ty, tx = np.ascontiguousarray(coords[:, 0]), np.ascontiguousarray(coords[:, 1])
    m = np.empty((3, 3))
    quo = np.empty((2, pts.shape[1]))

    def residual(p):
        m.ravel()[:8] = p
        m[2, 2] = 1.0
        prj = np.dot(m, pts)
        quo[...] = 0.0
        np.divide(prj[:2, :], prj[2, :], out=quo, where=prj[2, :] != 0)
        dy = ty - quo[1]                                # rows come back flipped: quo[1] is y, quo[0] is x
        dx = tx - quo[0]
        return np.sqrt(dy * dy + dx * dx)

    p0 = np.array([1., 0., 0., 0., 1., 0., 0., 0.])
    p = leastsq(residual, p0)[0]
    return np.append(p, 1.0).reshape(3, 3)


def fit_projective_plain(coords):
    """fit_projective written like the reference's residual (grid_fitter.py:139-164); kept to pin the fast form."""
    coords = np.asarray(coords, dtype=np.float64)
    dims = (int(coords[:, 2].max() + 1), int(coords[:, 3].max() + 1))
    if not (dims[0] > 3 and dims[1] > 3):
        return np.identity(3)
    grid = grid_points(dims)
    target = coords[:, :2]

    def residual(p):
        y, x = project(p, grid, flip_xy=True)
        return np.sqrt(np.sum((target - np.stack([y, x], axis=1)) ** 2, axis=1))

    p = leastsq(residual, np.array([1., 0., 0., 0., 1., 0., 0., 0.]))[0]
    return np.append(p, 1.0).reshape(3, 3)


def upper_factor(pmat):
    """K of GridFitter.decompose (grid_fitter.py:310-331): R of numpy's QR, scaled to K[2,2] = 1, sign-fixed."""
    _, k = np.linalg.qr(np.asarray(pmat, dtype=np.float64).reshape(3, 3), mode='reduced')
    k = k / k[2, 2]
    if k[0, 0] < 0:
        k = np.diag([-1., -1., 1.]) @ k
    return k


_PLAN_CACHE = {}          # per-camera plans: the fits and spline operators depend on the calibration, not on the exposure
_PLAN_CACHE_MAX = 4


def _cached(key, build):
    if key not in _PLAN_CACHE:
        if len(_PLAN_CACHE) >= _PLAN_CACHE_MAX:
            _PLAN_CACHE.pop(next(iter(_PLAN_CACHE)))
        _PLAN_CACHE[key] = build()
    return _PLAN_CACHE[key]


def projective_plan(centroids, img_shape, pat_type, hex_odd):
    """(tra_mat, oshape (rows, cols), pitch (py, px)) of LfpGlobalResampler.projective_alignment (:32-79); computed
    once per calibration (two Levenberg-Marquardt fits over all centroids) and reused for later exposures."""
    cent = np.ascontiguousarray(np.asarray(centroids, dtype=np.float64))
    key = ('prj', hash(cent.tobytes()), cent.shape, tuple(int(v) for v in img_shape[:2]), str(pat_type), int(bool(hex_odd)))
    tra, oshape, pitch = _cached(key, lambda: _projective_plan(cent, img_shape, pat_type, hex_odd))
    return tra.copy(), oshape.copy(), pitch.copy()


def _projective_plan(centroids, img_shape, pat_type, hex_odd):
    cent = np.asarray(centroids, dtype=np.float64)
    pmat_cent = fit_projective(cent)
    K = upper_factor(pmat_cent)
    target = np.array([np.floor(K[1, 1]), np.floor(K[0, 0])])
    if pat_type == 'rec':
        target += np.mod(target + 1, 2)                  # odd micro-image size
    elif pat_type == 'hex':
        target += np.mod(target, 2)                      # even micro-image size
    pitch = target.astype('int')
    scales = np.array([pitch[0] / K[1, 1], pitch[1] / K[0, 0]])
    oshape = np.round(np.array(img_shape[:2]) * scales).astype('int')
    kp = np.diag([float(pitch[0]), float(pitch[1]), K[2, 2]])
    if pat_type == 'hex':
        kp[0, 0] *= 2 / np.sqrt(3)
    dims = (int(cent[:, 2].max() + 1), int(cent[:, 3].max() + 1))
    grid = grid_points(dims, pat_type, hex_odd)
    gy, gx = project(kp, grid)
    crop = pitch // 2
    sy = gy.min() - crop[0] if gy.min() - crop[0] < 0 else 0.0
    sx = gx.min() - crop[1] if gx.min() - crop[1] < 0 else 0.0
    grid[:, 0], grid[:, 1] = gy - sy, gx - sx
    pmat_grid = fit_projective(grid)
    return pmat_grid @ np.linalg.inv(pmat_cent), oshape, pitch


# ------------------------------------------------------------------------------------------------------
# spline resize operators of misc.img_resize (misc/data_proc.py:57-92) in band form
# ------------------------------------------------------------------------------------------------------
def spline_resize_band(n_in, n_out, eps=1e-8, chunk=512):
    """Interpolating not-a-knot cubic spline through n_in samples, evaluated at linspace(0, n_in-1, n_out), as a
    band: (first (n_out,) int32, vals (n_out, K) float32).  Rows are solved chunk-wise against the banded
    collocation matrix, so nothing of size n_in x n_out is held (8750 x 9375 at Illum size)."""
    if n_in < 4:
        raise NotImplementedError("cubic spline resize needs at least 4 samples")
    x = np.arange(n_in, dtype=np.float64)
    t = np.concatenate([[x[0]] * 4, x[2:-2], [x[-1]] * 4])
    colloc = BSpline.design_matrix(x, t, 3).tocsc()
    ev = BSpline.design_matrix(np.linspace(0, n_in - 1, n_out), t, 3).tocsr()
    lu = spla.splu(colloc.T.tocsc())                      # W^T = colloc^-T ev^T
    firsts, rows, K = [], [], 1
    for r0 in range(0, n_out, chunk):
        Wt = lu.solve(ev[r0:r0 + chunk].T.toarray())      # (n_in, chunk)
        W = Wt.T
        sig = np.abs(W) > eps
        lo = np.argmax(sig, axis=1)
        hi = n_in - 1 - np.argmax(sig[:, ::-1], axis=1)
        K = max(K, int((hi - lo).max()) + 1)
        firsts.append(lo)
        rows.append(W)
    first = np.concatenate(firsts).astype(np.int64)
    first = np.minimum(first, max(n_in - K, 0))
    vals = np.zeros((n_out, K), dtype=np.float32)
    r = 0
    for W in rows:
        n = W.shape[0]
        cols = first[r:r + n, None] + np.arange(K)[None, :]
        ok = cols < n_in
        vals[r:r + n] = np.where(ok, W[np.arange(n)[:, None], np.minimum(cols, n_in - 1)], 0.0)
        r += n
    return first.astype(np.int32), vals


def spline_resize_dense(n_in, n_out):
    """Same operator as a dense (n_out, n_in) float64 matrix (small axes only)."""
    x = np.arange(n_in, dtype=np.float64)
    t = np.concatenate([[x[0]] * 4, x[2:-2], [x[-1]] * 4])
    colloc = BSpline.design_matrix(x, t, 3).tocsc()
    ev = BSpline.design_matrix(np.linspace(0, n_in - 1, n_out), t, 3).toarray()
    return spla.splu(colloc.T.tocsc()).solve(ev.T).T


def hex_plan(pitch, LY, LX, prj_cols, hex_odd):
    """Geometry and operators of LfpGlobalResampler.hexagonal_alignment / hex_stretch (:84-167); cached per geometry
    (the banded spline operator of a 8750 -> 9375 resize takes seconds to build)."""
    key = ('hex', int(pitch[0]), int(pitch[1]), int(LY), int(LX), int(prj_cols), int(bool(hex_odd)))
    return _cached(key, lambda: _hex_plan(pitch, LY, LX, prj_cols, hex_odd))


def _hex_plan(pitch, LY, LX, prj_cols, hex_odd):
    py, px = int(pitch[0]), int(pitch[1])
    if py < 4:
        raise NotImplementedError("micro-image pitch below 4 pixels (cubic spline resize of a lens row)")
    lens_new_x = int(round(LX * 2 / np.sqrt(3)))
    lnew = np.array([py, px]) + np.mod(np.array([py, px]) + 1, 2)
    if (LX - 1) * px + px - 1 >= prj_cols:
        raise ValueError("rectified image is narrower than the lens grid")     # pos[:LX] would run short (:162)
    m = lens_new_x * px                                   # stretched strip width before the resize
    x_len = int(round(m * (lnew[1] / px)))
    y_len = int(round(py * (lnew[0] / py)))
    coords = np.clip(np.linspace(0, LX, lens_new_x), 0, LX - 1)     # np.interp clamps to the end samples
    i0 = np.minimum(np.floor(coords).astype(np.int64), LX - 1)
    i1 = np.minimum(i0 + 1, LX - 1)
    wx_first, wx_vals = spline_resize_band(m, x_len)
    wy = spline_resize_dense(py, y_len).astype(np.float32)
    return dict(py=py, px=px, lens_new_x=lens_new_x, lnew=lnew, m=m, x_len=x_len, y_len=y_len,
                i0=i0.astype(np.int32), i1=i1.astype(np.int32), w=(coords - i0).astype(np.float32),
                wx_first=wx_first, wx_vals=np.ascontiguousarray(wx_vals.T), wy=wy,
                hex_odd=int(bool(hex_odd)), crop_lo=int(lnew[1] // 2 + 1), crop_hi=int((-lnew[1]) // 2),
                out_rows=int((LY - 1) * (py + 1)))
