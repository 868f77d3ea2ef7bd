"""CPU oracle: a numpy restatement of PlenoptiCam's data-parallel light-field hot path.

TEST INFRASTRUCTURE ONLY.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline /
`--impl reference` legs may import this module, and only as the checker or as the reported CPU
baseline.  The product package (`plenopticam_b200/`) never imports it and has no CPU fallback.

Parity pinning: the reference's own test-suite holds NO golden vectors / known-answer tests for this
path (its integration tests only assert `main()` is True).  The oracle is therefore
pinned against OUTPUTS OF THE REFERENCE ITSELF, produced in the build container by importing the
unmodified reference through `oracle/ref_shim.py` (see `oracle/gen_golden.py`; vectors are committed
under `tests/golden/`).  Two third-party functions stay "parity unpinned": `color_space_converter.rgb2gry`
(source absent; assumed BT.709 luma) which only feeds the lower contrast percentile and the white-image
conversion of the devignetter, and `skimage.transform.warp` (scikit-image not installed; restated from its
published algorithm in `warp_projective`), the one call of the global resampler.

All file:line citations are relative to /root/reference/plenopticam/.
All arithmetic is float64 unless the reference itself uses float32 (Normalizer).
"""
import numpy as np
import scipy.ndimage
from scipy.interpolate import make_interp_spline


# ----------------------------------------------------------------------------------------------------
# misc/type_checks.py:57-58  rint = int(round(val))  (Python round-half-even on float64)
# ----------------------------------------------------------------------------------------------------
def rint(val):
    return int(round(float(val)))


def rint_arr(a):
    """Vectorised `rint`: np.rint is round-half-even like Python's round()."""
    return np.rint(np.asarray(a, dtype=np.float64)).astype(np.int64)


# ----------------------------------------------------------------------------------------------------
# lfp_aligner/lfp_microlenses.py
# ----------------------------------------------------------------------------------------------------
def get_hex_direction(centroids):
    """lfp_microlenses.py:221-245 — True if the lens row below the upper-left lens is shifted right."""
    centroids = np.asarray(centroids, dtype=np.float64)
    first_mic = centroids[(centroids[:, 2] == 0) & (centroids[:, 3] == 0), :][0, :2]
    central_row_idx = int(centroids[:, 3].max() / 2)
    mean_pitch = np.mean(np.diff(centroids[centroids[:, 3] == central_row_idx, 0]))
    found = centroids[(centroids[:, 0] > first_mic[0] + mean_pitch / 2) &
                      (centroids[:, 0] < first_mic[0] + 3 * mean_pitch / 2) &
                      (centroids[:, 1] < first_mic[1]) &
                      (centroids[:, 1] > first_mic[1] - 3 * mean_pitch / 4)]
    return bool(found.size == 0)


def centroid_avg_pitch(centroids):
    """lfp_microlenses.py:159-174 — (mean_pitch_y, mean_pitch_x), ceil'ed ints."""
    centroids = np.asarray(centroids, dtype=np.float64)
    central_row_idx = int(max(centroids[:, 2]) / 2)
    central_col_idx = int(max(centroids[:, 3]) / 2)
    mean_pitch_x = int(np.ceil(np.mean(np.diff(centroids[centroids[:, 2] == central_row_idx, 1]))))
    mean_pitch_y = int(np.ceil(np.mean(np.diff(centroids[centroids[:, 3] == central_col_idx, 0]))))
    return np.array([mean_pitch_y, mean_pitch_x])


def safe_pitch(cavg_pitch, user_pitch):
    """lfp_microlenses.py:121-157 restricted to the case `limg_pitch == 0` (no previous alignment)."""
    safe = 3
    if safe <= user_pitch <= cavg_pitch + 2:
        safe = user_pitch
    elif user_pitch > cavg_pitch:
        safe = cavg_pitch
    elif user_pitch < safe < cavg_pitch:
        safe = cavg_pitch
    return int(safe)


def dense_centroid_table(centroids):
    """lfp_microlenses.py:49-53,113-119 — `get_coords_by_idx` for every (ly, lx) as one (LY,LX,2) array.

    The reference takes the FIRST table row matching (ly, lx); a missing lenslet raises IndexError."""
    c = np.asarray(centroids, dtype=np.float64)
    LY = int(c[:, 2].max() + 1)
    LX = int(c[:, 3].max() + 1)
    tab = np.full((LY, LX, 2), np.nan)
    # iterate backwards so that the first occurrence wins
    ly = c[::-1, 2].astype(np.int64)
    lx = c[::-1, 3].astype(np.int64)
    ok = (c[::-1, 2] == ly) & (c[::-1, 3] == lx) & (ly >= 0) & (lx >= 0)
    tab[ly[ok], lx[ok]] = c[::-1][ok, :2]
    if np.isnan(tab).any():
        raise IndexError("lenslet index missing from centroid table (lfp_microlenses.py:119)")
    return tab


# ----------------------------------------------------------------------------------------------------
# lfp_aligner/lfp_local_resampler.py — K1 (patch), rec tiling, K2 (hex stretch)
# ----------------------------------------------------------------------------------------------------
def patches(raw, centroids, M, flip=False):
    """lfp_local_resampler.py:44-65,91-95,122-127 — P[ly,lx,v,u,ch], method='linear'.

    Window = raw[ry-c-1 : ry+c+2, rx-c-1 : rx+c+2]; FITPACK k=1 spline evaluated at
    arange(M+2) + (m - rint(m)), cropped [1:-1, 1:-1].  For the surviving samples this is plain bilinear
    interpolation at (ry-c+v+fy, rx-c+u+fx).  A window that leaves the image (or an even M, for which the
    window is (M+3)^2) yields an all-zero patch (:50,:58-60)."""
    raw = np.asarray(raw, dtype=np.float64)
    if raw.ndim == 2:
        raw = raw[..., None]
    H, W, C = raw.shape
    tab = dense_centroid_table(centroids)
    LY, LX = tab.shape[:2]
    c = M // 2
    out = np.zeros((LY, LX, M, M, C))
    if M % 2 == 0:
        return out
    ry = rint_arr(tab[..., 0])
    rx = rint_arr(tab[..., 1])
    fy = tab[..., 0] - ry
    fx = tab[..., 1] - rx
    valid = (ry - c - 1 >= 0) & (ry + c + 2 <= H) & (rx - c - 1 >= 0) & (rx + c + 2 <= W)
    fly = np.floor(fy).astype(np.int64)   # -1 or 0
    flx = np.floor(fx).astype(np.int64)
    wy = fy - fly
    wx = fx - flx
    v = np.arange(M)
    for ly in range(LY):
        ok = valid[ly]
        if not ok.any():
            continue
        y0 = (ry[ly, ok] - c + fly[ly, ok])[:, None] + v[None, :]            # (n, M)
        x0 = (rx[ly, ok] - c + flx[ly, ok])[:, None] + v[None, :]            # (n, M)
        Y0 = y0[:, :, None]
        X0 = x0[:, None, :]
        a00 = raw[Y0, X0]
        a01 = raw[Y0, X0 + 1]
        a10 = raw[Y0 + 1, X0]
        a11 = raw[Y0 + 1, X0 + 1]
        WX = wx[ly, ok][:, None, None, None]
        WY = wy[ly, ok][:, None, None, None]
        top = a00 + (a01 - a00) * WX
        bot = a10 + (a11 - a10) * WX
        out[ly, ok] = top + (bot - top) * WY
    if flip:
        out = out[:, :, ::-1, ::-1]
    return out


def hex_columns(LX, ly, hex_odd):
    """lfp_local_resampler.py:110,134-136 — for output lens column k: source column x0 and lerp weight.

    interp_coords = linspace(0, LX, Xs) + .5*mod(ly+hex_odd, 2); np.interp clamps to [0, LX-1]."""
    Xs = int(np.round(2 * LX / np.sqrt(3)))
    coords = np.linspace(0, LX, int(np.round(LX * 2 / np.sqrt(3)))) + .5 * np.mod(ly + hex_odd, 2)
    xc = np.clip(coords, 0, LX - 1)
    x0 = np.floor(xc).astype(np.int64)        # == LX-1 only where xc == LX-1 (then w == 0: np.interp's fp[-1])
    w = xc - x0
    return Xs, x0, w


def resample_rec(raw, centroids, M, flip=False):
    """lfp_local_resampler.py:76-104 — aligned[ly*M+v, lx*M+u] = P[ly,lx,v,u]."""
    P = patches(raw, centroids, M, flip)
    LY, LX, _, _, C = P.shape
    return P.transpose(0, 2, 1, 3, 4).reshape(LY * M, LX * M, C)


def resample_hex(raw, centroids, M, flip=False, hex_odd=None):
    """lfp_local_resampler.py:106-148 — per lens row, np.interp along lx onto Xs = round(2*LX/sqrt(3))
    columns, then tile."""
    P = patches(raw, centroids, M, flip)
    LY, LX, _, _, C = P.shape
    if hex_odd is None:
        hex_odd = get_hex_direction(centroids)
    Xs = int(np.round(2 * LX / np.sqrt(3)))
    out = np.zeros((LY * M, Xs * M, C))
    for ly in range(LY):
        _, x0, w = hex_columns(LX, ly, hex_odd)
        p0 = P[ly, x0]                      # (Xs, M, M, C)
        p1 = P[ly, np.minimum(x0 + 1, LX - 1)]
        # np.interp arithmetic: slope*(x - xp[j]) + fp[j]
        st = (p1 - p0) * w[:, None, None, None] + p0
        out[ly * M:(ly + 1) * M] = st.transpose(1, 0, 2, 3).reshape(M, Xs * M, C)
    return out


def resample(raw, centroids, M, pat_type, flip=False):
    """lfp_local_resampler.py:29-42 dispatch."""
    if pat_type == 'rec':
        return resample_rec(raw, centroids, M, flip)
    if pat_type == 'hex':
        return resample_hex(raw, centroids, M, flip)
    raise ValueError(pat_type)


# ----------------------------------------------------------------------------------------------------
# misc/normalizer.py — float32 arithmetic
# ----------------------------------------------------------------------------------------------------
def norm_fun(data32, mn, mx):
    """normalizer.py:69-78.  `data32` float32; mn/mx keep their own type (NumPy promotion applies:
    a numpy float64 scalar promotes the expression to float64, a float32 scalar keeps float32)."""
    if mx != mn and mx != 0:
        norm = (data32 - mn) / (mx - mn)
    else:
        norm = data32.copy()
    norm[norm < 0] = 0
    norm[norm > 1] = 1
    return norm


def uint16_norm(aligned):
    """lfp_resampler.py:60 + normalizer.py:33-46 — global min/max of the float32 cast, round-half-even."""
    d = np.asarray(aligned, dtype='float32')
    mn, mx = d.min(), d.max()
    return np.asarray(np.round(norm_fun(d, mn, mx) * (2 ** 16 - 1)), dtype=np.uint16)


# ----------------------------------------------------------------------------------------------------
# lfp_aligner/lfp_rotator.py — K0
# ----------------------------------------------------------------------------------------------------
def estimate_rad(centroids):
    """lfp_rotator.py:68-118 — OLS slopes of middle lens row / every-other lens of middle column."""
    c = np.asarray(centroids, dtype=np.float64)
    mid_row_idx = int(c[:, 2].max() / 2)
    mid_col_idx = int(c[:, 3].max() / 2)
    mid_row = c[c[:, 2] == mid_row_idx][:, :2]
    mid_col = c[c[:, 3] == mid_col_idx][:, :2][0::2]
    A = np.vstack([np.ones(len(mid_row)), mid_row[:, 1]]).T
    slope_row = np.dot(np.linalg.pinv(A), mid_row[:, 0])[1]
    A = np.vstack([np.ones(len(mid_col)), mid_col[:, 0]]).T
    slope_col = np.dot(np.linalg.pinv(A), mid_col[:, 1])[1]
    return (np.arctan(slope_row) + np.arctan(slope_col) * -1) / 2


def rotate_img(img, rad):
    """lfp_rotator.py:129-145 — scipy.ndimage.rotate(order=3, reshape=False); skipped if the angle is 0.
    scipy is the reference's own third-party dependency and is installed, so it is called directly."""
    deg = rad * 180 / np.pi
    if deg != 0:
        return scipy.ndimage.rotate(np.asarray(img), angle=deg, reshape=False, order=3)
    return np.asarray(img)


def rotate_img_closed_form(img, rad):
    """Closed form of the call above, used to document what the CUDA kernel computes:
    coef = spline_filter(order 3, mirror); out[y,x] = sum_{p,q} B(Y-floor(Y))[p] B(X-floor(X))[q] coef[...]
    with (Y,X) = R (y,x - ctr) + ctr, zero outside [0,H-1]x[0,W-1]."""
    img = np.asarray(img, dtype=np.float64)
    squeeze = img.ndim == 2
    if squeeze:
        img = img[..., None]
    H, W, C = img.shape
    cs, sn = np.cos(rad), np.sin(rad)
    yy, xx = np.meshgrid(np.arange(H, dtype=np.float64), np.arange(W, dtype=np.float64), indexing='ij')
    cy, cx = (H - 1) / 2., (W - 1) / 2.
    Y = cs * (yy - cy) + sn * (xx - cx) + cy
    X = -sn * (yy - cy) + cs * (xx - cx) + cx
    out = np.zeros_like(img)
    for ch in range(C):
        out[..., ch] = scipy.ndimage.map_coordinates(img[..., ch], [Y, X], order=3, mode='constant', cval=0.0)
    return out[..., 0] if squeeze else out


def rotate_centroids(centroids, rad, shape):
    """lfp_rotator.py:147-172 — rotate (y,x) columns about (shape-1)/2."""
    c = np.array(centroids, dtype=np.float64)
    ctr = (np.asarray(shape[:2]) - 1) / 2.
    c[:, :2] -= ctr
    Rz = np.array([[np.cos(rad), -np.sin(rad)], [np.sin(rad), np.cos(rad)]])
    c[:, :2] = np.dot(Rz, c[:, :2].T).T
    c[:, :2] += ctr
    return c


# ----------------------------------------------------------------------------------------------------
# lfp_extractor — K4, K5, K6
# ----------------------------------------------------------------------------------------------------
def crop_micro_images(aligned, M_in, M_out):
    """lfp_cropper.py:34,46-62 — central crop of every micro image, margins k = (pitch - M_out)//2 per axis;
    `M_in` is the aligned pitch, a scalar or (My, Mx) (globally rectified hex light fields are not square)."""
    aligned = np.asarray(aligned)
    My, Mx = (int(M_in), int(M_in)) if np.ndim(M_in) == 0 else (int(M_in[0]), int(M_in[1]))
    if M_out == My and M_out == Mx:
        return aligned
    ky, kx = (My - M_out) // 2, (Mx - M_out) // 2
    LY, LX = aligned.shape[0] // My, aligned.shape[1] // Mx
    a = aligned[:LY * My, :LX * Mx].reshape(LY, My, LX, Mx, -1)
    # reference slice: [k + l*M_in : (l+1)*M_in - k] assigned into an M_out-long slot (:60-62);
    # shapes only agree when M_in - 2k == M_out, i.e. equal parity.
    a = a[:, ky:My - ky, :, kx:Mx - kx]
    return a.reshape(LY * M_out, LX * M_out, aligned.shape[2] if aligned.ndim == 3 else 1)


def compose_viewpoints(aligned, M):
    """lfp_rearranger.py:37-49,79-107 — vp[j,i] = aligned[j::M, i::M, :]; dtype preserved;
    shape (M, M, int(m/M), int(n/M), p)."""
    aligned = np.asarray(aligned)
    if aligned.ndim == 2:
        aligned = aligned[..., None]
    m, n, p = aligned.shape
    S, T = int(m / M), int(n / M)
    vp = np.zeros([M, M, S, T, p], dtype=aligned.dtype)
    for j in range(M):
        for i in range(M):
            vp[j, i] = aligned[j::M, i::M, :][:S, :T]
    return vp


def decompose_viewpoints(vp):
    """lfp_rearranger.py:51-68,109-137 — inverse scatter."""
    vp = np.asarray(vp)
    if vp.ndim == 4:
        vp = vp[..., None]
    Mj, Mi, S, T, p = vp.shape
    out = np.zeros([S * Mj, T * Mi, p], dtype=vp.dtype)
    for j in range(Mj):
        for i in range(Mi):
            out[j::Mj, i::Mi, :] = vp[j, i]
    return out


def aperture_mask(M):
    """lfp_viewpoints.py:229-250 — True where the view is ZEROED: int(np.round(sqrt(x^2+y^2))) > M//2."""
    r = M // 2
    mask = np.zeros([2 * r + 1, 2 * r + 1], dtype=bool)
    for x in range(-r, r + 1):
        for y in range(-r, r + 1):
            if int(np.round(np.sqrt(x ** 2 + y ** 2))) > r:
                mask[r + y][r + x] = True
    return mask


# ----------------------------------------------------------------------------------------------------
# lfp_refocuser/lfp_shiftandsum.py — K7 (integer), K8 (refinement)
# ----------------------------------------------------------------------------------------------------
def _apply_aperture(vp, M):
    vp = np.array(vp, dtype=np.float64)          # lfp_viewpoints.py:35 float64 cast (copy)
    mask = aperture_mask(M)
    r = M // 2
    # the reference indexes vp[coords] with coords in a (2r+1)^2 mask: identical to (M,M) for odd M
    for y in range(mask.shape[0]):
        for x in range(mask.shape[1]):
            if mask[y, x] and y < vp.shape[0] and x < vp.shape[1]:
                vp[y, x] = 0
    return vp


def refocus_int_slice_padadd(vp64, a, M):
    """Literal restatement of lfp_shiftandsum.py:94-124 for one `a` (np.pad 'edge' + add, then crop).
    `vp64` already divided by M and aperture-masked.  Kept for small cases; see refocus_int()."""
    S, T, C = vp64.shape[2:]
    overlap = abs(a) * (M - 1)
    img = np.zeros((S + overlap, T + overlap, 3))
    for j in range(M):
        for i in range(M):
            tb = (abs(a) * j, abs(a) * (M - 1 - j))
            lr = (abs(a) * i, abs(a) * (M - 1 - i))
            pad = (tb, lr, (0, 0)) if a >= 0 else (tb[::-1], lr[::-1], (0, 0))
            img = np.add(img, np.pad(vp64[j, i], pad, 'edge'))
    crop = int(overlap / 2)
    return img[crop:-crop, crop:-crop, :] if a != 0 else img


def refocus_int_slice(vp64, a, M, skip_zero_views=True):
    """Closed form of the above for odd M:
    E_a[y,x] = sum_{j,i} vp[j,i, clamp(y + a(c-j)), clamp(x + a(c-i))], float64, j-major/i-minor order."""
    S, T, C = vp64.shape[2:]
    c = M // 2
    ys = np.arange(S)
    xs = np.arange(T)
    out = np.zeros((S, T, C))
    mask = aperture_mask(M)
    for j in range(M):
        yi = np.clip(ys + a * (c - j), 0, S - 1)
        for i in range(M):
            if skip_zero_views and mask[j, i]:
                continue                      # adding an all-zero view is exact
            xi = np.clip(xs + a * (c - i), 0, T - 1)
            out += vp64[j, i][yi][:, xi]
    return out


def refocus_int(vp, ran_refo, M):
    """lfp_shiftandsum.py:48-139 without refinement: aperture mask, /M, one slice per a in range(*ran_refo).
    Returns float64 (N,S,T,C)."""
    vp64 = _apply_aperture(vp, M)
    vp64 /= M                                 # :89 (divisor is M, not M*M)
    if M % 2 == 0:
        return np.asarray([refocus_int_slice_padadd(vp64, a, M) for a in range(*ran_refo)])
    return np.asarray([refocus_int_slice(vp64, a, M) for a in range(*ran_refo)])


def refocus_from_scratch(aligned, ran_refo, M):
    """lfp_shiftandsum.py:141-221 (`refo_from_scratch`, the alternate used when only the aligned image is given and
    opt_refi is off).  Horizontal pass :168-186: hor[y, j] = sum_u img[y, u + s M] with s = j + a (c - i) [- overlap if
    a >= 0], i = u - c, EMPTY (zero) where the index leaves the image; vertical pass :189-206 likewise; symmetric crop by
    overlap / 2 (:209-210).  After the crop, output pixel (h, j) sums img[(h + a (c - v)) M + v, (j + a (c - u)) M + u]
    over ALL views (u, v) - no circular aperture (main() only applies it when vp_img_arr is given, :59) - i.e. the shifts
    of refo_from_vp with zeros instead of edge replication.  Returns float64 (N, H, J, P)."""
    img = np.asarray(aligned)
    if img.ndim == 2:
        img = img[..., None]
    H, J = img.shape[0] // M, img.shape[1] // M
    img = img[:H * M, :J * M] / M                                   # :159 (a float32 image is divided in float32)
    vp = compose_viewpoints(img, M)                                 # vp[v, u, h, j] = img[h M + v, j M + u]
    c = int((M - 1) / 2)
    stack = []
    for a in range(*ran_refo):
        hor = np.zeros((M, H, J, img.shape[2]))                     # per view row v: sum over u (the reference's order)
        for v in range(M):
            for u in range(M):
                d = a * (c - u)
                lo, hi = max(0, -d), min(J, J - d)
                if hi > lo:
                    hor[v, :, lo:hi] += vp[v, u, :, lo + d:hi + d]
        acc = np.zeros((H, J, img.shape[2]))
        for v in range(M):
            d = a * (c - v)
            lo, hi = max(0, -d), min(H, H - d)
            if hi > lo:
                acc[lo:hi] += hor[v, lo + d:hi + d]
        stack.append(acc)
    return np.asarray(stack)


def spline_resize_matrix(n_in, n_out):
    """misc/data_proc.py:57-92 — `interp2d(kind='cubic')` == FITPACK interpolating bicubic spline (s=0)
    == separable not-a-knot cubic spline.  Returns the (n_out, n_in) matrix W such that
    resized = W @ signal for samples at linspace(0, n_in-1, n_out)."""
    x = np.arange(n_in, dtype=np.float64)
    xn = np.linspace(0, n_in - 1, n_out)
    k = 3 if n_in >= 4 else n_in - 1
    spl = make_interp_spline(x, np.eye(n_in), k=k, axis=0)
    return spl(xn)


def img_resize(img, scale):
    """misc/data_proc.py:57-92 with method=None ('cubic'), x_scale=y_scale=scale."""
    img = np.asarray(img, dtype=np.float64)
    n, m = img.shape[:2]
    if scale == 1 or scale <= 0:
        return img
    y_len, x_len = int(round(n * scale)), int(round(m * scale))
    Wy = spline_resize_matrix(n, y_len)
    Wx = spline_resize_matrix(m, x_len)
    return np.einsum('yn,nmc,xm->yxc', Wy, img, Wx, optimize=True)


def refocus_refi_upscaled(vp, ran_refo, M, a_list=None):
    """The summed UP-SAMPLED image of every refined slice BEFORE the down-sampling (`final_img` at
    lfp_shiftandsum.py:123-124): what the reference histogram-aligns, sRGB-encodes and writes as upscale_*.png (:126-132).
    float64 (N, S*M, T*M, C)."""
    return refocus_refi(vp, ran_refo, M, a_list=a_list, _upscaled=True)


def uint8_norm(img):
    """misc/normalizer.py:48-51 — global min / max of the float32 cast, round-half-even."""
    d = np.asarray(img, dtype='float32')
    mn, mx = d.min(), d.max()
    return np.asarray(np.round(norm_fun(d, mn, mx) * (2 ** 8 - 1)), dtype=np.uint8)


def refocus_refi(vp, ran_refo, M, a_list=None, _upscaled=False):
    """lfp_shiftandsum.py:77-139 with opt_refi: every view upsampled xM (:102), shifts in upsampled
    pixels a in range(M*r0, M*r1) (:93-94), summed, cropped (:123-124), downsampled x1/M (:135).
    Materialises the upsampled views; use only at small sizes.  `a_list`: evaluate only these members of the
    range (the slices are independent; saves time in the large-M parity tests)."""
    vp64 = _apply_aperture(vp, M)
    vp64 /= M
    S, T, C = vp64.shape[2:]
    mask = aperture_mask(M)
    c = M // 2
    ups = {}
    for j in range(M):
        for i in range(M):
            if not mask[j, i]:
                ups[(j, i)] = img_resize(vp64[j, i], M)
    ys = np.arange(S * M)
    xs = np.arange(T * M)
    stack = []
    full = range(M * ran_refo[0], M * ran_refo[1])
    if a_list is not None:
        assert all(a in full for a in a_list)
    for a in (full if a_list is None else a_list):
        acc = np.zeros((S * M, T * M, C))
        for j in range(M):
            yi = np.clip(ys + a * (c - j), 0, S * M - 1)
            for i in range(M):
                if mask[j, i]:
                    continue
                xi = np.clip(xs + a * (c - i), 0, T * M - 1)
                acc += ups[(j, i)][yi][:, xi]
        stack.append(acc if _upscaled else img_resize(acc, 1. / M))
    return np.asarray(stack)


# ----------------------------------------------------------------------------------------------------
# K9 + K10: lfp_refocuser/top_level.py:68-70
# ----------------------------------------------------------------------------------------------------
LUMA_BT709 = np.array([0.2126, 0.7152, 0.0722])


def rgb2gry(img):
    """color_space_converter.rgb2gry — third-party, source absent, unpinned (`>=0.1.4`): assumed BT.709."""
    return (np.asarray(img)[..., :3] * LUMA_BT709).sum(-1, keepdims=True)


def hist_percentiles(ref_img):
    """lfp_contrast.py:252-255 (opt=True): lo = P0.005 of gray(ref), hi = P99.9 of ref (numpy linear)."""
    return np.percentile(rgb2gry(ref_img), 0.005), np.percentile(ref_img, 99.9)


def auto_hist_align(stack, ref_img):
    """lfp_contrast.py:249-263 + normalizer.py:33-41,53-78 (float branch of type_norm)."""
    lo, hi = hist_percentiles(ref_img)
    d = np.asarray(stack, dtype='float32')
    mn = d.min() if not lo else lo            # normalizer.py:40-41 (falsy min/max fall back to the data's)
    mx = d.max() if not hi else hi
    return np.asarray(norm_fun(d, mn, mx) * (1.0 - 0.0) + 0.0, dtype='float32')


def srgb_conv(img):
    """misc/gamma_converter.py:32-44 forward direction."""
    new = np.zeros_like(img)
    new[img >= 0.0031308] = 1.055 * img[img >= 0.0031308] ** (5 / 12) - 0.055
    new[img < 0.0031308] = img[img < 0.0031308] * 12.92
    return new


def refocuser_postprocess(stack):
    """lfp_refocuser/top_level.py:68-70."""
    return srgb_conv(auto_hist_align(stack, stack[0]))


# ----------------------------------------------------------------------------------------------------
# synthetic data (BASELINE.json configs)
# ----------------------------------------------------------------------------------------------------
from plenopticam_b200.misc.synth import synth_centroids, synth_raw  # noqa: E402,F401  (neutral data generators)


# ------------------------------------------------------------------------------------------------------
# de-vignetting (lfp_aligner/lfp_devignetter.py:33-93,160-176; misc/data_proc.py:31-42)
# ------------------------------------------------------------------------------------------------------
def gauss_kernel_1d(length, sigma=1.0):
    """1-D factor g of misc.create_gauss_kernel: kernel2d = outer(g, g) with g normalised to sum 1
    (exp(-(x^2+y^2)/2s^2) / sum is separable).  `length` is made odd like data_proc.py:34."""
    length = (length - 1) // 2 + (length + 1) // 2
    xs = np.arange(-length // 2 + 1, length // 2 + 1)
    g = np.exp(-(xs.astype(np.float64) ** 2) / (2 * sigma ** 2))
    return g / g.sum()


def devignette(lfp_img, wht_img, pitch_mean):
    """LfpDevignetter(lfp_img=..., wht_img=...).main() with patch_mode False (the reference's only reachable
    mode, lfp_devignetter.py:43): returns (lfp_img float64, wht_img float64 normalised, noise_lev)."""
    from scipy.signal import convolve2d
    lfp = np.asarray(lfp_img).astype('float64')
    wht = np.ones(lfp.shape) if wht_img is None else np.asarray(wht_img).astype('float64')
    if wht.ndim == 3 and wht.shape[2] == 3:
        wht = rgb2gry(wht)                                           # :53  -> (H, W, 1)
    if wht.ndim != lfp.ndim:
        wht = rgb2gry(wht)                                           # :56
    wht = wht / np.percentile(wht, q=99.9)                           # :64
    g = gauss_kernel_1d(int(pitch_mean))
    bw = wht[..., 0] if wht.ndim == 3 else wht                       # :165-168
    flt = convolve2d(bw, np.outer(g, g) / np.outer(g, g).sum(), 'same')
    noise = float(np.std(bw - flt))
    out = np.divide(lfp, wht, out=np.zeros_like(lfp), where=wht != 0)    # :87
    return out, wht, noise


# ------------------------------------------------------------------------------------------------------
# Scheimpflug composite (lfp_refocuser/lfp_scheimpflug.py:68-114)
# ------------------------------------------------------------------------------------------------------
PFLU_VALS = ('vertical', 'horizontal', 'skew up', 'skew down')


def scheimpflug_from_stack(stack, direction):
    """(float composite, Normalizer(composite).uint16_norm()) for one direction, including the reference's
    `== (PFLU_VALS[2] or PFLU_VALS[3])` quirk (:91: only 'skew up' takes the diagonal map)."""
    stack = np.asarray(stack)
    N = len(stack)
    m, n = stack[0].shape[:2]
    p = stack[0].shape[2] if stack[0].ndim == 3 else 1
    a_x = np.linspace(0, N, n + 2, dtype='int')[1:-1]
    a_y = np.linspace(0, N, m + 2, dtype='int')[1:-1]
    a_map_x = np.outer(np.ones(m, dtype='int'), a_x)
    a_map_y = np.outer(a_y, np.ones(n, dtype='int'))
    a_map = a_map_y
    if direction == PFLU_VALS[1]:
        a_map = a_map_x
    elif direction == (PFLU_VALS[2] or PFLU_VALS[3]):
        a_map = np.mean(np.stack([a_map_x, a_map_y]), dtype='int', axis=0)
    yy, xx = np.mgrid[0:m, 0:n]
    img = np.zeros([m, n, p])
    img[...] = stack[a_map, yy, xx].reshape(m, n, p)
    x32 = img.astype('float32')
    mn, mx = x32.min(), x32.max()
    q = np.asarray(np.round(norm_fun(x32, mn, mx) * 65535), dtype=np.uint16)
    return img, q


# ------------------------------------------------------------------------------------------------------
# Lytro raw bit-unpacking (lfp_reader/lfp_decoder.py:273-326)
# ------------------------------------------------------------------------------------------------------
def comp_bayer(img_buf, shape, bit_pac=10):
    buf = np.frombuffer(bytes(img_buf), dtype=np.uint8).astype(np.uint16)
    if bit_pac == 10:
        b = buf[:buf.size // 5 * 5].reshape(-1, 5)
        px = (b[:, :4] << 2) + ((b[:, 4:5] >> (2 * np.arange(4, dtype=np.uint16))) & 3)
    elif bit_pac == 12:
        b = buf[:buf.size // 3 * 3].reshape(-1, 3)
        px = np.stack([(b[:, 0] << 4) + (b[:, 1] >> 4), ((b[:, 1] & 15) << 8) + b[:, 2]], axis=1)
    else:
        raise ValueError(bit_pac)
    return px.reshape(shape[1], shape[0]).astype('float')


# ------------------------------------------------------------------------------------------------------
# global resampling (lfp_aligner/lfp_global_resampler.py:32-167, lfp_calibrator/grid_fitter.py)
# ------------------------------------------------------------------------------------------------------
def warp_projective(image, inv_matrix, output_shape, cval=0.0):
    """Restatement of skimage.transform.warp(image, ProjectiveTransform(m).inverse, output_shape=...) with its
    defaults (order=1, mode='constant', cval=0, clip=True, preserve_range=False) for float64 images.

    scikit-image is an unpinned requirement (`scikit-image>=0.18.0`) that is NOT installed here: PARITY UNPINNED.
    Published algorithm (skimage/transform/_warps.py, _warps_cy.pyx `_warp_fast`): for output pixel (r, c) the
    input position is (x, y, z) = inv_matrix @ (c, r, 1), (x/z, y/z); bilinear interpolation between
    floor/ceil neighbours, neighbours outside the image read `cval`; the result is clipped to the input's value
    range except pixels that are exactly `cval` when cval lies outside that range (`_clip_warp_output`)."""
    img = np.asarray(image, dtype=np.float64)
    squeeze = img.ndim == 2
    if squeeze:
        img = img[..., None]
    H, W, C = img.shape
    rows, cols = int(output_shape[0]), int(output_shape[1])
    rr, cc = np.mgrid[0:rows, 0:cols].astype(np.float64)
    m = np.asarray(inv_matrix, dtype=np.float64)
    z = m[2, 0] * cc + m[2, 1] * rr + m[2, 2]
    x = (m[0, 0] * cc + m[0, 1] * rr + m[0, 2]) / z
    y = (m[1, 0] * cc + m[1, 1] * rr + m[1, 2]) / z
    r0, c0 = np.floor(y), np.floor(x)
    r1, c1 = np.ceil(y), np.ceil(x)
    dr, dc = y - r0, x - c0

    def get(ri, ci):
        ok = (ri >= 0) & (ri < H) & (ci >= 0) & (ci < W)
        v = img[np.clip(ri, 0, H - 1).astype(np.int64), np.clip(ci, 0, W - 1).astype(np.int64)]
        return np.where(ok[..., None], v, cval)

    top = (1 - dc)[..., None] * get(r0, c0) + dc[..., None] * get(r0, c1)
    bot = (1 - dc)[..., None] * get(r1, c0) + dc[..., None] * get(r1, c1)
    out = (1 - dr)[..., None] * top + dr[..., None] * bot
    mn, mx = img.min(), img.max()
    preserve = not (mn <= cval <= mx)
    mask = out == cval
    out = np.clip(out, mn, mx)
    if preserve:
        out[mask] = cval
    return out[..., 0] if squeeze else out


def grid_gen(dims, pat_type='rec', hex_odd=0):
    """GridFitter.grid_gen (grid_fitter.py:215-256) without normalisation: rows [y, x, ly, lx], centred on 0."""
    y_range = np.linspace(0, dims[0] - 1, dims[0])
    x_range = np.linspace(0, dims[1] - 1, dims[1])
    y_grid, x_grid = np.meshgrid(y_range, x_range)
    y_idcs, x_idcs = np.meshgrid(y_range.copy(), x_range.copy())
    if pat_type == 'hex':
        y_grid = y_grid * (np.sqrt(3) / 2)
        x_grid[:, int(hex_odd)::2] += .5
    y_grid = y_grid - y_grid.max() / 2
    x_grid = x_grid - x_grid.max() / 2
    return np.vstack(np.dstack([y_grid.T, x_grid.T, y_idcs.T, x_idcs.T]))


def apply_projective(p, grid, flip_xy=False):
    """GridFitter.apply_transform (grid_fitter.py:166-194): returns the grid with projected (y, x) columns."""
    p = np.array([*p, 1.0]) if len(np.ravel(p)) == 8 else np.array(p, dtype=np.float64)
    pmat = p.reshape(3, 3).copy()
    pmat[-1, :] = np.array([*pmat[-1, :2], 1])
    grid = np.array(grid, dtype=np.float64)
    pts = np.concatenate((grid[:, :2].T, np.ones(len(grid))[np.newaxis, :]), axis=0)
    if flip_xy:
        pts[:2, :] = pts[:2, :][::-1]
    prj = np.dot(pmat, pts)
    if flip_xy:
        prj[:2, :] = prj[:2, :][::-1]
    grid[:, :2] = np.divide(prj[:2, :], prj[2, :], out=np.zeros(prj[:2, :].shape), where=prj[2, :] != 0).T[:, :2]
    return grid


def grid_fit_pmat(coords, pat_type='rec', hex_odd=0, flip_xy=True):
    """GridFitter(coords, flip_xy=True).main(); .pmat (grid_fitter.py:88-131,333-340) as used by the global resampler
    (lfp_global_resampler.py:35-37,72-75): MINPACK Levenberg-Marquardt on the per-point Euclidean distances between
    the centroids and the projected unit grid.  NOTE the reference constructs GridFitter without pat_type / hex_odd,
    so the fitted model grid is always RECTANGULAR, also for hexagonal MLAs (:35, :72)."""
    from scipy.optimize import leastsq
    coords = np.asarray(coords, dtype=np.float64)
    dims = [int(max(coords[:, 2]) + 1), int(max(coords[:, 3]) + 1)]
    if not (dims[0] > 3 and dims[1] > 3):
        return np.identity(3)

    def cost(p):
        g = apply_projective(p, grid_gen(dims, 'rec', 0), flip_xy)
        return np.sqrt(np.sum((coords[:, :2] - g[:, :2]) ** 2, axis=1)).flatten()

    p_init = np.diag([1.0, 1.0, 1.0])
    p_init[:2, -1] = np.array([0.0, 0.0])
    coeffs = leastsq(cost, p_init.flatten()[:8])[0]
    return np.array([*coeffs, 1.0]).reshape(3, 3)


def qr_decompose(pmat):
    """GridFitter.decompose (grid_fitter.py:310-331): (rmat, kmat) from numpy's QR, K scaled to K[2,2] = 1."""
    rmat, kmat = np.linalg.qr(np.asarray(pmat, dtype=np.float64).reshape(3, -1), mode='reduced')
    kmat = kmat / kmat[2, 2]
    if kmat[0, 0] < 0:
        D = np.diag([-1, -1, 1.0])
        kmat = np.dot(D, kmat)
        rmat = np.dot(rmat, D)
    return rmat, kmat


def global_projective_plan(centroids, img_shape, pat_type, hex_odd):
    """Host part of LfpGlobalResampler.projective_alignment (lfp_global_resampler.py:32-79):
    (tra_mat, oshape, limg_pitch)."""
    cent = np.asarray(centroids, dtype=np.float64)
    pmat_cent = grid_fit_pmat(cent.copy())
    _, K = qr_decompose(pmat_cent)
    target = np.array([np.floor(K[1, 1]), np.floor(K[0, 0])])
    if pat_type == 'rec':
        target += np.mod(target + 1, 2)
    elif pat_type == 'hex':
        target += np.mod(target, 2)
    pitch = target.astype('int')
    scales = np.array([pitch[0] / K[1, 1], pitch[1] / K[0, 0]])
    oshape = np.round(np.array(img_shape[:2]) * scales).astype('int')
    Kpts = np.diag([pitch[0], pitch[1], K[2, 2]]).astype(np.float64)
    if pat_type == 'hex':
        Kpts[0, 0] = Kpts[0, 0] * 2 / np.sqrt(3)
    dims = [int(max(cent[:, 2]) + 1), int(max(cent[:, 3]) + 1)]
    grid = apply_projective(Kpts.flatten(), grid_gen(dims, pat_type, hex_odd))
    shifts = np.zeros(2)
    crop = pitch // 2
    shifts[0] = min(grid[:, 0]) - crop[0] if min(grid[:, 0]) - crop[0] < 0 else 0
    shifts[1] = min(grid[:, 1]) - crop[1] if min(grid[:, 1]) - crop[1] < 0 else 0
    grid[:, :2] = grid[:, :2] - shifts
    pmat_grid = grid_fit_pmat(grid.copy())
    tra_mat = np.dot(pmat_grid, np.linalg.inv(pmat_cent))
    return tra_mat, oshape, pitch


def global_hex_alignment(prj, pitch, LY, LX, hex_odd):
    """LfpGlobalResampler.hexagonal_alignment + hex_stretch (lfp_global_resampler.py:84-167) for pat_type 'hex':
    returns (aligned float64, new odd pitch)."""
    prj = np.asarray(prj, dtype=np.float64)
    py, px = int(pitch[0]), int(pitch[1])
    C = prj.shape[2]
    lens_new_x = int(round(LX * 2 / np.sqrt(3)))
    pad_zro = np.zeros([py, px, C])
    pad_hlf = np.zeros([py, px // 2, C])
    lnew = np.array([py, px]) + np.mod(np.array([py, px]) + 1, 2)
    out = np.zeros([int((LY - 1) * (py + 1)), int(lens_new_x * (px + 1)), C])
    sw = bool(hex_odd)
    cut = (-px) // 2                                         # `:-self._limg_pitch[1]//2` (:113, :122)
    coords = np.linspace(0, LX, lens_new_x)
    Wy = spline_resize_matrix(py, int(round(py * (lnew[0] / py))))
    Wx = None
    for ly in range(LY - 1):
        if sw:
            row_a = np.hstack([pad_hlf, prj[ly * py:(ly + 1) * py, :cut, :]])
            row_b = prj[(ly + 1) * py:(ly + 2) * py, ...]
            row_c = np.hstack([row_a[:, px:, :], pad_zro])
        else:
            row_a = prj[ly * py:(ly + 1) * py, ...]
            row_b = np.hstack([pad_hlf, prj[(ly + 1) * py:(ly + 2) * py, :cut, :]])
            row_c = np.hstack([row_b[:, px:, :], pad_zro])
        row = np.mean(np.concatenate([row_a[..., None], row_b[..., None], row_c[..., None]], axis=-1), axis=-1)
        stack = np.zeros([py, lens_new_x * px, C])
        for y in range(py):
            for x in range(px):
                pos = row[y, x::px, :]
                for c in range(C):
                    stack[y, x::px, c] = np.interp(coords, range(LX), pos[:LX, c])
        if Wx is None:
            Wx = spline_resize_matrix(stack.shape[1], int(round(stack.shape[1] * (lnew[1] / px))))
        out[ly * (py + 1):(ly + 1) * (py + 1), ...] = np.einsum('yn,nmc,xm->yxc', Wy, stack, Wx, optimize=True)
        sw = not sw
    return out[:, lnew[1] // 2 + 1:(-lnew[1]) // 2, :], lnew


def resample_global(raw, centroids, pat_type, hex_odd=None):
    """LfpResampler with smp_meth='global' before the uint16 quantisation: (aligned float64, micro-image pitch)."""
    cent = np.asarray(centroids, dtype=np.float64)
    if hex_odd is None:
        hex_odd = get_hex_direction(cent)
    LY, LX = int(cent[:, 2].max() + 1), int(cent[:, 3].max() + 1)
    tra, oshape, pitch = global_projective_plan(cent, np.asarray(raw).shape, pat_type, hex_odd)
    prj = warp_projective(np.asarray(raw).astype('float64'), np.linalg.inv(tra), oshape)
    if pat_type == 'rec':
        return prj, pitch
    return global_hex_alignment(prj, pitch, LY, LX, hex_odd)


# ------------------------------------------------------------------------------------------------------
# Bayer white balance with highlight protection (lfp_aligner/cfa_processor.py:142-169,171-203,233-255,257-279)
# ------------------------------------------------------------------------------------------------------
def cfa_gains(awb, bay_pat):
    """CfaProcessor.set_gains for a 2-D Bayer image (cfa_processor.py:257-279): per Bayer-quad position gains."""
    if bay_pat == "GRBG":
        return np.array([awb[2], awb[1], awb[0], awb[3]], dtype=np.float64)
    if bay_pat == "BGGR":
        return np.array([awb[0], awb[2], awb[3], awb[1]], dtype=np.float64)
    raise ValueError("CfaProcessor.set_gains: Bayer pattern %r is not handled by the reference" % (bay_pat,))


def cfa_safe_bayer_awb(bay_img, wht_img, awb, bay_pat, exp_bias=None):
    """CfaProcessor.safe_bayer_awb (cfa_processor.py:233-250) on a 2-D Bayer image: float64 (H, W).
    The dtype flow of the reference is kept (float32 inputs, float64 gains and results)."""
    bay = np.asarray(bay_img).astype('float32')
    q = np.dstack((bay[0::2, 0::2], bay[0::2, 1::2], bay[1::2, 0::2], bay[1::2, 1::2]))
    gains = cfa_gains(awb, bay_pat)
    if wht_img is None:
        wq = np.ones(q.shape)
    else:
        w = np.asarray(wht_img).astype('float32')
        wq = np.dstack((w[0::2, 0::2], w[0::2, 1::2], w[1::2, 0::2], w[1::2, 1::2]))
    fact = gains.max()
    out = q.copy()
    min_sat, min_idx = q.min(2), q.argmin(2)
    min_sat_bal = min_sat * gains[min_idx]
    min_sat = min_sat * wq.mean(2)
    min_sat[min_sat > 1] = 1
    min_sat **= 2
    for i in range(4):
        ref = np.divide(.99, wq[..., i], out=np.ones_like(wq[..., i]) * float('Inf'), where=wq[..., i] != 0)
        ref[~np.isfinite(ref)] = 0
        sel = q[..., i] > ref
        inm = (min_sat_bal * (1 - min_sat) + q[..., i] * fact * min_sat) / gains[i]
        out[..., i][sel] = np.maximum(q[..., i][sel].astype('float64'), inm[sel])
    res = np.zeros(np.array(q.shape[:2]) * 2)
    res[0::2, 0::2], res[0::2, 1::2], res[1::2, 0::2], res[1::2, 1::2] = (out[..., k] for k in range(4))
    res[0::2, 0::2] *= gains[0]
    res[0::2, 1::2] *= gains[1]
    res[1::2, 0::2] *= gains[2]
    res[1::2, 1::2] *= gains[3]
    res[res < 0] = 0
    if exp_bias is not None:
        max_lum = 2 ** (-exp_bias)
        b = np.exp(7)
        res = np.log((1 + b) / (1 + b * np.exp(-7 * (res / max_lum)))) / np.log(1 + b) * max_lum
    return res
