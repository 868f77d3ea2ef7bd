"""Import shim that lets the UNMODIFIED Python reference (/root/reference) run in this container.

TEST INFRASTRUCTURE ONLY.  Used by `oracle/gen_golden.py` and by the (container-only) tests that pin
the numpy oracle against the real reference.  Nothing here is imported by the product package.

Why a shim is needed:
  * `plenopticam/__init__.py:26` imports the tkinter GUI,
  * `misc/data_proc.py:25`, `lfp_aligner/lfp_local_resampler.py:8` import `scipy.interpolate.interp2d`,
    which still imports in SciPy 1.18 but raises NotImplementedError when called,
  * imageio / colour_demosaicing / color_matcher / color_space_converter / skimage / depthy /
    tifffile / docutils / matplotlib are not installed.
The shim pre-seeds `sys.modules` with inert stubs and replaces `interp2d` with the replacement that
SciPy's own removal notice prescribes: `RectBivariateSpline(x, y, z.T, kx=k, ky=k, s=0)` (both are
FITPACK `regrid_smth`).  The reference sources themselves are untouched; `oracle/make_ref.sh` snapshots them byte for byte into the git-ignored
`oracle/_ref/` so that the benchmark's reference arm can run them on the GPU box (where /root/reference is absent).
"""
import importlib
import os
import sys
import types

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def _find_reference():
    """$PLENOPTICAM_REFERENCE, else the read-only checkout of the build container, else the byte-identical snapshot
    `sh oracle/make_ref.sh` leaves under oracle/_ref (git-ignored; it is what reaches the GPU box)."""
    env = os.environ.get("PLENOPTICAM_REFERENCE")
    if env:
        return env
    for cand in ("/root/reference", os.path.join(_HERE, "_ref")):
        if os.path.isdir(os.path.join(cand, "plenopticam")):
            return cand
    return "/root/reference"


REFERENCE_ROOT = _find_reference()


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "plenopticam"))


class _Anything(types.ModuleType):
    """Module stub: any attribute is a callable/namespace that does nothing."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        val = _Dummy(name)
        setattr(self, name, val)
        return val


class _Dummy:
    def __init__(self, name="dummy", *a, **k):
        self._name = name

    def __call__(self, *a, **k):
        return _Dummy(self._name)

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Dummy(name)

    def __mro_entries__(self, bases):  # usable as a base class: `class X(tk.Frame)`
        return (object,)

    def __iter__(self):
        return iter(())


def _stub(name, **attrs):
    mod = _Anything(name)
    mod.__path__ = []  # behave like a package so that sub-module imports resolve
    for k, v in attrs.items():
        setattr(mod, k, v)
    sys.modules[name] = mod
    return mod


def rgb2gry_bt709(img):
    """Stand-in for color_space_converter.rgb2gry (source absent from /root/reference; the package is an
    unpinned requirement `color-space-converter>=0.1.4`).  BT.709 luma, channel axis kept (H,W,1)."""
    img = np.asarray(img)
    w = np.array([0.2126, 0.7152, 0.0722])
    return (img[..., :3] * w).sum(-1, keepdims=True)


def _interp2d_replacement(x, y, z, kind="linear", **_):
    from scipy.interpolate import RectBivariateSpline

    k = {"linear": 1, "cubic": 3, "quintic": 5}[kind]
    spl = RectBivariateSpline(np.asarray(x, dtype=float), np.asarray(y, dtype=float), np.asarray(z).T, kx=k, ky=k, s=0)
    return lambda xn, yn: spl(np.asarray(xn, dtype=float), np.asarray(yn, dtype=float)).T


_LOADED = None


def load_reference():
    """Return the imported, unmodified `plenopticam` package of the reference."""
    global _LOADED
    if _LOADED is not None:
        return _LOADED
    if not reference_available():
        raise RuntimeError("reference sources not present at %s" % REFERENCE_ROOT)

    for name in ("tkinter", "tkinter.ttk", "tkinter.filedialog", "tkinter.messagebox", "tkinter.font",
                 "imageio", "colour_demosaicing", "colour", "color_matcher", "color_matcher.top_level",
                 "skimage", "skimage.transform", "depthy", "depthy.lightfield", "depthy.misc",
                 "tifffile", "docutils", "docutils.core", "matplotlib", "matplotlib.pyplot",
                 "mpl_toolkits", "mpl_toolkits.mplot3d", "requests"):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                _stub(name)
    csc = _stub("color_space_converter")
    csc.rgb2gry = rgb2gry_bt709
    for fn in ("yuv_conv", "hsv_conv", "lab_conv", "xyz_conv", "lms_conv"):
        setattr(csc, fn, lambda img, inverse=False, **k: img)

    import scipy.interpolate as si
    si.interp2d = _interp2d_replacement

    # skimage is absent: the global resampler's only use of it (lfp_global_resampler.py:81-82) gets the oracle's
    # restatement of transform.warp / ProjectiveTransform (PARITY UNPINNED for that one call, see oracle_np.py)
    sk = sys.modules.get("skimage.transform")
    if sk is not None and "_pcb_standin" not in vars(sk) and type(sk).__name__ == "_Anything":
        from oracle import oracle_np as _onp

        class _Projective(object):
            def __init__(self, matrix=None):
                self.params = np.eye(3) if matrix is None else np.asarray(matrix, dtype=np.float64)

            @property
            def inverse(self):
                return _Projective(np.linalg.inv(self.params))

        sk.ProjectiveTransform = _Projective
        sk.warp = lambda image, inverse_map, output_shape=None, **k: _onp.warp_projective(
            image, inverse_map.params, output_shape if output_shape is not None else image.shape[:2])
        sk._pcb_standin = True
        sys.modules["skimage"].transform = sk

    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        _LOADED = importlib.import_module("plenopticam")
    return _LOADED


def make_cfg_sta(tmpdir, mic_list, pat_type, ptc_leng, ran_refo=(0, 2), opt_refi=False, smp_meth="local"):
    """Reference cfg/sta objects pointing at a temp experiment dir (the resampler pickles there)."""
    pc = load_reference()
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        cfg = pc.cfg.PlenopticamConfig()
        cfg.default_values()
    cfg.params[cfg.lfp_path] = os.path.join(tmpdir, "exposure.png")
    cfg.params[cfg.smp_meth] = smp_meth
    cfg.params[cfg.ptc_leng] = int(ptc_leng)
    cfg.params[cfg.ran_refo] = list(ran_refo)
    cfg.params[cfg.opt_refi] = bool(opt_refi)
    cfg.params[cfg.opt_prnt] = False
    cfg.params[cfg.opt_dbug] = False
    cfg.calibs[cfg.mic_list] = np.asarray(mic_list).tolist()
    cfg.calibs[cfg.pat_type] = pat_type
    cfg.calibs[cfg.ptc_mean] = [float(ptc_leng), float(ptc_leng)]
    sta = pc.misc.PlenopticamStatus()
    return cfg, sta
