"""Sub-pixel "refocus refinement" (cfg.params['opt_refi']): host-side operator construction + device launch.

Reference chain per slice (lfp_refocuser/lfp_shiftandsum.py:93-135, misc/data_proc.py:57-92): every view is
up-sampled xM with a global interpolating bicubic spline, shifted by a(c-j), a(c-i) UP-SAMPLED pixels with edge
replication, summed, and the sum is down-sampled x1/M with the same kind of spline.  `interp2d(kind='cubic')` is
FITPACK's interpolating spline = a separable not-a-knot cubic spline, so the whole chain is linear and separable:

    R_a = (1/M) * sum_{j,i active}  A_{a(c-j)} . V_ji . B_{a(c-i)}^T
    A_s = Wdn_y . Shift_s . Wup_y   (S x S),      B_s = Wdn_x . Shift_s . Wup_x   (T x T)

with Wup (n*M x n) / Wdn (n x n*M) the spline sampling matrices and Shift_s the clamped row selection
`X[clamp(Y + s)]`.  The operators are narrow bands around the shifted diagonal (spline cardinal functions decay
by 0.268 per sample): they are built here ONCE per (n, M, set of shifts) in float64 with sparse algebra and
truncated at 1e-9, and the device evaluates two banded passes (csrc/refine.cuh).  Nothing of size (S*M x T*M) is
ever materialised (the reference's 6510 x 9375 x 3 float64 intermediates per view).
"""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla
from scipy.interpolate import BSpline

BAND_EPS = 1e-8
GROUP = 4           # outputs that share one sample window on the device


def spline_matrix(n_in, n_out):
    """Sparse (n_out x n_in) matrix W with  W @ f = not-a-knot cubic spline through f[0..n_in) sampled at
    linspace(0, n_in-1, n_out)  (== scipy.interpolate.make_interp_spline(k=3) == FITPACK s=0 for n_in >= 4)."""
    if n_in < 4:
        raise NotImplementedError("refocus refinement needs at least 4 samples per axis (cubic spline)")
    x = np.arange(n_in, dtype=np.float64)
    t = np.concatenate([[x[0]] * 4, x[2:-2], [x[-1]] * 4])                 # not-a-knot knot vector
    colloc = BSpline.design_matrix(x, t, 3).tocsc()                        # n_in x n_in, banded
    xe = np.linspace(0, n_in - 1, n_out)
    ev = BSpline.design_matrix(xe, t, 3).tocsr()                           # n_out x n_in
    # W = ev @ colloc^-1  ->  W^T = colloc^-T @ ev^T
    lu = spla.splu(colloc.T.tocsc())
    Wt = lu.solve(ev.T.toarray())
    W = Wt.T
    W[np.abs(W) < 1e-15] = 0.0
    return sp.csr_matrix(W)


class AxisOperators(object):
    """All A_s (or B_s) of one axis, as a band: values (n_shift, n, K) float32 + first column (n_shift, n) int32."""

    def __init__(self, n, M, shifts):
        self.n, self.M = int(n), int(M)
        self.shifts = sorted(set(int(s) for s in shifts))
        self.index = {s: k for k, s in enumerate(self.shifts)}
        nf = int(round(n * M))
        up = spline_matrix(n, nf)                       # img_resize(view, M):   n -> n*M
        nd = int(round(nf * (1.0 / M)))
        if nd != n:
            raise ValueError("down-sampled size %d != %d" % (nd, n))
        dn = spline_matrix(nf, nd)                      # img_resize(sum, 1/M):  n*M -> n
        rows = np.arange(nf)
        dense, lo_all, hi_all = [], [], []
        for s in self.shifts:
            A = (dn @ up[np.clip(rows + s, 0, nf - 1)]).toarray()          # (n, n) float64
            sig = np.abs(A) > BAND_EPS
            lo = np.argmax(sig, axis=1)
            hi = n - 1 - np.argmax(sig[:, ::-1], axis=1)
            dense.append(A)
            lo_all.append(lo)
            hi_all.append(hi)
        K = int(max((hi - lo).max() for lo, hi in zip(lo_all, hi_all))) + 1
        self.K = (K + 3) & ~3
        ns = len(self.shifts)
        self.first = np.zeros((ns, n), dtype=np.int32)
        self.vals = np.zeros((ns, n, self.K), dtype=np.float32)
        for k, (A, lo) in enumerate(zip(dense, lo_all)):
            lo = np.minimum(lo, max(n - self.K, 0)).astype(np.int64)       # keep the window inside [0, n)
            self.first[k] = lo
            cols = lo[:, None] + np.arange(self.K)[None, :]
            ok = cols < n
            self.vals[k] = np.where(ok, A[np.arange(n)[:, None], np.minimum(cols, n - 1)], 0.0)

    def grouped(self):
        """Device layout: outputs in groups of GROUP share one window start (the minimum of their band starts),
        weights (n_shift, G, Kg, GROUP) float32 with zeros where an output's band does not reach; first (n_shift, G)."""
        ns, n, K = self.vals.shape
        G = (n + GROUP - 1) // GROUP
        idx = np.minimum(np.arange(G * GROUP).reshape(G, GROUP), n - 1)            # last group repeats the last row
        fg = self.first[:, idx]                                                   # (ns, G, GROUP)
        start = fg.min(axis=2)
        spread = int((fg - start[:, :, None]).max())
        Kg = K + spread
        vals = np.zeros((ns, G, Kg, GROUP), dtype=np.float32)
        off = fg - start[:, :, None]                                              # (ns, G, GROUP)
        valid = (np.arange(G * GROUP).reshape(G, GROUP) < n)
        for q in range(GROUP):
            for d in range(spread + 1):
                sel = (off[:, :, q] == d) & valid[None, :, q]
                if sel.any():
                    si, gi = np.nonzero(sel)
                    vals[si, gi, d:d + K, q] = self.vals[si, idx[gi, q]]
        return vals, start.astype(np.int32), Kg

    def apply(self, s, f, axis=0):
        """numpy evaluation of the banded operator (used by the CPU tests)."""
        k = self.index[int(s)]
        f = np.moveaxis(np.asarray(f, dtype=np.float64), axis, 0)
        cols = np.minimum(self.first[k][:, None] + np.arange(self.K)[None, :], self.n - 1)
        out = np.einsum('nk,nk...->n...', self.vals[k].astype(np.float64), f[cols])
        return np.moveaxis(out, 0, axis)


_PLAN_CACHE = {}


def build_operators(S, T, M, a_list):
    """(Ay, Bx) for the light-field size, angular size and refocus parameters (up-sampled pixels)."""
    key = (int(S), int(T), int(M), tuple(int(a) for a in a_list))
    if key not in _PLAN_CACHE:
        c = M // 2
        shifts = sorted(set(int(a) * d for a in a_list for d in range(-c, c + 1)))
        _PLAN_CACHE.clear()                                 # one camera at a time; the tables can be large
        _PLAN_CACHE[key] = (AxisOperators(S, M, shifts), AxisOperators(T, M, shifts))
    return _PLAN_CACHE[key]


def refocus_refi(vp, a_list, view_zeroed, mm=None):
    """Device evaluation: vp (M,M,S,T,3) CUDA tensor (uint16 / uint8 / float32) -> (N,S,T,3) float32 stack."""
    import torch
    from .. import _native as nat
    from .. import ops
    import ctypes as C

    M, M2, S, T, Cn = vp.shape
    a_list = [int(a) for a in a_list]
    Ay, Bx = build_operators(S, T, M, a_list)
    dev = vp.device
    N = len(a_list)
    c = M // 2
    # per (slice, view-row / view-column): index of the operator to use
    iy = np.array([[Ay.index[a * (c - j)] for j in range(M)] for a in a_list], dtype=np.int32)
    ix = np.array([[Bx.index[a * (c - i)] for i in range(M)] for a in a_list], dtype=np.int32)
    t = lambda arr: torch.from_numpy(np.ascontiguousarray(arr)).to(dev)
    ayv, ayf, Kgy = Ay.grouped()
    bxv, bxf, Kgx = Bx.grouped()
    ay_v, ay_f, bx_v, bx_f = t(ayv), t(ayf), t(bxv), t(bxf)
    iy_d, ix_d = t(iy), t(ix)
    z_d = t(np.asarray(view_zeroed, dtype=np.uint8).reshape(-1))
    tmp = torch.empty((M, S, T, Cn), dtype=torch.float32, device=dev)      # horizontal pass of one slice
    out = torch.empty((N, S, T, Cn), dtype=torch.float32, device=dev)
    p = ops._ptr
    nat.check(nat.lib().pcb_refocus_refi(p(vp), ops._pcb_dtype(vp), M, S, T, Cn, N,
                                         p(ay_v), p(ay_f), Kgy, p(bx_v), p(bx_f), Kgx,
                                         p(iy_d), p(ix_d), p(z_d), p(tmp), p(out),
                                         p(mm) if mm is not None else C.c_void_p(0), ops._stream()))
    return out
