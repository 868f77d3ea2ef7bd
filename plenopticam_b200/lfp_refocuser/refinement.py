"""Sub-pixel "refocus refinement" (cfg.params['opt_refi']): host-side operator construction + device launch.

Reference chain per slice (lfp_refocuser/lfp_shiftandsum.py:93-135, misc/data_proc.py:57-92): every view is
up-sampled xM with a global interpolating bicubic spline, shifted by a(c-j), a(c-i) UP-SAMPLED pixels with edge
replication, summed, and the sum is down-sampled x1/M with the same kind of spline.  `interp2d(kind='cubic')` is
FITPACK's interpolating spline = a separable not-a-knot cubic spline, so the whole chain is linear and separable:

    R_a = (1/M) * sum_{j,i active}  A_{a(c-j)} . V_ji . B_{a(c-i)}^T
    A_s = Wdn_y . Shift_s . Wup_y   (S x S),      B_s = Wdn_x . Shift_s . Wup_x   (T x T)

with Wup (n*M x n) / Wdn (n x n*M) the spline sampling matrices and Shift_s the clamped row selection
`X[clamp(Y + s)]`.  The long tails of A_s come from ONE factor only: Wup = Ev . P, where P = colloc^-1 turns
samples into B-spline coefficients (an IIR-like prefilter whose rows decay by 0.268 per sample) and Ev evaluates
four basis functions per up-sampled point.  P does not depend on the shift, so

    A_s = D_s . P,     D_s = Wdn . Shift_s . Ev        (<= 5 taps at 1e-8: a down-sampled shifted spline is again
                                                         almost exactly a spline on the coarse knots)

and the device work splits into a prefilter of every view, done ONCE per light field (both axes), and two
narrow banded passes per slice (csrc/refine.cuh).  All operators are built here in float64 with sparse algebra
and truncated at BAND_EPS; nothing of size (S*M x T*M) is ever materialised (the reference's 6510 x 9375 x 3
float64 intermediates per view and slice).
"""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla
from scipy.interpolate import BSpline

BAND_EPS = 1e-8
GROUP = 8           # output rows that share one sample window in the vertical device pass (RF_GROUP)


def spline_factors(n_in, n_out):
    """(colloc, ev): not-a-knot cubic B-spline collocation matrix on the integer grid [0, n_in) (n_in x n_in) and
    evaluation matrix at linspace(0, n_in-1, n_out) (n_out x n_in).  Interpolation = ev @ colloc^-1."""
    if n_in < 4:
        raise NotImplementedError("refocus refinement needs at least 4 samples per axis (cubic spline)")
    x = np.arange(n_in, dtype=np.float64)
    t = np.concatenate([[x[0]] * 4, x[2:-2], [x[-1]] * 4])                 # not-a-knot knot vector
    colloc = BSpline.design_matrix(x, t, 3).tocsc()                        # n_in x n_in, banded
    xe = np.linspace(0, n_in - 1, n_out)
    ev = BSpline.design_matrix(xe, t, 3).tocsr()                           # n_out x n_in, 4 per row
    return colloc, ev


def spline_matrix(n_in, n_out):
    """Sparse (n_out x n_in) matrix W with  W @ f = not-a-knot cubic spline through f[0..n_in) sampled at
    linspace(0, n_in-1, n_out)  (== scipy.interpolate.make_interp_spline(k=3) == FITPACK s=0 for n_in >= 4)."""
    colloc, ev = spline_factors(n_in, n_out)
    # W = ev @ colloc^-1  ->  W^T = colloc^-T @ ev^T
    lu = spla.splu(colloc.T.tocsc())
    Wt = lu.solve(ev.T.toarray())
    W = Wt.T
    W[np.abs(W) < 1e-15] = 0.0
    return sp.csr_matrix(W)


def _band(A, eps, kmin=1):
    """Dense (n x n) -> (first (n,), vals (n, K)) keeping |a| > eps; windows are kept inside [0, n)."""
    n = A.shape[0]
    sig = np.abs(A) > eps
    lo = np.argmax(sig, axis=1)
    hi = n - 1 - np.argmax(sig[:, ::-1], axis=1)
    K = max(int((hi - lo).max()) + 1, kmin)
    lo = np.minimum(lo, max(n - K, 0)).astype(np.int64)
    cols = lo[:, None] + np.arange(K)[None, :]
    vals = np.where(cols < n, A[np.arange(n)[:, None], np.minimum(cols, n - 1)], 0.0)
    return lo.astype(np.int32), vals


def _group_band(first, vals, group):
    """Band (first (n,), vals (n, K)) -> outputs in groups of `group` sharing one window: start (G,) int32 and
    weights (G, Kg, group) float32, zero where an output's own band does not reach."""
    n, K = vals.shape
    G = (n + group - 1) // group
    idx = np.minimum(np.arange(G * group).reshape(G, group), n - 1)
    valid = np.arange(G * group).reshape(G, group) < n
    fg = first[idx].astype(np.int64)
    start = fg.min(axis=1)
    off = fg - start[:, None]
    Kg = K + int(off.max())
    out = np.zeros((G, Kg, group), dtype=np.float32)
    gi = np.arange(G)
    for q in range(group):
        for m in range(K):
            sel = valid[:, q]
            out[gi[sel], off[sel, q] + m, q] = vals[idx[sel, q], m]
    return start.astype(np.int32), out, Kg


PREFILTER_GROUP = 8     # output rows per thread in the vertical prefilter pass


class AxisOperators(object):
    """All A_s (or B_s) of one axis in factorised band form:

        prefilter  P    : p_first (n,) int32, p_vals (n, Kp) float32          (shift independent)
        per shift  D_s  : first (n_shift, n) int32, vals (n_shift, n, K) float32,   A_s = D_s . P
    """

    def __init__(self, n, M, shifts):
        self.n, self.M = int(n), int(M)
        self.shifts = sorted(set(int(s) for s in shifts))
        self.index = {s: k for k, s in enumerate(self.shifts)}
        nf = int(round(n * M))
        colloc, ev_up = spline_factors(n, nf)           # img_resize(view, M):   n -> n*M  =  ev_up @ colloc^-1
        nd = int(round(nf * (1.0 / M)))
        if nd != n:
            raise ValueError("down-sampled size %d != %d" % (nd, n))
        dn = spline_matrix(nf, nd)                      # img_resize(sum, 1/M):  n*M -> n
        # the four cubic B-spline taps of every up-sampled sample (rows of ev_up): first coefficient + weights, padded with
        # zeros where a row has fewer (the up-sampled image itself is only needed by `upscaled_slices`)
        ev = ev_up.tocsr()
        self.up_first = np.zeros(nf, dtype=np.int32)
        self.up_w = np.zeros((nf, 4), dtype=np.float32)
        for X in range(nf):
            cols = ev.indices[ev.indptr[X]:ev.indptr[X + 1]]
            vals = ev.data[ev.indptr[X]:ev.indptr[X + 1]]
            keep = vals != 0
            cols, vals = cols[keep], vals[keep]
            f0 = int(cols.min())
            assert cols.max() - f0 < 4
            self.up_first[X] = f0
            self.up_w[X, cols - f0] = vals
        P = np.linalg.inv(colloc.toarray())
        self.p_first, pv = _band(P, BAND_EPS)
        self.p_vals = pv.astype(np.float32)
        self.Kp = self.p_vals.shape[1]
        rows = np.arange(nf)
        firsts, dense = [], []
        for s in self.shifts:
            D = (dn @ ev_up[np.clip(rows + s, 0, nf - 1)]).toarray()       # (n, n) float64, ~5 per row
            dense.append(D)
        K = 1
        for D in dense:
            sig = np.abs(D) > BAND_EPS
            lo = np.argmax(sig, axis=1)
            hi = n - 1 - np.argmax(sig[:, ::-1], axis=1)
            K = max(K, int((hi - lo).max()) + 1)
        self.K = K
        ns = len(self.shifts)
        self.first = np.zeros((ns, n), dtype=np.int32)
        self.vals = np.zeros((ns, n, K), dtype=np.float32)
        for k, D in enumerate(dense):
            f, v = _band(D, BAND_EPS, kmin=K)
            self.first[k], self.vals[k] = f, v[:, :K]

    # -- device layouts ------------------------------------------------------------------------------------
    def tap_major(self):
        """Horizontal pass (one thread per output sample, lanes along the axis): vals (n_shift, K, n), first
        (n_shift, n), and per shift the smallest / largest sample offset (relative to the output index) a
        window touches."""
        vals = np.ascontiguousarray(self.vals.transpose(0, 2, 1))
        rel = self.first.astype(np.int64) - np.arange(self.n)[None, :]
        return vals, self.first, rel.min(axis=1).astype(np.int32), (rel.max(axis=1) + self.K - 1).astype(np.int32)

    def prefilter_tap_major(self):
        return np.ascontiguousarray(self.p_vals.T), self.p_first

    def prefilter_toeplitz(self):
        """Interior of P as ONE filter: (taps (Kp,) float32, c, lo, hi) such that every row x in [lo, hi] of the band
        equals `taps` exactly (as float32) with first[x] = x - c; (None, 0, 1, 0) when there is no such run.  Away from
        the border the inverse collocation matrix is the cubic B-spline prefilter sqrt(3) (sqrt(3) - 2)^|k|."""
        n, r0 = self.n, self.n // 2
        c = int(r0 - self.p_first[r0])
        same = (self.p_first == np.arange(n) - c) & np.all(self.p_vals == self.p_vals[r0], axis=1)
        if not same[r0]:
            return None, 0, 1, 0
        lo = hi = r0
        while lo > 0 and same[lo - 1]:
            lo -= 1
        while hi < n - 1 and same[hi + 1]:
            hi += 1
        return np.ascontiguousarray(self.p_vals[r0]), c, int(lo), int(hi)

    def prefilter_grouped(self):
        """Vertical prefilter pass: groups of PREFILTER_GROUP output rows share one window:
        vals (G, Kg, PREFILTER_GROUP) float32, first (G,) int32, Kg."""
        start, vals, Kg = _group_band(self.p_first, self.p_vals, PREFILTER_GROUP)
        return vals, start, Kg

    def grouped(self):
        """Vertical pass: outputs in groups of GROUP share one window start (the minimum of their band starts),
        weights (n_shift, G, Kg, GROUP) float32 with zeros where an output's band does not reach; first (n_shift, G)."""
        ns, n, K = self.vals.shape
        G = (n + GROUP - 1) // GROUP
        idx = np.minimum(np.arange(G * GROUP).reshape(G, GROUP), n - 1)            # last group repeats the last row
        fg = self.first[:, idx]                                                   # (ns, G, GROUP)
        start = fg.min(axis=2)
        spread = int((fg - start[:, :, None]).max())
        Kg = min(K + spread, n)
        # keep every window inside [0, n): the device pass then reads rows start .. start+Kg-1 without clamping
        start = np.minimum(start, n - Kg)
        vals = np.zeros((ns, G, Kg, GROUP), dtype=np.float32)
        off = fg - start[:, :, None]                                              # (ns, G, GROUP), off + K <= Kg
        valid = (np.arange(G * GROUP).reshape(G, GROUP) < n)
        for q in range(GROUP):
            for d in range(spread + 1):
                sel = (off[:, :, q] == d) & valid[None, :, q]
                if sel.any():
                    si, gi = np.nonzero(sel)
                    vals[si, gi, d:d + K, q] = self.vals[si, idx[gi, q]][:, :Kg - d]
        return vals, start.astype(np.int32), Kg

    # -- numpy evaluation of the truncated float32 bands (CPU tests) ---------------------------------------
    def prefilter(self, f, axis=0):
        f = np.moveaxis(np.asarray(f, dtype=np.float64), axis, 0)
        cols = np.minimum(self.p_first[:, None] + np.arange(self.Kp)[None, :], self.n - 1)
        out = np.einsum('nk,nk...->n...', self.p_vals.astype(np.float64), f[cols])
        return np.moveaxis(out, 0, axis)

    def apply_coef(self, s, coef, axis=0):
        """D_s applied to B-spline coefficients."""
        k = self.index[int(s)]
        f = np.moveaxis(np.asarray(coef, dtype=np.float64), axis, 0)
        cols = np.minimum(self.first[k][:, None] + np.arange(self.K)[None, :], self.n - 1)
        out = np.einsum('nk,nk...->n...', self.vals[k].astype(np.float64), f[cols])
        return np.moveaxis(out, 0, axis)

    def apply(self, s, f, axis=0):
        """A_s = D_s . P applied to samples."""
        return self.apply_coef(s, self.prefilter(f, axis), axis)


QUAD = 4            # consecutive slices that share one sample window in the horizontal device pass
QUAD_MAXW = 8       # longest shared window the quad kernel is built for


def quad_tables(Bx, ix, M):
    """Horizontal operators of the slices `ix` (n, M: operator index per slice and view column) regrouped in quads of
    four consecutive slices (csrc/refine.cuh::refine_hq_kernel, pcb_refi_quads_t).

    For quad g, view column i and output pixel x:  fu = min over the quad of the band starts (kept <= T - W_i),
    W_i = K + the largest spread of band starts seen for that view column, and the four bands written into the common
    window with zeros around them.  Returns dict(wq (float32, flat, float4 units of woff), fu (G, M, T) int32,
    woff (M,), W (M,), lo (M,), hi (M,), G) or None when a window would exceed QUAD_MAXW."""
    ix = np.asarray(ix)
    n, T, K = ix.shape[0], Bx.n, Bx.K
    G = (n + QUAD - 1) // QUAD
    pad = np.minimum(np.arange(G * QUAD), n - 1)
    live = (np.arange(G * QUAD) < n).reshape(G, QUAD)
    idx = ix[pad]                                               # (G*4, M)
    f = Bx.first[idx].astype(np.int64).reshape(G, QUAD, M, T)   # band start per slice, view column, pixel
    fu = f.min(axis=1)                                          # (G, M, T)
    W = (f.max(axis=1) + K - fu).max(axis=(0, 2))               # (M,)
    W = np.maximum(W, min(K, T)).astype(np.int64)
    if W.max() > QUAD_MAXW or W.max() > T:
        return None
    fu = np.minimum(fu, (T - W)[None, :, None])
    off = f - fu[:, None]                                       # (G, 4, M, T) >= 0, off + K <= W
    vals = Bx.vals[idx].reshape(G, QUAD, M, T, K).copy()        # float32
    vals[~live] = 0.0
    woff = np.zeros(M, dtype=np.int64)
    blocks = []
    pos = 0
    gi, xi = np.meshgrid(np.arange(G), np.arange(T), indexing='ij')
    for i in range(M):
        Wi = int(W[i])
        blk = np.zeros((G, Wi, T, QUAD), dtype=np.float32)      # [g][t][x] -> float4 {slice 0..3}
        for q in range(QUAD):
            for k in range(K):
                t = off[:, q, i, :] + k                         # (G, T)
                ok = t < Wi                                     # taps beyond the window are the band's zero padding
                assert np.all(vals[:, q, i, :, k][~ok] == 0.0)
                blk[gi[ok], t[ok], xi[ok], q] = vals[:, q, i, :, k][ok]
        woff[i] = pos
        pos += G * Wi * T
        blocks.append(blk.reshape(-1))
    rel = fu - np.arange(T)[None, None, :]
    return dict(wq=np.concatenate(blocks), fu=np.ascontiguousarray(fu.astype(np.int32)), woff=woff.astype(np.int32),
                W=W.astype(np.int32), lo=rel.min(axis=(0, 2)).astype(np.int32),
                hi=(rel.max(axis=(0, 2)) + W - 1).astype(np.int32), G=int(G))


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


class RefinementPlan(object):
    """Device-resident operators of one (S, T, M, a_list) configuration; reusable across light fields."""

    def __init__(self, S, T, M, a_list, view_zeroed, device, tmp_budget_bytes=3 << 30):
        import torch
        self.S, self.T, self.M = int(S), int(T), int(M)
        self.a_list = [int(a) for a in a_list]
        self.N = len(self.a_list)
        Ay, Bx = build_operators(S, T, M, self.a_list)
        c = M // 2
        t = lambda arr: torch.from_numpy(np.ascontiguousarray(arr)).to(device)
        self.z_host = np.ascontiguousarray(np.asarray(view_zeroed, dtype=np.uint8).reshape(-1))
        if self.z_host.size != M * M:
            raise ValueError("view mask must have M*M entries")
        self.z_dev = t(self.z_host)
        # prefilter
        pyv, pyf, Kpy = Ay.prefilter_grouped()              # vertical: 8 output rows per thread, uniform weights
        pxv, pxf = Bx.prefilter_tap_major()                 # horizontal: lanes along x -> (Kp, T)
        self.py_v, self.py_f, self.Kpy = t(pyv), t(pyf), int(Kpy)
        self.px_v, self.px_f, self.Kpx = t(pxv), t(pxf), int(Bx.Kp)
        self.tx, self.ty = Bx.prefilter_toeplitz(), Ay.prefilter_toeplitz()      # interior filters (host)
        # per-slice bands
        dyv, dyf, self.Kgy = Ay.grouped()
        self._dy_first_host = np.ascontiguousarray(dyf)
        dxv, dxf, xlo, xhi = Bx.tap_major()
        self.dy_v, self.dy_f = t(dyv), t(dyf)
        self.dx_v, self.dx_f, self.Kx = t(dxv), t(dxf), int(Bx.K)
        self.iy = np.array([[Ay.index[a * (c - j)] for j in range(M)] for a in self.a_list], dtype=np.int32)
        self.ix = np.array([[Bx.index[a * (c - i)] for i in range(M)] for a in self.a_list], dtype=np.int32)
        self.iy_d, self.ix_d = t(self.iy), t(self.ix)
        self._xlo, self._xhi = xlo, xhi
        self._Bx = Bx
        self._quads = {}
        per_slice = M * S * T * 3 * 4                       # horizontal pass of one slice: (M, S, T, 3) float32
        self.batch = int(max(1, min(self.N, tmp_budget_bytes // per_slice, 1024 // M)))
        self.device = device

    def rows(self, n_lo, n_hi, y_lo, y_hi):
        """pcb_refi_rows_t for the output rows [y_lo, y_hi) of slices n_lo .. n_hi-1 (a rank's band of a row-sharded stack):
        the rows of the horizontal pass the vertical windows of those outputs reach, and the rows of the prefilter's x
        pass behind them (with slack for the whole strips / groups the y pass works in)."""
        from .. import _native as nat
        S = self.S
        y_lo, y_hi = int(y_lo), min(int(y_hi), S)
        if y_lo % GROUP or (y_hi % GROUP and y_hi != S) or not 0 <= y_lo < y_hi:
            raise ValueError("row band must start / end on multiples of %d inside the image" % GROUP)
        act = [j for j in range(self.M) if not self.z_host.reshape(self.M, self.M)[j].all()]
        sidx = np.unique(self.iy[n_lo:n_hi][:, act])
        first = self._dy_first_host[sidx][:, y_lo // GROUP:(y_hi + GROUP - 1) // GROUP]
        h_lo, h_hi = int(first.min()), int(first.max()) + self.Kgy
        # coefficient rows are produced in blocks of 6 (horizontal pass) out of strips of 16 / groups of 8 (y prefilter)
        c_lo, c_hi = (h_lo // 6) * 6, min(-(-h_hi // 6) * 6, S)
        slack = 16 + self.Kpy                    # whole strips of 16 rows + the longest (grouped) prefilter window
        r = nat.RefiRows()
        r.out_lo, r.out_hi, r.h_lo, r.h_hi = y_lo, y_hi, c_lo, c_hi
        r.tmp_lo, r.tmp_hi = max(c_lo - slack, 0), min(c_hi + slack, S)
        return r

    def halo_rows(self):
        """Rows of the horizontal pass a band needs beyond each inner edge (largest reach of a vertical window)."""
        G = self._dy_first_host.shape[1]
        rel = self._dy_first_host - np.arange(G)[None, :] * GROUP            # window start relative to the group's first row
        up = int(max(-rel.min(), 0))
        down = int(max((rel + self.Kgy).max() - GROUP, 0))
        return max(up, down)

    def quads(self, n0, nb):
        """pcb_refi_quads_t of slices n0 .. n0+nb-1 (device tables, cached per slice range), or None."""
        import ctypes as C
        import torch
        from .. import _native as nat
        key = (int(n0), int(nb))
        if key not in self._quads:
            if len(self._quads) >= 4:
                self._quads.pop(next(iter(self._quads)))
            tab = quad_tables(self._Bx, self.ix[n0:n0 + nb], self.M) if self._Bx.K <= QUAD_MAXW else None
            if tab is None:
                self._quads[key] = None
            else:
                wq = torch.from_numpy(tab["wq"]).to(self.device)
                fu = torch.from_numpy(tab["fu"]).to(self.device)
                st = nat.RefiQuads()
                st.wq, st.fu, st.G = wq.data_ptr(), fu.data_ptr(), tab["G"]
                for i in range(self.M):
                    st.woff[i], st.W[i], st.lo[i], st.hi[i] = (int(tab["woff"][i]), int(tab["W"][i]), int(tab["lo"][i]),
                                                               int(tab["hi"][i]))
                self._quads[key] = (st, wq, fu)
        ent = self._quads[key]
        return None if ent is None else ent[0]

    def window(self, n0, nb):
        """Per view column: smallest / largest sample offset any slice of the batch reads, relative to x."""
        sl = self.ix[n0:n0 + nb]                             # (nb, M)
        lo = self._xlo[sl].min(axis=0).astype(np.int32)
        hi = self._xhi[sl].max(axis=0).astype(np.int32)
        return np.ascontiguousarray(lo), np.ascontiguousarray(hi)


_DEVICE_PLANS = {}


def device_plan(S, T, M, a_list, view_zeroed, device):
    """RefinementPlan of the configuration, kept for the next light field of the same camera (2 entries)."""
    z = np.ascontiguousarray(np.asarray(view_zeroed, dtype=np.uint8).reshape(-1))
    key = (int(S), int(T), int(M), tuple(int(a) for a in a_list), z.tobytes(), str(device))
    plan = _DEVICE_PLANS.get(key)
    if plan is None:
        if len(_DEVICE_PLANS) >= 2:
            _DEVICE_PLANS.pop(next(iter(_DEVICE_PLANS)))
        plan = _DEVICE_PLANS[key] = RefinementPlan(S, T, M, a_list, view_zeroed, device)
    return plan


def refocus_refi(vp, a_list, view_zeroed, mm=None, plan=None, out=None, fanout=None, coef=None, n_range=None,
                 row_range=None):
    """Device evaluation: vp (M,M,S,T,3) CUDA tensor (uint16 / uint8 / float32) -> (N,S,T,3) float32 stack.

    out    : optional (N,S,T,3) float32 destination (e.g. this rank's block of a larger stack)
    fanout : optional list of (address of the first slice, address of the min/max keys or None) - the finished slices
             are stored into every listed buffer by the kernel that produces them (slice-sharded stacks: the peers'
             stacks through peer-mapped pointers, see parallel.PeerStack); `out` / `mm` are then ignored
    coef   : optional B-spline coefficients of the light field from an earlier call (`prefilter`)
    n_range: (lo, hi) - evaluate only slices lo..hi-1 of `a_list` / `plan` (a rank's block of a sharded stack; the
             operator tables are those of the whole range, so a block equals the same rows of the unsharded stack bit
             for bit); `out` / `fanout` then address the first slice of the block
    row_range: (y_lo, y_hi) - produce only these output rows of every slice (multiples of 8; a rank's band of a ROW-sharded
             stack); the other rows of `out` / the fan-out destinations are not touched.  `coef` may then hold only the
             coefficient rows the band needs (`prefilter(vp, plan, rows=plan.rows(...))`)"""
    import ctypes as C
    import torch
    from .. import _native as nat
    from .. import ops

    M, M2, S, T, Cn = vp.shape
    if plan is None:
        plan = device_plan(S, T, M, a_list, view_zeroed, vp.device)
    dev = vp.device
    n_lo, n_hi = (0, plan.N) if n_range is None else (int(n_range[0]), int(n_range[1]))
    if not 0 <= n_lo <= n_hi <= plan.N:
        raise ValueError("n_range outside the plan")
    N = n_hi - n_lo
    L, p, st = nat.lib(), ops._ptr, ops._stream()
    if N == 0:
        return out if out is not None else torch.empty((0, S, T, Cn), dtype=torch.float32, device=dev)
    rows = plan.rows(n_lo, n_hi, row_range[0], row_range[1]) if row_range is not None else None
    if coef is None:
        coef = prefilter(vp, plan, rows=rows)
    if fanout is None:
        if out is None:
            out = torch.empty((N, S, T, Cn), dtype=torch.float32, device=dev)
        elif tuple(out.shape) != (N, S, T, Cn) or out.dtype != torch.float32 or not out.is_contiguous():
            raise ValueError("out must be a contiguous float32 (N, S, T, C) tensor")
        fanout = [(out.data_ptr(), mm.data_ptr() if mm is not None else None)]
    if not 1 <= len(fanout) <= nat.PCB_MAX_PEERS:
        raise ValueError("1..%d fan-out destinations" % nat.PCB_MAX_PEERS)
    hbuf = torch.empty((min(plan.batch, N), M, S, T, Cn), dtype=torch.float32, device=dev)
    z_ptr = plan.z_host.ctypes.data_as(C.c_void_p)
    for n0 in range(0, N, plan.batch):
        nb = min(plan.batch, N - n0)
        lo, hi = plan.window(n_lo + n0, nb)
        qd = plan.quads(n_lo + n0, nb)
        dst = nat.Fanout()
        dst.n = len(fanout)
        for d, (o_ptr, m_ptr) in enumerate(fanout):
            dst.out[d] = int(o_ptr) + n0 * S * T * Cn * 4
            dst.mm[d] = m_ptr if m_ptr else None
        nat.check(L.pcb_refocus_refi_fanout(p(coef), M, S, T, Cn, nb,
                                            p(plan.dy_v), p(plan.dy_f), plan.Kgy, p(plan.dx_v), p(plan.dx_f), plan.Kx,
                                            C.c_void_p(plan.iy_d.data_ptr() + (n_lo + n0) * M * 4),
                                            C.c_void_p(plan.ix_d.data_ptr() + (n_lo + n0) * M * 4),
                                            lo.ctypes.data_as(C.c_void_p), hi.ctypes.data_as(C.c_void_p), z_ptr,
                                            p(plan.z_dev), p(hbuf), C.byref(dst),
                                            C.byref(qd) if qd is not None else None,
                                            C.byref(rows) if rows is not None else None, st))
    return out


def upscaled_slices(vp, a_values, view_zeroed, plan, coef=None, srgb=True):
    """Generator over the summed UP-SAMPLED images of the refined slices `a_values` (shifts in up-sampled pixels, members
    of the plan's range): float32 CUDA tensors (S*M, T*M, 3) - `final_img` of lfp_shiftandsum.py:123-124 and, with
    srgb=True, the `upscale_img` the reference hands to LfpExporter.save_refo_slice (:127-131: histogram-aligned against
    itself, sRGB-encoded).  One buffer is reused: consume (or clone) a slice before asking for the next."""
    import ctypes as C
    import torch
    from .. import _native as nat
    from .. import ops
    M, M2, S, T, Cn = (int(v) for v in vp.shape)
    Ay, Bx = build_operators(S, T, M, plan.a_list)
    dev = vp.device
    if coef is None:
        coef = prefilter(vp, plan)
    t = lambda arr: torch.from_numpy(np.ascontiguousarray(arr)).to(dev)
    fx, wx, fy, wy = t(Bx.up_first), t(Bx.up_w), t(Ay.up_first), t(Ay.up_w)
    G = torch.empty((M, S, T * M, Cn), dtype=torch.float32, device=dev)
    U = torch.empty((S * M, T * M, Cn), dtype=torch.float32, device=dev)
    out = torch.empty_like(U) if srgb else None
    L, p = nat.lib(), ops._ptr
    z_ptr = plan.z_host.ctypes.data_as(C.c_void_p)
    for a in a_values:
        nat.check(L.pcb_refi_upscaled_slice(p(coef), M, S, T, Cn, int(a), p(fx), p(wx), p(fy), p(wy), z_ptr, p(G), p(U),
                                            ops._stream()))
        if not srgb:
            yield int(a), U
            continue
        lohi = ops.contrast_bounds(U)                       # auto_hist_align(final_img.copy(), ref_img=final_img, opt=True)
        yield int(a), ops.contrast_srgb(U, lohi, mm=ops.minmax_f32(U), out=out)


def prefilter(vp, plan, rows=None):
    """B-spline coefficients of every unmasked view, float32 (M,M,S,T,3): the shift-independent half of the refinement,
    done once per light field.  `rows` (pcb_refi_rows_t from `plan.rows`): only the coefficient rows a row band needs."""
    import ctypes as C
    import torch
    from .. import _native as nat
    from .. import ops
    M, M2, S, T, Cn = vp.shape
    L, p, st = nat.lib(), ops._ptr, ops._stream()
    coef = torch.empty((M, M, S, T, Cn), dtype=torch.float32, device=vp.device)
    tmp = torch.empty((M, M, S, T, Cn), dtype=torch.float32, device=vp.device)
    (hx, cx, xlo, xhi), (hy, cy, ylo, yhi) = plan.tx, plan.ty
    hptr = lambda h: h.ctypes.data_as(C.c_void_p) if h is not None else C.c_void_p(0)
    nat.check(L.pcb_refi_prefilter_toeplitz(p(vp), ops._pcb_dtype(vp), M, S, T, Cn, p(plan.py_v), p(plan.py_f), plan.Kpy,
                                            p(plan.px_v), p(plan.px_f), plan.Kpx, p(plan.z_dev), p(tmp), p(coef),
                                            hptr(hx), 0 if hx is None else int(hx.size), cx, xlo, xhi,
                                            hptr(hy), 0 if hy is None else int(hy.size), cy, ylo, yhi,
                                            C.byref(rows) if rows is not None else None, st))
    return coef
