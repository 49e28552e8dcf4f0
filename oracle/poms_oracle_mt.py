"""ORACLE / TEST INFRASTRUCTURE ONLY -- multi-threaded CPU port of the MG-PCG path (numba prange).

Same algorithm, same operation order as oracle/poms_oracle.py (which is pinned against the golden
vectors); only the three array kernels -- band pass along an axis, no-pivot banded line solves,
row-compressed transfer along an axis -- are compiled and threaded over all host cores.  Used by
bench.py as the timed CPU baseline (`cpu_baseline`, `--impl reference`); checked against
poms_oracle.py by tests/test_oracle_mt.py.  Never imported by the product.
"""
import numpy as np
import numba as nb

from . import poms_oracle as po


CB = 1 << 30   # one task per (outer, row): finer blocking was measured 100x slower in numba


@nb.njit(parallel=True, cache=True, fastmath=False)
def _band_pass(band, X3, Y3):
    """Y[o, j, c] = sum_k band[j, k] X[o, j+k-p, c] on a (outer, n, inner) view."""
    no, n, ni = X3.shape
    w = band.shape[1]
    p = (w - 1) // 2
    if ni == 1:
        for o in nb.prange(no):
            for j in range(n):
                s = 0.0
                for k in range(w):
                    jj = j + k - p
                    if jj >= 0 and jj < n:
                        s += band[j, k] * X3[o, jj, 0]
                Y3[o, j, 0] = s
    else:
        nb_ = (ni + CB - 1) // CB
        for t in nb.prange(no * n * nb_):
            o = t // (n * nb_)
            r = t - o * (n * nb_)
            j = r // nb_
            c0 = (r - j * nb_) * CB
            c1 = min(ni, c0 + CB)
            for c in range(c0, c1):
                Y3[o, j, c] = 0.0
            for k in range(w):
                jj = j + k - p
                if jj >= 0 and jj < n:
                    b = band[j, k]
                    if b != 0.0:
                        for c in range(c0, c1):
                            Y3[o, j, c] += b * X3[o, jj, c]


@nb.njit(parallel=True, cache=True, fastmath=False)
def _band_solve(ab, kl, ku, X3):
    """In-place no-pivot dgbtrs (row-oriented substitution) along axis 1 of (outer, n, inner)."""
    no, n, ni = X3.shape
    kd = kl + ku
    if ni == 1:
        for o in nb.prange(no):
            for j in range(n):
                s = X3[o, j, 0]
                for m in range(kl, 0, -1):
                    if j - m >= 0:
                        s -= ab[kd + m, j - m] * X3[o, j - m, 0]
                X3[o, j, 0] = s
            for j in range(n - 1, -1, -1):
                s = X3[o, j, 0]
                for m in range(ku, 0, -1):
                    if j + m < n:
                        s -= ab[kd - m, j + m] * X3[o, j + m, 0]
                X3[o, j, 0] = s / ab[kd, j]
    else:
        SB = max(64, (ni * no + 63) // 64 // max(no, 1))   # ~64 tasks in total
        nb_ = (ni + SB - 1) // SB
        for t in nb.prange(no * nb_):
            o = t // nb_
            c0 = (t - o * nb_) * SB
            c1 = min(ni, c0 + SB)
            for j in range(n):
                for m in range(kl, 0, -1):
                    if j - m >= 0:
                        l = ab[kd + m, j - m]
                        for c in range(c0, c1):
                            X3[o, j, c] -= l * X3[o, j - m, c]
            for j in range(n - 1, -1, -1):
                for m in range(ku, 0, -1):
                    if j + m < n:
                        u = ab[kd - m, j + m]
                        for c in range(c0, c1):
                            X3[o, j, c] -= u * X3[o, j + m, c]
                d = ab[kd, j]
                for c in range(c0, c1):
                    X3[o, j, c] = X3[o, j, c] / d


@nb.njit(parallel=True, cache=True, fastmath=False)
def _rows_apply(start, coef, X3, Y3, accumulate):
    """Y[o, i, c] (+)= sum_w coef[i, w] X[o, start[i]+w, c]."""
    no, n_in, ni = X3.shape
    n_out, W = coef.shape
    nb_ = (ni + CB - 1) // CB
    for t in nb.prange(no * n_out * nb_):
        o = t // (n_out * nb_)
        r = t - o * (n_out * nb_)
        i = r // nb_
        c0 = (r - i * nb_) * CB
        c1 = min(ni, c0 + CB)
        if not accumulate:
            for c in range(c0, c1):
                Y3[o, i, c] = 0.0
        for w in range(W):
            j = start[i] + w
            cf = coef[i, w]
            if cf != 0.0 and j >= 0 and j < n_in:
                for c in range(c0, c1):
                    Y3[o, i, c] += cf * X3[o, j, c]


def _view3(X, axis):
    shp = X.shape
    no = int(np.prod(shp[:axis])) if axis > 0 else 1
    ni = int(np.prod(shp[axis + 1:])) if axis + 1 < len(shp) else 1
    return X.reshape(no, shp[axis], ni)


def apply_band(band, X, axis):
    X = np.ascontiguousarray(X)
    Y = np.empty_like(X)
    _band_pass(np.ascontiguousarray(band), _view3(X, axis), _view3(Y, axis))
    return Y


class KronSumOperatorMT(po.KronSumOperator):
    def dot(self, X):
        Y = None
        for t in self.terms:
            Z = X
            for ax in range(self.ndim - 1, -1, -1):
                Z = apply_band(t[ax], Z, ax)
            Y = Z if Y is None else Y + Z
        return Y


def kron_solve_banded(factors, Y):
    X = np.array(Y, dtype=float, order="C")
    for ax, (lub, la, ua, piv) in enumerate(factors):
        assert np.array_equal(piv, np.arange(len(piv))), "threaded port: no-pivot factors only"
        _band_solve(np.ascontiguousarray(lub), int(la), int(ua), _view3(X, ax))
    return X


def _rows(P):
    n_rows, n_cols = P.shape
    nz = P != 0.0
    lo = np.where(nz.any(axis=1), nz.argmax(axis=1), 0).astype(np.int64)
    hi = np.where(nz.any(axis=1), n_cols - 1 - nz[:, ::-1].argmax(axis=1), 0)
    W = int((hi - lo).max()) + 1
    coef = np.zeros((n_rows, W))
    for i in range(n_rows):
        e = min(n_cols, lo[i] + W)
        coef[i, :e - lo[i]] = P[i, lo[i]:e]
    return lo, coef


def transfer(rows_per_axis, X, order):
    """Apply per-axis row-compressed matrices in the given axis order."""
    for ax in order:
        st, cf = rows_per_axis[ax]
        X = np.ascontiguousarray(X)
        shp = list(X.shape)
        shp[ax] = cf.shape[0]
        Y = np.empty(shp)
        _rows_apply(st, cf, _view3(X, ax), _view3(Y, ax), False)
        X = Y
    return X


class MGHierarchyMT(po.MGHierarchy):
    """poms_oracle.MGHierarchy with the array kernels swapped for the threaded ones."""

    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        for lv in self.levels:
            lv["A"] = KronSumOperatorMT(lv["A"].terms)
            if "P1s" in lv:
                lv["Prows"] = [_rows(P) for P in lv["P1s"]]
                lv["Rrows"] = [_rows(np.ascontiguousarray(P.T)) for P in lv["P1s"]]

    def smooth(self, lv, b, x, zero_guess):
        A = lv["A"]
        theta = 0.5 * (lv["lmax"] + lv["lmin"])
        if self.smoother == "glt_poly":          # poms_oracle.MGHierarchy.smooth, threaded band passes
            for k in range(self.nu):
                z = b if (k == 0 and zero_guess) else b - A.dot(x)
                for ax in range(A.ndim):
                    c = lv["qc"][ax]
                    y = c[-1] * z
                    for cf in c[-2::-1]:
                        y = apply_band(lv["gband"][ax], y, ax) + cf * z
                    z = y
                x = x + z / theta
            return x
        delta = 0.5 * (lv["lmax"] - lv["lmin"])
        sigma = theta / delta
        rho = 1.0 / sigma
        d = np.zeros_like(b)
        for k in range(self.nu):
            r = b if (k == 0 and zero_guess) else b - A.dot(x)
            z = kron_solve_banded(lv["glt"], r)
            if k == 0:
                c1, c2 = 0.0, 1.0 / theta
            else:
                rho_n = 1.0 / (2.0 * sigma - rho)
                c1, c2 = rho_n * rho, 2.0 * rho_n / delta
                rho = rho_n
            d = c1 * d + c2 * z
            x = x + d
        return x

    def vcycle(self, l, b):
        lv = self.levels[l]
        if l == len(self.levels) - 1:
            return self.coarse_solve(b)
        nd = b.ndim
        x = self.smooth(lv, b, np.zeros_like(b), True)
        r = b - lv["A"].dot(x)
        rc = transfer(lv["Rrows"], r, range(nd))
        ec = self.vcycle(l + 1, rc)
        x = x + transfer(lv["Prows"], ec, range(nd - 1, -1, -1))
        return self.smooth(lv, b, x, False)


def warmup():
    """Trigger the numba compilation (not part of any timed region)."""
    h = MGHierarchyMT(2, [16, 16])
    h.mg_pcg(np.ones(h.levels[0]["A"].npts), tol=1e-6, maxiter=3)
    h = MGHierarchyMT(2, [8, 8, 8])
    h.mg_pcg(np.ones(h.levels[0]["A"].npts), tol=1e-6, maxiter=3)
