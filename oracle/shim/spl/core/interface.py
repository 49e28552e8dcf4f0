"""ORACLE / TEST INFRASTRUCTURE ONLY.  Stand-in for `spl.core.interface`.

Pinned by exact definition:
* make_open_knots(p, n): clamped uniform knot vector with n basis functions, length n+p+1
  (usage sources/multilevels.py:10-22: interior knots are T[p+1 .. n-1]).  Interior knots
  are the correctly rounded quotients i/(n-p), so nested uniform meshes share knots
  bit-for-bit (sources/multilevels.py:22 compares with `==`).
* matrix_multi_stages(ts, n, p, knots): product of single-knot Boehm insertion matrices,
  shape (n+len(ts), n) (shapes forced by sources/mg_jac.py:67-70,94,102).
* compute_spans(p, n, T): 1-based span index per element (sources/utils.py:118-127,
  sources/matrix_assembler.py:129: first non-zero basis function is span-p-1).

UNPINNED (no definition in the reference tree): collocation_cardinal_splines(p, n).  We take
the literal meaning of its name: the n x n collocation matrix of the degree-p cardinal
B-spline at its integer (p odd) / half-integer (p even) abscissae, i.e. the symmetric
Toeplitz matrix t_k = phi_p((p+1)/2 + k).  `pcg_glt` receives M1, M2 as arguments
(sources/solvers.py:239), so they are inputs to the hot path either way.
"""
import numpy as np
from scipy.interpolate import BSpline


def make_open_knots(p, n):
    N = n - p
    T = np.zeros(n + p + 1)
    for i in range(1, N):
        T[p + i] = i / N
    T[n:] = 1.0
    return T


def _insert_one(t, p, T):
    """Boehm: (len(T)-p) x (len(T)-p-1) matrix mapping coefficients on T to T + {t}."""
    n = len(T) - p - 1
    k = int(np.searchsorted(T, t, side="right")) - 1  # T[k] <= t < T[k+1]
    Q = np.zeros((n + 1, n))
    for i in range(n + 1):
        if i <= k - p:
            a = 1.0
        elif i <= k:
            a = (t - T[i]) / (T[i + p] - T[i])
        else:
            a = 0.0
        if i < n:
            Q[i, i] += a
        if i >= 1:
            Q[i, i - 1] += 1.0 - a
    Tn = np.concatenate([T[: k + 1], [t], T[k + 1:]])
    return Q, Tn


def matrix_multi_stages(ts, n, p, knots):
    T = np.array(knots, dtype=float)
    M = np.eye(n)
    for t in ts:
        Q, T = _insert_one(float(t), p, T)
        M = Q @ M
    return M


def compute_spans(p, n, T):
    T = np.asarray(T)
    spans = np.zeros(n, dtype=int)
    ie = 0
    for k in range(p, n):
        if T[k] != T[k + 1]:
            spans[ie] = k + 1
            ie += 1
    return spans


def cardinal_bspline(p, x):
    """phi_p: cardinal B-spline of degree p on the knots 0..p+1."""
    b = BSpline.basis_element(np.arange(p + 2, dtype=float), extrapolate=False)
    return np.nan_to_num(b(np.asarray(x, dtype=float)))


def collocation_cardinal_splines(p, n):
    k = np.arange(-(p // 2) - 1, p // 2 + 2)
    t = cardinal_bspline(p, (p + 1) / 2.0 + k)
    C = np.zeros((n, n))
    for kk, v in zip(k, t):
        if v != 0.0:
            C += v * np.eye(n, k=int(kk))
    return C
