"""ORACLE / TEST INFRASTRUCTURE ONLY -- CPU restatement of the POMS multigrid solve path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product (poms_b200/) never does: it has no CPU fallback.

Parity status: PINNED.  tests/test_oracle_golden.py checks every function below against
tests/golden/*.npz, which oracle/gen_golden.py produced by running the UNMODIFIED reference
modules (over the serial spl stand-in in oracle/shim/) on the reference's own fixtures.
Third-party arithmetic that is absent from /root/reference (package `spl`, unpinned,
requirements.txt:4) is restated from its published definitions: clamped uniform knot
vectors, Boehm knot insertion, Gauss-Legendre B-spline mass/stiffness.  3-D has no
reference implementation except the banded Kronecker solve
(pyccel/pyccel_functions.py:174-248); 3-D operators here are the dimension-generic form
of the 2-D reference and are marked EXTENSION.

Every vector is a NumPy fp64 array of shape (n1, ..., nd), C order, axis 1 slowest
(sources/mg_jac.py:106-108).  A 1-D banded matrix is an (n, 2p+1) array `band` with
band[i, k] = A[i, i+k-p] (column k <-> diagonal offset k-p; pyccel/pyccel_functions.py:15,19),
entries that fall outside the matrix are zero (`remove_spurious_entries`).
"""
from math import sqrt

import numpy as np
from scipy.linalg.lapack import dgbtrf, dgbtrs, dgetrf, dgetrs

# =============================================================================================
# 1-D spline setup (what `spl` provides to sources/mg_jac.py:27-46,67 and matrix_assembler.py)
# =============================================================================================


def make_open_knots(p, n):
    """Clamped uniform knots, n basis functions (sources/mg_jac.py:28-29 call sites)."""
    N = n - p
    T = np.zeros(n + p + 1)
    T[p + 1:n] = np.arange(1, N) / N
    T[n:] = 1.0
    return T


def knots_to_insert(Tf, nf, pf, Tc, nc, pc):
    """sources/multilevels.py:7-33: interior fine knots Tf[pf+1..nf-1] that are not bit-equal
    to any interior coarse knot Tc[pc+1..nc]."""
    out = []
    for i in range(pf + 1, nf):
        t = Tf[i]
        j = pc + 1
        found = False
        while True:
            if t < Tc[j]:
                break
            if t == Tc[j]:
                found = True
            j += 1
            if j > nc:
                break
        # reference line 29: the while loop always ends with condition False, so the knot
        # is kept iff it was never found
        if not found:
            out.append(t)
    return np.array(out)


def fine_knots(Tc, ts):
    """sources/mg_jac.py:32-35: sorted union of the coarse knots and the inserted ones."""
    return np.sort(np.concatenate([np.asarray(Tc, float), np.asarray(ts, float)]), kind="stable")


def insertion_matrix(ts, n, p, knots):
    """`matrix_multi_stages(ts, n, p, knots)` (sources/mg_jac.py:67): product of single-knot
    Boehm insertion matrices; (n+len(ts)) x n, maps coarse to fine coefficients."""
    T = np.array(knots, dtype=float)
    M = np.eye(n)
    for t in ts:
        m = len(T) - p - 1
        k = int(np.searchsorted(T, t, side="right")) - 1
        Q = np.zeros((m + 1, m))
        for i in range(m + 1):
            if i <= k - p:
                a = 1.0
            elif i <= k:
                a = (t - T[i]) / (T[i + p] - T[i])
            else:
                a = 0.0
            if i < m:
                Q[i, i] += a
            if i >= 1:
                Q[i, i - 1] += 1.0 - a
        M = Q @ M
        T = np.concatenate([T[:k + 1], [t], T[k + 1:]])
    return M


def _basis_and_ders(T, p, span, x):
    """Cox-de Boor values and first derivatives of the p+1 non-zero B-splines on
    [T[span], T[span+1]) at points x.  Returns (p+1, len(x)) arrays."""
    x = np.asarray(x, float)
    N = np.zeros((p + 1, p + 1, len(x)))  # N[d][j]: degree d, j-th function
    N[0, 0] = 1.0
    for d in range(1, p + 1):
        for j in range(d + 1):
            i = span - d + j  # global index of the degree-d function
            v = 0.0
            if j >= 1:
                den = T[i + d] - T[i]
                if den > 0:
                    v = v + (x - T[i]) / den * N[d - 1, j - 1]
            if j <= d - 1:
                den = T[i + d + 1] - T[i + 1]
                if den > 0:
                    v = v + (T[i + d + 1] - x) / den * N[d - 1, j]
            N[d, j] = v
    vals = N[p].copy()
    ders = np.zeros_like(vals)
    for j in range(p + 1):
        i = span - p + j
        if j >= 1:
            den = T[i + p] - T[i]
            if den > 0:
                ders[j] += p / den * N[p - 1, j - 1]
        if j <= p - 1:
            den = T[i + p + 1] - T[i + 1]
            if den > 0:
                ders[j] -= p / den * N[p - 1, j]
    return vals, ders


def assemble_1d(p, T):
    """1-D mass M and stiffness K (dense n x n) on knot vector T: the integrals that
    sources/matrix_assembler.py:59-74 accumulate (v_m, v_s), Gauss-Legendre p+1 points/element."""
    T = np.asarray(T, float)
    n = len(T) - p - 1
    u, w = np.polynomial.legendre.leggauss(p + 1)
    M = np.zeros((n, n))
    K = np.zeros((n, n))
    for span in range(p, n):
        a, b = T[span], T[span + 1]
        if b <= a:
            continue
        x = 0.5 * (a + b) + 0.5 * (b - a) * u
        wq = 0.5 * (b - a) * w
        v, dv = _basis_and_ders(T, p, span, x)
        i0 = span - p
        M[i0:i0 + p + 1, i0:i0 + p + 1] += (v * wq) @ v.T
        K[i0:i0 + p + 1, i0:i0 + p + 1] += (dv * wq) @ dv.T
    return M, K


def dense_to_band(A, p):
    """(n, 2p+1) band of a dense matrix (what sources/utils.py:105-134 extracts)."""
    n = A.shape[0]
    band = np.zeros((n, 2 * p + 1))
    for k in range(-p, p + 1):
        i = np.arange(max(0, -k), min(n, n - k))
        band[i, k + p] = A[i, i + k]
    return band


def band_to_dense(band):
    n, w = band.shape
    p = (w - 1) // 2
    A = np.zeros((n, n))
    for k in range(-p, p + 1):
        i = np.arange(max(0, -k), min(n, n - k))
        A[i, i + k] = band[i, k + p]
    return A


def glt_band(p, n, degree=None):
    """Symmetric Toeplitz band t_k = phi_q((q+1)/2 + k) of the cardinal B-spline of degree
    q (default q=p), stored with half-bandwidth p.  `collocation_cardinal_splines` itself is
    UNPINNED (absent third-party code); M1, M2 are *inputs* of pcg_glt (solvers.py:239)."""
    from scipy.interpolate import BSpline
    q = p if degree is None else degree
    b = BSpline.basis_element(np.arange(q + 2, dtype=float), extrapolate=False)
    band = np.zeros((n, 2 * p + 1))
    for k in range(-p, p + 1):
        v = float(np.nan_to_num(b((q + 1) / 2.0 + k)))
        i = np.arange(max(0, -k), min(n, n - k))
        band[i, k + p] = v
    return band


# =============================================================================================
# Operators
# =============================================================================================


def apply_band(band, X, axis):
    """Y = A applied along `axis` of X (one stage of sources/kron_product.py:80-86)."""
    n, w = band.shape
    p = (w - 1) // 2
    X = np.moveaxis(X, axis, 0)
    assert X.shape[0] == n
    Y = np.zeros_like(X)
    bshape = (-1,) + (1,) * (X.ndim - 1)
    for k in range(-p, p + 1):
        lo, hi = max(0, -k), min(n, n - k)
        if hi > lo:
            Y[lo:hi] += band[lo:hi, k + p].reshape(bshape) * X[lo + k:hi + k]
    return np.moveaxis(Y, 0, axis)


def kron_dot(A, B, X):
    """kron_dot_v2(A, B, X) = A X B^T (sources/kron_product.py:56-89): stage 1 contracts the
    last axis with B, stage 2 the first axis with A."""
    return apply_band(A, apply_band(B, X, 1), 0)


class KronSumOperator:
    """A = sum_t A_1^t (x) ... (x) A_d^t with banded 1-D factors.  For the reference weak form
    -Lap(u)+u (sources/matrix_assembler.py:82-83,173) the full 2-D StencilMatrix equals
    K(x)M + M(x)(K+M); 3-D (EXTENSION): K(x)M(x)M + M(x)K(x)M + M(x)M(x)(K+M)."""

    def __init__(self, terms):
        self.terms = [tuple(np.asarray(b, float) for b in t) for t in terms]
        self.npts = tuple(b.shape[0] for b in self.terms[0])
        self.pads = tuple((b.shape[1] - 1) // 2 for b in self.terms[0])
        self.ndim = len(self.npts)
        n = int(np.prod(self.npts))
        self.shape = (n, n)

    def dot(self, X):
        Y = np.zeros_like(X)
        for t in self.terms:
            Z = X
            for ax in range(self.ndim - 1, -1, -1):
                Z = apply_band(t[ax], Z, ax)
            Y += Z
        return Y

    def diagonal(self):
        D = np.zeros(self.npts)
        for t in self.terms:
            d = np.ones(())
            for b, p in zip(t, self.pads):
                d = np.multiply.outer(d, b[:, p])
            D += d
        return D

    def to_stencil(self):
        """Full (n1,n2,2p1+1,2p2+1) stencil array of the 2-D operator (spl StencilMatrix)."""
        assert self.ndim == 2
        S = 0.0
        for a, b in self.terms:
            S = S + a[:, None, :, None] * b[None, :, None, :]
        return S

    def tocsr(self):
        from scipy.sparse import kron, csr_matrix
        out = None
        for t in self.terms:
            m = csr_matrix(band_to_dense(t[0]))
            for b in t[1:]:
                m = kron(m, csr_matrix(band_to_dense(b)), format="csr")
            out = m if out is None else out + m
        return out.tocsr()


class StencilOperator2D:
    """The reference's full 2-D StencilMatrix (a3): v[i1,i2] = sum M[i1,i2,k1,k2] u[i1+k1,i2+k2]
    (slides/content.tex:285-290); data as stored by the shim, (n1, n2, 2p1+1, 2p2+1)."""

    def __init__(self, data):
        self.data = np.asarray(data, float)
        n1, n2, w1, w2 = self.data.shape
        self.npts = (n1, n2)
        self.pads = ((w1 - 1) // 2, (w2 - 1) // 2)
        self.shape = (n1 * n2, n1 * n2)

    def dot(self, X):
        n1, n2 = self.npts
        p1, p2 = self.pads
        Xp = np.zeros((n1 + 2 * p1, n2 + 2 * p2))
        Xp[p1:p1 + n1, p2:p2 + n2] = X
        Y = np.zeros((n1, n2))
        for k1 in range(2 * p1 + 1):
            for k2 in range(2 * p2 + 1):
                Y += self.data[:, :, k1, k2] * Xp[k1:k1 + n1, k2:k2 + n2]
        return Y

    def diagonal(self):
        return self.data[:, :, self.pads[0], self.pads[1]].copy()


def poisson_operator(p, knots_per_axis):
    """-Lap(u)+u on a tensor-product spline space as a Kronecker sum (see KronSumOperator)."""
    MK = [assemble_1d(p, T) for T in knots_per_axis]
    Mb = [dense_to_band(M, p) for M, K in MK]
    Kb = [dense_to_band(K, p) for M, K in MK]
    d = len(MK)
    terms = []
    for a in range(d):
        t = []
        for b in range(d):
            if b == a:
                t.append(Kb[b] + Mb[b] if a == d - 1 else Kb[b])
            else:
                t.append(Mb[b])
        terms.append(tuple(t))
    return KronSumOperator(terms), Mb, Kb


# =============================================================================================
# Kronecker solves (sources/kron_product.py:93-239, pyccel/pyccel_functions.py:114-248)
# =============================================================================================


def to_bnd(A):
    """LAPACK general-band storage (sources/tests/test_kron_solve_bnd.py:30-42; the copy in
    sources/kron_product.py:175-187 raises NameError): A_bnd[la+ua+i-j, j] = A[i, j]."""
    n = A.shape[0]
    nz = np.nonzero(A)
    la = int(max(0, (nz[0] - nz[1]).max()))
    ua = int(max(0, (nz[1] - nz[0]).max()))
    A_bnd = np.zeros((1 + ua + 2 * la, n))
    i, j = nz
    A_bnd[la + ua + i - j, j] = A[i, j]
    return A_bnd, la, ua


def kron_solve_dense(mats, Y):
    """kron_solve_serial / kron_solve_par (sources/kron_product.py:93-170): dense dgetrf of each
    factor, then per-axis dgetrs sweeps, axis 1 first."""
    X = np.array(Y, dtype=float)
    for ax, A in enumerate(mats):
        lu, piv, info = dgetrf(A)
        Z = np.moveaxis(X, ax, 0)
        shp = Z.shape
        sol, info = dgetrs(lu, piv, np.ascontiguousarray(Z.reshape(shp[0], -1)))
        X = np.moveaxis(sol.reshape(shp), 0, ax)
    return np.ascontiguousarray(X)


def kron_solve_banded(factors, Y):
    """kron_solve_bnd_par (sources/kron_product.py:191-239) and its 3-D analogue
    (pyccel/pyccel_functions.py:174-248): `factors[a] = (lu_band, la, ua, piv)` as returned by
    dgbtrf; dgbtrs along axis 1, then 2, then 3."""
    X = np.array(Y, dtype=float)
    for ax, (lub, la, ua, piv) in enumerate(factors):
        Z = np.moveaxis(X, ax, 0)
        shp = Z.shape
        sol, info = dgbtrs(lub, la, ua, np.ascontiguousarray(Z.reshape(shp[0], -1)), piv)
        X = np.moveaxis(sol.reshape(shp), 0, ax)
    return np.ascontiguousarray(X)


def band_factor(band):
    """dgbtrf of a (n, 2p+1) band matrix -> (lu_band, la, ua, piv) with la = ua = p."""
    n, w = band.shape
    p = (w - 1) // 2
    ab = np.zeros((3 * p + 1, n))
    for k in range(-p, p + 1):
        i = np.arange(max(0, -k), min(n, n - k))
        ab[2 * p - k, i + k] = band[i, k + p]
    lub, piv, info = dgbtrf(ab, p, p)
    assert info == 0
    return lub, p, p, piv


# =============================================================================================
# Iterative solvers: verbatim operation order of sources/solvers.py
# =============================================================================================


def _dot(a, b, log):
    v = float(np.dot(a.ravel(), b.ravel()))
    if log is not None:
        log.append(v)
    return v


def jacobi(A, b, log=None):
    """sources/solvers.py:139-163: x = b / diag(A)."""
    return b / A.diagonal()


def damped_jacobi(A, b, x0=None, tol=1e-6, maxiter=10, log=None):
    """sources/solvers.py:167-235.  omega = 2/3 (193); dr = omega*r/diag (213); x = x + dr
    (217); early exit when dr.dr < tol**2, tested AFTER the update (219-222); returns x only."""
    omega = 2.0 / 3
    D = A.diagonal()
    x = 0.0 * b.copy() if x0 is None else x0.copy()
    tol_sqr = tol ** 2
    for k in range(1, maxiter + 1):
        r = b - A.dot(x)
        dr = omega * r / D
        x = x + dr
        nrmr = _dot(dr, dr, log)
        if nrmr < tol_sqr:
            break
    return x


def pcg(A, psolve, b, x0=None, tol=1e-6, maxiter=100, log=None):
    """sources/solvers.py:69-135, including its quirks: stopping rule
    r.r < tol*sqrt(r0.r0) (87,111-113); niter decremented on break (114); the discarded
    mat-vec of line 109 has no observable effect and is omitted."""
    x = 0.0 * b.copy() if x0 is None else x0.copy()
    r = b - A.dot(x)
    nrmr0 = sqrt(_dot(r, r, log))
    s = psolve(A, r)
    p = s
    sr = _dot(s, r, log)
    hist = []
    k = 0
    nrmr = nrmr0 ** 2
    for k in range(1, maxiter + 1):
        q = A.dot(p)
        alpha = sr / _dot(p, q, log)
        x = x + alpha * p
        r = r - alpha * q
        nrmr = _dot(r, r, log)
        hist.append(nrmr)
        if nrmr < tol * nrmr0:
            k -= 1
            break
        s = psolve(A, r)
        srold = sr
        sr = _dot(s, r, log)
        beta = sr / srold
        p = s + beta * p
    info = {"niter": k, "success": nrmr < tol * nrmr0, "res_norm": sqrt(nrmr),
            "history": np.sqrt(np.array(hist))}
    return x, info


def pcg_glt(A, M1, M2, b, x0=None, tol=1e-6, maxiter=100, log=None, factors=None):
    """sources/solvers.py:239-306: pcg whose preconditioner is kron_solve_par(M2, M1, r)
    (260, 288): the FIRST argument acts along axis 1.  M1, M2 are (n, 2p+1) bands."""
    mats = [band_to_dense(M2), band_to_dense(M1)]

    def psolve(A_, r):
        return kron_solve_dense(mats, r)

    return pcg(A, psolve, b, x0=x0, tol=tol, maxiter=maxiter, log=log)


def crl(A, b, x0=None, tol=1e-5, maxiter=1000, log=None):
    """sources/solvers.py:3-65 (conjugate residuals)."""
    x = 0.0 * b.copy() if x0 is None else x0.copy()
    r = b - A.dot(x)
    p = r.copy()
    q = A.dot(p)
    s = q.copy()
    sr = _dot(s, r, log)
    tol_sqr = tol ** 2
    k = 0
    for k in range(1, maxiter + 1):
        if sr < tol_sqr:
            k -= 1
            break
        alpha = sr / _dot(q, q, log)
        x = x + alpha * p
        r = r - alpha * q
        s = A.dot(r)
        srold = sr
        sr = _dot(s, r, log)
        beta = sr / srold
        p = r + beta * p
        q = s + beta * q
    info = {"niter": k, "success": sr < tol_sqr, "res_norm": sqrt(sr)}
    return x, info


# =============================================================================================
# Grid transfer and the two-grid cycle (sources/mg_jac.py, sources/mg_glt.py)
# =============================================================================================


def restrict(P1s, R):
    """rc = (P1^T (x) ... (x) P1^T) rf  (sources/mg_jac.py:68-69,94), applied axis by axis."""
    for ax, P1 in enumerate(P1s):
        R = np.moveaxis(np.tensordot(P1.T, R, axes=([1], [ax])), 0, ax)
    return R


def prolong(P1s, E):
    """ef = (P1 (x) ... (x) P1) ec  (sources/mg_jac.py:70,102)."""
    for ax, P1 in enumerate(P1s):
        E = np.moveaxis(np.tensordot(P1, E, axes=([1], [ax])), 0, ax)
    return E


def galerkin_dense(A_csr, P1s):
    """Ac = R * Af * P (sources/mg_jac.py:81), dense result."""
    from scipy.sparse import kron, csr_matrix
    P = csr_matrix(P1s[0])
    for P1 in P1s[1:]:
        P = kron(P, csr_matrix(P1), format="csr")
    return (P.T @ A_csr @ P).toarray()


def two_grid(A, P1s, Ac_dense, b, post="jac", M1=None, M2=None, p=None, log_pre=None,
             log_post=None):
    """One two-grid cycle exactly as sources/mg_jac.py:85-119 (post='jac') or
    sources/mg_glt.py:84-123 (post='glt'): PCG-Jacobi pre-smoothing (tol 1e-6, 10 its),
    residual, restriction, direct coarse solve (splu, 98-99), prolongation + correction,
    post-smoothing (PCG-Jacobi 10 its, or pcg_glt p+1 its)."""
    from scipy.sparse import csc_matrix
    from scipy.sparse.linalg import splu
    out = {}

    def dj_pre(A_, r):
        return damped_jacobi(A_, r, log=log_pre)

    def dj_post(A_, r):
        return damped_jacobi(A_, r, log=log_post)

    xf, info_pre = pcg(A, dj_pre, b, tol=1e-6, maxiter=10, log=log_pre)
    rf = b - A.dot(xf)
    rc = restrict(P1s, rf)
    xc = splu(csc_matrix(Ac_dense)).solve(rc.ravel()).reshape(rc.shape)
    xf1 = xf + prolong(P1s, xc)
    if post == "jac":
        xf2, info_post = pcg(A, dj_post, b, x0=xf1, tol=1e-6, maxiter=10, log=log_post)
    else:
        xf2, info_post = pcg_glt(A, M1, M2, b, x0=xf1, tol=1e-6, maxiter=p + 1, log=log_post)
    out.update(x_pre=xf, info_pre=info_pre, r_f=rf, r_c=rc, x_c=xc, x_corr=xf1, x_post=xf2,
               info_post=info_post)
    return out


# =============================================================================================
# EXTENSION (no reference counterpart): multi-level V-cycle + MG-preconditioned CG.
# Restates poms_b200/mg.py on the CPU with independent NumPy/SciPy building blocks so that the
# GPU path can be checked for identical iteration counts and histories.  The outer driver is
# the reference's pcg (above); only the stopping rule can be switched to the true relative one.
# =============================================================================================


def pcg_relative(A, psolve, b, x0=None, tol=1e-10, maxiter=200, log=None, abs_thresh=None):
    """pcg with the BASELINE metric's rule ||r|| <= tol*||r0|| instead of the reference's."""
    x = 0.0 * b.copy() if x0 is None else x0.copy()
    r = b - A.dot(x)
    nrmr0 = sqrt(_dot(r, r, log))
    thresh = (tol * nrmr0) ** 2
    if abs_thresh is not None:
        thresh = abs_thresh
        if nrmr0 * nrmr0 <= thresh:
            return x, {"niter": 0, "success": True, "res_norm": nrmr0, "history": np.zeros(0),
                       "res_norm0": nrmr0}
    s = psolve(A, r)
    p = s
    sr = _dot(s, r, log)
    hist = []
    k = 0
    nrmr = nrmr0 ** 2
    for k in range(1, maxiter + 1):
        q = A.dot(p)
        alpha = sr / _dot(p, q, log)
        x = x + alpha * p
        r = r - alpha * q
        nrmr = _dot(r, r, log)
        hist.append(nrmr)
        if nrmr <= thresh:
            k -= 1
            break
        s = psolve(A, r)
        srold = sr
        sr = _dot(s, r, log)
        p = s + (sr / srold) * p
    return x, {"niter": k, "success": nrmr <= thresh, "res_norm": sqrt(nrmr),
               "history": np.sqrt(np.array(hist)), "res_norm0": nrmr0}


def _symbol_range(band):
    n, w = band.shape
    p = (w - 1) // 2
    row = band[n // 2]
    th = np.linspace(0.0, np.pi, 4097)
    m = row[p] + 2.0 * sum(row[p + k] * np.cos(k * th) for k in range(1, p + 1))
    return float(m.min()), float(m.max())


def _cheb_inverse_poly(lmin, lmax, degree):
    """q(t) = (1 - T_{k+1}((a-t)/d)/T_{k+1}(a/d))/t in monomial coefficients c_0..c_k."""
    from numpy.polynomial import chebyshev as C, polynomial as P
    a, d = 0.5 * (lmax + lmin), 0.5 * (lmax - lmin)
    mono = C.Chebyshev.basis(degree + 1).convert(kind=np.polynomial.Polynomial).coef
    out, powr = np.zeros(1), np.ones(1)
    for c in mono:
        out = P.polyadd(out, c * powr)
        powr = P.polymul(powr, np.array([a / d, -1.0 / d]))
    out = out / C.Chebyshev.basis(degree + 1)(a / d)
    return P.polysub(np.ones(1), out)[1:]


class MGHierarchy:
    def __init__(self, p, N, Nc=8, smoother="glt", nu=1, ratio=4.0, safety=1.1, lengths=None,
                 coarsen="semi"):
        from scipy.linalg import eigh
        lengths = [1.0] * len(N) if lengths is None else [float(v) for v in lengths]
        self.p, self.smoother, self.nu, self.ratio = p, smoother, nu, ratio
        N = list(N)
        d = len(N)
        self.levels = []
        while True:
            knots = [make_open_knots(p, n + p) * L for n, L in zip(N, lengths)]
            A, Mb, Kb = poisson_operator(p, knots)
            self.levels.append(dict(N=list(N), knots=knots, A=A, Mb=Mb, Kb=Kb))
            if all(n <= Nc for n in N):
                break
            # 'uniform': stop as soon as one axis cannot be halved (elements keep their shape on every
            # level); 'semi': keep halving the longer axes alone
            if coarsen == "uniform" and not all(n > Nc and n % 2 == 0 for n in N):
                break
            nxt = [n // 2 if (n > Nc and n % 2 == 0) else n for n in N]
            if nxt == N:
                break
            N = nxt
        for f, c in zip(self.levels[:-1], self.levels[1:]):
            P1s = []
            for a in range(d):
                nf, nc = f["N"][a] + p, c["N"][a] + p
                if nf == nc:
                    P1s.append(np.eye(nf))
                    continue
                ts = knots_to_insert(f["knots"][a], nf, p, c["knots"][a], nc, p)
                P1s.append(insertion_matrix(ts, nc, p, c["knots"][a]))
            f["P1s"] = P1s
        # coarsest level: the reference's exact solve (splu(csc_matrix(Ac)).solve,
        # /root/reference/sources/mg_jac.py:98-99), restated as a dense inverse while that is small.
        # Above DENSE_COARSE_MAX unknowns (the bench's own C3 hierarchy stops at 35^3 = 42 875: a
        # 14.7 GB inverse, 13 minutes on 8 cores) the same exact solve goes through the 1-D
        # generalised eigenpairs K_a Q_a = M_a Q_a L_a, Q_a^T M_a Q_a = I:
        #   A^-1 = (x)Q_a . diag(1 / (1 + sum_a l_a)) . (x)Q_a^T        (A = sum_a K_a (x) M_others + (x)M)
        cl = self.levels[-1]
        if int(np.prod(cl["A"].npts)) <= self.DENSE_COARSE_MAX:
            self.Ainv_c = np.linalg.inv(cl["A"].tocsr().toarray())
            self._fd = None
        else:
            from scipy.linalg import eigh
            self.Ainv_c = None
            Qs, D = [], 1.0
            for a, (Mb_, Kb_) in enumerate(zip(cl["Mb"], cl["Kb"])):
                Md, Kd = band_to_dense(Mb_), band_to_dense(Kb_)
                w, Q = eigh(0.5 * (Kd + Kd.T), 0.5 * (Md + Md.T))
                Qs.append(Q)
                shp = [1] * d
                shp[a] = len(w)
                D = D + w.reshape(shp)
            self._fd = (Qs, D)
        for lv in self.levels[:-1]:
            A = lv["A"]
            if smoother == "glt":
                q = max(2 * p - 1, 1)
                bands = [glt_band(p, n, degree=q) for n in A.npts]
                lv["glt"] = [band_factor(b) for b in bands]
                muM = [eigh(band_to_dense(lv["Mb"][a]), band_to_dense(bands[a]),
                            eigvals_only=True)[-1] for a in range(d)]
                best = 0.0
                for a in range(d):
                    KM = band_to_dense(lv["Kb"][a]) + band_to_dense(lv["Mb"][a])
                    muK = eigh(KM, band_to_dense(bands[a]), eigvals_only=True)[-1]
                    best = max(best, muK * float(np.prod([muM[c] for c in range(d) if c != a])))
                lv["lmax"] = safety * best
            elif smoother == "glt_poly":
                # restates poms_b200/mg.py 'glt_poly': B^-1 ~ q(T) per axis, q = degree-3 Chebyshev
                # approximation of 1/t on the symbol range of T widened by 2 %
                q = max(2 * p - 1, 1)
                bands = [glt_band(p, n, degree=q) for n in A.npts]
                lv["gband"] = bands
                lv["qc"] = []
                for b_ in bands:
                    lo, hi = _symbol_range(b_)
                    lv["qc"].append(_cheb_inverse_poly(lo * 0.98, hi * 1.02, 3))
                muM = [eigh(band_to_dense(lv["Mb"][a]), band_to_dense(bands[a]),
                            eigvals_only=True)[-1] for a in range(d)]
                best = 0.0
                for a in range(d):
                    KM = band_to_dense(lv["Kb"][a]) + band_to_dense(lv["Mb"][a])
                    muK = eigh(KM, band_to_dense(bands[a]), eigvals_only=True)[-1]
                    best = max(best, muK * float(np.prod([muM[c] for c in range(d) if c != a])))
                lv["lmax"] = safety * best      # q(t) t = 0.9 at the high-frequency end: no inflation needed
            else:
                raise NotImplementedError
            lv["lmin"] = lv["lmax"] / ratio

    def smooth(self, lv, b, x, zero_guess):
        A = lv["A"]
        theta = 0.5 * (lv["lmax"] + lv["lmin"])
        if self.smoother == "glt_poly":
            for k in range(self.nu):
                z = b if (k == 0 and zero_guess) else b - A.dot(x)
                for ax in range(A.ndim):        # Horner evaluation of q(T) along every axis
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

    DENSE_COARSE_MAX = 6000

    def coarse_solve(self, b):
        if self._fd is None:
            return (self.Ainv_c @ b.ravel()).reshape(b.shape)
        Qs, D = self._fd
        z = b
        for a, Q in enumerate(Qs):
            z = np.moveaxis(np.tensordot(Q.T, z, axes=(1, a)), 0, a)
        z = z / D
        for a, Q in enumerate(Qs):
            z = np.moveaxis(np.tensordot(Q, z, axes=(1, a)), 0, a)
        return z

    def vcycle(self, l, b):
        lv = self.levels[l]
        if l == len(self.levels) - 1:
            return self.coarse_solve(b)
        x = self.smooth(lv, b, np.zeros_like(b), True)
        r = b - lv["A"].dot(x)
        x = x + prolong(lv["P1s"], self.vcycle(l + 1, restrict(lv["P1s"], r)))
        return self.smooth(lv, b, x, False)

    def mg_pcg(self, b, tol=1e-10, maxiter=200, log=None, max_restarts=3):
        A = self.levels[0]["A"]

        def psolve(A_, r):
            return self.vcycle(0, r)

        x, info = pcg_relative(A, psolve, b, tol=tol, maxiter=maxiter, log=log)
        target = (tol * info["res_norm0"]) ** 2
        info["restarts"] = 0
        while info["restarts"] < max_restarts and info["niter"] < maxiter:
            x2, i2 = pcg_relative(A, psolve, b, x0=x, tol=tol, maxiter=maxiter - info["niter"],
                                  log=log, abs_thresh=target)
            info["res_norm"] = i2["res_norm"] if i2["niter"] else i2["res_norm0"]
            if i2["niter"] == 0:
                break
            x = x2
            info["niter"] += i2["niter"]
            info["history"] = np.concatenate([info["history"], i2["history"]])
            info["restarts"] += 1
        info["success"] = bool(info["res_norm"] ** 2 <= target)
        return x, info


# ==========================================================================================
# EXTENSION checkers: the solvers the reference names as future work (slides/content.tex:393-394)
# and the full 3-D stencil (dimension-generic form of StencilOperator2D)
# ==========================================================================================
class StencilOperatorND:
    """Full d-dim stencil v[i] = sum_k S[i, k] u[i + k - p]; data (n1..nd, 2p1+1..2pd+1)."""

    def __init__(self, data):
        self.data = np.asarray(data, float)
        d = self.data.ndim // 2
        self.ndim = d
        self.npts = tuple(self.data.shape[:d])
        self.pads = tuple((w - 1) // 2 for w in self.data.shape[d:])
        n = int(np.prod(self.npts))
        self.shape = (n, n)

    def dot(self, X):
        d = self.ndim
        Xp = np.zeros(tuple(n + 2 * p for n, p in zip(self.npts, self.pads)))
        Xp[tuple(slice(p, p + n) for n, p in zip(self.npts, self.pads))] = X
        Y = np.zeros(self.npts)
        for ks in np.ndindex(*[2 * p + 1 for p in self.pads]):
            Y += self.data[(Ellipsis,) + ks] * Xp[tuple(slice(k, k + n) for k, n in zip(ks, self.npts))]
        return Y

    def diagonal(self):
        return self.data[(Ellipsis,) + tuple(self.pads)].copy()


def gmres(A, b, x0=None, tol=1e-6, maxiter=100, restart=30, psolve=None):
    """CPU restatement of poms_b200.solvers.gmres: restarted GMRES, right preconditioning, CGS2
    Arnoldi, Givens least squares, stop at ||r|| <= tol ||r0||.  Returns (x, info)."""
    x = np.zeros_like(b) if x0 is None else x0.copy()
    niter, nrm0, res, hist = 0, None, None, []
    while True:
        r = b - A.dot(x)
        beta = float(np.sqrt(np.vdot(r, r)))
        if nrm0 is None:
            nrm0 = beta
        res = beta
        if beta <= tol * nrm0 or niter >= maxiter or beta == 0.0:
            break
        basis = [r / beta]
        H = np.zeros((restart + 1, restart))
        cs, sn = np.zeros(restart), np.zeros(restart)
        g = np.zeros(restart + 1)
        g[0] = beta
        k_used = 0
        for k in range(restart):
            z = psolve(A, basis[k]) if psolve is not None else basis[k]
            w = A.dot(z)
            h = np.array([np.vdot(w, v) for v in basis])
            for hi, v in zip(h, basis):
                w = w - hi * v
            h2 = np.array([np.vdot(w, v) for v in basis])
            for hi, v in zip(h2, basis):
                w = w - hi * v
            h = h + h2
            hk1 = float(np.sqrt(max(np.vdot(w, w), 0.0)))
            H[:k + 1, k] = h
            H[k + 1, k] = hk1
            basis.append(w / hk1 if hk1 > 0 else w)
            for i in range(k):
                t = cs[i] * H[i, k] + sn[i] * H[i + 1, k]
                H[i + 1, k] = -sn[i] * H[i, k] + cs[i] * H[i + 1, k]
                H[i, k] = t
            den = np.hypot(H[k, k], H[k + 1, k])
            cs[k], sn[k] = (H[k, k] / den, H[k + 1, k] / den) if den > 0 else (1.0, 0.0)
            H[k, k] = cs[k] * H[k, k] + sn[k] * H[k + 1, k]
            H[k + 1, k] = 0.0
            g[k + 1] = -sn[k] * g[k]
            g[k] = cs[k] * g[k]
            niter += 1
            k_used = k + 1
            res = abs(g[k + 1])
            hist.append(res)
            if res <= tol * nrm0 or niter >= maxiter or hk1 == 0.0:
                break
        yk = np.linalg.solve(np.triu(H[:k_used, :k_used]), g[:k_used]) if k_used else np.zeros(0)
        u = np.zeros_like(b)
        for yi, v in zip(yk, basis):
            u = u + yi * v
        x = x + (psolve(A, u) if psolve is not None else u)
        if res <= tol * nrm0 or niter >= maxiter:
            r = b - A.dot(x)
            res = float(np.sqrt(np.vdot(r, r)))
            if res <= tol * nrm0 or niter >= maxiter:
                break
    return x, {"niter": niter, "success": bool(res <= tol * nrm0), "res_norm": res, "res_norm0": nrm0,
               "history": hist}


def rb_jacobi(A, b, x0=None, tol=1e-6, maxiter=10, omega=2.0 / 3.0):
    """CPU restatement of poms_b200.solvers.rb_jacobi (two-colour damped Jacobi)."""
    D = A.diagonal()
    x = np.zeros_like(b) if x0 is None else x0.copy()
    idx = np.indices(b.shape).sum(axis=0)
    tol_sqr = tol ** 2
    for k in range(1, maxiter + 1):
        tot = 0.0
        for colour in (0, 1):
            d = omega * (b - A.dot(x)) / D
            tot += 0.5 * float(np.vdot(d, d))
            mask = (idx & 1) == colour
            x = x + np.where(mask, d, 0.0)
        if tot < tol_sqr:
            break
    return x
