"""ORACLE-side design experiment (not product code): V-cycle variants as PCG preconditioner
for the 3-D/2-D degree-p operator, to choose the MG-PCG defaults.  Usage:
    python oracle/mg_experiments.py d p N"""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np
from math import sqrt
from oracle import poms_oracle as po
from scipy.sparse.linalg import eigsh, LinearOperator

def build(d, p, N, Nc=8):
    levels = []
    n = N
    while True:
        T = po.make_open_knots(p, n + p)
        A, Mb, Kb = po.poisson_operator(p, [T] * d)
        levels.append(dict(N=n, T=T, A=A, Mb=Mb, Kb=Kb))
        if n <= Nc: break
        n //= 2
    for f, c in zip(levels[:-1], levels[1:]):
        ts = po.knots_to_insert(f["T"], f["N"] + p, p, c["T"], c["N"] + p, p)
        f["P1"] = po.insertion_matrix(ts, c["N"] + p, p, c["T"])
    return levels

def lam_max(A):
    D = A.diagonal(); n = A.npts; Nn = int(np.prod(n)); Dh = 1 / np.sqrt(D)
    op = LinearOperator((Nn, Nn), matvec=lambda v: (Dh * A.dot((Dh.ravel() * v).reshape(n))).ravel(), dtype=float)
    return eigsh(op, k=1, which="LA", return_eigenvectors=False, tol=1e-3)[0]

def jac(A, b, x, nu, omega):
    D = A.diagonal()
    for _ in range(nu):
        x = x + omega * (b - A.dot(x)) / D
    return x

def mass_solver(lv, d, p, q=None):
    bands = [po.glt_band(p, lv["A"].npts[a], degree=q) if q is not None else lv["Mb"][a] for a in range(d)]
    f = [po.band_factor(b) for b in bands]
    return lambda r: po.kron_solve_banded(f, r)

def pcg_smooth(A, minv, b, x, nu):
    # nu steps of CG preconditioned by minv starting from x (the reference's pcg_glt used as smoother)
    r = b - A.dot(x); s = minv(r); pp = s; sr = np.vdot(s, r)
    for k in range(nu):
        q = A.dot(pp); al = sr / np.vdot(pp, q); x = x + al * pp; r = r - al * q
        s = minv(r); sro = sr; sr = np.vdot(s, r); pp = s + sr / sro * pp
    return x

def vcycle(levels, l, b, smooth):
    lv = levels[l]; A = lv["A"]
    if l == len(levels) - 1:
        if "Ainv" not in lv: lv["Ainv"] = np.linalg.inv(A.tocsr().toarray())
        return (lv["Ainv"] @ b.ravel()).reshape(b.shape)
    x = smooth(lv, b, np.zeros_like(b), True)
    r = b - A.dot(x)
    P1s = [lv["P1"]] * A.ndim
    x = x + po.prolong(P1s, vcycle(levels, l + 1, po.restrict(P1s, r), smooth))
    return smooth(lv, b, x, False)

def run(d, p, N, smooth, tol=1e-10, maxiter=200, label=""):
    levels = build(d, p, N)
    A = levels[0]["A"]; b = np.ones(A.npts)
    x = np.zeros_like(b); r = b.copy(); n0 = sqrt(np.vdot(r, r))
    s = vcycle(levels, 0, r, smooth); pp = s; sr = np.vdot(s, r); hist = []
    for k in range(1, maxiter + 1):
        q = A.dot(pp); al = sr / np.vdot(pp, q); x = x + al * pp; r = r - al * q
        nr = sqrt(np.vdot(r, r)); hist.append(nr / n0)
        if nr <= tol * n0: break
        s = vcycle(levels, 0, r, smooth); sro = sr; sr = np.vdot(s, r); pp = s + sr / sro * pp
    print("%-40s d=%d p=%d N=%d: iters=%d  final=%.2e  rho=%.3f" % (label, d, p, N, k, hist[-1], hist[-1] ** (1.0 / k)))
    return k

if __name__ == "__main__":
    d, p, N = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    lm = {}
    def sm_jac(nu, scale):
        def f(lv, b, x, pre):
            key = lv["N"]
            if key not in lm: lm[key] = lam_max(lv["A"])
            return jac(lv["A"], b, x, nu, scale / lm[key])
        return f
    def sm_mass(nu, q=None):
        def f(lv, b, x, pre):
            if "minv%s" % q not in lv: lv["minv%s" % q] = mass_solver(lv, d, p, q)
            return pcg_smooth(lv["A"], lv["minv%s" % q], b, x, nu)
        return f
    for nu in (1, 2, 3):
        run(d, p, N, sm_jac(nu, 4.0 / 3), label="V(%d,%d) Jacobi w=4/(3 lmax)" % (nu, nu))
    for nu in (1, 2):
        run(d, p, N, sm_mass(nu), label="V(%d,%d) PCG-mass(M_p) smoother" % (nu, nu))
    if p > 1:
        run(d, p, N, sm_mass(1, 2 * p - 1), label="V(1,1) PCG-T[m_{p-1}] (deg 2p-1 card.)")
        run(d, p, N, sm_mass(2, 2 * p - 1), label="V(2,2) PCG-T[m_{p-1}]")
    run(d, p, N, sm_mass(1, p), label="V(1,1) PCG-colloc(p) smoother")

def cheb(A, minv, b, x, nu, lmin, lmax):
    theta = 0.5 * (lmax + lmin); delta = 0.5 * (lmax - lmin); sigma = theta / delta; rho = 1.0 / sigma
    r = b - A.dot(x); dd = minv(r) / theta
    for k in range(nu):
        x = x + dd
        if k == nu - 1: break
        r = r - A.dot(dd)
        rho_n = 1.0 / (2 * sigma - rho)
        dd = rho_n * rho * dd + 2 * rho_n / delta * minv(r)
        rho = rho_n
    return x

def lam_max_prec(A, minv):
    n = A.npts; Nn = int(np.prod(n))
    # power iteration on minv*A
    v = np.random.default_rng(0).standard_normal(n)
    for _ in range(40):
        w = minv(A.dot(v)); lam = np.vdot(v, w) / np.vdot(v, v); v = w / np.linalg.norm(w)
    return lam

def experiments2(d, p, N):
    cache = {}
    def sm_cheb(nu, kind, ratio):
        def f(lv, b, x, pre):
            key = (lv["N"], kind)
            if key not in cache:
                if kind == "jac":
                    D = lv["A"].diagonal(); minv = lambda r: r / D
                else:
                    minv = mass_solver(lv, d, p, None if kind == "mass" else kind)
                cache[key] = (minv, lam_max_prec(lv["A"], minv) * 1.05)
            minv, lmax = cache[key]
            return cheb(lv["A"], minv, b, x, nu, lmax / ratio, lmax)
        return f
    for kind in ("jac", "mass", 2 * p - 1):
        for nu, ratio in ((1, 4), (2, 4), (2, 10), (3, 10), (4, 30)):
            run(d, p, N, sm_cheb(nu, kind, ratio), label="V(%d,%d) Cheb-%s ratio %g" % (nu, nu, kind, ratio))

if __name__ == "__main__" and len(sys.argv) > 4:
    experiments2(int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]))
