"""Drop-in for /root/reference/sources/solvers.py: crl, pcg, jacobi, damped_jacobi, pcg_glt.

Same signatures, same operation order, same stopping rules and `info` keys as the reference
(SURVEY.md Appendix A lists the quirks that are reproduced on purpose).  The vectors stay on
the device; per iteration the host reads back exactly one scalar (the quantity the
reference's break test needs).  alpha and beta never visit the host: the fused kernels read
their numerator / denominator from device scalars.

`info` additionally carries 'history' (sqrt of every r.r the driver computed) so that
residual histories can be compared with the oracle; the reference only prints them.
"""
from math import sqrt

import numpy as np

import torch

from . import _lib
from . import profiling
from .stencil import (StencilVector, DeviceContext, dot_into, _stream, EPI_STORE, EPI_RESID,
                      EPI_JACOBI, EPI_DINV)

__all__ = ["crl", "pcg", "jacobi", "damped_jacobi", "pcg_glt", "gmres", "rb_jacobi"]

# device scalar slots (DeviceContext.scal)
S_TMP, S_RR, S_PQ, S_SR0, S_SR1, S_DR, S_QQ = 0, 1, 2, 3, 4, 5, 6


def _read(ctx, V, slot):
    """Value of a device scalar on the host (sum over slabs when partitioned)."""
    v = ctx.scal[slot:slot + 1]
    if V.slab is not None and V.slab.size > 1:
        V.slab.allreduce_sum(v)
    return float(v.item())


def _reduce(ctx, V, slot):
    """Make a device scalar global (no host sync)."""
    if V.slab is not None and V.slab.size > 1:
        V.slab.allreduce_sum(ctx.scal[slot:slot + 1])


def _check_shapes(A, b, x0):
    n = A.shape[0]
    assert A.shape == (n, n)
    assert b.shape == (n,)
    if x0 is not None:
        assert x0.shape == (n,)


def _residual(A, b, x0, ctx):
    """x, r = b - A x, r.r -> S_RR  (/root/reference/sources/solvers.py:79-87)."""
    V = b.space
    if x0 is None:
        # reference: x = 0.0*b.copy(); r = b - A.dot(x).  A.0 = 0 exactly, so r = b bit for bit.
        x = StencilVector(V)
        r = b.copy()
        dot_into(r, r, ctx.sptr(S_RR), ctx)
    else:
        x = x0.copy()
        r = StencilVector(V, zero=False)
        A.apply(x, r, EPI_RESID, b=b, dot_ptr=ctx.sptr(S_RR))
    return x, r


def _pcg_driver(A, psolve, b, x0, tol, maxiter, verbose, title, relative=False, abs_thresh=None):
    _check_shapes(A, b, x0)
    V = b.space
    ctx = DeviceContext.get(V.device)
    L = _lib.lib()
    x, r = _residual(A, b, x0, ctx)
    nrmr0 = sqrt(_read(ctx, V, S_RR))
    # reference rule (lines 111-113): r.r < tol*||r0|| (squared vs unsquared norm);
    # relative=True (EXTENSION, the BASELINE metric): ||r|| <= tol*||r0||
    thresh = (tol * nrmr0) ** 2 if relative else tol * nrmr0
    if abs_thresh is not None:      # restart of mg_pcg: absolute target on r.r (EXTENSION)
        thresh = abs_thresh
        if nrmr0 * nrmr0 <= thresh:
            return x, {"niter": 0, "success": True, "res_norm": nrmr0, "history": [],
                       "res_norm0": nrmr0}
    s = psolve(A, r)
    # The reference aliases p = s (line 90) and later REBINDS p and r; here p and r are updated in
    # place, so p needs storage of its own: psolve may return r itself (unpreconditioned CG) or the
    # same pre-allocated vector on every call.
    p = s.copy()
    cur = S_SR0
    dot_into(s, r, ctx.sptr(cur), ctx)
    _reduce(ctx, V, cur)
    q = StencilVector(V, zero=False)
    if verbose:
        print(title)
        print("+---------+---------------------+")
        print("+ Iter. # | L2-norm of residual |")
        print("+---------+---------------------+")
    template = "| {:7d} | {:19.2e} |"
    hist = []
    k = 0
    nrmr = nrmr0 * nrmr0
    for k in range(1, maxiter + 1):
        # q = A p ; p.q fused into the mat-vec epilogue (lines 103-104)
        A.apply(p, q, EPI_STORE, dot_ptr=ctx.sptr(S_PQ))
        _reduce(ctx, V, S_PQ)
        # x += alpha p ; r -= alpha q ; r.r   (lines 106-111; alpha = sr / p.q on the device)
        with profiling.region("cg_update", 48 * x.n_owned):
            _lib.check(L.poms_cg_update(x.ptr, r.ptr, p.ptr, q.ptr, x.n_owned, ctx.sptr(cur),
                                        ctx.sptr(S_PQ), ctx.sptr(S_RR), ctx.ws_ptr, _stream()),
                       "poms_cg_update")
        # (line 109 `s = A.dot(r)` is dead: overwritten at 117 or unused after the break)
        nrmr = _read(ctx, V, S_RR)
        hist.append(sqrt(nrmr))
        if (nrmr <= thresh) if relative else (nrmr < thresh):
            k -= 1
            break
        s = psolve(A, r)
        nxt = S_SR1 if cur == S_SR0 else S_SR0
        with profiling.region("dot", 16 * x.n_owned):
            dot_into(s, r, ctx.sptr(nxt), ctx)
        _reduce(ctx, V, nxt)
        # p = s + (sr/srold) p   (lines 119-124)
        with profiling.region("p_update", 24 * x.n_owned):
            _lib.check(L.poms_p_update(p.ptr, s.ptr, p.n_owned, ctx.sptr(nxt), ctx.sptr(cur),
                                       _stream()), "poms_p_update")
        cur = nxt
        if verbose:
            print(template.format(k, sqrt(nrmr)))
    if verbose:
        print("+---------+---------------------+")
    info = {"niter": k, "success": bool(nrmr <= thresh) if relative else bool(nrmr < thresh),
            "res_norm": sqrt(nrmr), "history": hist, "res_norm0": nrmr0}
    return x, info


def pcg(A, psolve, b, x0=None, tol=1e-6, maxiter=100, verbose=False):
    """Preconditioned CG (/root/reference/sources/solvers.py:69-135).  `psolve(A, r) -> s`."""
    return _pcg_driver(A, psolve, b, x0, tol, maxiter, verbose, "CG solver:")


def pcg_glt(A, M1, M2, b, x0=None, tol=1e-6, maxiter=100, verbose=False):
    """PCG preconditioned by the Kronecker solve kron_solve_par(M2, M1, r): the FIRST argument
    acts along axis 1 (/root/reference/sources/solvers.py:239-306, lines 260/288)."""
    from .kron_product import kron_solve_par

    def psolve(A_, r):
        return kron_solve_par(M2, M1, r)

    return _pcg_driver(A, psolve, b, x0, tol, maxiter, verbose, "CG-GL-GLTT solver:")


def jacobi(A, b):
    """x = b / diag(A) (/root/reference/sources/solvers.py:139-163)."""
    _check_shapes(A, b, None)
    x = StencilVector(b.space, zero=False)
    A.jacobi_first(x, b, 1.0, None)
    return x


def damped_jacobi(A, b, x0=None, tol=1e-6, maxiter=10, verbose=False, omega=2.0 / 3):
    """Weighted Jacobi (/root/reference/sources/solvers.py:167-235): up to `maxiter` sweeps
    x += omega*(b - A x)/diag(A), early exit when dr.dr < tol**2 (tested after the update,
    lines 217-222), returns x only (line 235).  omega = 2/3 is hard-coded in the reference
    (line 193); it is a keyword here because that value diverges for p >= 3."""
    _check_shapes(A, b, x0)
    V = b.space
    ctx = DeviceContext.get(V.device)
    tol_sqr = tol ** 2
    first = 1
    if x0 is None:
        if maxiter < 1:
            return StencilVector(V)
        # sweep 1 from x = 0: r = b - A.0 = b, so x = omega*b/diag (one 16 B/DOF pass)
        x = StencilVector(V, zero=False)
        A.jacobi_first(x, b, omega, ctx.sptr(S_DR))
        if _read(ctx, V, S_DR) < tol_sqr:
            return x
        first = 2
    else:
        x = x0.copy()
    y = StencilVector(V, zero=False)
    for k in range(first, maxiter + 1):
        # one fused pass: y = x + omega*(b - A x)/diag ; dr.dr   (lines 209-219)
        A.apply(x, y, EPI_JACOBI, b=b, omega=omega, dot_ptr=ctx.sptr(S_DR))
        x, y = y, x
        nrmr = _read(ctx, V, S_DR)
        if verbose:
            print("| {:7d} | {:19.2e} |".format(k, sqrt(nrmr)))
        if nrmr < tol_sqr:
            break
    return x


def crl(A, b, x0=None, tol=1e-5, maxiter=1000, verbose=False):
    """Conjugate residuals (/root/reference/sources/solvers.py:3-65)."""
    _check_shapes(A, b, x0)
    V = b.space
    ctx = DeviceContext.get(V.device)
    L = _lib.lib()
    x, r = _residual(A, b, x0, ctx)
    p = r.copy()
    q = A.dot(p)
    s = q.copy()
    cur = S_SR0
    dot_into(s, r, ctx.sptr(cur), ctx)
    sr = _read(ctx, V, cur)
    tol_sqr = tol ** 2
    k = 0
    for k in range(1, maxiter + 1):
        if sr < tol_sqr:
            k -= 1
            break
        dot_into(q, q, ctx.sptr(S_QQ), ctx)
        _reduce(ctx, V, S_QQ)
        # x += alpha p ; r -= alpha q   with alpha = sr / q.q
        _lib.check(L.poms_axpy_dev(x.ptr, p.ptr, x.n_owned, ctx.sptr(cur), ctx.sptr(S_QQ), 1.0,
                                   _stream()), "poms_axpy_dev")
        _lib.check(L.poms_axpy_dev(r.ptr, q.ptr, r.n_owned, ctx.sptr(cur), ctx.sptr(S_QQ), -1.0,
                                   _stream()), "poms_axpy_dev")
        A.apply(r, s, EPI_STORE)
        nxt = S_SR1 if cur == S_SR0 else S_SR0
        dot_into(s, r, ctx.sptr(nxt), ctx)
        sr = _read(ctx, V, nxt)
        # p = r + beta p ; q = s + beta q
        _lib.check(L.poms_p_update(p.ptr, r.ptr, p.n_owned, ctx.sptr(nxt), ctx.sptr(cur),
                                   _stream()), "poms_p_update")
        _lib.check(L.poms_p_update(q.ptr, s.ptr, q.n_owned, ctx.sptr(nxt), ctx.sptr(cur),
                                   _stream()), "poms_p_update")
        cur = nxt
        if verbose:
            print("| {:7d} | {:19.2e} |".format(k, sqrt(sr)))
    info = {"niter": k, "success": sr < tol_sqr, "res_norm": sqrt(max(sr, 0.0))}
    return x, info


# ==========================================================================================
# EXTENSION: the two solvers the reference names as future work (slides/content.tex:393-394)
# ==========================================================================================
S_GM0 = 16          # first of the device scalar slots used by gmres (one per Krylov vector)


def gmres(A, b, x0=None, tol=1e-6, maxiter=100, restart=30, psolve=None, verbose=False):
    """Restarted GMRES(restart) with RIGHT preconditioning (x = x0 + M^-1 V y, so the monitored
    norm is the true residual's), Arnoldi by classical Gram-Schmidt applied twice (CGS2: all
    inner products of a pass are queued on the device and read back together, two host reads per
    step instead of one per basis vector).  `psolve(A, v) -> M^-1 v` must be a fixed linear map.
    Stops when ||r|| <= tol * ||r0||.  Returns (x, info) like pcg; niter counts mat-vecs of the
    Arnoldi process.  The test suite checks it against a NumPy restatement of the same operations."""
    _check_shapes(A, b, x0)
    assert 1 <= restart <= 40
    V = b.space
    ctx = DeviceContext.get(V.device)
    L = _lib.lib()
    st = _stream

    def axpby(z, a_, x_, b_, y_):
        _lib.check(L.poms_axpby(z.ptr, float(a_), x_.ptr, float(b_), y_.ptr if y_ is not None else None,
                                z.n_owned, st()), "poms_axpby")

    def dots(w, basis):
        for i, v in enumerate(basis):
            dot_into(w, v, ctx.sptr(S_GM0 + i), ctx)
        h = ctx.scal[S_GM0:S_GM0 + len(basis)]
        if V.slab is not None and V.slab.size > 1:
            V.slab.allreduce_sum(h)
        return h.cpu().numpy().copy()

    x = StencilVector(V) if x0 is None else x0.copy()
    r = StencilVector(V, zero=False)
    w = StencilVector(V, zero=False)
    basis = [StencilVector(V, zero=False) for _ in range(restart + 1)]
    niter, nrm0, res = 0, None, None
    hist = []
    while True:
        A.apply(x, r, EPI_RESID, b=b, dot_ptr=ctx.sptr(S_RR))
        beta = sqrt(_read(ctx, V, S_RR))
        if nrm0 is None:
            nrm0 = beta
        res = beta
        if beta <= tol * nrm0 or niter >= maxiter or beta == 0.0:
            break
        axpby(basis[0], 1.0 / beta, r, 0.0, None)
        H = np.zeros((restart + 1, restart))
        cs, sn = np.zeros(restart), np.zeros(restart)
        g = np.zeros(restart + 1)
        g[0] = beta
        k_used = 0
        for k in range(restart):
            z = psolve(A, basis[k]) if psolve is not None else basis[k]
            A.apply(z, w, EPI_STORE)
            h = dots(w, basis[:k + 1])
            for i in range(k + 1):
                axpby(w, 1.0, w, -h[i], basis[i])
            h2 = dots(w, basis[:k + 1])
            for i in range(k + 1):
                axpby(w, 1.0, w, -h2[i], basis[i])
            h = h + h2
            hk1 = sqrt(max(w.dot(w), 0.0))
            H[:k + 1, k] = h
            H[k + 1, k] = hk1
            if hk1 > 0.0:
                axpby(basis[k + 1], 1.0 / hk1, w, 0.0, None)
            for i in range(k):                                # earlier Givens rotations
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
            if verbose:
                print("| {:7d} | {:19.2e} |".format(niter, res))
            if res <= tol * nrm0 or niter >= maxiter or hk1 == 0.0:
                break
        yk = np.linalg.solve(np.triu(H[:k_used, :k_used]), g[:k_used]) if k_used else np.zeros(0)
        axpby(w, yk[0], basis[0], 0.0, None)
        for i in range(1, k_used):
            axpby(w, 1.0, w, yk[i], basis[i])
        upd = psolve(A, w) if psolve is not None else w
        axpby(x, 1.0, x, 1.0, upd)
        if res <= tol * nrm0 or niter >= maxiter:
            A.apply(x, r, EPI_RESID, b=b, dot_ptr=ctx.sptr(S_RR))
            res = sqrt(_read(ctx, V, S_RR))
            if res <= tol * nrm0 or niter >= maxiter:
                break
    info = {"niter": niter, "success": bool(res <= tol * nrm0), "res_norm": res, "res_norm0": nrm0,
            "history": hist}
    return x, info


def rb_jacobi(A, b, x0=None, tol=1e-6, maxiter=10, omega=2.0 / 3.0, verbose=False):
    """Two-colour (red-black) damped Jacobi: every sweep updates the points with even i1+i2(+i3)
    from the current residual, then the odd ones from the refreshed residual,
        x_c += omega D^-1 (b - A x)_c ,   c = red, black
    (two fused residual passes per sweep; each half sweep is a Jacobi step on half of the
    unknowns, so omega = 2/3 is stable where the plain sweep of damped_jacobi is not, DESIGN.md
    section 2).  Stops like damped_jacobi when the update of a sweep has ||d||^2 < tol^2; returns x
    only, like damped_jacobi (/root/reference/sources/solvers.py:235).  Checked against a NumPy
    restatement by the test suite."""
    _check_shapes(A, b, x0)
    V = b.space
    ctx = DeviceContext.get(V.device)
    L = _lib.lib()
    x = StencilVector(V) if x0 is None else x0.copy()
    d = StencilVector(V, zero=False)
    shp = tuple(V.local_shape)
    n1, n2, n3 = (1,) * (3 - len(shp)) + shp
    pld = x.pld if len(shp) == 3 else n2 * x.ld
    off = V.starts[0]                      # global index of the first owned plane / row
    tol_sqr = tol ** 2
    for k in range(1, maxiter + 1):
        tot = 0.0
        for colour in (0, 1):
            A.apply(x, d, EPI_DINV, b=b, omega=omega, dot_ptr=ctx.sptr(S_DR))
            tot += 0.5 * _read(ctx, V, S_DR)
            _lib.check(L.poms_color_add(x.ptr, d.ptr, n1, n2, n3, x.ld, pld, off, colour, _stream()),
                       "poms_color_add")
        if verbose:
            print("| {:7d} | {:19.2e} |".format(k, sqrt(tot)))
        if tot < tol_sqr:
            break
    return x
