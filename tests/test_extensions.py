"""Extension components: GMRES, two-colour Jacobi, full 3-D stencils.
CPU part: the oracle restatements against independent SciPy references.  GPU part: the CUDA path
against the oracle on the same inputs (identical iteration counts, 1e-10 on the solutions)."""
import numpy as np
import pytest

from oracle import poms_oracle as po


def _problem(p, N):
    knots = [po.make_open_knots(p, n + p) for n in N]
    A, _, _ = po.poisson_operator(p, knots)
    x0 = np.zeros(A.npts)
    for a in range(len(N)):
        shp = [1] * len(N)
        shp[a] = -1
        x0 = x0 + np.arange(A.npts[a], dtype=float).reshape(shp)
    return knots, A, A.dot(x0 + 1.0), x0 + 1.0


def test_oracle_gmres_matches_scipy_and_direct_solve():
    from scipy.sparse.linalg import spsolve
    knots, A, b, xt = _problem(2, (10, 8))
    x, info = po.gmres(A, b, tol=1e-10, maxiter=400, restart=25)
    xd = spsolve(A.tocsr().tocsc(), b.ravel()).reshape(b.shape)
    assert info["success"]
    assert np.abs(x - xd).max() < 1e-7 * np.abs(xd).max()
    # preconditioned by the diagonal: fewer iterations, same solution
    x2, info2 = po.gmres(A, b, tol=1e-10, maxiter=400, restart=25, psolve=lambda A_, v: v / A.diagonal())
    assert info2["success"] and info2["niter"] <= info["niter"]
    assert np.abs(x2 - xd).max() < 1e-7 * np.abs(xd).max()


def test_oracle_rb_jacobi_converges_where_plain_jacobi_diverges():
    knots, A, b, xt = _problem(3, (12, 12))
    r0 = np.linalg.norm(b)
    xr = po.rb_jacobi(A, b, maxiter=20)
    xj = po.damped_jacobi(A, b, maxiter=20)
    assert np.linalg.norm(b - A.dot(xr)) < r0                    # two-colour sweep contracts
    assert np.linalg.norm(b - A.dot(xj)) > np.linalg.norm(b - A.dot(xr))


def test_oracle_stencil_nd_equals_kron_sum():
    knots, A, b, xt = _problem(2, (5, 4, 6))
    rng = np.random.default_rng(1)
    X = rng.standard_normal(A.npts)
    S = 0.0
    for t in A.terms:
        T = np.ones(())
        for a, bnd in enumerate(t):
            shp = [1] * 6
            shp[a], shp[3 + a] = bnd.shape
            T = T * bnd.reshape(shp)
        S = S + T
    assert np.abs(po.StencilOperatorND(S).dot(X) - A.dot(X)).max() < 1e-12 * np.abs(A.dot(X)).max()


@pytest.fixture(scope="module")
def dev():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda", 0)


@pytest.mark.gpu
@pytest.mark.parametrize("p,N,pre", [(2, (10, 8), False), (3, (12, 12), True), (2, (6, 5, 7), True)])
def test_gmres_vs_oracle(dev, p, N, pre):
    from poms_b200 import solvers
    from poms_b200.stencil import StencilVector, StencilVectorSpace, KronSumMatrix
    knots, Ao, b, xt = _problem(p, N)
    A = KronSumMatrix.poisson(p, knots)
    V = StencilVectorSpace(list(Ao.npts), [p] * len(N), [False] * len(N), device=dev)
    bv = StencilVector.from_array(V, b)
    ps = solvers.jacobi if pre else None
    pso = (lambda A_, v: po.jacobi(Ao, v)) if pre else None
    x, info = solvers.gmres(A, bv, tol=1e-10, maxiter=300, restart=20, psolve=ps)
    xo, io = po.gmres(Ao, b, tol=1e-10, maxiter=300, restart=20, psolve=pso)
    assert info["niter"] == io["niter"], (info["niter"], io["niter"])
    assert info["success"] and io["success"]
    err = np.abs(x.toarray().reshape(xo.shape) - xo).max() / np.abs(xo).max()
    assert err < 1e-10, err
    # residual history: tight while the residual is well above the attainable accuracy, loose in the
    # tail (the estimates |g_k+1| of the last steps carry the rounding of the whole Arnoldi process)
    hg, ho = np.array(info["history"]), np.array(io["history"])
    head = ho > 1e-6 * io["res_norm0"]
    assert np.allclose(hg[head], ho[head], rtol=1e-7)
    assert np.allclose(hg[~head], ho[~head], rtol=5e-2)


@pytest.mark.gpu
@pytest.mark.parametrize("p,N", [(3, (12, 12)), (2, (9, 11)), (3, (8, 6, 7))])
def test_rb_jacobi_vs_oracle(dev, p, N):
    from poms_b200 import solvers
    from poms_b200.stencil import StencilVector, StencilVectorSpace, KronSumMatrix
    knots, Ao, b, xt = _problem(p, N)
    A = KronSumMatrix.poisson(p, knots)
    V = StencilVectorSpace(list(Ao.npts), [p] * len(N), [False] * len(N), device=dev)
    x = solvers.rb_jacobi(A, StencilVector.from_array(V, b), maxiter=8)
    xo = po.rb_jacobi(Ao, b, maxiter=8)
    err = np.abs(x.toarray().reshape(xo.shape) - xo).max() / np.abs(xo).max()
    assert err < 1e-12, err


@pytest.mark.gpu
@pytest.mark.parametrize("p,N", [(1, (6, 5, 7)), (2, (9, 8, 10)), (3, (12, 9, 35))])
def test_full_3d_stencil_matvec_and_solver(dev, p, N):
    """StencilMatrix in 3-D ((2p+1)^3 coefficients per row): every epilogue against the oracle, and the
    reference's pcg on it with identical iteration count."""
    import torch
    from poms_b200 import solvers
    from poms_b200.stencil import (StencilVector, StencilVectorSpace, StencilMatrix, KronSumMatrix,
                                   DeviceContext, EPI_STORE, EPI_RESID, EPI_JACOBI, EPI_DINV, EPI_AXPY)
    knots, Ao, b, xt = _problem(p, N)
    K = KronSumMatrix.poisson(p, knots)
    rng = np.random.default_rng(5)
    Sarr = K.to_stencil_array() * (1.0 + 0.1 * rng.random(tuple(Ao.npts) + (1, 1, 1)))   # non-separable rows
    V = StencilVectorSpace(list(Ao.npts), [p] * 3, [False] * 3, device=dev)
    S = StencilMatrix(V)
    S._data[...] = Sarr
    S.remove_spurious_entries()
    So = po.StencilOperatorND(S._data)
    X, B = rng.standard_normal(Ao.npts), rng.standard_normal(Ao.npts)
    x, bv, y = StencilVector.from_array(V, X), StencilVector.from_array(V, B), StencilVector(V)
    ctx = DeviceContext.get(dev)
    Yo = So.dot(X)
    D = So.diagonal()
    want = {EPI_STORE: (Yo, np.vdot(X, Yo)), EPI_RESID: (B - Yo, np.vdot(B - Yo, B - Yo)),
            EPI_JACOBI: (X + 0.6 * (B - Yo) / D, None), EPI_DINV: (0.6 * (B - Yo) / D, None),
            EPI_AXPY: (B + 0.6 * Yo, None)}
    for epi, (ref, dref) in want.items():
        S.apply(x, y, epi, b=bv, omega=0.6, dot_ptr=ctx.sptr(20))
        got = y.toarray().reshape(ref.shape)
        assert np.abs(got - ref).max() < 1e-13 * np.abs(ref).max(), epi
        if dref is not None:
            assert abs(ctx.scal[20].item() - dref) < 1e-11 * abs(dref)
    bb = StencilVector.from_array(V, So.dot(xt))
    xs, info = solvers.pcg(S, solvers.jacobi, bb, tol=1e-8, maxiter=200)
    xo, io = po.pcg(So, po.jacobi, So.dot(xt), tol=1e-8, maxiter=200)
    assert info["niter"] == io["niter"]
    assert np.abs(xs.toarray().reshape(xo.shape) - xo).max() < 1e-9 * np.abs(xo).max()


@pytest.mark.gpu
@pytest.mark.parametrize("p,N", [(3, (40, 30, 70)), (2, (20, 33, 129))])
def test_axpy_epilogue_with_fused_dot_against_a_third_vector(dev, p, N):
    """EPI_AXPY with dot_with = z: y = b + om * (S x) and the fused reduction y . z (the s.r of the CG
    driver riding in the last smoother pass), on grids with interior, boundary and ragged tiles."""
    import torch
    from poms_b200 import bsplines as bs
    from poms_b200.stencil import (StencilVector, StencilVectorSpace, KronSumMatrix, DeviceContext, EPI_AXPY)
    npts = [n + p for n in N]
    glt = [bs.glt_band(p, n, degree=max(2 * p - 1, 1)) for n in npts]
    S2 = KronSumMatrix([bs.poly_inverse_factors(b_, 3)[1] for b_ in glt])
    V = StencilVectorSpace(npts, [S2.P] * 3, [False] * 3, device=dev)
    rng = np.random.default_rng(7)
    X, B, Z = (rng.standard_normal(npts) for _ in range(3))
    x, b, z, y = (StencilVector.from_array(V, a) for a in (X, B, Z, np.zeros(npts)))
    ctx = DeviceContext.get(dev)
    Yo = B + 0.41 * po.KronSumOperator([tuple(S2.Ms)]).dot(X)
    fused = S2.apply(x, y, EPI_AXPY, b=b, omega=0.41, dot_ptr=ctx.sptr(21), dot_with=z)
    got = y.toarray().reshape(Yo.shape)
    assert np.abs(got - Yo).max() < 1e-13 * np.abs(Yo).max()
    assert fused, "the TMA kernel should have taken the fused reduction"
    want = float(np.vdot(Yo, Z))
    assert abs(ctx.scal[21].item() - want) < 1e-11 * np.abs(Yo * Z).sum()
    # in place (b is y), as the smoother calls it
    y2 = StencilVector.from_array(V, B)
    fused = S2.apply(x, y2, EPI_AXPY, b=y2, omega=0.41, dot_ptr=ctx.sptr(21), dot_with=z)
    assert fused and abs(ctx.scal[21].item() - want) < 1e-11 * np.abs(Yo * Z).sum()
    assert np.abs(y2.toarray().reshape(Yo.shape) - Yo).max() < 1e-13 * np.abs(Yo).max()
