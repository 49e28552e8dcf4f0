"""GPU parity on the code path the bench times: Toeplitz (uniform-knot)
operators at shapes that hit interior tiles, ragged last tiles, boundary fix-up columns and
multi-chunk CTAs at once -- i.e. the constant-bank / steady-plane-loop path of the TMA kernels --
compared with the CPU oracle (`oracle.poms_oracle.KronSumOperator`), every epilogue, both smoother
factors, every kernel variant; and full MG-PCG solves at BASELINE size C3 against the threaded CPU
port with the SAME hierarchy parameters (identical iteration counts, 1e-10 on x)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda", 0)


def rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def _space(npts, pads, dev):
    from poms_b200.stencil import StencilVectorSpace
    return StencilVectorSpace(list(npts), list(pads), [False] * len(npts), device=dev)


def _vec(V, arr):
    from poms_b200.stencil import StencilVector
    return StencilVector.from_array(V, arr)


def _arr(v):
    return v.toarray().reshape(v.space.npts)


def _poisson(p, N, dev):
    from poms_b200 import bsplines as bs
    from poms_b200.stencil import KronSumMatrix
    from oracle import poms_oracle as po
    knots = [bs.make_open_knots(p, n + p) for n in N]
    A = KronSumMatrix.poisson(p, knots)
    Ao, _, _ = po.poisson_operator(p, knots)
    return A, Ao


# n = N + p basis functions per axis: (40,150,300) interior + ragged tiles in both tile axes and
# 2 chunks; (70,131,515) = C3/C5 extents, 9 column tiles, ragged last row tile; 2-D p=5 C4-shaped
SHAPES = [(3, (37, 147, 297)), (2, (68, 129, 513)), (3, (128, 35, 77)), (4, (21, 60, 200)),
          (5, (295, 8192)), (3, (300, 2048))]


@pytest.mark.parametrize("variant", [1, 0])
@pytest.mark.parametrize("p,N", SHAPES)
def test_toeplitz_operator_all_epilogues_vs_oracle(dev, p, N, variant):
    from poms_b200 import _lib
    from poms_b200.stencil import (StencilVector, DeviceContext, EPI_STORE, EPI_RESID, EPI_JACOBI,
                                   EPI_DINV, EPI_AXPY)
    if len(N) == 2 and variant == 0:
        pytest.skip("kernel variants exist for the 3-D TMA path only")
    L = _lib.lib()
    L.poms_set_matvec3d_variant(variant)
    try:
        A, Ao = _poisson(p, N, dev)
        V = _space([n + p for n in N], [p] * len(N), dev)
        if len(N) == 3:
            lo, hi = A._toeplitz()[1][4:6]
            assert hi - lo > 64, "the shape must contain Toeplitz-interior tiles"
        rng = np.random.default_rng(101)
        Xh, Bh = rng.standard_normal(V.npts), rng.standard_normal(V.npts)
        X, B = _vec(V, Xh), _vec(V, Bh)
        ctx = DeviceContext.get(dev)
        Yo = Ao.dot(Xh)
        D = Ao.diagonal()
        sc = np.abs(Yo).max()
        Y = StencilVector(V)
        A.apply(X, Y, EPI_STORE, dot_ptr=ctx.sptr(10))
        assert rel(_arr(Y), Yo) < 1e-13
        assert abs(ctx.scal[10].item() - np.vdot(Xh, Yo)) < 1e-12 * np.vdot(np.abs(Xh), np.abs(Yo))
        A.apply(X, Y, EPI_RESID, b=B, dot_ptr=ctx.sptr(11))
        Ro = Bh - Yo
        assert np.abs(_arr(Y) - Ro).max() < 1e-13 * sc
        assert abs(ctx.scal[11].item() - np.vdot(Ro, Ro)) < 1e-12 * np.vdot(Ro, Ro)
        om = 0.61
        dr = om * Ro / D
        A.apply(X, Y, EPI_JACOBI, b=B, omega=om, dot_ptr=ctx.sptr(12))
        assert np.abs(_arr(Y) - (Xh + dr)).max() < 1e-12 * np.abs(Xh + dr).max()
        assert abs(ctx.scal[12].item() - np.vdot(dr, dr)) < 1e-11 * np.vdot(dr, dr)
        A.apply(X, Y, EPI_DINV, b=B, omega=1.0)
        assert np.abs(_arr(Y) - Ro / D).max() < 1e-12 * np.abs(Ro / D).max()
        A.apply(X, Y, EPI_AXPY, b=B, omega=om)
        assert np.abs(_arr(Y) - (Bh + om * Yo)).max() < 1e-13 * (sc + np.abs(Bh).max())
        A.apply(X, Y, EPI_AXPY, b=None, omega=om)
        assert np.abs(_arr(Y) - om * Yo).max() < 1e-13 * sc
        # pad column (odd n_last) stays zero: BLAS-1 kernels run over it
        if V.ld != V.local_shape[-1]:
            assert float(Y.flat[..., V.local_shape[-1]:].abs().max()) == 0.0
        if len(N) == 3:
            # and the generic (non-TMA) kernel computes the same thing
            L.poms_set_force_generic(1)
            Y2 = StencilVector(V)
            A.apply(X, Y2, EPI_STORE)
            L.poms_set_force_generic(0)
            assert rel(_arr(Y2), Yo) < 1e-13
    finally:
        L.poms_set_force_generic(0)
        L.poms_set_matvec3d_variant(1)


@pytest.mark.parametrize("variant", [1, 0])
@pytest.mark.parametrize("p,N", [(3, (37, 147, 297)), (2, (68, 129, 513)), (3, (128, 128, 128))])
def test_smoother_factors_vs_oracle(dev, p, N, variant):
    """S1 = (x) F1, S2 = (x) F2 with q3(T) = F2 F1 (glt_poly): the single-product form of the same
    kernel with half-bandwidths q and 2q, STORE and AXPY epilogues, then the whole smoothing step
    x <- x + S2 S1 (b - A x) / theta against the oracle's Horner evaluation of q3(T) per axis."""
    from poms_b200 import _lib, bsplines as bs
    from poms_b200.stencil import (StencilVector, KronSumMatrix, EPI_STORE, EPI_AXPY, EPI_RESID)
    from oracle import poms_oracle as po
    L = _lib.lib()
    L.poms_set_matvec3d_variant(variant)
    try:
        A, Ao = _poisson(p, N, dev)
        q = max(2 * p - 1, 1)
        glt = [bs.glt_band(p, n, degree=q) for n in A.npts]
        F = [bs.poly_inverse_factors(b_, 3) for b_ in glt]
        S1 = KronSumMatrix([f[0] for f in F])
        S2 = KronSumMatrix([f[1] for f in F])
        S1o = po.KronSumOperator([tuple(f[0] for f in F)])
        S2o = po.KronSumOperator([tuple(f[1] for f in F)])
        gp = max(p, S2.P)
        V = _space(A.npts, [gp, p, p], dev)
        rng = np.random.default_rng(5)
        Xh, Bh = rng.standard_normal(V.npts), rng.standard_normal(V.npts)
        X, B = _vec(V, Xh), _vec(V, Bh)
        T1 = StencilVector(V)
        S1.apply(X, T1, EPI_STORE)
        T1o = S1o.dot(Xh)
        assert rel(_arr(T1), T1o) < 1e-13
        Y = StencilVector(V)
        S2.apply(T1, Y, EPI_AXPY, b=B, omega=0.3)
        Yo = Bh + 0.3 * S2o.dot(T1o)
        assert rel(_arr(Y), Yo) < 1e-13
        S2.apply(T1, Y, EPI_AXPY, b=None, omega=0.3)
        assert rel(_arr(Y), 0.3 * S2o.dot(T1o)) < 1e-13
        # one smoothing step, product path vs Horner form of the oracle
        theta = 1.7
        R = StencilVector(V)
        A.apply(X, R, EPI_RESID, b=B)
        S1.apply(R, T1, EPI_STORE)
        Xn = X.copy()
        S2.apply(T1, Xn, EPI_AXPY, b=Xn, omega=1.0 / theta)
        z = Bh - Ao.dot(Xh)
        for ax in range(3):
            lo, hi = po._symbol_range(glt[ax])
            c = po._cheb_inverse_poly(lo * 0.98, hi * 1.02, 3)
            y = c[-1] * z
            for cf in c[-2::-1]:
                y = po.apply_band(glt[ax], y, ax) + cf * z
            z = y
        assert rel(_arr(Xn), Xh + z / theta) < 1e-12
    finally:
        L.poms_set_matvec3d_variant(1)


@pytest.mark.parametrize("smoother,rhs", [("glt_poly", "manufactured"), ("glt_poly", "ones"),
                                          ("glt", "manufactured")])
def test_c3_mg_pcg_vs_cpu_port_same_hierarchy(dev, smoother, rhs):
    """BASELINE C3 (3-D, p=3, 128^3 elements = 2.25 M DOF), the bench's own hierarchy parameters
    (coarsest grid 32, uniform coarsening): the GPU path and the threaded CPU port must take the SAME
    number of iterations, agree on the residual history and on x to 1e-10."""
    from poms_b200.mg import Hierarchy, mg_pcg
    from poms_b200.stencil import StencilVector
    mt = pytest.importorskip("oracle.poms_oracle_mt")
    p, N, Nc = 3, [128, 128, 128], 32
    h = Hierarchy(p, N, device=dev, smoother=smoother, nu=1, Nc=Nc, coarsen="uniform")
    ho = mt.MGHierarchyMT(p, N, Nc=Nc, smoother=smoother, nu=1, coarsen="uniform")
    assert len(h.levels) == len(ho.levels) == 3
    for a, b_ in zip(h.levels[:-1], ho.levels[:-1]):
        assert abs(a.lmax - b_["lmax"]) < 1e-9 * b_["lmax"]
    V = h.levels[0].V
    Ao = ho.levels[0]["A"]
    if rhs == "ones":
        bh = np.ones(V.npts)
    else:
        i = [np.arange(n, dtype=float) for n in V.npts]
        x0 = i[0][:, None, None] + i[1][None, :, None] + i[2][None, None, :] + 1.0
        bh = Ao.dot(x0)
    b = StencilVector.from_array(V, bh)
    x, info = mg_pcg(h, b, tol=1e-10, maxiter=100)
    xo, io = ho.mg_pcg(bh, tol=1e-10, maxiter=100)
    print("C3 %s/%s: GPU %d iterations (%d restarts), CPU port %d (%d)"
          % (smoother, rhs, info["niter"], info["restarts"], io["niter"], io["restarts"]))
    assert info["niter"] == io["niter"] and info["restarts"] == io["restarts"]
    assert info["success"] and io["success"]
    assert np.allclose(info["history"], io["history"], rtol=1e-6)
    assert np.allclose(info["history"][:6], io["history"][:6], rtol=1e-10)
    assert rel(_arr(x), xo) < 1e-10


def test_unpreconditioned_cg_identity_psolve(dev):
    """pcg(A, lambda A, r: r, b) is valid with the reference (it rebinds r); here the in-place
    CG update must not alias p, s and r."""
    from poms_b200.solvers import pcg
    from oracle import poms_oracle as po
    A, Ao = _poisson(2, (12, 14), dev)
    V = _space(A.npts, (2, 2), dev)
    bh = np.random.default_rng(2).standard_normal(V.npts)
    x, info = pcg(A, lambda A_, r: r, _vec(V, bh), tol=1e-12, maxiter=400)
    xo, io = po.pcg(Ao, lambda A_, r: r, bh, tol=1e-12, maxiter=400)
    assert info["niter"] == io["niter"]
    assert rel(_arr(x), xo) < 1e-9


def test_setitem_getitem_odd_last_axis(dev):
    """spl-style x[:, :] = c / ndarray and x[:, :] on a space whose last extent is odd (pitched
    storage with a zero pad column): the pad must stay zero so that dots are right."""
    from poms_b200.stencil import StencilVector
    V = _space((6, 7), (2, 2), dev)
    x = StencilVector(V)
    x[:, :] = 1.0
    assert x.dot(x) == 42.0
    assert x[:, :].shape == (6, 7)
    a = np.arange(42, dtype=float).reshape(6, 7)
    x[:, :] = a
    assert np.array_equal(x[:, :], a)
    assert abs(x.dot(x) - float((a * a).sum())) < 1e-9
    x[2:4, 1:3] = 0.0
    a[2:4, 1:3] = 0.0
    assert np.array_equal(_arr(x), a)
    assert x[3, 5] == a[3, 5]
    assert np.array_equal(x._data[2:-2, 2:-2], a)      # spl `_data`: padded by the ghost widths
