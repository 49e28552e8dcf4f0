#!/usr/bin/env python
"""ORACLE / TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz.

Runs the UNMODIFIED reference modules (/root/reference/sources/{kron_product,solvers,
multilevels,utils,matrix_assembler}.py and /root/reference/pyccel/pyccel_functions.py) over
the serial `spl`/`mpi4py` stand-in in oracle/shim/ and records inputs and outputs of every
hot-path function on the reference's own test fixtures (section 8c) plus
seeded random cases.  The reference cannot travel to the GPU box, so the vectors are
committed; this script is the provenance.  Run in the dev container only:

    python oracle/gen_golden.py          # writes tests/golden/*.npz

`dots` arrays are the complete sequence of StencilVector.dot results a solver made (the
shim logs them): they pin the residual history AND the op order of the reference drivers.
"""
import os
import sys
import io
import contextlib
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("POMS_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(HERE, "shim"))
sys.path.insert(0, os.path.join(REF, "sources"))
OUT = os.path.join(HERE, "..", "tests", "golden")

from spl.linalg import stencil as _st  # noqa: E402
from spl.linalg.stencil import StencilVectorSpace, StencilVector, StencilMatrix  # noqa: E402
from spl.fem.splines import SplineSpace  # noqa: E402
from spl.fem.tensor import TensorFemSpace  # noqa: E402
from spl.core.interface import (  # noqa: E402
    make_open_knots, matrix_multi_stages, collocation_cardinal_splines)
import utils  # noqa: E402
import kron_product  # noqa: E402
import solvers  # noqa: E402
import multilevels  # noqa: E402
import matrix_assembler  # noqa: E402
from scipy.linalg.lapack import dgbtrf  # noqa: E402
from scipy.sparse import kron, csc_matrix  # noqa: E402
from scipy.sparse.linalg import splu  # noqa: E402

# --- dot logging hook (monkey-patch of OUR shim, not of the reference) -------------------
_LOG = []
_orig_dot = StencilVector.dot


def _logged_dot(self, other):
    v = _orig_dot(self, other)
    _LOG.append(v)
    return v


StencilVector.dot = _logged_dot


def save(name, **kw):
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **kw)
    print("wrote", name, {k: np.shape(v) for k, v in kw.items()})


def mat1d(n, p, band=None):
    V = StencilVectorSpace([n], [p], [False])
    A = StencilMatrix(V, V)
    if band is not None:
        A._data[...] = band
        A.remove_spurious_entries()
    return A


def vec(npts, pads, arr=None):
    V = StencilVectorSpace(list(npts), list(pads), [False] * len(npts))
    X = StencilVector(V)
    if arr is not None:
        X._data[tuple(slice(p, -p) for p in pads)] = arr
    return X


def interior(X):
    return X._data[tuple(slice(p, -p) for p in X.space.pads)].copy()


def to_bnd(A_dense):
    # restated from sources/tests/test_kron_solve_bnd.py:30-42 (sources/kron_product.py:175-187
    # raises NameError: dia_matrix is never imported there).
    from scipy.sparse import dia_matrix
    dmat = dia_matrix(A_dense)
    la = abs(dmat.offsets.min())
    ua = dmat.offsets.max()
    cmat = dmat.tocsr()
    A_bnd = np.zeros((1 + ua + 2 * la, cmat.shape[1]))
    for i, j in zip(*cmat.nonzero()):
        A_bnd[la + ua + i - j, j] = cmat[i, j]
    return A_bnd, int(la), int(ua)


# ------------------------------------------------------------------------------------------
def gen_kron_dot():
    # fixture of sources/tests/test_kron_dot.py:125-126 with utils.populate_* (utils.py:7-39)
    n1, n2, p1, p2 = 8, 4, 2, 1
    A = mat1d(n1, p1)
    B = mat1d(n2, p2)
    utils.populate_1d_matrix(A, 5.0)
    utils.populate_1d_matrix(B, 6.0)
    X = vec((n1, n2), (p1, p2))
    utils.populate_2d_vector(X)
    Y2 = kron_product.kron_dot_v2(A, B, X)
    Y1 = kron_product.kron_dot_v1(A, B, X)
    Yr = utils.kron_dot_ref(A, B, X)
    save("kron_dot_fixture", A=A._data, B=B._data, X=interior(X), Y_v2=interior(Y2),
         Y_v1=interior(Y1), Y_ref=Yr.reshape(n1, n2))
    # seeded random, non-symmetric bands, odd sizes
    rng = np.random.default_rng(20181)
    n1, n2, p1, p2 = 13, 9, 3, 2
    A = mat1d(n1, p1, rng.standard_normal((n1, 2 * p1 + 1)))
    B = mat1d(n2, p2, rng.standard_normal((n2, 2 * p2 + 1)))
    X = vec((n1, n2), (p1, p2), rng.standard_normal((n1, n2)))
    Y2 = kron_product.kron_dot_v2(A, B, X)
    Yr = utils.kron_dot_ref(A, B, X)
    save("kron_dot_random", A=A._data, B=B._data, X=interior(X), Y_v2=interior(Y2),
         Y_ref=Yr.reshape(n1, n2))


def gen_kron_solve():
    # sources/tests/test_kron_solve.py:132-133 fixture
    n1, n2, p1, p2 = 4, 4, 1, 1
    A = mat1d(n1, p1)
    B = mat1d(n2, p2)
    utils.populate_1d_matrix(A, 5.0)
    utils.populate_1d_matrix(B, 6.0)
    Y = vec((n1, n2), (p1, p2))
    utils.populate_2d_vector(Y)
    Xs = kron_product.kron_solve_serial(A, B, Y)
    Xp = kron_product.kron_solve_par(A, B, Y)
    Xr = utils.kron_solve_ref(A, B, Y)
    save("kron_solve_fixture", A=A._data, B=B._data, Y=interior(Y), X_serial=interior(Xs),
         X_par=interior(Xp), X_ref=Xr.reshape(n1, n2))
    rng = np.random.default_rng(20182)
    n1, n2, p1, p2 = 12, 7, 2, 3
    a = rng.standard_normal((n1, 2 * p1 + 1))
    a[:, p1] += 6.0
    b = rng.standard_normal((n2, 2 * p2 + 1))
    b[:, p2] += 8.0
    A = mat1d(n1, p1, a)
    B = mat1d(n2, p2, b)
    Y = vec((n1, n2), (p1, p2), rng.standard_normal((n1, n2)))
    Xs = kron_product.kron_solve_serial(A, B, Y)
    Xp = kron_product.kron_solve_par(A, B, Y)
    Xr = utils.kron_solve_ref(A, B, Y)
    save("kron_solve_random", A=A._data, B=B._data, Y=interior(Y), X_serial=interior(Xs),
         X_par=interior(Xp), X_ref=Xr.reshape(n1, n2))


def _tri(n, p, lo, d, up):
    A = mat1d(n, p)
    A[:, -p:0] = lo
    A[:, 0:1] = d
    A[:, 1:p + 1] = up
    A.remove_spurious_entries()
    return A


def gen_kron_solve_bnd():
    # sources/tests/test_kron_solve_bnd.py:71-93,121-124 (symmetric) and the non-symmetric
    # variants of pyccel/test_kron_solve.py:141-152
    for name, n1, n2, p1, p2, up1, up2 in (("sym64", 64, 64, 2, 2, -4, -1),
                                             ("nonsym10", 10, 10, 1, 1, -2, -2),
                                             ("nonsym_rect", 17, 9, 3, 2, -2, -2)):
        A1 = _tri(n1, p1, -4, 10 * p1, up1)
        A2 = _tri(n2, p2, -1, 2 * p2, up2)
        b1, la1, ua1 = to_bnd(A1.toarray())
        b2, la2, ua2 = to_bnd(A2.toarray())
        f1, piv1, _ = dgbtrf(b1, la1, ua1)
        f2, piv2, _ = dgbtrf(b2, la2, ua2)
        Yg = np.array([[(i1 + 1) * 10.0 + (i2 + 1) for i2 in range(n2)] for i1 in range(n1)])
        Y = vec((n1, n2), (p1, p2), Yg)
        X, _ = kron_product.kron_solve_bnd_par([f1, la1, ua1, piv1], [f2, la2, ua2, piv2], Y)
        C = csc_matrix(kron(A1.tocsr(), A2.tocsr()))
        Xr = splu(C).solve(Yg.flatten()).reshape(n1, n2)
        save("kron_solve_bnd_" + name, A1=A1._data, A2=A2._data, A1_bnd=b1, A2_bnd=b2,
             A1_lu=f1, A2_lu=f2, piv1=piv1, piv2=piv2, lu=np.array([la1, ua1, la2, ua2]),
             Y=Yg, X=interior(X), X_splu=Xr)


def gen_kron_solve_bnd_pivot():
    # random non-dominant bands so that dgbtrf really interchanges rows (the fixtures above
    # never do): pins the ipiv handling of dgbtrs (sources/kron_product.py:226,232).
    rng = np.random.default_rng(20183)
    n1, n2, p1, p2 = 15, 11, 2, 3
    A1 = mat1d(n1, p1, rng.standard_normal((n1, 2 * p1 + 1)))
    A2 = mat1d(n2, p2, rng.standard_normal((n2, 2 * p2 + 1)))
    b1, la1, ua1 = to_bnd(A1.toarray())
    b2, la2, ua2 = to_bnd(A2.toarray())
    f1, piv1, _ = dgbtrf(b1, la1, ua1)
    f2, piv2, _ = dgbtrf(b2, la2, ua2)
    Yg = rng.standard_normal((n1, n2))
    Y = vec((n1, n2), (p1, p2), Yg)
    X, _ = kron_product.kron_solve_bnd_par([f1, la1, ua1, piv1], [f2, la2, ua2, piv2], Y)
    C = csc_matrix(kron(A1.tocsr(), A2.tocsr()))
    Xr = splu(C).solve(Yg.flatten()).reshape(n1, n2)
    assert (piv1 != np.arange(n1)).any() and (piv2 != np.arange(n2)).any()
    save("kron_solve_bnd_pivot", A1=A1._data, A2=A2._data, A1_bnd=b1, A2_bnd=b2,
         A1_lu=f1, A2_lu=f2, piv1=piv1, piv2=piv2, lu=np.array([la1, ua1, la2, ua2]),
         Y=Yg, X=interior(X), X_splu=Xr)


def gen_pyccel_bnd():
    # pyccel/pyccel_functions.py:114-168 (2-D) and 174-248 (3-D): raw padded F-order arrays.
    sys.path.insert(0, os.path.join(REF, "pyccel"))
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "pyccel_functions", os.path.join(REF, "pyccel", "pyccel_functions.py"))
    pf = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(pf)
    from mpi4py import MPI

    def run3d(n, p, bands, tag):
        n1, n2, n3 = n
        p1, p2, p3 = p
        mats = [_tri(nn, pp, lo, d, up).toarray() for nn, pp, (lo, d, up) in zip(n, p, bands)]
        bnds = [to_bnd(m) for m in mats]
        Yg = np.array([[[(i1 + 1) * 100.0 + (i2 + 1) * 10 + (i3 + 1) for i3 in range(n3)]
                        for i2 in range(n2)] for i1 in range(n1)])
        Y = vec(n, p, Yg)
        X = vec(n, p)
        Xd = X._data.copy(order="F")
        Yd = Y._data.copy(order="F")
        sub = np.array([MPI.Comm(), MPI.Comm(), MPI.Comm()])
        args = []
        for b, la, ua in bnds:
            args += [b.copy(order="F"), la, ua]
        out = pf.kron_solve_par_bnd_pyccel_3d(
            *args, Xd, Yd, np.array(n), np.array(p), np.array([0, 0, 0]),
            np.array([n1 - 1, n2 - 1, n3 - 1]), sub,
            np.array([n1]), np.array([0]), np.array([n2]), np.array([0]),
            np.array([n3]), np.array([0]))
        Xi = out[p1:-p1, p2:-p2, p3:-p3]
        from scipy.linalg import solve
        Xr = solve(np.kron(np.kron(mats[0], mats[1]), mats[2]), Yg.flatten()).reshape(n)
        save("kron_solve_bnd3d_" + tag, A1=mats[0], A2=mats[1], A3=mats[2],
             A1_bnd=bnds[0][0], A2_bnd=bnds[1][0], A3_bnd=bnds[2][0],
             lu=np.array([[b[1], b[2]] for b in bnds]), Y=Yg, X=np.ascontiguousarray(Xi),
             X_dense=Xr)

    # pyccel/test_kron_solve.py:241-251,293-294 (n=10, p=1; A3 = -2 / 3*p2 / -2)
    run3d((10, 10, 10), (1, 1, 1), ((-4, 10, -4), (-1, 2, -1), (-2, 3, -2)), "fixture")
    run3d((7, 6, 5), (2, 1, 2), ((-4, 20, -2), (-1, 2, -2), (-2, 9, -1)), "nonsym")

    def run2d(n, p, bands, tag):
        n1, n2 = n
        p1, p2 = p
        mats = [_tri(nn, pp, lo, d, up).toarray() for nn, pp, (lo, d, up) in zip(n, p, bands)]
        bnds = [to_bnd(m) for m in mats]
        Yg = np.array([[(i1 + 1) * 10.0 + (i2 + 1) for i2 in range(n2)] for i1 in range(n1)])
        Y = vec(n, p, Yg)
        X = vec(n, p)
        Xd = X._data.copy(order="F")
        Yd = Y._data.copy(order="F")
        sub = np.array([MPI.Comm(), MPI.Comm()])
        args = []
        for b, la, ua in bnds:
            args += [b.copy(order="F"), la, ua]
        out = pf.kron_solve_par_bnd_pyccel_2d(
            *args, Xd, Yd, np.array(n), np.array(p), np.array([0, 0]),
            np.array([n1 - 1, n2 - 1]), sub, np.array([n1]), np.array([0]),
            np.array([n2]), np.array([0]))
        save("kron_solve_bnd2d_" + tag, A1=mats[0], A2=mats[1], A1_bnd=bnds[0][0],
             A2_bnd=bnds[1][0], lu=np.array([[b[1], b[2]] for b in bnds]), Y=Yg,
             X=np.ascontiguousarray(out[p1:-p1, p2:-p2]))

    # pyccel/test_kron_solve.py:141-152 (non-symmetric) at n=10, p=1 (293-294)
    run2d((10, 10), (1, 1), ((-4, 10, -2), (-1, 2, -2)), "nonsym")


def fem_problem(p, ne, knots=None):
    if knots is None:
        grid = np.linspace(0.0, 1.0, ne + 1)
        S1 = SplineSpace(p, grid=grid)
        S2 = SplineSpace(p, grid=grid)
    else:
        S1 = SplineSpace(p, knots=knots)
        S2 = SplineSpace(p, knots=knots)
    S = TensorFemSpace(S1, S2)
    A = matrix_assembler.assembly_2d(S)  # sources/matrix_assembler.py:84 (the intended `assembly`)
    return S, A


def ramp(V):
    # x0[i1,i2] = i1 + i2 + 1 (sources/tests/test_pcg.py:52-56)
    x0 = StencilVector(V)
    (s1, s2), (e1, e2) = V.starts, V.ends
    for i1 in range(s1, e1 + 1):
        for i2 in range(s2, e2 + 1):
            x0[i1, i2] = i1 + i2 + 1.0
    return x0


def logged(fn, *a, **k):
    del _LOG[:]
    with contextlib.redirect_stdout(io.StringIO()):
        out = fn(*a, **k)
    return out, np.array(_LOG)


def info_arr(info):
    return np.array([info["niter"], float(info["success"]), info["res_norm"]])


def gen_solvers():
    # --- sources/tests/test_pcg.py:23-24,52-64: p=1, 16x16 elements, b=A*x0, tol 1e-8
    for tag, p, ne, tol, maxiter in (("p1_ne16", 1, 16, 1e-8, 1000), ("p3_ne12", 3, 12, 1e-8, 40),
                                      ("p2_ne10", 2, 10, 1e-6, 100)):
        S, A = fem_problem(p, ne)
        V = S.vector_space
        x0 = ramp(V)
        b = A.dot(x0)
        (x, info), dots = logged(solvers.pcg, A, solvers.damped_jacobi, b, tol=tol, maxiter=maxiter)
        save("pcg_jacobi_" + tag, A=A._data, p=p, ne=ne, b=interior(b), x_true=interior(x0),
             x=interior(x), info=info_arr(info), dots=dots, tol=tol, maxiter=maxiter)
        # psolve=jacobi (sources/solvers.py:139-163)
        (x, info), dots = logged(solvers.pcg, A, solvers.jacobi, b, tol=tol, maxiter=maxiter)
        save("pcg_diag_" + tag, A=A._data, p=p, ne=ne, b=interior(b), x=interior(x),
             info=info_arr(info), dots=dots, tol=tol, maxiter=maxiter)
        # standalone smoothers
        xj = solvers.jacobi(A, b)
        xd, dots = logged(solvers.damped_jacobi, A, b)
        xd2, dots2 = logged(solvers.damped_jacobi, A, b, x0=xj, tol=1e-3, maxiter=25)
        save("jacobi_" + tag, A=A._data, p=p, ne=ne, b=interior(b), x_jacobi=interior(xj),
             x_damped=interior(xd), dots_damped=dots, x_damped2=interior(xd2),
             dots_damped2=dots2)
        # crl (sources/solvers.py:3-65)
        (x, info), dots = logged(solvers.crl, A, b, tol=1e-5, maxiter=60)
        save("crl_" + tag, A=A._data, p=p, ne=ne, b=interior(b), x=interior(x),
             info=info_arr(info), dots=dots)
        # pcg_glt (sources/tests/test_glt.py:63-70 pattern)
        n1, n2 = V.npts
        M1 = utils.array_to_mat_stencil(n1, p, collocation_cardinal_splines(p, n1))
        M2 = utils.array_to_mat_stencil(n2, p, collocation_cardinal_splines(p, n2))
        (x, info), dots = logged(solvers.pcg_glt, A, M1, M2, b, tol=tol, maxiter=100)
        save("pcg_glt_" + tag, A=A._data, p=p, ne=ne, M1=M1._data, M2=M2._data, b=interior(b),
             x=interior(x), info=info_arr(info), dots=dots, tol=tol)
    # --- sources/tests/test_glt.py:25-26: p=1, 4x4
    S, A = fem_problem(1, 4)
    V = S.vector_space
    n1, n2 = V.npts
    x0 = ramp(V)
    b = A.dot(x0)
    M1 = utils.array_to_mat_stencil(n1, 1, collocation_cardinal_splines(1, n1))
    M2 = utils.array_to_mat_stencil(n2, 1, collocation_cardinal_splines(1, n2))
    (x, info), dots = logged(solvers.pcg_glt, A, M1, M2, b, tol=1e-8, maxiter=100)
    save("pcg_glt_p1_ne4", A=A._data, p=1, ne=4, M1=M1._data, M2=M2._data, b=interior(b),
         x=interior(x), info=info_arr(info), dots=dots, tol=1e-8)


def two_grid(p, nf, nc, post):
    """sources/mg_jac.py:27-119 / sources/mg_glt.py:26-123 restated as a function over the
    reference's own pieces (the scripts import a non-existent `assembly`; assembly_2d is the
    function they mean).  nc is a parameter here (scripts hard-code 8 / 64)."""
    Tc = make_open_knots(p, nc)
    Tf = make_open_knots(p, nf)
    Ts = multilevels.knots_to_insert(Tf, nf, p, Tc, nc, p)
    T = list(Ts)
    for i in range(len(Tc)):
        T.insert(i, Tc[i])
    T.sort()
    S, Af = fem_problem(p, None, knots=T)
    V = S.vector_space
    n1, n2 = V.npts
    bf = StencilVector(V)
    bf[0:n1, 0:n2] = 1.0
    P1 = matrix_multi_stages(Ts, nc, p, Tc)
    R1 = P1.transpose()
    R = kron(R1, R1)
    P = kron(P1, P1)
    Ac = R * Af.tocoo() * P
    (xf, info_pre), dots_pre = logged(solvers.pcg, Af, solvers.damped_jacobi, bf,
                                      tol=1e-6, maxiter=10)
    rf = bf - Af.dot(xf)
    rc = R.dot(rf.toarray())
    xc = splu(csc_matrix(Ac)).solve(rc)
    xc_p = P.dot(xc)
    corr = utils.array_to_vect_stencil(V, xc_p.reshape(n1, n2))
    xf1 = xf + corr
    if post == "jac":
        (xf2, info_pos), dots_pos = logged(solvers.pcg, Af, solvers.damped_jacobi, bf, x0=xf1,
                                           tol=1e-6, maxiter=10)
        extra = {}
    else:
        M1 = utils.array_to_mat_stencil(n1, p, collocation_cardinal_splines(p, n1))
        M2 = utils.array_to_mat_stencil(n2, p, collocation_cardinal_splines(p, n2))
        (xf2, info_pos), dots_pos = logged(solvers.pcg_glt, Af, M1, M2, bf, x0=xf1,
                                           tol=1e-6, maxiter=p + 1)
        extra = dict(M1=M1._data, M2=M2._data)
    return dict(p=p, nf=nf, nc=nc, Tc=Tc, Tf=Tf, Ts=Ts, T=np.array(T), n=np.array([n1, n2]),
                A=Af._data, P1=P1, Ac=np.asarray(Ac.todense()), x_pre=interior(xf),
                info_pre=info_arr(info_pre), dots_pre=dots_pre, r_f=interior(rf),
                r_c=rc, x_c=xc, x_corr=interior(xf1), x_post=interior(xf2),
                info_post=info_arr(info_pos), dots_post=dots_pos, **extra)


def gen_two_grid():
    # nc=8 as in sources/mg_jac.py:25; nf=8+... nested (5 | 20 elements) and a non-nested union mesh
    save("mg_jac_p3_nc8_nf23", **two_grid(3, 23, 8, "jac"))
    save("mg_jac_p2_nc8_nf14", **two_grid(2, 14, 8, "jac"))   # 6 vs 12 elements: nested
    save("mg_jac_p3_nc8_nf16", **two_grid(3, 16, 8, "jac"))   # 5 vs 13 elements: union mesh
    save("mg_glt_p3_nc8_nf23", **two_grid(3, 23, 8, "glt"))
    save("mg_glt_p2_nc10_nf18", **two_grid(2, 18, 10, "glt"))
    if "c1" in sys.argv:
        # BASELINE config 1: p=3, 64x64 elements (n=67), coarse 8x8 elements (nc=11)
        save("mg_jac_c1_p3_nc11_nf67", **two_grid(3, 67, 11, "jac"))


def gen_knots():
    cases = []
    for (pf, nf, pc, nc) in ((3, 12, 3, 8), (3, 23, 3, 8), (3, 16, 3, 8), (2, 14, 2, 8),
                             (1, 10, 1, 8), (3, 67, 3, 35), (5, 21, 5, 13), (3, 515, 3, 11)):
        Tc = make_open_knots(pc, nc)
        Tf = make_open_knots(pf, nf)
        ts = multilevels.knots_to_insert(Tf, nf, pf, Tc, nc, pc)
        P1 = matrix_multi_stages(ts, nc, pc, Tc) if nf <= 70 else np.zeros((0, 0))
        cases.append((pf, nf, pc, nc, Tc, Tf, ts, P1))
    kw = {}
    for i, (pf, nf, pc, nc, Tc, Tf, ts, P1) in enumerate(cases):
        kw["case%d_params" % i] = np.array([pf, nf, pc, nc])
        kw["case%d_Tc" % i] = Tc
        kw["case%d_Tf" % i] = Tf
        kw["case%d_ts" % i] = ts
        kw["case%d_P1" % i] = P1
    kw["ncases"] = len(cases)
    save("knots_to_insert", **kw)


def gen_assembly_1d():
    # sources/matrix_assembler.py:10-77 returns only the mass matrix (defect, Appendix B);
    # the stiffness is recovered from assembly_2d = K(x)M + M(x)K + M(x)M in the tests.
    kw = {}
    for p, ne in ((1, 6), (2, 7), (3, 9), (5, 8)):
        S1 = SplineSpace(p, grid=np.linspace(0.0, 1.0, ne + 1))
        kw["mass_p%d_ne%d" % (p, ne)] = matrix_assembler.assembly_1d(S1)._data
    save("assembly_1d", **kw)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    which = [a for a in sys.argv[1:] if a != "c1"] or ["dot", "solve", "bnd", "pyccel", "knots", "asm", "solvers", "mg"]
    if "dot" in which:
        gen_kron_dot()
    if "solve" in which:
        gen_kron_solve()
    if "bnd" in which:
        gen_kron_solve_bnd()
        gen_kron_solve_bnd_pivot()
    if "pyccel" in which:
        gen_pyccel_bnd()
    if "knots" in which:
        gen_knots()
    if "asm" in which:
        gen_assembly_1d()
    if "solvers" in which:
        gen_solvers()
    if "mg" in which:
        gen_two_grid()
