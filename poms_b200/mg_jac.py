"""Function form of the script /root/reference/sources/mg_jac.py (argv: p nf; nc = 8 hard-coded
there, a parameter here): one two-grid cycle with PCG-Jacobi pre- and post-smoothing."""
import numpy as np

from . import bsplines as bs
from .multilevels import knots_to_insert
from .stencil import StencilVectorSpace, StencilVector, KronSumMatrix
from .mg import Transfer, CoarseSolver, two_grid, fine_knots

__all__ = ["mg_jac", "setup_two_grid"]


def setup_two_grid(p, nf, nc, device="cuda", ndim=2):
    """Lines 27-81 of the script: knot vectors, fine space (sorted union mesh), operator,
    transfer, coarse operator.  `nf`/`nc` are numbers of basis functions as in the script."""
    Tc = bs.make_open_knots(p, nc)
    Tf = bs.make_open_knots(p, nf)
    Ts = knots_to_insert(Tf, nf, p, Tc, nc, p)
    T = fine_knots(Tc, Ts)
    n = len(T) - p - 1
    V = StencilVectorSpace([n] * ndim, [p] * ndim, [False] * ndim, device=device)
    Vc = StencilVectorSpace([nc] * ndim, [p] * ndim, [False] * ndim, device=device)
    Af = KronSumMatrix.poisson(p, [T] * ndim)        # = assembly(S), matrix_assembler.py:84
    Ac = KronSumMatrix.poisson(p, [Tc] * ndim)       # = R*Af*P for nested spaces (line 81)
    transfer = Transfer([Tc] * ndim, [T] * ndim, p, V.device)
    coarse = CoarseSolver(Ac, V.device)
    return dict(Tc=Tc, Tf=Tf, Ts=Ts, T=T, V=V, Vc=Vc, Af=Af, Ac=Ac, transfer=transfer,
                coarse=coarse)


def mg_jac(p, nf, nc=8, device="cuda", ndim=2, verbose=False):
    """Returns the dict of `two_grid` (x_pre, info_pre, r_f, r_c, x_c, x_corr, x_post, info_post)
    plus the setup; b = 1 on every DOF (script lines 57-62)."""
    s = setup_two_grid(p, nf, nc, device, ndim)
    bf = StencilVector(s["V"])
    bf.data.fill_(1.0)
    out = two_grid(s["Af"], s["transfer"], s["coarse"], bf, s["Vc"], post="jac")
    out.update(s)
    if verbose:
        print("rank= ", 0, (p, nc, s["V"].npts[0]))
        print("PRES: ", {k: v for k, v in out["info_pre"].items() if k != "history"})
        print("POST: ", {k: v for k, v in out["info_post"].items() if k != "history"})
    return out
