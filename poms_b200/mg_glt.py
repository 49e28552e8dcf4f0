"""Function form of the script /root/reference/sources/mg_glt.py (nc = 64 hard-coded there): the
two-grid cycle of mg_jac with the GLT post-smoother pcg_glt(Af, M1, M2, bf, x0=xf, maxiter=p+1)
(lines 113-123)."""
from . import bsplines as bs
from .stencil import StencilVector
from .utils import array_to_mat_stencil
from .mg import two_grid
from .mg_jac import setup_two_grid

__all__ = ["mg_glt", "collocation_cardinal_splines"]


def collocation_cardinal_splines(p, n):
    """Dense n x n symmetric Toeplitz collocation matrix of the degree-p cardinal B-spline.
    The spl function of this name is absent and UNPINNED; M1, M2 are inputs of
    pcg_glt, so callers may pass any banded SPD matrices."""
    return bs.band_to_dense(bs.glt_band(p, n))


def mg_glt(p, nf, nc=64, device="cuda", ndim=2, M1=None, M2=None, verbose=False):
    s = setup_two_grid(p, nf, nc, device, ndim)
    n1, n2 = s["V"].npts[:2]
    if M1 is None:
        M1 = array_to_mat_stencil(n1, p, collocation_cardinal_splines(p, n1))
    if M2 is None:
        M2 = array_to_mat_stencil(n2, p, collocation_cardinal_splines(p, n2))
    bf = StencilVector(s["V"])
    bf.data.fill_(1.0)
    out = two_grid(s["Af"], s["transfer"], s["coarse"], bf, s["Vc"], post="glt", M1=M1, M2=M2,
                   p=p)
    out.update(s)
    out.update(M1=M1, M2=M2)
    if verbose:
        print("rank= ", 0, (p, nc, n1))
        print("PRES: ", {k: v for k, v in out["info_pre"].items() if k != "history"})
        print("POST: ", {k: v for k, v in out["info_post"].items() if k != "history"})
    return out
