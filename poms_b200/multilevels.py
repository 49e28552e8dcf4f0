"""Drop-in for /root/reference/sources/multilevels.py."""
import numpy as np

__all__ = ["knots_to_insert"]


def knots_to_insert(Tf, nf, pf, Tc, nc, pc):
    """Interior fine knots Tf[pf+1 .. nf-1] that are not bit-equal to any interior coarse knot
    Tc[pc+1 .. nc] (/root/reference/sources/multilevels.py:7-33).  The reference scans the
    sorted coarse knots and keeps t iff it never hit `t == Tc[j]` (line 22-29); on sorted
    input that is exactly set membership with exact float comparison."""
    Tf = np.asarray(Tf, dtype=float)
    Tc = np.asarray(Tc, dtype=float)
    t = Tf[pf + 1:nf]
    return t[~np.isin(t, Tc[pc + 1:nc + 1])].copy()
