"""ORACLE / TEST INFRASTRUCTURE ONLY.  Stand-in for `spl.fem.splines.SplineSpace`
(attributes read by sources/matrix_assembler.py:18-21,93-97).

basis[il, d, g, ie] = d-th derivative of the il-th non-zero basis function of element ie at
Gauss-Legendre point g; weights[g, ie] the quadrature weights scaled to the element.  The
rule has p+1 points (exact for the degree-2p integrands), so the assembled entries are the
exact integrals up to rounding whatever rule spl really used.
"""
import numpy as np
from scipy.interpolate import BSpline
from spl.linalg.stencil import StencilVectorSpace
from spl.core.interface import compute_spans


class SplineSpace:
    def __init__(self, degree, knots=None, grid=None):
        p = int(degree)
        if knots is None:
            grid = np.asarray(grid, dtype=float)
            knots = np.concatenate([[grid[0]] * p, grid, [grid[-1]] * p])
        T = np.asarray(knots, dtype=float)
        n = len(T) - p - 1
        self.degree = p
        self.knots = T
        self.nbasis = n
        self.vector_space = StencilVectorSpace([n], [p], [False])
        self.quad_order = p + 1
        self.spans = compute_spans(p, n, T)
        breaks = np.unique(T)
        ne = len(breaks) - 1
        self.ncells = ne
        u, w = np.polynomial.legendre.leggauss(self.quad_order)
        k = self.quad_order
        self.points = np.zeros((k, ne))
        self.weights = np.zeros((k, ne))
        self.basis = np.zeros((p + 1, 2, k, ne))
        full = BSpline(T, np.eye(n), p)
        dfull = full.derivative()
        for ie in range(ne):
            a, b = breaks[ie], breaks[ie + 1]
            x = 0.5 * (a + b) + 0.5 * (b - a) * u
            self.points[:, ie] = x
            self.weights[:, ie] = 0.5 * (b - a) * w
            first = self.spans[ie] - p - 1
            v = full(x)[:, first:first + p + 1]
            dv = dfull(x)[:, first:first + p + 1]
            self.basis[:, 0, :, ie] = v.T
            self.basis[:, 1, :, ie] = dv.T
