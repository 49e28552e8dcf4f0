"""Diagnostic (not a pytest test): iteration counts and history deviations of the GPU solvers vs the
golden reference runs, to decide the tolerances of tests/test_gpu_parity.py:
    python tests/gpu_diag_counts.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from poms_b200 import bsplines as bs
from poms_b200.stencil import StencilVectorSpace, StencilVector, StencilMatrix, KronSumMatrix
from poms_b200.solvers import pcg, pcg_glt, jacobi, crl
from oracle import poms_oracle as po

dev = torch.device("cuda", 0)
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def mat1d(band):
    n, w = band.shape
    V = StencilVectorSpace([n], [(w - 1) // 2], [False], device=dev)
    M = StencilMatrix(V, V)
    M._data[...] = band
    return M


def rel(a, b):
    return np.abs(a - b).max() / np.abs(b).max()


for tag in ["p1_ne4", "p1_ne16", "p2_ne10", "p3_ne12"]:
    g = np.load(os.path.join(G, "pcg_glt_%s.npz" % tag))
    p, ne = int(g["p"]), int(g["ne"])
    T = bs.make_open_knots(p, ne + p)
    A = KronSumMatrix.poisson(p, [T, T])
    V = StencilVectorSpace(list(A.npts), [p, p], [False, False], device=dev)
    S = StencilMatrix(V, V)
    S._data[...] = g["A"]
    b = StencilVector.from_array(V, g["b"])
    _, io = po.pcg_glt(po.StencilOperator2D(g["A"]), g["M1"], g["M2"], g["b"], tol=float(g["tol"]), maxiter=100)
    ho = np.array(io["history"])
    thr = float(g["tol"]) * np.sqrt(np.vdot(g["b"], g["b"]))
    for name, op in (("stencil", S), ("kronsum", A)):
        x, info = pcg_glt(op, mat1d(g["M1"]), mat1d(g["M2"]), b, tol=float(g["tol"]), maxiter=100)
        h = np.array(info["history"])
        m = min(len(h), len(ho))
        dev_h = np.abs(h[:m] / ho[:m] - 1)
        print("pcg_glt %-8s %-8s niter gpu %d ref %d  hist maxrel %.1e (first>1e-8 at %s)  x rel %.1e  thresh %.3e  last rr gpu %s ref %s"
              % (tag, name, info["niter"], int(g["info"][0]), dev_h.max(),
                 np.argmax(dev_h > 1e-8) if (dev_h > 1e-8).any() else None,
                 rel(x.toarray().reshape(g["x"].shape), g["x"]), thr, h[-2:] ** 2, ho[-2:] ** 2))
for tag in ["p1_ne16", "p2_ne10", "p3_ne12"]:
    g = np.load(os.path.join(G, "jacobi_%s.npz" % tag))
    gd = np.load(os.path.join(G, "pcg_diag_%s.npz" % tag))
    gc = np.load(os.path.join(G, "crl_%s.npz" % tag))
    p, ne = int(g["p"]), int(g["ne"])
    T = bs.make_open_knots(p, ne + p)
    A = KronSumMatrix.poisson(p, [T, T])
    V = StencilVectorSpace(list(A.npts), [p, p], [False, False], device=dev)
    S = StencilMatrix(V, V)
    S._data[...] = g["A"]
    b = StencilVector.from_array(V, g["b"])
    Ao = po.StencilOperator2D(g["A"])
    _, io = po.pcg(Ao, po.jacobi, g["b"], tol=float(gd["tol"]), maxiter=int(gd["maxiter"]))
    ho = np.array(io["history"])
    for name, op in (("stencil", S), ("kronsum", A)):
        x, info = pcg(op, jacobi, b, tol=float(gd["tol"]), maxiter=int(gd["maxiter"]))
        h = np.array(info["history"])
        m = min(len(h), len(ho))
        dv = np.abs(h[:m] / ho[:m] - 1)
        print("pcg+jacobi %-8s %-8s niter gpu %d ref %d oracle %d hist maxrel %.1e x rel %.1e thr %.3e last rr gpu %s ref %s"
              % (tag, name, info["niter"], int(gd["info"][0]), io["niter"], dv.max(),
                 rel(x.toarray().reshape(gd["x"].shape), gd["x"]),
                 float(gd["tol"]) * np.sqrt(np.vdot(g["b"], g["b"])), h[-2:] ** 2, ho[-2:] ** 2))
        x, info = crl(op, b, tol=1e-5, maxiter=60)
        xo, ico = po.crl(Ao, g["b"], tol=1e-5, maxiter=60)
        print("crl        %-8s %-8s niter gpu %d ref %d oracle %d  x rel vs golden %.1e  res_norm gpu %.6e oracle %.6e"
              % (tag, name, info["niter"], int(gc["info"][0]), ico["niter"],
                 rel(x.toarray().reshape(gc["x"].shape), gc["x"]), info["res_norm"], ico["res_norm"]))
