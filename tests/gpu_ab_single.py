"""Timing of the single-term Kronecker pass (smoother factors) at 512^3: python tests/gpu_ab_single.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from poms_b200 import _lib, bsplines as bs
from poms_b200.stencil import (StencilVectorSpace, StencilVector, KronSumMatrix, EPI_STORE, EPI_AXPY, EPI_RESID)
N, p = 512, 3
dev = torch.device("cuda", 0)
n = N + p
V = StencilVectorSpace([n] * 3, [p] * 3, [False] * 3, device=dev)
x, y, b = StencilVector(V), StencilVector(V), StencilVector(V)
x.data.normal_(); b.data.normal_()
F1, F2 = bs.poly_inverse_factors(bs.glt_band(p, n, degree=5), 3)
ops = {"S1 (P=2) store": (KronSumMatrix([F1] * 3), EPI_STORE, 16), "S2 (P=4) axpy": (KronSumMatrix([F2] * 3), EPI_AXPY, 24),
       "A (P=3 sum) resid": (KronSumMatrix.poisson(p, [bs.make_open_knots(p, n)] * 3), EPI_RESID, 24)}
for name, (op, epi, nb) in ops.items():
    for _ in range(3):
        op.apply(x, y, epi, b=b, omega=0.3)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        op.apply(x, y, epi, b=b, omega=0.3)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print("%-20s %.3f ms  %.0f GB/s algorithmic (%d B/DOF)" % (name, ms, nb * V.local_size / ms / 1e6, nb))
