"""Perf figure of the full (non-separable) 3-D stencil mat-vec at C3 size
(131^3, p = 3: 343 coefficients per row, 6.2 GB of coefficients) against the Kronecker-sum kernel on the
same operator.    python tests/gpu_ab_stencil3d.py [N=128] [p=3]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from poms_b200 import bsplines as bs
from poms_b200.stencil import StencilVectorSpace, StencilVector, StencilMatrix, KronSumMatrix, EPI_STORE

N = int(sys.argv[1]) if len(sys.argv) > 1 else 128
p = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda", 0)
knots = [bs.make_open_knots(p, N + p)] * 3
K = KronSumMatrix.poisson(p, knots)
V = StencilVectorSpace([N + p] * 3, [p] * 3, [False] * 3, device=dev)
S = StencilMatrix(V)
S._data[...] = K.to_stencil_array()
x, y1, y2 = StencilVector(V), StencilVector(V), StencilVector(V)
g = torch.Generator(device=dev).manual_seed(0)
x.data.copy_(torch.randn(V.npts, generator=g, dtype=torch.float64, device=dev))
dof = V.local_size
W = (2 * p + 1) ** 3
for name, op, y, nbytes in (("full stencil ((2p+1)^3 = %d coefficients per row)" % W, S, y1, 8 * W * dof + 16 * dof),
                            ("Kronecker sum (same operator, 16 B/DOF)", K, y2, 16 * dof)):
    for _ in range(2):
        op.apply(x, y, EPI_STORE)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        op.apply(x, y, EPI_STORE)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print("%-58s %8.3f ms  %8.1f GB/s algorithmic (%4.1f %% of 6417)  %.3e DOF/s"
          % (name, ms, nbytes / ms / 1e6, 100 * nbytes / ms / 1e6 / 6416.7, dof / ms * 1e3))
err = ((y1.data - y2.data).abs().max() / y2.data.abs().max()).item()
print("max rel diff full stencil vs Kronecker sum: %.1e  (%d^3 points, p = %d)" % (err, N + p, p))
