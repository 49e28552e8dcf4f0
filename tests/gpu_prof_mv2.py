"""Profiling driver for the 2-D Kronecker mat-vec (run under ncu): C4 operator (p = 5, 8192^2), one
warm-up and one measured launch of store+dot and of the fused residual.
    ncu --set full --clock-control none --import-source on -k regex:kron_matvec2d_tma --launch-skip 2 \\
        --launch-count 2 -o gpurun_out/mv2 python tests/gpu_prof_mv2.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from poms_b200 import bsplines as bs
from poms_b200.stencil import (StencilVectorSpace, StencilVector, KronSumMatrix, DeviceContext,
                               EPI_STORE, EPI_RESID)

p, N = 5, 8192
dev = torch.device("cuda", 0)
A = KronSumMatrix.poisson(p, [bs.make_open_knots(p, N + p)] * 2)
V = StencilVectorSpace([N + p] * 2, [p, p], [False] * 2, device=dev)
x, b, y = StencilVector(V), StencilVector(V), StencilVector(V)
g = torch.Generator(device=dev).manual_seed(0)
x.data.copy_(torch.randn(V.npts, generator=g, dtype=torch.float64, device=dev))
b.data.copy_(torch.randn(V.npts, generator=g, dtype=torch.float64, device=dev))
ctx = DeviceContext.get(dev)
for rep in range(2):
    A.apply(x, y, EPI_STORE, dot_ptr=ctx.sptr(30))
    A.apply(x, y, EPI_RESID, b=b)
    torch.cuda.synchronize()
print("done")
