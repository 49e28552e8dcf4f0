"""Profiling driver for the 3-D Kronecker mat-vec (run under ncu): one warm-up and ONE measured launch
of each fine-level pass of the C5 bench (operator STORE+dot / RESID, smoother factors S1, S2).
    ncu --set full --clock-control none --import-source on -k regex:kron_matvec3d --launch-skip 5 \\
        --launch-count 5 -o gpurun_out/mv3 python tests/gpu_prof_mv3.py [N=512] [variant=1]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from poms_b200 import _lib, bsplines as bs
from poms_b200.stencil import (StencilVectorSpace, StencilVector, KronSumMatrix, DeviceContext,
                               EPI_STORE, EPI_RESID, EPI_AXPY)

N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
var = int(sys.argv[2]) if len(sys.argv) > 2 else 1
p = 3
dev = torch.device("cuda", 0)
knots = [bs.make_open_knots(p, N + p)] * 3
A = KronSumMatrix.poisson(p, knots)
glt = [bs.glt_band(p, n, degree=max(2 * p - 1, 1)) for n in A.npts]
F = [bs.poly_inverse_factors(b_, 3) for b_ in glt]
S1 = KronSumMatrix([f[0] for f in F])
S2 = KronSumMatrix([f[1] for f in F])
V = StencilVectorSpace([N + p] * 3, [max(p, S2.P), p, p], [False] * 3, device=dev)
x, b, y = StencilVector(V), StencilVector(V), StencilVector(V)
g = torch.Generator(device=dev).manual_seed(0)
x.data.copy_(torch.randn(V.npts, generator=g, dtype=torch.float64, device=dev))
b.data.copy_(torch.randn(V.npts, generator=g, dtype=torch.float64, device=dev))
ctx = DeviceContext.get(dev)
_lib.lib().poms_set_matvec3d_variant(var)
cases = [(A, EPI_STORE, None, True), (A, EPI_RESID, b, False), (S1, EPI_STORE, None, False),
         (S2, EPI_AXPY, None, False), (S2, EPI_AXPY, b, False)]
for rep in range(2):
    for op, epi, rhs, dot in cases:
        op.apply(x, y, epi, b=rhs, omega=0.37, dot_ptr=ctx.sptr(30) if dot else None)
    torch.cuda.synchronize()
print("done")
