"""A/B timing of the 3-D Kronecker mat-vec kernels (generic vs TMA) -- run on the GPU box:
    python tests/gpu_ab_matvec.py [N] [p]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from poms_b200 import _lib, bsplines as bs
from poms_b200.stencil import (StencilVectorSpace, StencilVector, KronSumMatrix, DeviceContext,
                               EPI_STORE, EPI_RESID, EPI_JACOBI)

N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
p = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda", 0)
knots = [bs.make_open_knots(p, N + p)] * 3
A = KronSumMatrix.poisson(p, knots)
V = StencilVectorSpace([N + p] * 3, [p] * 3, [False] * 3, device=dev)
x, y, y2, b = (StencilVector(V) for _ in range(4))
g = torch.Generator(device=dev).manual_seed(0)
x.data.copy_(torch.randn(V.npts, generator=g, dtype=torch.float64, device=dev))
b.data.copy_(torch.randn(V.npts, generator=g, dtype=torch.float64, device=dev))
ctx = DeviceContext.get(dev)
L = _lib.lib()
dof = V.local_size
for epi, name, nb in ((EPI_STORE, "store+dot", 16), (EPI_RESID, "resid", 24), (EPI_JACOBI, "jacobi", 24)):
    res = {}
    for force in (1, 2):
        L.poms_set_force_generic(1 if force == 1 else 0)
        out = y if force == 1 else y2
        for _ in range(3):
            A.apply(x, out, epi, b=b, omega=0.5, dot_ptr=ctx.sptr(20 + force))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        reps = 10
        for _ in range(reps):
            A.apply(x, out, epi, b=b, omega=0.5, dot_ptr=ctx.sptr(20 + force))
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        res[force] = ms
        print("%-10s %-8s %8.3f ms  %7.1f GB/s (alg %d B/DOF)  dot=%.15e" % (
            name, {1: "generic", 2: "tma"}[force], ms, nb * dof / ms / 1e6, nb, ctx.scal[20 + force].item()))
    diff = (y.data - y2.data).abs().max().item() / y.data.abs().max().item()
    print("   max rel diff generic vs tma: %.2e   speedup %.2fx" % (diff, res[1] / res[2]))
L.poms_set_force_generic(0)
