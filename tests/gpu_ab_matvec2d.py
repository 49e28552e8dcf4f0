"""A/B timing + cross-check of the 2-D Kronecker mat-vec kernels (round-1 kernel = variant 0, the
warp-autonomous TMA kernel = variant 1) on the operators of BASELINE configs C2 (p=3, 2048^2) and C4
(p=5, 8192^2):   python tests/gpu_ab_matvec2d.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from poms_b200 import _lib, bsplines as bs
from poms_b200.stencil import (StencilVectorSpace, StencilVector, KronSumMatrix, DeviceContext,
                               EPI_STORE, EPI_RESID, EPI_JACOBI)

dev = torch.device("cuda", 0)
L = _lib.lib()
peak = 6416.7
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float64, device=dev)
for p, N in ((3, 2048), (5, 8192), (2, 4096)):
    knots = [bs.make_open_knots(p, N + p)] * 2
    A = KronSumMatrix.poisson(p, knots)
    V = StencilVectorSpace([N + p] * 2, [p, p], [False] * 2, device=dev)
    x, b = StencilVector(V), StencilVector(V)
    g = torch.Generator(device=dev).manual_seed(0)
    x.data.copy_(torch.randn(V.npts, generator=g, dtype=torch.float64, device=dev))
    b.data.copy_(torch.randn(V.npts, generator=g, dtype=torch.float64, device=dev))
    ctx = DeviceContext.get(dev)
    dof = V.local_size
    for name, epi, rhs, nb, dot in (("STORE+dot", EPI_STORE, None, 16, True), ("RESID", EPI_RESID, b, 24, False),
                                    ("JACOBI", EPI_JACOBI, b, 24, False)):
        ref = None
        for var in (0, 1):
            L.poms_set_matvec2d_variant(var)
            y = StencilVector(V)
            dp = ctx.sptr(30) if dot else None
            for _ in range(3):
                A.apply(x, y, epi, b=rhs, omega=0.37, dot_ptr=dp)
            torch.cuda.synchronize()
            times = []
            for _ in range(7):
                flush.fill_(1.0)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                A.apply(x, y, epi, b=rhs, omega=0.37, dot_ptr=dp)
                e1.record()
                torch.cuda.synchronize()
                times.append(e0.elapsed_time(e1))
            ms = float(np.median(times))
            dv = ctx.scal[30].item() if dot else float("nan")
            if ref is None:
                ref, diff = y.data.clone(), 0.0
            else:
                diff = ((y.data - ref).abs().max() / ref.abs().max()).item()
            print("2-D p=%d %5d^2 %-10s var %d  %8.3f ms  %7.1f GB/s alg (%4.1f %% of %.0f)  relerr vs var0 %.1e  dot %.15e"
                  % (p, N, name, var, ms, nb * dof / ms / 1e6, 100 * nb * dof / ms / 1e6 / peak, peak, diff, dv),
                  flush=True)
L.poms_set_matvec2d_variant(1)
