"""A/B timing + cross-check of the 3-D Kronecker mat-vec kernel variants on the operators the C5
bench applies on its fine level (Toeplitz bands: constant-bank path, interior tiles):
    python tests/gpu_ab_mv3_variants.py [N=512] [variants=0,1]
Rows: operator (p=3, 3 terms) STORE+dot / RESID, smoother factor S1 (single product, q=2) STORE,
smoother factor S2 (2q=4) AXPY with and without rhs.  Every variant's result is compared with
variant 0 (the round-1 kernel) and variant 0 with the generic (non-TMA) kernel."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from poms_b200 import _lib, bsplines as bs
from poms_b200.stencil import (StencilVectorSpace, StencilVector, KronSumMatrix, DeviceContext,
                               EPI_STORE, EPI_RESID, EPI_AXPY)

N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
variants = [int(v) for v in (sys.argv[2] if len(sys.argv) > 2 else "0,1").split(",")]
p = 3
dev = torch.device("cuda", 0)
knots = [bs.make_open_knots(p, N + p)] * 3
A = KronSumMatrix.poisson(p, knots)
q = max(2 * p - 1, 1)
glt = [bs.glt_band(p, n, degree=q) for n in A.npts]
F = [bs.poly_inverse_factors(b_, 3) for b_ in glt]
S1 = KronSumMatrix([f[0] for f in F])
S2 = KronSumMatrix([f[1] for f in F])
gp = max(p, S2.P)
V = StencilVectorSpace([N + p] * 3, [gp, p, p], [False] * 3, device=dev)
x, b = StencilVector(V), StencilVector(V)
g = torch.Generator(device=dev).manual_seed(0)
x.data.copy_(torch.randn(V.npts, generator=g, dtype=torch.float64, device=dev))
b.data.copy_(torch.randn(V.npts, generator=g, dtype=torch.float64, device=dev))
ctx = DeviceContext.get(dev)
L = _lib.lib()
dof = V.local_size
peak = 6416.7
cases = [("A p=3 sum STORE+dot", A, EPI_STORE, None, 16, True),
         ("A p=3 sum RESID", A, EPI_RESID, b, 24, False),
         ("S1 P=%d single STORE" % S1.P, S1, EPI_STORE, None, 16, False),
         ("S2 P=%d single AXPY(b=0)" % S2.P, S2, EPI_AXPY, None, 16, False),
         ("S2 P=%d single AXPY+b" % S2.P, S2, EPI_AXPY, b, 24, False)]
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float64, device=dev)   # 512 MB > L2
ref = {}
tot = {v: 0.0 for v in variants}
for name, op, epi, rhs, nb, dot in cases:
    for var in variants:
        L.poms_set_matvec3d_variant(var)
        y = StencilVector(V)
        dp = ctx.sptr(30) if dot else None
        for _ in range(3):
            op.apply(x, y, epi, b=rhs, omega=0.37, dot_ptr=dp)
        torch.cuda.synchronize()
        times = []
        for _ in range(10):
            flush.fill_(1.0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            op.apply(x, y, epi, b=rhs, omega=0.37, dot_ptr=dp)
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        ms = float(np.median(times))
        tot[var] += ms * {0: 1, 1: 2, 2: 2, 3: 1, 4: 1}[[c[0] for c in cases].index(name)]
        dv = ctx.scal[30].item() if dot else float("nan")
        if var == variants[0]:
            ref[name] = (y.data.clone(), dv)
            diff = 0.0
        else:
            diff = ((y.data - ref[name][0]).abs().max() / ref[name][0].abs().max()).item()
        print("%-26s var %d  %7.3f ms  %7.1f GB/s alg (%4.1f %% of %.0f)  min %.3f  relerr vs var%d %.1e  dot %.15e"
              % (name, var, ms, nb * dof / ms / 1e6, 100 * nb * dof / ms / 1e6 / peak, peak,
                 min(times), variants[0], diff, dv), flush=True)
# generic kernel cross-check of the first variant
L.poms_set_matvec3d_variant(variants[0])
for name, op, epi, rhs, nb, dot in cases[:2]:
    L.poms_set_force_generic(1)
    y = StencilVector(V)
    op.apply(x, y, epi, b=rhs, omega=0.37)
    L.poms_set_force_generic(0)
    diff = ((y.data - ref[name][0]).abs().max() / ref[name][0].abs().max()).item()
    print("%-26s generic vs var%d relerr %.1e" % (name, variants[0], diff))
print("per-iteration fine-level mix (3 A: 1 STORE + 2 RESID, 2 S1, S2 zero-guess + S2 with rhs):")
L.poms_set_matvec3d_variant(1)
for var in variants:
    print("   variant %d: %.3f ms per PCG iteration on the fine level" % (var, tot[var]))
