"""Timing of the banded Kronecker solve kernels at a given size (run on the GPU box):
    python tests/gpu_ab_bandsolve.py [N] [p]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from poms_b200 import bsplines as bs
from poms_b200.stencil import StencilVectorSpace, StencilVector
from poms_b200.kron_product import BandLU, _solve_axis

N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
p = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda", 0)
n = N + p
V = StencilVectorSpace([n] * 3, [p] * 3, [False] * 3, device=dev)
y, x = StencilVector(V), StencilVector(V)
g = torch.Generator(device=dev).manual_seed(0)
y.data.copy_(torch.randn(V.npts, generator=g, dtype=torch.float64, device=dev))
lu = BandLU.from_band(bs.glt_band(p, n, degree=2 * p - 1), dev)
for ax in range(3):
    for _ in range(2):
        _solve_axis(lu, y, x, ax)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        _solve_axis(lu, y, x, ax)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print("axis %d: %.3f ms  %.0f GB/s moved (32 B/DOF)  %.0f GB/s algorithmic (16 B/DOF)"
          % (ax + 1, ms, 32 * V.local_size / ms / 1e6, 16 * V.local_size / ms / 1e6))
